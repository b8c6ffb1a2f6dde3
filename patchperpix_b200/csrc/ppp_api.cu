// C-ABI plumbing: last-error string, version.
#include <cstdio>
#include <cstring>
#include "ppp_api.cuh"
#include "../../include/ppp_b200.h"

static thread_local char g_err[512] = "";

int ppp_fail(int code, const char* msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

int ppp_check(const char* where)
{
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}

// ---------------------------------------------------------------------------
// launch counter: every kernel this library launches (its own kernels and the CUB
// kernels compiled into it) goes through the runtime's cudaLaunchKernel; the library
// is linked with -Bsymbolic and the definition is hidden, so those calls (and only
// those) bind to it; it counts and
// forwards to the real runtime entry.  bench.py reports the count as `gpu_launches`.
// ---------------------------------------------------------------------------
#include <atomic>
#include <dlfcn.h>
static std::atomic<long long> g_launches{0};

extern "C" __attribute__((visibility("hidden")))
cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** args,
                             size_t smem, cudaStream_t stream)
{
    typedef cudaError_t (*fn_t)(const void*, dim3, dim3, void**, size_t, cudaStream_t);
    static fn_t real = (fn_t)dlsym(RTLD_NEXT, "cudaLaunchKernel");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return real(func, grid, block, args, smem, stream);
}

extern "C" int64_t ppp_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" const char* ppp_last_error(void) { return g_err; }
extern "C" int ppp_version(void) { return 100; }
