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

extern "C" const char* ppp_last_error(void) { return g_err; }
extern "C" int ppp_version(void) { return 100; }
