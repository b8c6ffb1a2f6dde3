// TEST INFRASTRUCTURE ONLY (oracle/): host shim that lets the UNMODIFIED
// reference kernels (/root/reference/PatchPerPix/vote_instances/cuda/*.cu,
// read where they lie, never copied) compile as plain serial C++.
// SURVEY.md Appendix D documents the recipe.  Built by oracle/build_ref.py /
// oracle/ref_runner.py into oracle/_ref/ (git-ignored, travels with gpurun).
#pragma once
#include <cstdint>
#include <cmath>
#include <cstdlib>
#define __device__
#define __global__
struct ppp_dim3_ { unsigned x, y, z; };
static thread_local ppp_dim3_ blockIdx, blockDim, threadIdx;
#ifdef PPP_SHIM_ATOMIC
// multi-threaded baseline timing: a real atomic float add (CAS loop)
static inline void atomicAdd(float* p, float v) {
    uint32_t* ip = reinterpret_cast<uint32_t*>(p);
    uint32_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), nw;
    do {
        float f; __builtin_memcpy(&f, &old, 4); f += v;
        __builtin_memcpy(&nw, &f, 4);
    } while (!__atomic_compare_exchange_n(ip, &old, nw, true, __ATOMIC_RELAXED,
                                          __ATOMIC_RELAXED));
}
static inline void atomicAdd(float* p, int v) { atomicAdd(p, (float)v); }
#else
template <class T, class U> static inline void atomicAdd(T* p, U v) { *p += (T)v; }
#endif
// rankPatches.cu:147, computePatchGraph.cu:132 call max(int, unsigned)
static inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
