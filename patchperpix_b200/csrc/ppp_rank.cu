// Step 2: patch ranking (rankPatches.cu:1-161) and the stable descending
// sort of ranked_patches.py:21-30.
//
// One CTA per patch centre.  The class-folded patch row gives the list of
// voting pixels (high: D > 0, background: D < 0; the gate of the pixel is
// already folded in).  Every unordered pixel pair {p1 < p2} with at least one
// high pixel reads the consensus slot (base p1, offset p2 - p1):
//     both high          acc += cons          (or +-1 with COUNT_POS_NEG)
//     high / background  acc -= cons
// fgCnt of rankPatches.cu:139 has the closed form
//     nH * nG - nH - nH (nH - 1) / 2
// (nH high pixels, nG gated pixels in the window).  Accumulation is in double
// in a fixed order, so the score is deterministic.
#include <cub/cub.cuh>
#include "ppp_common.cuh"
#include "ppp_api.cuh"

__global__ void rank_fill_kernel(ppp_cfg cfg, float* __restrict__ score)
{
    Geo g = make_geo(cfg);
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.V) return;
    int z, y, x;
    vox_decode(g, (int)v, z, y, x);
    bool interior = x >= g.rx && x < g.X - g.rx && y >= g.ry && y < g.Y - g.ry &&
                    z >= g.rz && z < g.Z - g.rz;
    // rankPatches.cu:152-160
    score[v] = interior ? 0.0f : ((cfg.rank_flags & 1) ? -1.0f : -9999999.0f);
}

#define RANK_THREADS 256

__global__ void __launch_bounds__(RANK_THREADS)
rank_kernel(const float* __restrict__ dp, const uint8_t* __restrict__ flags,
            const int32_t* __restrict__ fgidx, const int32_t* __restrict__ rowvox,
            const float* __restrict__ cons, int64_t F, ppp_cfg cfg, float* __restrict__ score)
{
    Geo g = make_geo(cfg);
    extern __shared__ unsigned char smem_raw[];
    // per voting pixel: lin position, consensus row, class sign
    int32_t* s_lin = (int32_t*)smem_raw;             // [P]
    int32_t* s_row = s_lin + g.P;                    // [P]
    int8_t* s_sgn = (int8_t*)(s_row + g.P);          // [P]
    __shared__ int s_n, s_nH, s_nG;
    __shared__ int s_wcnt[RANK_THREADS / 32];
    __shared__ double s_red[RANK_THREADS / 32];

    const int64_t row = blockIdx.x;
    const int vc = rowvox[row];
    if (!(flags[vc] & PPP_FLAG_CENTRE)) return;     // keeps the fill value
    int cz, cy, cx;
    vox_decode(g, vc, cz, cy, cx);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_n = 0; s_nH = 0; s_nG = 0; }
    __syncthreads();

    // ordered compaction of the voting pixels (po order == lexicographic order)
    int nH = 0, nG = 0;
    for (int base = 0; base < g.P; base += RANK_THREADS) {
        int po = base + threadIdx.x;
        float d = 0.0f;
        int qz = 0, qy = 0, qx = 0, pv = 0;
        bool gated = false;
        if (po < g.P) {
            d = dp[dp_index(g, F, row, po)];
            po_decode(g, po, qz, qy, qx);
            pv = ((cz + qz - g.rz) * g.Y + (cy + qy - g.ry)) * g.X + (cx + qx - g.rx);
            gated = (flags[pv] & PPP_FLAG_GATED) != 0;
        }
        bool vote = d != 0.0f;
        unsigned bal = __ballot_sync(0xffffffffu, vote);
        nH += (d > 0.0f);
        nG += gated;
        if (lane == 0) s_wcnt[w] = __popc(bal);
        __syncthreads();
        int off = s_n;
        for (int i = 0; i < w; i++) off += s_wcnt[i];
        if (vote) {
            int idx = off + __popc(bal & ((1u << lane) - 1));
            s_lin[idx] = po_lin(g, qz, qy, qx);
            s_row[idx] = fgidx[pv];
            s_sgn[idx] = d > 0.0f ? 1 : -1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int i = 0; i < RANK_THREADS / 32; i++) t += s_wcnt[i];
            s_n += t;
        }
        __syncthreads();
    }
    nH = warp_sum_i(nH);
    nG = warp_sum_i(nG);
    if (lane == 0) { atomicAdd(&s_nH, nH); atomicAdd(&s_nG, nG); }
    __syncthreads();
    const int n = s_n;
    const bool count_mode = (cfg.rank_flags & 2) != 0;

    double acc = 0.0;
    for (int i = 0; i < n; i++) {
        const int li = s_lin[i], si = s_sgn[i];
        const float* crow = cons + (int64_t)s_row[i] * g.K - li - 1;
        for (int j = i + 1 + threadIdx.x; j < n; j += RANK_THREADS) {
            int sj = s_sgn[j];
            if (si < 0 && sj < 0) continue;
            float v3 = crow[s_lin[j]];
            float c;
            if (count_mode) c = (v3 != 0.0f) ? copysignf(1.0f, v3) : ((si > 0 && sj > 0) ? -1.0f : 1.0f);
            else c = v3;
            // both high: += ; mixed: -=   (count mode, v3 == 0: always -1)
            acc += (si > 0 && sj > 0) ? (double)c : -(double)c;
        }
    }
    acc = warp_sum_d(acc);
    if (lane == 0) s_red[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < RANK_THREADS / 32; i++) t += s_red[i];
        unsigned h = (unsigned)s_nH, gg = (unsigned)s_nG;
        unsigned fgCnt = h * gg - h - (h * (h - 1)) / 2;
        float a = (float)t;
        score[vc] = (cfg.rank_flags & 1) ? a / (float)(fgCnt > 1 ? fgCnt : 1) : a;
    }
}

extern "C" int ppp_rank(const float* dp, const uint8_t* flags, const int32_t* fgidx,
                        const int32_t* rowvox, int64_t F, const float* cons,
                        const ppp_cfg* cfg, float* score, void* stream)
{
    Geo g = make_geo(*cfg);
    cudaStream_t s = (cudaStream_t)stream;
    rank_fill_kernel<<<(unsigned)((g.V + 255) / 256), 256, 0, s>>>(*cfg, score);
    if (F > 0) {
        size_t smem = (size_t)g.P * 9 + 16;
        rank_kernel<<<(unsigned)F, RANK_THREADS, smem, s>>>(dp, flags, fgidx, rowvox,
                                                            cons, F, *cfg, score);
    }
    return ppp_check("ppp_rank");
}

// ---------------------------------------------------------------------------
// stable descending sort: 64-bit key = (~orderable(score) << 32) | position
// ---------------------------------------------------------------------------
__global__ void rank_keys_kernel(const float* __restrict__ score,
                                 const int32_t* __restrict__ cand, int64_t n,
                                 uint64_t* __restrict__ keys)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = score[cand[i]];
    if (s == 0.0f) s = 0.0f;                    // -0.0 == +0.0 for python's sort
    uint32_t u = __float_as_uint(s);
    u ^= (u >> 31) ? 0xffffffffu : 0x80000000u; // ascending-orderable
    keys[i] = ((uint64_t)(~u) << 32) | (uint32_t)i;
}

__global__ void rank_order_kernel(const uint64_t* __restrict__ keys,
                                  const int32_t* __restrict__ cand, int64_t n,
                                  int32_t* __restrict__ order)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    order[i] = cand[(uint32_t)(keys[i] & 0xffffffffu)];
}

static size_t sort_temp_bytes(int64_t n)
{
    size_t tb = 0;
    cub::DeviceRadixSort::SortKeys((void*)nullptr, tb, (const uint64_t*)nullptr,
                                   (uint64_t*)nullptr, (int)n);
    return tb;
}

extern "C" int64_t ppp_rank_sort_scratch_bytes(int64_t n)
{
    if (n <= 0) return 256;
    size_t tb = sort_temp_bytes(n);
    return (int64_t)(((tb + 255) / 256) * 256 + 2 * ((n * 8 + 255) / 256) * 256);
}

extern "C" int ppp_rank_sort(const float* score, const int32_t* cand, int64_t n,
                             int32_t* order, void* scratch, void* stream)
{
    if (n <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    size_t tb = sort_temp_bytes(n);
    size_t tb_al = ((tb + 255) / 256) * 256;
    size_t kb = ((n * 8 + 255) / 256) * 256;
    char* base = (char*)scratch;
    uint64_t* k_in = (uint64_t*)(base + tb_al);
    uint64_t* k_out = (uint64_t*)(base + tb_al + kb);
    unsigned nb = (unsigned)((n + 255) / 256);
    rank_keys_kernel<<<nb, 256, 0, s>>>(score, cand, n, k_in);
    cub::DeviceRadixSort::SortKeys(scratch, tb, k_in, k_out, (int)n, 0, 64, s);
    rank_order_kernel<<<nb, 256, 0, s>>>(k_out, cand, n, order);
    return ppp_check("ppp_rank_sort");
}
