// Step 1: consensus voting in GATHER form.
//
// The reference scatters: one thread per patch centre walks all P x P pixel
// pairs of its window and float-atomicAdds a vote into the dense
// [NSZ][NSY][NSX][Z][Y][X] array (fillConsensusArray.cu:36-173), then a second
// kernel divides by the vote counter (normConsensusArray.cu:19-26).
// Here every output slot (base voxel b, positive offset o) has exactly ONE
// writer, which walks the centres c whose window contains both b and b+o:
//     pos  = #{c : high_c(b)  and high_c(b+o)}
//     neg  = #{c : exactly one of them high, the other background}
//     sum  = SUM_c  D_c(b) * D_c(b+o)            (D = class-folded patch, see
//            minus the both-background products    include/ppp_b200.h)
// so the counters are integers, the float sum has a fixed order (deterministic,
// unlike the reference's atomics) and normalisation is fused into the epilogue.
#include <cuda.h>
#include "ppp_common.cuh"
#include "ppp_api.cuh"

__device__ __forceinline__ float consensus_epilogue(const ppp_cfg& cfg, float sum,
                                                    int pos, int neg)
{
    // fillConsensusArray.cu:104-109,127-133 summed over the votes of one slot
    double s;
    if (cfg.prod_mode == 2) s = ((double)sum - cfg.th2 * (double)(pos - neg)) / cfg.one_m_th2;
    else if (cfg.prod_mode == 1) s = (double)sum;
    else s = (double)(pos - neg);
    int cnt = pos + neg;
    if (cfg.norm_aff && cnt != 0) s = s / (double)cnt;   // normConsensusArray.cu:22-23
    return (float)s;
}

// ---------------------------------------------------------------------------
// v1 "naive": one CTA per base row, threads over offsets, global gathers.
// Kept as the simple cross-check for the tiled kernels.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
consensus_naive_kernel(const float* __restrict__ dp, const uint8_t* __restrict__ flags,
                       const int32_t* __restrict__ fgidx, const int32_t* __restrict__ rowvox,
                       int64_t F, ppp_cfg cfg, float* __restrict__ cons,
                       uint32_t* __restrict__ cnt)
{
    Geo g = make_geo(cfg);
    const int64_t row = blockIdx.x;
    const int vb = rowvox[row];
    int bz, by, bx;
    vox_decode(g, vb, bz, by, bx);
    const bool gated = (flags[vb] & PPP_FLAG_GATED) != 0;
    for (int k = threadIdx.x; k < g.K; k += blockDim.x) {
        float out = 0.0f;
        uint32_t outc = 0;
        int lin = k + g.K + 1;
        int ox = lin % g.nx - (g.psx - 1);
        int t = lin / g.nx;
        int oy = t % g.ny - (g.psy - 1);
        int oz = t / g.ny - (g.psz - 1);
        int pz = bz + oz, py = by + oy, px = bx + ox;
        if (gated && pz >= 0 && pz < g.Z && py >= 0 && py < g.Y && px >= 0 && px < g.X) {
            int vp = (pz * g.Y + py) * g.X + px;
            if (flags[vp] & PPP_FLAG_GATED) {
                int pos = 0, neg = 0;
                float sum = 0.0f;
                int cz0 = max(max(bz, pz) - g.rz, g.rz), cz1 = min(min(bz, pz) + g.rz, g.Z - 1 - g.rz);
                int cy0 = max(max(by, py) - g.ry, g.ry), cy1 = min(min(by, py) + g.ry, g.Y - 1 - g.ry);
                int cx0 = max(max(bx, px) - g.rx, g.rx), cx1 = min(min(bx, px) + g.rx, g.X - 1 - g.rx);
                for (int cz = cz0; cz <= cz1; cz++)
                for (int cy = cy0; cy <= cy1; cy++)
                for (int cx = cx0; cx <= cx1; cx++) {
                    int vc = (cz * g.Y + cy) * g.X + cx;
                    int rc = fgidx[vc];
                    if (rc < 0) continue;       // not fg => not a centre (interior by the clamps)
                    int po1 = ((bz - cz + g.rz) * g.psy + (by - cy + g.ry)) * g.psx + (bx - cx + g.rx);
                    int po2 = ((pz - cz + g.rz) * g.psy + (py - cy + g.ry)) * g.psx + (px - cx + g.rx);
                    float d1 = dp[dp_index(g, F, rc, po1)];
                    float d2 = dp[dp_index(g, F, rc, po2)];
                    bool h1 = d1 > 0.0f, h2 = d2 > 0.0f, l1 = d1 < 0.0f, l2 = d2 < 0.0f;
                    pos += (h1 && h2);
                    neg += (h1 && l2) || (l1 && h2);
                    if (!(l1 && l2)) sum = fmaf(d1, d2, sum);
                }
                out = consensus_epilogue(cfg, sum, pos, neg);
                outc = ((uint32_t)neg << 16) | (uint32_t)pos;
            }
        }
        cons[row * g.K + k] = out;
        if (cnt != nullptr) cnt[row * g.K + k] = outc;
    }
}

// ---------------------------------------------------------------------------
// vote counters from the "received" class bits (ppp_prep.cu): for the slot
// (b, o) and every centre line (dz,dy) shared by the two windows,
//   pos += popc(H_b & H_b'>>o),  neg += popc(H_b & L_b'>>o) + popc(L_b & H_b'>>o)
// where the partner's word is shifted by ox so that equal bit positions mean
// the same centre.  One thread per slot; also zero-fills `cons`, so the sum
// kernel below only has to touch slots that can be non-zero.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
consensus_count_kernel(const unsigned long long* __restrict__ rbits,
                       const uint8_t* __restrict__ flags, const int32_t* __restrict__ fgidx,
                       const int32_t* __restrict__ rowvox, int64_t F, ppp_cfg cfg,
                       float* __restrict__ cons, uint32_t* __restrict__ cnt)
{
    Geo g = make_geo(cfg);
    extern __shared__ __align__(16) unsigned char cc_smem[];
    const int nrw = g.psz * g.psy;
    unsigned long long* s_rb = (unsigned long long*)cc_smem;     // [nrw][2] bits of this row
    int32_t* s_prow = (int32_t*)(s_rb + 2 * nrw);                // [K] partner row of a live slot
    uint16_t* s_k = (uint16_t*)(s_prow + g.K);                   // [K] its slot index
    __shared__ int s_n;
    const int64_t row = blockIdx.x;
    const int vb = rowvox[row];
    int bz, by, bx;
    vox_decode(g, vb, bz, by, bx);
    const bool gated = (flags[vb] & PPP_FLAG_GATED) != 0;
    for (int i = threadIdx.x; i < 2 * nrw; i += blockDim.x)      // rbits is [line][row][2]
        s_rb[i] = rbits[((int64_t)(i >> 1) * F + row) * 2 + (i & 1)];
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // pass 1: slots whose partner is gated are compacted (dense warps in pass 2), all
    // others are written as zeros right away
    const int lane = threadIdx.x & 31;
    for (int k0 = 0; k0 < g.K; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        int prow = -1;
        if (k < g.K && gated) {
            int lin = k + g.K + 1;
            int ox = lin % g.nx - (g.psx - 1);
            int t = lin / g.nx;
            int oy = t % g.ny - (g.psy - 1);
            int oz = t / g.ny - (g.psz - 1);
            int pz = bz + oz, py = by + oy, px = bx + ox;
            if (pz >= 0 && pz < g.Z && py >= 0 && py < g.Y && px >= 0 && px < g.X) {
                int vp = (pz * g.Y + py) * g.X + px;
                if (flags[vp] & PPP_FLAG_GATED) prow = fgidx[vp];
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, prow >= 0);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&s_n, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (prow >= 0) {
            int idx = base + __popc(bal & ((1u << lane) - 1u));
            s_prow[idx] = prow;
            s_k[idx] = (uint16_t)k;
        } else if (k < g.K) {
            cnt[row * g.K + k] = 0;
            cons[row * g.K + k] = 0.0f;
        }
    }
    __syncthreads();
    // pass 2: counters of the live slots
    const int n = s_n;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int k = s_k[idx];
        const unsigned long long* rp = rbits + (int64_t)s_prow[idx] * 2;   // + line * 2F
        int lin = k + g.K + 1;
        int ox = lin % g.nx - (g.psx - 1);
        int t = lin / g.nx;
        int oy = t % g.ny - (g.psy - 1);
        int oz = t / g.ny - (g.psz - 1);
        int pos = 0, neg = 0;
        // centre offsets d (from b) with d - o inside the partner's window
        int dz0 = max(-g.rz, oz - g.rz), dz1 = min(g.rz, oz + g.rz);
        int dy0 = max(-g.ry, oy - g.ry), dy1 = min(g.ry, oy + g.ry);
        for (int dz = dz0; dz <= dz1; dz++)
        for (int dy = dy0; dy <= dy1; dy++) {
            int w1 = (dz + g.rz) * g.psy + (dy + g.ry);
            int w2 = (dz - oz + g.rz) * g.psy + (dy - oy + g.ry);
            unsigned long long h1 = s_rb[2 * w1], l1 = s_rb[2 * w1 + 1];
            const ulonglong2 hl = *(const ulonglong2*)(rp + (int64_t)w2 * 2 * F);
            unsigned long long h2 = hl.x, l2 = hl.y;
            if (ox >= 0) { h2 <<= ox; l2 <<= ox; } else { h2 >>= -ox; l2 >>= -ox; }
            pos += __popcll(h1 & h2);
            neg += __popcll(h1 & l2) + __popcll(l1 & h2);
        }
        const uint32_t outc = ((uint32_t)neg << 16) | (uint32_t)pos;
        cnt[row * g.K + k] = outc;
        // plain vote counter (no probability product): the counters are the result
        cons[row * g.K + k] = (cfg.prod_mode == 0 && outc)
            ? consensus_epilogue(cfg, 0.0f, (int)(outc & 0xffffu), (int)(outc >> 16)) : 0.0f;
    }
}

// ---------------------------------------------------------------------------
// small patches (psx < 16, e.g. the 7^3 patches of the flylight setup on sparse
// 3-D data): counters AND sums in one kernel, one thread per slot.  The
// received class bits say exactly which centres vote for the slot, so the
// thread only visits those (a handful on thin neurites) instead of the whole
// window intersection; the slot is then written once, coalesced.  Centres are
// visited in raster order: bit-identical to the simple kernel.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
consensus_bits_kernel(const float* __restrict__ dp, const unsigned long long* __restrict__ rbits,
                      const uint8_t* __restrict__ flags, const int32_t* __restrict__ fgidx,
                      const int32_t* __restrict__ rowvox, int64_t F, ppp_cfg cfg,
                      float* __restrict__ cons, uint32_t* __restrict__ cnt)
{
    Geo g = make_geo(cfg);
    extern __shared__ __align__(16) unsigned char cb_smem[];
    const int nrw = g.psz * g.psy;
    unsigned long long* s_rb = (unsigned long long*)cb_smem;     // [nrw][2] bits of this row
    int32_t* s_prow = (int32_t*)(s_rb + 2 * nrw);                // [K] partner row of a live slot
    uint16_t* s_k = (uint16_t*)(s_prow + g.K);                   // [K] its slot index
    __shared__ int s_n;
    const int64_t row = blockIdx.x;
    const int vb = rowvox[row];
    int bz, by, bx;
    vox_decode(g, vb, bz, by, bx);
    const bool gated = (flags[vb] & PPP_FLAG_GATED) != 0;
    for (int i = threadIdx.x; i < 2 * nrw; i += blockDim.x)      // rbits is [line][row][2]
        s_rb[i] = rbits[((int64_t)(i >> 1) * F + row) * 2 + (i & 1)];
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // pass 1: slots with a gated partner are compacted, the others written as zeros
    const int lane = threadIdx.x & 31;
    for (int k0 = 0; k0 < g.K; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        int prow = -1;
        if (k < g.K && gated) {
            int lin = k + g.K + 1;
            int ox = lin % g.nx - (g.psx - 1);
            int t = lin / g.nx;
            int oy = t % g.ny - (g.psy - 1);
            int oz = t / g.ny - (g.psz - 1);
            int pz = bz + oz, py = by + oy, px = bx + ox;
            if (pz >= 0 && pz < g.Z && py >= 0 && py < g.Y && px >= 0 && px < g.X) {
                int vp = (pz * g.Y + py) * g.X + px;
                if (flags[vp] & PPP_FLAG_GATED) prow = fgidx[vp];
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, prow >= 0);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&s_n, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (prow >= 0) {
            int idx = base + __popc(bal & ((1u << lane) - 1u));
            s_prow[idx] = prow;
            s_k[idx] = (uint16_t)k;
        } else if (k < g.K) {
            cnt[row * g.K + k] = 0;
            cons[row * g.K + k] = 0.0f;
        }
    }
    __syncthreads();
    // pass 2: counters from the class bits, sums over the centres that vote
    const int n = s_n;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int k = s_k[idx];
        const unsigned long long* rp = rbits + (int64_t)s_prow[idx] * 2;   // + line * 2F
        int lin = k + g.K + 1;
        int ox = lin % g.nx - (g.psx - 1);
        int t = lin / g.nx;
        int oy = t % g.ny - (g.psy - 1);
        int oz = t / g.ny - (g.psz - 1);
        const int px = bx + ox;
        int pos = 0, neg = 0;
        float sum = 0.0f;
        int dz0 = max(-g.rz, oz - g.rz), dz1 = min(g.rz, oz + g.rz);
        int dy0 = max(-g.ry, oy - g.ry), dy1 = min(g.ry, oy + g.ry);
        for (int dz = dz0; dz <= dz1; dz++)
        for (int dy = dy0; dy <= dy1; dy++) {
            int w1 = (dz + g.rz) * g.psy + (dy + g.ry);
            int w2 = (dz - oz + g.rz) * g.psy + (dy - oy + g.ry);
            unsigned long long h1 = s_rb[2 * w1], l1 = s_rb[2 * w1 + 1];
            const ulonglong2 hl = *(const ulonglong2*)(rp + (int64_t)w2 * 2 * F);
            unsigned long long h2 = hl.x, l2 = hl.y;
            if (ox >= 0) { h2 <<= ox; l2 <<= ox; } else { h2 >>= -ox; l2 >>= -ox; }
            pos += __popcll(h1 & h2);
            neg += __popcll(h1 & l2) + __popcll(l1 & h2);
            unsigned long long m = (h1 & (h2 | l2)) | (l1 & h2);   // centres that vote
            if (!m || cfg.prod_mode == 0) continue;
            const int64_t cline = ((int64_t)(bz + dz) * g.Y + (by + dy)) * g.X;
            // patch rows that talk about b and about b + o, seen from this centre line
            const int64_t pr1 = (int64_t)((g.rz - dz) * g.psy + (g.ry - dy)) * F;
            const int64_t pr2 = (int64_t)((g.rz - dz + oz) * g.psy + (g.ry - dy + oy)) * F;
            while (m) {
                int tb = __ffsll((long long)m) - 1;
                m &= m - 1;
                int cx = bx - g.rx + tb;
                int64_t rc = fgidx[cline + cx];
                float d1 = dp[(pr1 + rc) * g.rsg + DP_GUARD + (bx - cx + g.rx)];
                float d2 = dp[(pr2 + rc) * g.rsg + DP_GUARD + (px - cx + g.rx)];
                sum = fmaf(d1, d2, sum);
            }
        }
        cons[row * g.K + k] = consensus_epilogue(cfg, sum, pos, neg);
        cnt[row * g.K + k] = ((uint32_t)neg << 16) | (uint32_t)pos;
    }
}

// ---------------------------------------------------------------------------
// small patches from the RECEIVED tables (ppp_received): same slot-per-thread
// walk as consensus_bits_kernel (same centre order, bit-identical results), but
// the partner's class bits are ONE contiguous 16-bit row (the 49 words of a 7^3
// window = 98 bytes) and the two values of a voting centre are rv[b][d] (own
// row, staged in shared memory) and rv[b'][d - o] (one load from the partner's
// contiguous row): no fgidx lookups, no dependent gathers into the centre-major
// patch array.  `need` (u8 [F], may be NULL): rows with 0 are skipped altogether
// (face regions of the blockwise stitcher only read the rows of their candidates).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
consensus_small_kernel(const float* __restrict__ rv, const uint16_t* __restrict__ rb16, int rbw,
                       const uint8_t* __restrict__ flags, const int32_t* __restrict__ fgidx,
                       const int32_t* __restrict__ rowvox, const uint8_t* __restrict__ need,
                       int64_t F, ppp_cfg cfg, float* __restrict__ cons,
                       uint32_t* __restrict__ cnt)
{
    Geo g = make_geo(cfg);
    extern __shared__ __align__(16) unsigned char cs_smem[];
    const int nrw = g.psz * g.psy;
    float* s_rv = (float*)cs_smem;                               // [P] own received values
    int32_t* s_prow = (int32_t*)(s_rv + ((g.P + 3) & ~3));       // [K] partner row of a live slot
    uint16_t* s_k = (uint16_t*)(s_prow + g.K);                   // [K] its slot index
    uint16_t* s_rb = s_k + g.K;                                  // [nrw] own class bits
    __shared__ int s_n;
    const int64_t row = blockIdx.x;
    if (need != nullptr && !need[row]) return;
    const int vb = rowvox[row];
    int bz, by, bx;
    vox_decode(g, vb, bz, by, bx);
    const bool gated = (flags[vb] & PPP_FLAG_GATED) != 0;
    if (gated) {
        for (int i = threadIdx.x; i < g.P; i += blockDim.x) s_rv[i] = rv[row * g.P + i];
        for (int i = threadIdx.x; i < nrw; i += blockDim.x) s_rb[i] = rb16[row * rbw + i];
    }
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int k0 = 0; k0 < g.K; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        int prow = -1;
        if (k < g.K && gated) {
            int lin = k + g.K + 1;
            int ox = lin % g.nx - (g.psx - 1);
            int t = lin / g.nx;
            int oy = t % g.ny - (g.psy - 1);
            int oz = t / g.ny - (g.psz - 1);
            int pz = bz + oz, py = by + oy, px = bx + ox;
            if (pz >= 0 && pz < g.Z && py >= 0 && py < g.Y && px >= 0 && px < g.X) {
                int vp = (pz * g.Y + py) * g.X + px;
                if (flags[vp] & PPP_FLAG_GATED) prow = fgidx[vp];
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, prow >= 0);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&s_n, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (prow >= 0) {
            int idx = base + __popc(bal & ((1u << lane) - 1u));
            s_prow[idx] = prow;
            s_k[idx] = (uint16_t)k;
        } else if (k < g.K) {
            cnt[row * g.K + k] = 0;
            cons[row * g.K + k] = 0.0f;
        }
    }
    __syncthreads();
    const int n = s_n;
    const unsigned xmask = (1u << g.psx) - 1u;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int k = s_k[idx];
        const int64_t pr = s_prow[idx];
        const uint16_t* __restrict__ rbp = rb16 + pr * rbw;
        const float* __restrict__ rvp = rv + pr * g.P;
        int lin = k + g.K + 1;
        int ox = lin % g.nx - (g.psx - 1);
        int t = lin / g.nx;
        int oy = t % g.ny - (g.psy - 1);
        int oz = t / g.ny - (g.psz - 1);
        int pos = 0, neg = 0;
        float sum = 0.0f;
        int dz0 = max(-g.rz, oz - g.rz), dz1 = min(g.rz, oz + g.rz);
        int dy0 = max(-g.ry, oy - g.ry), dy1 = min(g.ry, oy + g.ry);
        for (int dz = dz0; dz <= dz1; dz++)
        for (int dy = dy0; dy <= dy1; dy++) {
            const int w1 = (dz + g.rz) * g.psy + (dy + g.ry);
            const int w2 = (dz - oz + g.rz) * g.psy + (dy - oy + g.ry);
            const unsigned a = s_rb[w1], b = __ldg(rbp + w2);
            const unsigned h1 = a & 0xffu, l1 = a >> 8;
            unsigned h2 = b & 0xffu, l2 = b >> 8;
            if (ox >= 0) { h2 = (h2 << ox) & xmask; l2 = (l2 << ox) & xmask; }
            else { h2 >>= -ox; l2 >>= -ox; }
            pos += __popc(h1 & h2);
            neg += __popc(h1 & l2) + __popc(l1 & h2);
            unsigned m = (h1 & (h2 | l2)) | (l1 & h2);           // centres that vote
            if (!m || cfg.prod_mode == 0) continue;
            const float* __restrict__ r1 = s_rv + w1 * g.psx;
            const float* __restrict__ r2 = rvp + w2 * g.psx - ox;
            // all partner values of the line first (independent, predicated loads in
            // flight together), then the multiply-adds in centre order
            float p2v[8];
#pragma unroll
            for (int tb = 0; tb < 8; tb++) p2v[tb] = ((m >> tb) & 1u) ? __ldg(r2 + tb) : 0.0f;
#pragma unroll
            for (int tb = 0; tb < 8; tb++)
                if ((m >> tb) & 1u) sum = fmaf(r1[tb], p2v[tb], sum);
        }
        cons[row * g.K + k] = consensus_epilogue(cfg, sum, pos, neg);
        cnt[row * g.K + k] = ((uint32_t)neg << 16) | (uint32_t)pos;
    }
}

extern "C" int ppp_consensus_small(const float* rv, const uint16_t* rb16, const uint8_t* flags,
                                   const int32_t* fgidx, const int32_t* rowvox,
                                   const uint8_t* need, int64_t F, const ppp_cfg* cfg,
                                   float* cons, uint32_t* cnt, void* stream)
{
    if (F <= 0) return 0;
    Geo g = make_geo(*cfg);
    if (g.psx > 8 || g.psz * g.psy > 64)
        return ppp_fail(-1, "ppp_consensus_small: window too large (psx <= 8, psz*psy <= 64)");
    if (g.K > 65535) return ppp_fail(-1, "ppp_consensus_small: more than 65535 offsets");
    const int rbw = ((g.psz * g.psy + 7) / 8) * 8;
    size_t sm = (size_t)((g.P + 3) & ~3) * 4 + (size_t)g.K * 6 + (size_t)g.psz * g.psy * 2 + 16;
    cudaError_t e = cudaFuncSetAttribute(consensus_small_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return ppp_fail((int)e, "ppp_consensus_small: smem attribute");
    const int nthr = ((cfg->reserved >> 22) & 3) == 1 ? 64 : ((cfg->reserved >> 22) & 3) == 2 ? 96 : 128;   // (tuning)
    consensus_small_kernel<<<(unsigned)F, nthr, sm, (cudaStream_t)stream>>>(
        rv, rb16, rbw, flags, fgidx, rowvox, need, F, *cfg, cons, cnt);
    return ppp_check("ppp_consensus_small");
}

// ---------------------------------------------------------------------------
// sums.  CTA = (base line, group of NOY consecutive offset rows (oz,oy)).
//
// Rows of `dp` are in raster order, so the valid centres of a line are a
// contiguous row range (rows_before) and, dp being patch-row major, one patch
// x-row of NCCH consecutive centres is one contiguous block: it is fetched by
// ONE bulk copy on the TMA engine (cp.async.bulk) that signals an mbarrier,
// into a double buffer, so the copy of the next chunk overlaps the arithmetic
// on the current one and costs no LSU issue slots.
// The gated voxels of the base line and of every partner line are covered
// greedily by T-wide x-windows ("tiles"); a work item is a (base tile, partner
// tile) pair within reach, i.e. T*T accumulators in registers.  For every
// centre line around the base line the patch row that talks about the base
// line (A1) and the NOY rows that talk about the partner lines (A2) are
// staged, and every thread walks the staged centres that can reach both of
// its tiles:
//     acc[j][m] += D1[j] * max(D2[m],0) + max(D1[j],0) * min(D2[m],0)
// (D = class-folded patch value; pairs that are background on both sides add
// exact zeros: they do not vote).  The zero guards of the dp layout let a tile
// be read without range checks.  Every slot has one writer and a fixed
// summation order -> deterministic, and bit-identical to the simple kernel.
// ---------------------------------------------------------------------------
#define CT_T 8
// consumer threads per CTA (+ one producer warp), centres per stage buffer; tuned on the
// bench image together with CT_STAGES / CT_MINB below (two CTAs per SM)
#ifndef CT_THREADS
#define CT_THREADS 192
#endif
#ifndef CT_NCCH
#define CT_NCCH 40
#endif
#define CT_XMAX 2048
#define CT_MAXTILES 512
#ifndef CT_MAXITEMS
#define CT_MAXITEMS 6144
#endif
#define CT_MAXLINES 128
#define CT_MAXSTAGES 8

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(a), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" :: "r"(a), "r"(parity) : "memory");
}
// one contiguous block of dp -> shared memory through the TMA engine (1-D bulk
// copy, 16-byte aligned on both sides), completion signalled on `bar`
__device__ __forceinline__ void bulk_load(void* smem, const void* gmem, unsigned bytes,
                                          uint64_t* bar)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    unsigned ba = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
        :: "r"(sa), "l"(gmem), "r"(bytes), "r"(ba) : "memory");
}

// ---------------------------------------------------------------------------
// pre-pass: greedy cover of the gated voxels of every line by T-wide x-windows
// ("tiles"), once per line instead of once per (line, offset-row group) CTA.
// One warp per line; tiles i16 [lines][ts] (ts = ceil(X/T) + 1), ntiles i32 [lines].
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
consensus_tiles_kernel(const uint8_t* __restrict__ flags, ppp_cfg cfg, int ts,
                       int16_t* __restrict__ tiles, int32_t* __restrict__ ntiles)
{
    Geo g = make_geo(cfg);
    const int lane = threadIdx.x & 31;
    const int64_t line = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (line >= (int64_t)g.Z * g.Y) return;
    const uint8_t* fl = flags + line * g.X;
    int16_t* out = tiles + line * ts;
    int n = 0, covered = -1;                       // last x covered by a tile so far
    for (int x0 = 0; x0 < g.X; x0 += 32) {
        const int x = x0 + lane;
        unsigned m = __ballot_sync(0xffffffffu, x < g.X && (fl[x] & PPP_FLAG_GATED));
        while (m) {                                // uniform over the warp
            const int b = __ffs(m) - 1;
            const int xs = x0 + b;
            if (xs > covered) {
                if (lane == 0) out[n] = (int16_t)xs;
                n++;
                covered = xs + CT_T - 1;
            }
            // drop the bits this tile covers
            const int upto = covered - x0;         // last covered bit of this word
            m &= upto >= 31 ? 0u : ~((2u << upto) - 1u);
        }
    }
    if (lane == 0) ntiles[line] = n;
}

#ifdef CT_PROFILE
__device__ unsigned long long ct_prof[8];
extern "C" int ppp_debug_ct_prof(unsigned long long* out, int reset)
{
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(ct_prof, z, sizeof(z)); return 0; }
    return (int)cudaMemcpyFromSymbol(out, ct_prof, 8 * sizeof(unsigned long long));
}
#define CT_TICK(i) if (tid == 0) { long long now_ = clock64(); atomicAdd(&ct_prof[i], (unsigned long long)(now_ - t_prof)); t_prof = now_; }
#else
#define CT_TICK(i)
#endif

#ifndef CT_MINB
#define CT_MINB 2
#endif
template <int NOY>
__global__ void __launch_bounds__(CT_THREADS + 32, CT_MINB)
consensus_rows_kernel(const float* __restrict__ dp, const uint8_t* __restrict__ flags,
                      const int32_t* __restrict__ fgidx, const int32_t* __restrict__ rowvox,
                      int F, ppp_cfg cfg, const uint32_t* __restrict__ cnt,
                      float* __restrict__ cons, int S, const int16_t* __restrict__ tiles,
                      const int32_t* __restrict__ ntiles, int ts)
{
    constexpr int T = CT_T;
    Geo g = make_geo(cfg);
    const int RS = g.rsg;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sA = (float*)smem_raw;                                // [S][(1+NOY)][NCCH][RS]
    const int substride = CT_NCCH * RS;                          // one staged patch row
    const int bufstride = (1 + NOY) * substride;
    int32_t* s_ra = (int32_t*)(sA + S * bufstride);              // [MAXLINES] first row per centre line
    int32_t* s_rb = s_ra + CT_MAXLINES;                          // [MAXLINES]
    uint32_t* s_items = (uint32_t*)(s_rb + CT_MAXLINES);         // [MAXITEMS]
    int16_t* s_bt = (int16_t*)(s_items + CT_MAXITEMS);           // [MAXTILES] base tile starts
    int16_t* s_pt = s_bt + CT_MAXTILES;                          // [NOY][MAXTILES]
    int16_t* s_tmp = s_pt + NOY * CT_MAXTILES;                   // [XMAX]
    __shared__ int s_nitems;
    __shared__ __align__(8) uint64_t s_full[CT_MAXSTAGES], s_empty[CT_MAXSTAGES];
    __shared__ int2 s_info[CT_MAXSTAGES];                        // (centre line, centres) of a stage

    const int line = blockIdx.y;
    const int bz = line / g.Y, by = line % g.Y;
    const int64_t bline = (int64_t)line * g.X;
    const int nrows_off = (g.nz * g.ny - 1) / 2 + 1;             // offset rows >= (0,0)
    const int rlin_c = (g.nz * g.ny - 1) / 2;
    const int tid = threadIdx.x;
#ifdef CT_PROFILE
    long long t_prof = clock64();
#endif

    // ---- tiles of the base line (from the pre-pass) ---------------------------
    const int nbt = min(ntiles[line], CT_MAXTILES);
    if (nbt == 0) return;
    for (int i = tid; i < nbt; i += CT_THREADS + 32) s_bt[i] = tiles[(int64_t)line * ts + i];
    // ---- offset rows of this group, their partner lines and tiles --------------
    int oz_t[NOY], oy_t[NOY];
    int64_t pline_t[NOY];
    bool row_ok[NOY];
    int np_t[NOY];
#pragma unroll
    for (int t = 0; t < NOY; t++) {
        int orow = blockIdx.x * NOY + t;
        int rlin = rlin_c + orow;
        oz_t[t] = rlin / g.ny - (g.psz - 1);
        oy_t[t] = rlin % g.ny - (g.psy - 1);
        int pz = bz + oz_t[t], py = by + oy_t[t];
        row_ok[t] = orow < nrows_off && pz >= 0 && pz < g.Z && py >= 0 && py < g.Y;
        const int64_t pl = row_ok[t] ? (int64_t)pz * g.Y + py : 0;
        pline_t[t] = pl * g.X;
        np_t[t] = row_ok[t] ? min(ntiles[pl], CT_MAXTILES) : 0;
        for (int i = tid; i < np_t[t]; i += CT_THREADS + 32)
            s_pt[t * CT_MAXTILES + i] = tiles[pl * ts + i];
    }
    __syncthreads();
    // ---- work items, base-tile major so that a warp works on one x-neighbourhood:
    // every (base tile, offset row) counts its partner tiles within reach in parallel,
    // thread 0 turns the counts into offsets, then the items are written in parallel --
    int32_t* s_cntoff = (int32_t*)s_tmp;                         // [nbt * NOY + 1] (XMAX i16 = 1024 i32)
    const int nent = nbt * NOY;
    auto reach = [&](int e, int& lo, int& hi) {                  // partner tiles of entry e
        const int bt = e / NOY, t = e - bt * NOY;
        const int b0 = s_bt[bt], np = np_t[t];
        const int16_t* pt = s_pt + t * CT_MAXTILES;
        // partner windows [p0, p0+T) with some |p - b| < psx for b in [b0, b0+T)
        int a = 0, z = np;
        while (a < z) { int mid = (a + z) >> 1; if (pt[mid] + T - 1 < b0 - (g.psx - 1)) a = mid + 1; else z = mid; }
        lo = a;
        a = lo; z = np;
        while (a < z) { int mid = (a + z) >> 1; if (pt[mid] <= b0 + T - 1 + (g.psx - 1)) a = mid + 1; else z = mid; }
        hi = a;
        if ((blockIdx.x * NOY + t) == 0) {                       // offset row (0,0): only p > b is stored
            a = lo; z = hi;
            while (a < z) { int mid = (a + z) >> 1; if (pt[mid] + T - 1 <= b0) a = mid + 1; else z = mid; }
            lo = a;
        }
    };
    for (int e = tid; e < nent; e += CT_THREADS + 32) {
        int lo, hi;
        reach(e, lo, hi);
        s_cntoff[e] = max(hi - lo, 0);
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int e = 0; e < nent; e++) {
            int c = s_cntoff[e];
            if (n + c > CT_MAXITEMS) c = CT_MAXITEMS - n;        // cannot happen: bound checked on the host
            s_cntoff[e] = n;
            n += c;
        }
        s_cntoff[nent] = n;
        s_nitems = n;
        for (int i = 0; i < S; i++) {
            mbar_init(&s_full[i], 1);                             // the producer's expect_tx arrive
            mbar_init(&s_empty[i], CT_THREADS / 32);              // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    for (int e = tid; e < nent; e += CT_THREADS + 32) {
        int lo, hi;
        reach(e, lo, hi);
        const int bt = e / NOY, t = e - bt * NOY;
        const int off = s_cntoff[e], c = s_cntoff[e + 1] - off;
        for (int q = 0; q < c; q++)
            s_items[off + q] = ((uint32_t)t << 24) | ((uint32_t)bt << 12) | (uint32_t)(lo + q);
    }
    // ---- row ranges of the centre lines around the base line -------------------
    const int cza = max(bz - g.rz, g.rz), czb = min(bz + g.rz, g.Z - 1 - g.rz);
    const int cya = max(by - g.ry, g.ry), cyb = min(by + g.ry, g.Y - 1 - g.ry);
    const int ncy = cyb - cya + 1, nlines = max(czb - cza + 1, 0) * max(ncy, 0);
    __shared__ int s_clo, s_chi;
    __syncthreads();
    const int nitems = s_nitems;
    CT_TICK(0)
#ifdef CT_PROFILE
    if (tid == 0) { atomicAdd(&ct_prof[4], 1ull); atomicAdd(&ct_prof[5], (unsigned long long)nitems);
                    atomicAdd(&ct_prof[6], (unsigned long long)((nitems + CT_THREADS - 1) / CT_THREADS)); }
#endif
    if (nitems == 0 || nlines <= 0) return;
    unsigned stage_no = 0;           // stages published (producer) / consumed (consumers) so far

    for (int ibase = 0; ibase < nitems; ibase += CT_THREADS) {
        const bool have = tid < CT_THREADS && ibase + tid < nitems;
        int my_t = 0, b0 = -30000, p0 = -30000;
        if (have) {
            uint32_t it = s_items[ibase + tid];
            my_t = it >> 24;
            b0 = s_bt[(it >> 12) & 0xfff];
            p0 = s_pt[my_t * CT_MAXTILES + (it & 0xfff)];
        }
        // centres that can see a voxel of both windows
        const int my_clo = max(b0, p0) - g.rx, my_chi = min(b0, p0) + T - 1 + g.rx;
        // ---- row ranges of the centre lines, restricted to the centres this batch
        // of items can use (items are base-tile major: one x-neighbourhood) ----------
        if (tid == 0) { s_clo = 0x7fffffff; s_chi = -0x7fffffff; }
        __syncthreads();
        if (have && my_chi >= my_clo) { atomicMin(&s_clo, my_clo); atomicMax(&s_chi, my_chi); }
        __syncthreads();
        const int cxa = max(s_clo, g.rx), cxb = min(s_chi, g.X - 1 - g.rx);
        for (int l = tid; l < nlines; l += CT_THREADS + 32) {
            int cz = cza + l / ncy, cy = cya + l % ncy;
            int64_t cline = ((int64_t)cz * g.Y + cy) * g.X;
            int ra = 0, rb = 0;
            if (cxb >= cxa) {
                ra = rows_before(fgidx, cline + cxa, g.V, F);
                rb = rows_before(fgidx, cline + cxb + 1, g.V, F);
            }
            s_ra[l] = ra; s_rb[l] = rb;
        }
        __syncthreads();
        float acc[T][T];
#pragma unroll
        for (int j = 0; j < T; j++)
#pragma unroll
            for (int m = 0; m < T; m++) acc[j][m] = 0.0f;

        // ---- pipeline over (centre line, chunk) stages: a ring of S buffers, filled by
        // the producer warp through the TMA engine (full barriers), released by the
        // consumer warps one by one (empty barriers) — no CTA-wide barrier per stage, so
        // warps whose tiles see no centre in a stage run ahead to the next ----------------
        // offset rows that have a patch row in centre line l (bit t), 0 = none
        auto act_mask = [&](int l) -> unsigned {
            int cz = cza + l / ncy, cy = cya + l % ncy;
            int q1z = bz - cz + g.rz, q1y = by - cy + g.ry;
            unsigned m = 0;
#pragma unroll
            for (int t = 0; t < NOY; t++) {
                int q2z = q1z + oz_t[t], q2y = q1y + oy_t[t];
                if (row_ok[t] && q2z >= 0 && q2z < g.psz && q2y >= 0 && q2y < g.psy) m |= 1u << t;
            }
            return m;
        };
        if (tid >= CT_THREADS) {
            if (tid == CT_THREADS) {
                int st_l = 0, st_c = s_ra[0];
                while (true) {
                    while (st_l < nlines && !(act_mask(st_l) && st_c < s_rb[st_l])) {
                        st_l++;
                        if (st_l < nlines) st_c = s_ra[st_l];
                    }
                    const int slot = (int)(stage_no % (unsigned)S);
                    const unsigned use = stage_no / (unsigned)S;
                    if (use > 0) mbar_wait(&s_empty[slot], (use - 1) & 1);
                    stage_no++;
                    if (st_l >= nlines) {                         // end of this batch
                        s_info[slot] = make_int2(0, 0);
                        mbar_arrive(&s_full[slot]);
                        break;
                    }
                    const int l = st_l, c0 = st_c;
                    const int nc = min(CT_NCCH, s_rb[l] - c0);
                    s_info[slot] = make_int2(l, nc);
                    const unsigned am = act_mask(l);
                    const int cz = cza + l / ncy, cy = cya + l % ncy;
                    const int q1z = bz - cz + g.rz, q1y = by - cy + g.ry;
                    float* dst = sA + slot * bufstride;
                    const unsigned bytes = (unsigned)(nc * RS * 4);
                    mbar_expect_tx(&s_full[slot], (1 + __popc(am)) * bytes);
                    bulk_load(dst, dp + ((int64_t)(q1z * g.psy + q1y) * F + c0) * RS, bytes,
                              &s_full[slot]);
#pragma unroll
                    for (int t = 0; t < NOY; t++)
                        if (am & (1u << t))
                            bulk_load(dst + (1 + t) * substride,
                                      dp + ((int64_t)((q1z + oz_t[t]) * g.psy + (q1y + oy_t[t])) * F + c0) * RS,
                                      bytes, &s_full[slot]);
                    st_c += nc;
                }
            }
        } else {
            while (true) {
                const int slot = (int)(stage_no % (unsigned)S);
                const unsigned use = stage_no / (unsigned)S;
                mbar_wait(&s_full[slot], use & 1);
                const int2 info = s_info[slot];
                stage_no++;
                const int cur_l = info.x, cur_nc = info.y;
                // ---- accumulate over the staged centres ---------------------------------
                bool mine = cur_nc > 0 && have && my_chi >= my_clo &&
                            ((act_mask(cur_l) >> my_t) & 1u);
                if (mine) {
                    const float* A1 = sA + slot * bufstride;
                    const int nc = cur_nc;
                    const int cx_first = __float_as_int(A1[0]);
                    const int cx_last = __float_as_int(A1[(nc - 1) * RS]);
                    if (!(cx_first > my_chi || cx_last < my_clo)) {
                        int lo = 0, hi = nc;
                        while (lo < hi) {
                            int mid = (lo + hi) >> 1;
                            if (__float_as_int(A1[mid * RS]) < my_clo) lo = mid + 1; else hi = mid;
                        }
                        const float* A2 = A1 + (1 + my_t) * substride;
                        for (int ci = lo; ci < nc; ci++) {
                            const int cx = __float_as_int(A1[ci * RS]);
                            if (cx > my_chi) break;
                            const float* p1 = A1 + ci * RS + DP_GUARD + g.rx + (b0 - cx);
                            const float* p2 = A2 + ci * RS + DP_GUARD + g.rx + (p0 - cx);
                            // exactly one of the two products below is non-zero per pair:
                            //   a1 * max(a2,0)          high-high (+) and background-high (-)
                            //   max(a1,0) * min(a2,0)   high-background (-)
                            // background-background pairs add exact zeros (they do not vote)
                            float a1[T], h1[T], h2[T], l2[T];
#pragma unroll
                            for (int j = 0; j < T; j++) {
                                a1[j] = p1[j];
                                float a2 = p2[j];
                                h1[j] = fmaxf(a1[j], 0.0f);
                                h2[j] = fmaxf(a2, 0.0f);
                                l2[j] = fminf(a2, 0.0f);
                            }
#pragma unroll
                            for (int j = 0; j < T; j++)
#pragma unroll
                                for (int m = 0; m < T; m++)
                                    acc[j][m] = fmaf(h1[j], l2[m], fmaf(a1[j], h2[m], acc[j][m]));
                        }
                    }
                }
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&s_empty[slot]);  // this warp is done with the buffer
                if (cur_nc == 0) break;
            }
        }
        // the producer's other lanes and every consumer agree on the stage count: the
        // producer thread counted the same stages (one per buffer fill + the end marker)
        stage_no = __shfl_sync(0xffffffffu, stage_no, 0);
        __syncthreads();
        CT_TICK(1)
        // ---- epilogue: normalise with the integer counters and store -----------
        if (have) {
            int rlin = rlin_c + blockIdx.x * NOY + my_t;
            const int kbase = rlin * g.nx - g.K - 1 + (g.psx - 1);   // k = kbase + ox
            const bool same = (blockIdx.x * NOY + my_t) == 0;
            int64_t pline = 0;
#pragma unroll
            for (int t = 0; t < NOY; t++) if (my_t == t) pline = pline_t[t];
#pragma unroll
            for (int j = 0; j < T; j++) {
                const int b = b0 + j;
                if (b >= g.X || !(flags[bline + b] & PPP_FLAG_GATED)) continue;
                const int64_t orow = (int64_t)fgidx[bline + b] * g.K;
#pragma unroll
                for (int m = 0; m < T; m++) {
                    const int p = p0 + m;
                    const int ox = p - b;
                    if (p >= g.X || ox <= -g.psx || ox >= g.psx) continue;
                    if (same && ox <= 0) continue;
                    if (!(flags[pline + p] & PPP_FLAG_GATED)) continue;
                    const int64_t o = orow + kbase + ox;
                    const uint32_t c = cnt[o];
                    if (c == 0) continue;
                    cons[o] = consensus_epilogue(cfg, acc[j][m], (int)(c & 0xffffu), (int)(c >> 16));
                }
            }
        }
        __syncthreads();
        CT_TICK(2)
    }
}

#ifndef CT_NOY
#define CT_NOY 3
#endif

static size_t rows_smem(const Geo& g, int stages)
{
    return (size_t)stages * (1 + CT_NOY) * CT_NCCH * g.rsg * 4 + 2 * CT_MAXLINES * 4 +
           (size_t)CT_MAXITEMS * 4 + (size_t)(1 + CT_NOY) * CT_MAXTILES * 2 + CT_XMAX * 2 + 128;
}
// ring depth: as many stage buffers as fit next to the static tables (2..CT_STAGES)
#ifndef CT_STAGES
#define CT_STAGES 2
#endif
static int rows_stages(const Geo& g)
{
    int s = CT_STAGES;
    while (s > 2 && rows_smem(g, s) > 112 * 1024) s--;
    return s;
}

static int tiles_stride(const Geo& g) { return (g.X + CT_T - 1) / CT_T + 1; }

extern "C" int64_t ppp_consensus_scratch_bytes(const ppp_cfg* cfg)
{
    // tile lists of the pre-pass: i16 [lines][ts] + i32 [lines]
    Geo g = make_geo(*cfg);
    const int64_t lines = (int64_t)g.Z * g.Y;
    return ((lines * tiles_stride(g) * 2 + 255) / 256) * 256 + lines * 4 + 256;
}

extern "C" int ppp_consensus(const float* dp, const uint64_t* rbits, const uint8_t* flags,
                             const int32_t* fgidx, const int32_t* rowvox, int64_t F,
                             const ppp_cfg* cfg, float* cons, uint32_t* cnt, int32_t impl,
                             void* scratch, void* stream)
{
    (void)scratch;
    if (F <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    Geo g = make_geo(*cfg);
    if (impl == 1) {
        consensus_naive_kernel<<<(unsigned)F, 128, 0, s>>>(dp, flags, fgidx, rowvox, F, *cfg,
                                                           cons, cnt);
        return ppp_check("ppp_consensus(naive)");
    }
    if (rbits == nullptr || cnt == nullptr)
        return ppp_fail(-1, "ppp_consensus: tiled path needs rbits and cnt");
    if (g.psx > 64) return ppp_fail(-1, "ppp_consensus: psx > 64 unsupported by the bit path");
    if (g.X > CT_XMAX) return ppp_fail(-1, "ppp_consensus: X > 2048 unsupported, use blocks");
    if (g.psz * g.psy > CT_MAXLINES)
        return ppp_fail(-1, "ppp_consensus: more than 128 centre lines per window");
    // worst-case work items of one CTA of the tiled kernel: base tiles x partner tiles
    // within reach x offset rows; beyond the table the bit-guided gather takes over
    {
        const int nbt = (g.X + CT_T - 1) / CT_T;
        int reach = (2 * (g.psx - 1) + 2 * CT_T - 1) / CT_T + 1;
        if (reach > nbt) reach = nbt;
        if ((int64_t)nbt * reach * CT_NOY > CT_MAXITEMS) {
            if (impl == 3) return ppp_fail(-1, "ppp_consensus: line too long for the tiled kernel");
            if (impl == 0) impl = 2;
        }
    }
    if (impl == 2 || (impl == 0 && g.psx < 16)) {
        if (g.psx > 64) return ppp_fail(-1, "ppp_consensus: psx > 64 needs the tiled kernel");
        if (g.K > 65535) return ppp_fail(-1, "ppp_consensus: more than 65535 offsets");
        size_t bsm = (size_t)g.psz * g.psy * 16 + (size_t)g.K * 6 + 16;
        cudaError_t be = cudaFuncSetAttribute(consensus_bits_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)bsm);
        if (be != cudaSuccess) return ppp_fail((int)be, "ppp_consensus: smem attribute (bits)");
        consensus_bits_kernel<<<(unsigned)F, 128, bsm, s>>>(
            dp, (const unsigned long long*)rbits, flags, fgidx, rowvox, F, *cfg, cons, cnt);
        return ppp_check("ppp_consensus(bits)");
    }
    {
        if (g.K > 65535) return ppp_fail(-1, "ppp_consensus: more than 65535 offsets");
        size_t csm = (size_t)g.psz * g.psy * 16 + (size_t)g.K * 6 + 16;
        cudaError_t ce = cudaFuncSetAttribute(consensus_count_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)csm);
        if (ce != cudaSuccess) return ppp_fail((int)ce, "ppp_consensus: smem attribute (count)");
        consensus_count_kernel<<<(unsigned)F, 256, csm, s>>>(
            (const unsigned long long*)rbits, flags, fgidx, rowvox, F, *cfg, cons, cnt);
    }
    if (cfg->prod_mode == 0) return ppp_check("ppp_consensus(count)");   // no float sums needed
    const int stages = rows_stages(g);
    size_t smem = rows_smem(g, stages);
    cudaError_t e = cudaFuncSetAttribute(consensus_rows_kernel<CT_NOY>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ppp_fail((int)e, "ppp_consensus: smem attribute");
    int nrows_off = (g.nz * g.ny - 1) / 2 + 1;
    // offset-row groups on x (scheduled first): the groups of one line share its A1 rows in L2
    dim3 grid((unsigned)((nrows_off + CT_NOY - 1) / CT_NOY), (unsigned)(g.Z * g.Y));
    if (scratch == nullptr) return ppp_fail(-1, "ppp_consensus: scratch required");
    const int64_t lines = (int64_t)g.Z * g.Y;
    const int ts = tiles_stride(g);
    int16_t* tiles = (int16_t*)scratch;
    int32_t* ntiles = (int32_t*)((char*)scratch + ((lines * ts * 2 + 255) / 256) * 256);
    consensus_tiles_kernel<<<(unsigned)((lines + 3) / 4), 128, 0, s>>>(flags, *cfg, ts, tiles, ntiles);
    consensus_rows_kernel<CT_NOY><<<grid, CT_THREADS + 32, smem, s>>>(
        dp, flags, fgidx, rowvox, (int)F, *cfg, cnt, cons, stages, tiles, ntiles, ts);
    return ppp_check("ppp_consensus(rows)");
}
