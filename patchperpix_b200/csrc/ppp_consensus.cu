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
                       ppp_cfg cfg, float* __restrict__ cons, uint32_t* __restrict__ cnt)
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
                    float d1 = dp[(int64_t)rc * g.P + po1];
                    float d2 = dp[(int64_t)rc * g.P + po2];
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

extern "C" int ppp_consensus(const float* dp, const uint8_t* flags, const int32_t* fgidx,
                             const int32_t* rowvox, int64_t F, const ppp_cfg* cfg,
                             float* cons, uint32_t* cnt, void* stream)
{
    if (F <= 0) return 0;
    consensus_naive_kernel<<<(unsigned)F, 128, 0, (cudaStream_t)stream>>>(
        dp, flags, fgidx, rowvox, *cfg, cons, cnt);
    return ppp_check("ppp_consensus");
}
