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

// ---------------------------------------------------------------------------
// reference-order ranking.  rankPatches.cu accumulates its ~P^2/2 terms in ONE
// float, serially; for large patches the rounding of that running sum is
// systematic (1e-3 relative at 41x41) and decides the rank order, hence the
// cover, hence the labels.  To stay label-identical these kernels reproduce the
// exact sequence of float additions of the reference's loop nest (po1 over the
// high pixels, po2 over the other gated pixels):
//   rank_lists_kernel  compacts, per patch centre, its voting pixels in po
//                      order: e_lin (position in the offset raster | high bit),
//                      h_idx (list index of every high pixel) and, from both
//                      ends of x_lin / x_row, the offset-raster position and the
//                      consensus row of the background pixels (front) and of the
//                      high pixels (back);
//   rank_ref_kernel    one WARP owns RR_CPW centres.  Lane s < RR_CPW performs
//                      the serial float adds of centre s; the values are
//                      gathered by the whole warp, 32 consecutive terms of one
//                      centre per coalesced request (4 requests in flight per
//                      lane), and handed over through a transposition buffer.
//                      A high pixel p1 contributes two segments: the background
//                      pixels before it (reversed slot, row of p2) and all
//                      voting pixels after it (row of p1).  Missing terms are
//                      fed as 0.0f, which leaves a float sum unchanged.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
rank_lists_kernel(const float* __restrict__ dp, const uint8_t* __restrict__ flags,
                  const int32_t* __restrict__ fgidx, const int32_t* __restrict__ rowvox,
                  int64_t F, ppp_cfg cfg, uint16_t* __restrict__ e_lin,
                  uint16_t* __restrict__ h_idx, uint16_t* __restrict__ x_lin,
                  int32_t* __restrict__ x_row, int32_t* __restrict__ meta)
{
    Geo g = make_geo(cfg);
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= F) return;
    const int vc = rowvox[row];
    const int64_t loff = row * g.P;
    int n = 0, nH = 0, nG = 0;
    if (flags[vc] & PPP_FLAG_CENTRE) {
        int cz, cy, cx;
        vox_decode(g, vc, cz, cy, cx);
        for (int base = 0; base < g.P; base += 32) {
            int po = base + lane;
            float d = 0.0f;
            bool gated = false;
            int lin = 0, prow = -1;
            if (po < g.P) {
                int qz, qy, qx;
                po_decode(g, po, qz, qy, qx);
                d = dp[dp_index(g, F, row, po)];
                int pv = ((cz + qz - g.rz) * g.Y + (cy + qy - g.ry)) * g.X + (cx + qx - g.rx);
                gated = (flags[pv] & PPP_FLAG_GATED) != 0;
                lin = po_lin(g, qz, qy, qx);
                if (d != 0.0f) prow = fgidx[pv];
            }
            unsigned bal = __ballot_sync(0xffffffffu, d != 0.0f);
            unsigned balh = __ballot_sync(0xffffffffu, d > 0.0f);
            unsigned lt = (1u << lane) - 1u;
            if (d != 0.0f) {
                int idx = n + __popc(bal & lt);
                e_lin[loff + idx] = (uint16_t)(lin | (d > 0.0f ? 0x8000 : 0));
                if (d > 0.0f) {
                    int h = nH + __popc(balh & lt);
                    h_idx[loff + h] = (uint16_t)idx;
                    x_lin[loff + g.P - 1 - h] = (uint16_t)lin;
                    x_row[loff + g.P - 1 - h] = prow;
                } else {
                    int q = (n - nH) + __popc((bal & ~balh) & lt);
                    x_lin[loff + q] = (uint16_t)lin;
                    x_row[loff + q] = prow;
                }
            }
            n += __popc(bal);
            nH += __popc(balh);
            nG += __popc(__ballot_sync(0xffffffffu, gated));
        }
    } else n = -1;
    if (lane == 0) { meta[row * 4] = n; meta[row * 4 + 1] = nH; meta[row * 4 + 2] = nG; }
}

#define RR_WARPS 4
#define RR_ST 36            // floats per staging row (16-byte aligned, conflict-free for 8 lanes)

// the segment of terms every lane gathers next for one centre: term t of the
// segment is +-cons[base + xrow[rp + t] * K + sgn * (lst[lp + t] & 0x7fff)].
// Indices, not pointers: loads through them stay LDG (predicated, no branches).
// IdxT = int32_t when every index space fits (one 16-byte descriptor), else int64_t.
template <typename IdxT>
struct __align__(16) RankSeg {
    IdxT base;              // "after": row(p1)*K - lin(p1) - 1;  "before": lin(p1) - 1
    IdxT lp;                // next entry in lst (e_lin part for "after", x_lin part for "before")
    IdxT rp;                // "before": next entry in x_row (rows of the background pixels)
    int cb;                 // terms left in this segment (<= 0: nothing to gather) * 2 + before
};

// RR_CPW centres per warp, RR_T requests of 32 terms per centre and round, RR_B
// centres gathered at a time.  lst = [e_lin | h_idx | x_lin] (three u16 arrays of
// F*P entries, `lb16` entries apart).
template <int RR_CPW, int RR_B, int RR_T, typename IdxT>
__global__ void __launch_bounds__(RR_WARPS * 32)
rank_ref_kernel(const int32_t* __restrict__ rowvox, const float* __restrict__ cons,
                const uint16_t* __restrict__ lst, int64_t lb16,
                const int32_t* __restrict__ x_row,
                const int32_t* __restrict__ meta, const uint32_t* __restrict__ perm,
                int64_t F, ppp_cfg cfg, float* __restrict__ score)
{
    constexpr int ST = 32 * RR_T + 4;     // staging row: 16-byte aligned, conflict-free
    constexpr int TPR = 32 * RR_T;        // terms per round
    typedef RankSeg<IdxT> Seg;
    Geo g = make_geo(cfg);
    __shared__ Seg s_seg[RR_WARPS][RR_CPW];
    __shared__ __align__(16) float s_stage[RR_WARPS][RR_CPW][ST];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    Seg* segs = s_seg[w];
    float (*stage)[ST] = s_stage[w];
    const bool count_mode = (cfg.rank_flags & 2) != 0;
    const int64_t ngroups = (F + RR_CPW - 1) / RR_CPW;
    const IdxT K = (IdxT)g.K;
    const IdxT H_IDX = (IdxT)lb16, X_LIN = (IdxT)(2 * lb16);

    for (int64_t grp = (int64_t)blockIdx.x * RR_WARPS + w; grp < ngroups;
         grp += (int64_t)gridDim.x * RR_WARPS) {
        // lane s < RR_CPW owns the centre with the (grp*RR_CPW + s)-th largest work
        // (perm: rows sorted by term count, so that the centres of a warp finish
        // together and the heavy warps start first)
        const int64_t slot = grp * RR_CPW + lane;
        const int64_t myrow = (lane < RR_CPW && slot < F) ? (int64_t)perm[slot] : F;
        int n_s = -1, nH_s = 0, nG_s = 0;
        if (myrow < F) {
            n_s = meta[myrow * 4]; nH_s = meta[myrow * 4 + 1]; nG_s = meta[myrow * 4 + 2];
        }
        const IdxT loff = (IdxT)((myrow < F ? myrow : 0) * g.P);
        float acc = 0.0f;                 // running float sum of my centre
        int cur_h = 0;                    // current high pixel (index into h_idx)
        int pend_after = 0, pend_i = 0;   // "after" segment still to come for cur_h
        IdxT pend_base = 0;
        bool alive = n_s > 0 && nH_s > 0;
        auto open_after = [&](Seg& r) {
            r.base = pend_base;
            r.lp = loff + pend_i + 1;
            r.rp = 0;                     // unused
            r.cb = pend_after * 2;
            pend_after = 0;
        };
        // descriptor of the NEXT high pixel, fetched one row ahead so that opening a
        // row never waits on memory
        int nx_i = 0, nx_li = 0, nx_row = 0;
        auto fetch = [&](int h) {
            if (h < nH_s) {
                nx_i = lst[H_IDX + loff + h];
                nx_li = lst[X_LIN + loff + g.P - 1 - h];
                nx_row = x_row[loff + g.P - 1 - h];
            }
        };
        auto open_row = [&]() {           // owner lane: publish the next non-empty segment
            Seg r;
            r.base = 0; r.lp = 0; r.rp = 0; r.cb = 0;
            while (cur_h < nH_s) {
                const int i = nx_i, li = nx_li;
                const IdxT hrow = nx_row;
                fetch(cur_h + 1);
                const int lb = i - cur_h, na = n_s - 1 - i;
                if (lb + na > 0) {
                    pend_base = hrow * K - li - 1;
                    pend_after = na; pend_i = i;
                    if (lb > 0) {
                        r.base = li - 1;
                        r.lp = X_LIN + loff; r.rp = loff; r.cb = lb * 2 + 1;
                    } else open_after(r);
                    break;
                }
                cur_h++;
            }
            if (cur_h >= nH_s) { alive = false; r.cb = 0; }
            segs[lane] = r;
        };
        if (alive) fetch(0);
        if (lane < RR_CPW) {
            if (alive) open_row();
            else { Seg r; r.base = 0; r.lp = 0; r.rp = 0; r.cb = 0; segs[lane] = r; }
        }
        __syncwarp();
        while (__ballot_sync(0xffffffffu, alive)) {
            // ---- gather TPR consecutive terms of every centre, RR_B centres at a time,
            // in three phases (list entries, consensus values, hand-over) so that
            // RR_B * RR_T independent loads are in flight per lane; branch-free ---------
#pragma unroll
            for (int u0 = 0; u0 < RR_CPW; u0 += RR_B) {
                int e_[RR_B][RR_T], r_[RR_B][RR_T], bf_[RR_B];
                IdxT b_[RR_B];
                float v_[RR_B][RR_T];
#pragma unroll
                for (int u = 0; u < RR_B; u++) {
                    const Seg d = segs[u0 + u];                   // broadcast
                    const int cnt = d.cb >> 1;
                    b_[u] = d.base; bf_[u] = d.cb & 1;
#pragma unroll
                    for (int t = 0; t < RR_T; t++) {
                        const bool ok = lane + 32 * t < cnt;
                        e_[u][t] = ok ? (int)lst[d.lp + lane + 32 * t] : -1;
                        r_[u][t] = (ok && bf_[u]) ? x_row[d.rp + lane + 32 * t] : 0;
                    }
                }
#pragma unroll
                for (int u = 0; u < RR_B; u++)
#pragma unroll
                    for (int t = 0; t < RR_T; t++) {
                        const int le = e_[u][t] & 0x7fff;
                        const IdxT idx = b_[u] + (IdxT)r_[u][t] * K + (bf_[u] ? -le : le);
                        v_[u][t] = e_[u][t] >= 0 ? cons[idx] : 0.0f;
                    }
#pragma unroll
                for (int u = 0; u < RR_B; u++)
#pragma unroll
                    for (int t = 0; t < RR_T; t++) {
                        // rankPatches.cu:88-100 (both high), :102-137 (high, background),
                        // :109-126 (background before the high pixel: reversed slot).  A
                        // "before" entry never carries the high bit.
                        const bool hj = (e_[u][t] & 0x8000) != 0;
                        float v3 = v_[u][t];
                        if (count_mode)
                            v3 = (v3 != 0.0f) ? copysignf(1.0f, v3) : (hj ? -1.0f : 1.0f);
                        float val = hj ? v3 : -v3;
                        stage[u0 + u][lane + 32 * t] = e_[u][t] >= 0 ? val : 0.0f;
                    }
            }
            __syncwarp();
            // ---- serial float adds of my centre, in the reference's order --------------
            if (alive) {
                const float4* st = (const float4*)stage[lane];
#pragma unroll
                for (int q = 0; q < 8 * RR_T; q++) {
                    float4 t = st[q];
                    acc += t.x; acc += t.y; acc += t.z; acc += t.w;
                }
                Seg& r = segs[lane];
                if ((r.cb >> 1) > TPR) {
                    r.cb -= 2 * TPR; r.lp += TPR; r.rp += TPR;
                } else if ((r.cb & 1) && pend_after > 0) {
                    open_after(r);
                } else {
                    cur_h++;
                    open_row();
                }
            }
            __syncwarp();
        }
        if (n_s >= 0) {
            unsigned h = (unsigned)nH_s, gg = (unsigned)nG_s;
            unsigned fgCnt = h * gg - h - (h * (h - 1)) / 2;
            score[rowvox[myrow]] = (cfg.rank_flags & 1) ? acc / (float)(fgCnt > 1 ? fgCnt : 1) : acc;
        }
    }
}

// sort key: number of float additions of a centre, descending
__global__ void rank_work_kernel(const int32_t* __restrict__ meta, int64_t F, int shift,
                                 uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= F) return;
    int n = meta[r * 4], nH = meta[r * 4 + 1];
    uint32_t work = (n > 0 && nH > 0)
        ? (uint32_t)nH * (uint32_t)(n - 1) - ((uint32_t)nH * (uint32_t)(nH - 1)) / 2 : 0u;
    keys[r] = ~(work >> shift);
    vals[r] = (uint32_t)r;
}

static size_t work_sort_bytes(int64_t F)
{
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs((void*)nullptr, tb, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)F);
    return tb;
}

extern "C" int64_t ppp_rank_scratch_bytes(const ppp_cfg* cfg, int64_t F)
{
    Geo g = make_geo(*cfg);
    if (F < 1) F = 1;
    // e_lin, h_idx, x_lin (u16) + x_row (i32) + meta + keys/vals x4 + sort temp
    return 5 * ((F * g.P * 2 + 255) / 256) * 256 + ((F * 16 + 255) / 256) * 256 +
           4 * ((F * 4 + 255) / 256) * 256 + ((work_sort_bytes(F) + 255) / 256) * 256 + 256;
}

extern "C" int ppp_rank(const float* dp, const uint8_t* flags, const int32_t* fgidx,
                        const int32_t* rowvox, int64_t F, const float* cons,
                        const ppp_cfg* cfg, float* score, void* scratch, void* stream)
{
    Geo g = make_geo(*cfg);
    cudaStream_t s = (cudaStream_t)stream;
    rank_fill_kernel<<<(unsigned)((g.V + 255) / 256), 256, 0, s>>>(*cfg, score);
    if (F <= 0) return ppp_check("ppp_rank");
    if (cfg->rank_flags & 4) {
        // fast path: parallel sum in double (NOT the reference's rounding)
        size_t smem = (size_t)g.P * 9 + 16;
        rank_kernel<<<(unsigned)F, RANK_THREADS, smem, s>>>(dp, flags, fgidx, rowvox,
                                                            cons, F, *cfg, score);
        return ppp_check("ppp_rank(fast)");
    }
    if (2 * g.K + 1 >= 32768) return ppp_fail(-1, "ppp_rank: patch too large");
    if (scratch == nullptr) return ppp_fail(-1, "ppp_rank: scratch required");
    size_t lb = ((F * g.P * 2 + 255) / 256) * 256;
    uint16_t* e_lin = (uint16_t*)scratch;
    uint16_t* h_idx = (uint16_t*)((char*)scratch + lb);
    uint16_t* x_lin = (uint16_t*)((char*)scratch + 2 * lb);
    int32_t* x_row = (int32_t*)((char*)scratch + 3 * lb);
    int32_t* meta = (int32_t*)((char*)scratch + 5 * lb);
    size_t mb = ((F * 16 + 255) / 256) * 256, fb = ((F * 4 + 255) / 256) * 256;
    uint32_t* keys = (uint32_t*)((char*)scratch + 5 * lb + mb);
    uint32_t* keys_out = keys + fb / 4;
    uint32_t* vals = keys_out + fb / 4;
    uint32_t* perm = vals + fb / 4;
    void* sort_tmp = (char*)scratch + 5 * lb + mb + 4 * fb;
    size_t stb = work_sort_bytes(F);
    rank_lists_kernel<<<(unsigned)((F + 3) / 4), 128, 0, s>>>(dp, flags, fgidx, rowvox, F, *cfg,
                                                              e_lin, h_idx, x_lin, x_row, meta);
    rank_work_kernel<<<(unsigned)((F + 255) / 256), 256, 0, s>>>(meta, F, cfg->reserved & 31, keys, vals);
    cub::DeviceRadixSort::SortPairs(sort_tmp, stb, keys, keys_out, vals, perm, (int)F, 0, 32, s);
    // 32-bit indices when the list and consensus index spaces fit
    const bool idx32 = 3 * (int64_t)(lb / 2) < 0x7fffffffLL &&
                       (F + 1) * (int64_t)g.K + 0x8000 < 0x7fffffffLL;
#define RR_LAUNCH(CPW, B, T)                                                                \
    {                                                                                       \
        int64_t ngroups = (F + CPW - 1) / CPW;                                              \
        int64_t nblk = (ngroups + RR_WARPS - 1) / RR_WARPS;                                 \
        if (grid_cap > 0 && nblk > grid_cap) nblk = grid_cap;                               \
        if (idx32)                                                                          \
            rank_ref_kernel<CPW, B, T, int32_t><<<(unsigned)nblk, RR_WARPS * 32, 0, s>>>(   \
                rowvox, cons, e_lin, (int64_t)(lb / 2), x_row, meta, perm, F, *cfg, score); \
        else                                                                                \
            rank_ref_kernel<CPW, B, T, int64_t><<<(unsigned)nblk, RR_WARPS * 32, 0, s>>>(   \
                rowvox, cons, e_lin, (int64_t)(lb / 2), x_row, meta, perm, F, *cfg, score); \
    }
    // tuning: resident CTAs per SM (0 = one CTA per group of centres, no grid-stride)
    const int64_t grid_cap = (int64_t)((cfg->reserved >> 12) & 15) * 148;
    // small windows (7^3: a high pixel's segments hold ~60 terms): 64 terms per centre and
    // round waste less padding than 128 (1.50 -> 1.26 ms on a 47 k-row flylight block)
    int variant = (cfg->reserved >> 8) & 7;
    if (variant == 0 && g.P <= 1024) variant = 1;
    switch (variant) {
    case 1: RR_LAUNCH(4, 4, 2); break;
    case 2: RR_LAUNCH(8, 4, 1); break;
    case 3: RR_LAUNCH(4, 4, 1); break;
    case 4: RR_LAUNCH(4, 2, 2); break;
    case 5: RR_LAUNCH(2, 2, 2); break;
    case 6: RR_LAUNCH(8, 4, 2); break;
    case 7: RR_LAUNCH(4, 2, 4); break;
    default: RR_LAUNCH(4, 4, 4); break;
    }
    return ppp_check("ppp_rank(reference order)");
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
