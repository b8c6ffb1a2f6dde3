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
//   rank_lists_kernel  compacts, per patch centre, the voting pixels (in po
//                      order) and the sub-list of high pixels into scratch;
//   rank_ref_kernel    one WARP owns RR_CPW centres.  Lane s < RR_CPW performs
//                      the serial float adds of centre s; the values are
//                      gathered by the whole warp, 32 consecutive terms of one
//                      centre per coalesced request (all RR_CPW requests in
//                      flight together), and handed over through a
//                      transposition buffer.  Skipped terms are fed as +0.0f,
//                      which leaves a float sum unchanged.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
rank_lists_kernel(const float* __restrict__ dp, const uint8_t* __restrict__ flags,
                  const int32_t* __restrict__ rowvox, int64_t F, ppp_cfg cfg,
                  uint16_t* __restrict__ lists, uint16_t* __restrict__ hlists,
                  uint16_t* __restrict__ llists, int32_t* __restrict__ meta)
{
    Geo g = make_geo(cfg);
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= F) return;
    const int vc = rowvox[row];
    int n = 0, nH = 0, nG = 0;
    if (flags[vc] & PPP_FLAG_CENTRE) {
        int cz, cy, cx;
        vox_decode(g, vc, cz, cy, cx);
        for (int base = 0; base < g.P; base += 32) {
            int po = base + lane;
            float d = 0.0f;
            bool gated = false;
            if (po < g.P) {
                int qz, qy, qx;
                po_decode(g, po, qz, qy, qx);
                d = dp[dp_index(g, F, row, po)];
                int pv = ((cz + qz - g.rz) * g.Y + (cy + qy - g.ry)) * g.X + (cx + qx - g.rx);
                gated = (flags[pv] & PPP_FLAG_GATED) != 0;
            }
            unsigned bal = __ballot_sync(0xffffffffu, d != 0.0f);
            unsigned balh = __ballot_sync(0xffffffffu, d > 0.0f);
            unsigned lt = (1u << lane) - 1u;
            if (d != 0.0f) {
                int idx = n + __popc(bal & lt);
                lists[row * g.P + idx] = (uint16_t)(po | (d > 0.0f ? 0x8000 : 0));
                if (d > 0.0f) hlists[row * g.P + nH + __popc(balh & lt)] = (uint16_t)idx;
                else llists[row * g.P + (n - nH) + __popc((bal & ~balh) & lt)] = (uint16_t)po;
            }
            n += __popc(bal);
            nH += __popc(balh);
            nG += __popc(__ballot_sync(0xffffffffu, gated));
        }
    } else n = -1;
    if (lane == 0) { meta[row * 4] = n; meta[row * 4 + 1] = nH; meta[row * 4 + 2] = nG; }
}

#define RR_WARPS 4
#define RR_CPW 4

// what every lane needs to know about the current high row of one centre
struct __align__(16) RankRow {
    const float* cbase;     // cons + row(p1)*K - lin(p1) - 1 : slot of (p1, p2) = cbase[lin(p2)]
    int64_t loff;           // row * P: offset of this centre in lists / llists
    int i;                  // list index of the high pixel p1
    int t0;                 // first term of this round
    int lb;                 // background entries before p1 (reversed terms come first)
    int nslots;             // lb + (n - 1 - i)
    int li;                 // lin(p1)
    int vc;                 // centre voxel
    int pad0, pad1;
};

__global__ void __launch_bounds__(RR_WARPS * 32)
rank_ref_kernel(const uint8_t* __restrict__ flags, const int32_t* __restrict__ fgidx,
                const int32_t* __restrict__ rowvox, const float* __restrict__ cons,
                const uint16_t* __restrict__ lists, const uint16_t* __restrict__ hlists,
                const uint16_t* __restrict__ llists, const int32_t* __restrict__ meta,
                const uint32_t* __restrict__ perm, int64_t F, ppp_cfg cfg,
                float* __restrict__ score)
{
    Geo g = make_geo(cfg);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RankRow* s_rows = (RankRow*)smem_raw;                       // [RR_WARPS][RR_CPW]
    int32_t* tab_dv = (int32_t*)(s_rows + RR_WARPS * RR_CPW);   // [P] voxel delta of patch pixel
    uint16_t* tab_lin = (uint16_t*)(tab_dv + g.P);              // [P] position in the offset raster
    float* stage = (float*)(tab_lin + g.P + (g.P & 1));         // [RR_WARPS][RR_CPW][33]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int po = threadIdx.x; po < g.P; po += blockDim.x) {
        int qz, qy, qx;
        po_decode(g, po, qz, qy, qx);
        tab_dv[po] = ((qz - g.rz) * g.Y + (qy - g.ry)) * g.X + (qx - g.rx);
        tab_lin[po] = (uint16_t)po_lin(g, qz, qy, qx);
    }
    __syncthreads();
    float* mystage = stage + w * RR_CPW * 33;
    RankRow* myrows = s_rows + w * RR_CPW;
    const bool count_mode = (cfg.rank_flags & 2) != 0;
    const int64_t ngroups = (F + RR_CPW - 1) / RR_CPW;

    for (int64_t grp = (int64_t)blockIdx.x * RR_WARPS + w; grp < ngroups;
         grp += (int64_t)gridDim.x * RR_WARPS) {
        // lane s < RR_CPW owns the centre with the (grp*RR_CPW + s)-th largest work
        // (perm: rows sorted by term count, so that the centres of a warp finish
        // together and the heavy warps start first)
        const int64_t slot = grp * RR_CPW + lane;
        const int64_t myrow = (lane < RR_CPW && slot < F) ? (int64_t)perm[slot] : F;
        int n_s = -1, nH_s = 0, nG_s = 0, vc_s = 0;
        if (lane < RR_CPW && myrow < F) {
            n_s = meta[myrow * 4]; nH_s = meta[myrow * 4 + 1]; nG_s = meta[myrow * 4 + 2];
            vc_s = rowvox[myrow];
        }
        float acc = 0.0f;                 // running float sum of my centre
        int cur_h = 0;                    // current high row (index into hlist)
        bool alive = n_s > 0 && nH_s > 0;
        auto open_row = [&]() {           // owner lane: publish high row cur_h
            const int i = hlists[myrow * g.P + cur_h];
            const int ei = lists[myrow * g.P + i] & 0x7fff;
            const int li = tab_lin[ei];
            RankRow r;
            r.cbase = cons + (int64_t)fgidx[vc_s + tab_dv[ei]] * g.K - li - 1;
            r.loff = myrow * g.P;
            r.i = i; r.t0 = 0; r.lb = i - cur_h; r.nslots = (i - cur_h) + (n_s - 1 - i);
            r.li = li; r.vc = vc_s; r.pad0 = 0; r.pad1 = 0;
            myrows[lane] = r;
        };
        if (alive) open_row();
        __syncwarp();
        while (true) {
            const unsigned live = __ballot_sync(0xffffffffu, alive);
            if (!live) break;
            // ---- gather 32 consecutive terms of every live centre, in three phases
            // (list entries, consensus values, hand-over) so that RR_CPW independent
            // loads are in flight per lane ---------------------------------------------
            int ej_[RR_CPW], li_[RR_CPW], vc_[RR_CPW];
            const float* cb_[RR_CPW];
            int st_[RR_CPW];                                     // 0 none, 1 after, 2 before
            float val_[RR_CPW];
#pragma unroll
            for (int u = 0; u < RR_CPW; u++) {
                const RankRow r = myrows[u];                     // broadcast
                const int tt = r.t0 + lane;
                const bool ok = ((live >> u) & 1u) && tt < r.nslots;
                const bool before = tt < r.lb;
                st_[u] = ok ? (before ? 2 : 1) : 0;
                cb_[u] = r.cbase; li_[u] = r.li; vc_[u] = r.vc;
                ej_[u] = !ok ? 0 : (before ? (int)llists[r.loff + tt]
                                           : (int)lists[r.loff + r.i + 1 + (tt - r.lb)]);
            }
#pragma unroll
            for (int u = 0; u < RR_CPW; u++) {
                const int pj = ej_[u] & 0x7fff;
                const int lj = tab_lin[pj];
                float v3 = 0.0f;
                if (st_[u] == 1) v3 = cb_[u][lj];
                else if (st_[u] == 2)
                    v3 = cons[(int64_t)fgidx[vc_[u] + tab_dv[pj]] * g.K + li_[u] - lj - 1];
                val_[u] = v3;
            }
#pragma unroll
            for (int u = 0; u < RR_CPW; u++) {
                const bool hj = (ej_[u] & 0x8000) != 0;
                float v3 = val_[u], val = 0.0f;
                if (st_[u] == 1) {
                    // rankPatches.cu:88-100 (both high) / :102-137 (high, background)
                    if (count_mode) v3 = (v3 != 0.0f) ? copysignf(1.0f, v3) : (hj ? -1.0f : 1.0f);
                    val = hj ? v3 : -v3;
                } else if (st_[u] == 2) {
                    // background pixel before the high one: reversed slot (:109-126)
                    if (count_mode) v3 = (v3 != 0.0f) ? copysignf(1.0f, v3) : 1.0f;
                    val = -v3;
                }
                mystage[u * 33 + lane] = val;
            }
            __syncwarp();
            // ---- serial float adds of my centre, in the reference's order --------------
            if (alive) {
                const float* st = mystage + lane * 33;
#pragma unroll
                for (int q = 0; q < 32; q++) acc += st[q];
                const int t1 = myrows[lane].t0 + 32;
                if (t1 >= myrows[lane].nslots) {                 // next high row
                    cur_h++;
                    if (cur_h >= nH_s) alive = false;
                    else open_row();
                } else myrows[lane].t0 = t1;
            }
            __syncwarp();
        }
        if (n_s >= 0) {
            unsigned h = (unsigned)nH_s, gg = (unsigned)nG_s;
            unsigned fgCnt = h * gg - h - (h * (h - 1)) / 2;
            score[vc_s] = (cfg.rank_flags & 1) ? acc / (float)(fgCnt > 1 ? fgCnt : 1) : acc;
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
    return 3 * ((F * g.P * 2 + 255) / 256) * 256 + ((F * 16 + 255) / 256) * 256 +
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
    if (g.P >= 32768) return ppp_fail(-1, "ppp_rank: patch too large");
    if (scratch == nullptr) return ppp_fail(-1, "ppp_rank: scratch required");
    size_t lb = ((F * g.P * 2 + 255) / 256) * 256;
    uint16_t* lists = (uint16_t*)scratch;
    uint16_t* hlists = (uint16_t*)((char*)scratch + lb);
    uint16_t* llists = (uint16_t*)((char*)scratch + 2 * lb);
    int32_t* meta = (int32_t*)((char*)scratch + 3 * lb);
    size_t mb = ((F * 16 + 255) / 256) * 256, fb = ((F * 4 + 255) / 256) * 256;
    uint32_t* keys = (uint32_t*)((char*)scratch + 3 * lb + mb);
    uint32_t* keys_out = keys + fb / 4;
    uint32_t* vals = keys_out + fb / 4;
    uint32_t* perm = vals + fb / 4;
    void* sort_tmp = (char*)scratch + 3 * lb + mb + 4 * fb;
    size_t stb = work_sort_bytes(F);
    rank_lists_kernel<<<(unsigned)((F + 3) / 4), 128, 0, s>>>(dp, flags, rowvox, F, *cfg, lists,
                                                              hlists, llists, meta);
    rank_work_kernel<<<(unsigned)((F + 255) / 256), 256, 0, s>>>(meta, F, cfg->reserved & 31, keys, vals);
    cub::DeviceRadixSort::SortPairs(sort_tmp, stb, keys, keys_out, vals, perm, (int)F, 0, 32, s);
    size_t smem = (size_t)RR_WARPS * RR_CPW * sizeof(RankRow) + (size_t)g.P * 4 +
                  (size_t)(g.P + 1) * 2 + (size_t)RR_WARPS * RR_CPW * 33 * 4 + 32;
    cudaError_t e = cudaFuncSetAttribute(rank_ref_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ppp_fail((int)e, "ppp_rank: smem attribute");
    int64_t ngroups = (F + RR_CPW - 1) / RR_CPW;
    int64_t nblk = (ngroups + RR_WARPS - 1) / RR_WARPS;
    rank_ref_kernel<<<(unsigned)nblk, RR_WARPS * 32, smem, s>>>(flags, fgidx, rowvox, cons, lists,
                                                               hlists, llists, meta, perm, F, *cfg,
                                                               score);
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
