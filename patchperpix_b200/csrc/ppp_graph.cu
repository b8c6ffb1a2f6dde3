// Steps 5+6: patch affinity graph (computePatchGraph.cu:3-136), connected
// components over positive edges and painting (aff_patch_graph.py:31-40,
// graph_to_labeling.py:50-84).
#include <cub/cub.cuh>
#include "ppp_common.cuh"
#include "ppp_api.cuh"

// ---------------------------------------------------------------------------
// patch graph.  The reference runs ONE THREAD per patch pair over all P x P
// pixel pairs, 512 pairs per launch with a host sync in between
// (aff_patch_graph.py:137-159).  Here one CTA owns a pair: the voting pixels of
// both patches are compacted (in raster order) into shared memory and the
// |L1| x |L2| products are spread over the threads.
//
// The 20 % sub-sampling inside the window intersection uses a serial LCG that
// advances once per intersection pixel pair in loop order
// (computePatchGraph.cu:75-86).  The k-th such pair (1-based) sees
// rnd0 * a^k mod 2^32, and k = idx1 * n2 + idx2 + 1 with idx1 / idx2 the rank
// of the pixel among the intersection pixels of its patch and n2 the number of
// intersection pixels of patch 2, so every thread can jump to its state.
// ---------------------------------------------------------------------------
#define PG_THREADS 256
#define LCG_A 1103515245u

__device__ __forceinline__ uint32_t pow_u32(uint32_t a, uint32_t e)
{
    uint32_t r = 1u;
    while (e) { if (e & 1u) r *= a; a *= a; e >>= 1; }
    return r;
}

// ordered compaction of the voting pixels of patch (cz,cy,cx) into smem
// (computePatchGraph.cu:41-52): pred[mid][p] > TH and pred[po][c] > TH.
// Also ranks the pixels that lie inside the other patch's window.
template <class Src>
__device__ int pg_build_list(const Geo& g, const ppp_cfg& cfg, const Src& pred,
                             const uint8_t* __restrict__ flags, int cz, int cy, int cx,
                             int oz, int oy, int ox,       // the other centre
                             int16_t* s_po, int16_t* s_ii, int* s_scratch, int* n_inter)
{
    // s_scratch: [0..7] warp counts, [8..15] warp inter counts, [16] n, [17] n_inter
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t vc = ((int64_t)cz * g.Y + cy) * g.X + cx;
    if (threadIdx.x == 0) { s_scratch[16] = 0; s_scratch[17] = 0; }
    __syncthreads();
    for (int base = 0; base < g.P; base += (int)blockDim.x) {
        int po = base + threadIdx.x;
        bool vote = false, inter = false;
        if (po < g.P) {
            int qz, qy, qx;
            po_decode(g, po, qz, qy, qx);
            int z = cz + qz - g.rz, y = cy + qy - g.ry, x = cx + qx - g.rx;
            if (z >= 0 && z < g.Z && y >= 0 && y < g.Y && x >= 0 && x < g.X) {
                int pv = (z * g.Y + y) * g.X + x;
                vote = (flags[pv] & PPP_FLAG_FG) && pred.at(po, vc) > cfg.th_gt;
                inter = vote && abs(x - ox) <= g.rx && abs(y - oy) <= g.ry && abs(z - oz) <= g.rz;
            }
        }
        unsigned bv = __ballot_sync(0xffffffffu, vote);
        unsigned bi = __ballot_sync(0xffffffffu, inter);
        if (lane == 0) { s_scratch[w] = __popc(bv); s_scratch[8 + w] = __popc(bi); }
        __syncthreads();
        int off = s_scratch[16], offi = s_scratch[17];
        for (int i = 0; i < w; i++) { off += s_scratch[i]; offi += s_scratch[8 + i]; }
        if (vote) {
            unsigned lt = (1u << lane) - 1u;
            int idx = off + __popc(bv & lt);
            s_po[idx] = (int16_t)po;
            s_ii[idx] = inter ? (int16_t)(offi + __popc(bi & lt)) : (int16_t)-1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0, ti = 0;
            for (int i = 0; i < (int)blockDim.x / 32; i++) { t += s_scratch[i]; ti += s_scratch[8 + i]; }
            s_scratch[16] += t; s_scratch[17] += ti;
        }
        __syncthreads();
    }
    *n_inter = s_scratch[17];
    return s_scratch[16];
}

template <class Src>
__global__ void __launch_bounds__(PG_THREADS)
patch_graph_kernel(Src pred, const uint8_t* __restrict__ flags,
                   const int32_t* __restrict__ fgidx, const float* __restrict__ cons,
                   const uint32_t* __restrict__ pairs, ppp_cfg cfg, float* __restrict__ aff,
                   const int32_t* __restrict__ pair_org)
{
    Geo g = make_geo(cfg);
    extern __shared__ unsigned char smem_raw[];
    int16_t* s_po1 = (int16_t*)smem_raw;          // [P]
    int16_t* s_ii1 = s_po1 + g.P;                 // [P]
    int16_t* s_po2 = s_ii1 + g.P;                 // [P]
    int16_t* s_ii2 = s_po2 + g.P;                 // [P]
    uint32_t* s_pw2 = (uint32_t*)(s_ii2 + g.P);              // [P] a^(ii2+1)
    __shared__ int s_scr1[18], s_scr2[18];
    __shared__ double s_red[PG_THREADS / 32];
    __shared__ unsigned s_redc[PG_THREADS / 32];

    const int64_t id = blockIdx.x;
    const int z1c = pairs[id * 6], y1c = pairs[id * 6 + 1], x1c = pairs[id * 6 + 2];
    const int z2c = pairs[id * 6 + 3], y2c = pairs[id * 6 + 4], x2c = pairs[id * 6 + 5];
    const int oz_ = pair_org ? pair_org[id * 3] : 0, oy_ = pair_org ? pair_org[id * 3 + 1] : 0,
              ox_ = pair_org ? pair_org[id * 3 + 2] : 0;
    const uint32_t rnd0 = (uint32_t)(z1c - oz_) * (uint32_t)(z2c - oz_) * (uint32_t)(y1c - oy_) *
                          (uint32_t)(y2c - oy_) * (uint32_t)(x1c - ox_) * (uint32_t)(x2c - ox_);
    int ni1, ni2;
    const int n1 = pg_build_list(g, cfg, pred, flags, z1c, y1c, x1c, z2c, y2c, x2c,
                                 s_po1, s_ii1, s_scr1, &ni1);
    const int n2 = pg_build_list(g, cfg, pred, flags, z2c, y2c, x2c, z1c, y1c, x1c,
                                 s_po2, s_ii2, s_scr2, &ni2);
    const uint32_t a_n2 = pow_u32(LCG_A, (uint32_t)ni2);     // a^(n2)
    for (int j = threadIdx.x; j < n2; j += PG_THREADS)
        s_pw2[j] = s_ii2[j] >= 0 ? pow_u32(LCG_A, (uint32_t)s_ii2[j] + 1u) : 0u;
    __syncthreads();

    double acc = 0.0;
    unsigned cnt = 0;
    for (int i = 0; i < n1; i++) {
        int po1 = s_po1[i], ii1 = s_ii1[i];
        int qz, qy, qx;
        po_decode(g, po1, qz, qy, qx);
        const int z1 = z1c + qz - g.rz, y1 = y1c + qy - g.ry, x1 = x1c + qx - g.rx;
        const int64_t g1 = ((int64_t)z1 * g.Y + y1) * g.X + x1;
        const int row1 = fgidx[g1];
        const bool r1ok = row1 >= 0;
        // LCG state before this row of the loop nest: rnd0 * a^(ii1*n2)
        const uint32_t base_rnd = ii1 >= 0 ? rnd0 * pow_u32(a_n2, (uint32_t)ii1) : 0u;
        for (int j = threadIdx.x; j < n2; j += PG_THREADS) {
            int po2 = s_po2[j], ii2 = s_ii2[j];
            int pz, py, px;
            po_decode(g, po2, pz, py, px);
            const int z2 = z2c + pz - g.rz, y2 = y2c + py - g.ry, x2 = x2c + px - g.rx;
            if (ii1 >= 0 && ii2 >= 0) {
                uint32_t rnd = base_rnd * s_pw2[j];
                float rndT = (float)rnd / 4294967296.0f;
                if ((double)rndT > 0.2) continue;
            }
            const int64_t g2 = ((int64_t)z2 * g.Y + y2) * g.X + x2;
            int dz, dy, dx, rowb;
            if (g1 <= g2) { dz = z2 - z1; dy = y2 - y1; dx = x2 - x1; rowb = row1; }
            else { dz = z1 - z2; dy = y1 - y2; dx = x1 - x2; rowb = fgidx[g2]; }
            // computePatchGraph.cu:98-101 / 116-119 (index = offset + ps - 1 in [0, 2ps))
            if (dz < -(g.psz - 1) || dz > g.psz || dy < -(g.psy - 1) || dy > g.psy ||
                dx < -(g.psx - 1) || dx > g.psx) continue;
            cnt++;
            int k = k_of_offset(g, dz, dy, dx);
            if (k >= 0 && rowb >= 0 && (g1 <= g2 ? r1ok : true))
                acc += (double)cons[(int64_t)rowb * g.K + k];
        }
    }
    acc = warp_sum_d(acc);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { s_red[w] = acc; s_redc[w] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        unsigned c = 0;
        for (int i = 0; i < PG_THREADS / 32; i++) { t += s_red[i]; c += s_redc[i]; }
        float a = (float)t;
        aff[id] = (cfg.graph_flags & 1) ? a / (float)(c > 1 ? c : 1) : a;
    }
}

// ---------------------------------------------------------------------------
// reference-order patch affinity.  computePatchGraph.cu adds its up to P^2 terms
// into ONE float, serially (po1 outer, po2 inner); like in rankPatches.cu the
// rounding of that running sum is systematic (1e-3 relative at 41x41).  Only the
// sign matters for the connected-components labelling, but the mutex watershed
// (the flylight default) orders the edges by |aff|, so the default kernel
// reproduces the exact sequence of float additions: the CTA (2 warps) produces
// the terms of the flattened (i, j) sequence in chunks of PGR_CH into a double
// buffer (skipped terms as 0.0f, which leaves a float sum unchanged) while
// thread 0 adds the previous chunk in order.  The chain of ~n1*n2 dependent
// FADDs of a pair bounds its latency; many light CTAs per SM run side by side.
// ---------------------------------------------------------------------------
#ifndef PGR_THREADS
#define PGR_THREADS 160         // warp 0 adds, the other four warps produce (0.77 vs 0.81 ms with two)
#endif
#define PGR_STR2(x) #x
#define PGR_STR(x) PGR_STR2(x)
#define PGR_NT PGR_STR(PGR_THREADS)
#define PGR_PROD (PGR_THREADS - 32)
#define PGR_CH 256
#define PGR_BIAS 512

// ordered compaction of the voting pixels of patch (cz,cy,cx) (computePatchGraph.cu:41-52):
// s_q = coordinates relative to (rz0,ry0,rx0), biased, packed z<<20|y<<10|x; s_row =
// consensus row of the pixel; s_ii = rank among the pixels inside the other window, or -1
template <class Src>
__device__ int pgr_build_list(const Geo& g, const ppp_cfg& cfg, const Src& pred,
                              const uint8_t* __restrict__ flags, const int32_t* __restrict__ fgidx,
                              int cz, int cy, int cx, int oz, int oy, int ox,
                              int rz0, int ry0, int rx0,
                              int32_t* s_q, int32_t* s_row, int32_t* s_ii, int* s_scratch,
                              int* n_inter)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = PGR_THREADS / 32;
    const int64_t vc = ((int64_t)cz * g.Y + cy) * g.X + cx;
    if (threadIdx.x == 0) { s_scratch[16] = 0; s_scratch[17] = 0; }
    __syncthreads();
    for (int base = 0; base < g.P; base += PGR_THREADS) {
        int po = base + threadIdx.x;
        bool vote = false, inter = false;
        int q = 0, row = -1;
        if (po < g.P) {
            int qz, qy, qx;
            po_decode(g, po, qz, qy, qx);
            int z = cz + qz - g.rz, y = cy + qy - g.ry, x = cx + qx - g.rx;
            if (z >= 0 && z < g.Z && y >= 0 && y < g.Y && x >= 0 && x < g.X) {
                int pv = (z * g.Y + y) * g.X + x;
                vote = (flags[pv] & PPP_FLAG_FG) && pred.at(po, vc) > cfg.th_gt;
                inter = vote && abs(x - ox) <= g.rx && abs(y - oy) <= g.ry && abs(z - oz) <= g.rz;
                if (vote) {
                    row = fgidx[pv];
                    q = ((z - rz0 + PGR_BIAS) << 20) | ((y - ry0 + PGR_BIAS) << 10) |
                        (x - rx0 + PGR_BIAS);
                }
            }
        }
        unsigned bv = __ballot_sync(0xffffffffu, vote);
        unsigned bi = __ballot_sync(0xffffffffu, inter);
        if (lane == 0) { s_scratch[w] = __popc(bv); s_scratch[8 + w] = __popc(bi); }
        __syncthreads();
        int off = s_scratch[16], offi = s_scratch[17];
        for (int i = 0; i < w; i++) { off += s_scratch[i]; offi += s_scratch[8 + i]; }
        if (vote) {
            unsigned lt = (1u << lane) - 1u;
            int idx = off + __popc(bv & lt);
            s_q[idx] = q;
            s_row[idx] = row;
            if (s_ii != nullptr) s_ii[idx] = inter ? offi + __popc(bi & lt) : -1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0, ti = 0;
            for (int i = 0; i < nw; i++) { t += s_scratch[i]; ti += s_scratch[8 + i]; }
            s_scratch[16] += t; s_scratch[17] += ti;
        }
        __syncthreads();
    }
    *n_inter = s_scratch[17];
    return s_scratch[16];
}

// LCG = false: every pair has a zero coordinate product (2-D data, z = 0): the sub-sampling
// state stays 0 and nothing is ever skipped, so the factor tables are not needed
template <bool LCG, class Src>
__global__ void __launch_bounds__(PGR_THREADS)
patch_graph_ref_kernel(Src pred, const uint8_t* __restrict__ flags,
                       const int32_t* __restrict__ fgidx, const float* __restrict__ cons,
                       const uint32_t* __restrict__ pairs, ppp_cfg cfg, float* __restrict__ aff,
                       int32_t* __restrict__ lists, const int32_t* __restrict__ pair_org)
{
    Geo g = make_geo(cfg);
    // the two pixel lists of this pair live in GLOBAL scratch (written once, then read
    // through L1/L2 by the producer warps: the same i for consecutive terms, consecutive
    // j): the CTA needs almost no shared memory, so enough pairs are resident per SM to
    // run every serial chain of the launch side by side
    int32_t* s_q1 = lists + (int64_t)blockIdx.x * (LCG ? 6 : 4) * g.P;   // [P] packed coordinates,
    int32_t* s_row1 = s_q1 + g.P;                 //     consensus row,
    int32_t* s_q2 = s_row1 + g.P;
    int32_t* s_row2 = s_q2 + g.P;
    uint32_t* s_pw1 = (uint32_t*)(s_row2 + g.P);  //     LCG factor (0 = outside the intersection)
    uint32_t* s_pw2 = s_pw1 + g.P;                //     (only with LCG)
    __shared__ int s_scr1[18], s_scr2[18];
    __shared__ __align__(16) float s_val[2][PGR_CH];
    __shared__ unsigned s_cnt[PGR_PROD / 32];

    const int64_t id = blockIdx.x;
    const int tid = threadIdx.x;
    const int z1c = pairs[id * 6], y1c = pairs[id * 6 + 1], x1c = pairs[id * 6 + 2];
    const int z2c = pairs[id * 6 + 3], y2c = pairs[id * 6 + 4], x2c = pairs[id * 6 + 5];
    // pair_org: origin the reference's coordinates of this pair are relative to (its seed is
    // the product of the coordinates it was handed, computePatchGraph.cu:24-27)
    const int oz_ = pair_org ? pair_org[id * 3] : 0, oy_ = pair_org ? pair_org[id * 3 + 1] : 0,
              ox_ = pair_org ? pair_org[id * 3 + 2] : 0;
    const uint32_t rnd0 = (uint32_t)(z1c - oz_) * (uint32_t)(z2c - oz_) * (uint32_t)(y1c - oy_) *
                          (uint32_t)(y2c - oy_) * (uint32_t)(x1c - ox_) * (uint32_t)(x2c - ox_);
    // patches further apart than 2*ps on an axis share no slot: every term is skipped
    // by the offset test (:98-101), sum and count stay 0
    // (two window pixels are at least |delta| - 2*(ps//2) apart; beyond ps - 1 the slot index
    // falls outside the cube and the term is zero, :98-101, so the sum is exactly 0)
    if (abs(z2c - z1c) > 2 * (g.psz - 1) || abs(y2c - y1c) > 2 * (g.psy - 1) ||
        abs(x2c - x1c) > 2 * (g.psx - 1)) {
        if (tid == 0) aff[id] = 0.0f;
        return;
    }
    int ni1, ni2;
    const int n1 = pgr_build_list(g, cfg, pred, flags, fgidx, z1c, y1c, x1c, z2c, y2c, x2c,
                                  z1c, y1c, x1c, s_q1, s_row1, LCG ? (int32_t*)s_pw1 : nullptr,
                                  s_scr1, &ni1);
    const int n2 = pgr_build_list(g, cfg, pred, flags, fgidx, z2c, y2c, x2c, z1c, y1c, x1c,
                                  z1c, y1c, x1c, s_q2, s_row2, LCG ? (int32_t*)s_pw2 : nullptr,
                                  s_scr2, &ni2);
    if (LCG) {
        const uint32_t a_n2 = pow_u32(LCG_A, (uint32_t)ni2);
        // rank -> LCG factor, in place.  a is odd, so a^k is never 0: 0 marks "outside".
        // k-th intersection pair (1-based) sees rnd0 * a^k, k = ii1 * ni2 + ii2 + 1
        for (int i = tid; i < n1; i += PGR_THREADS) {
            int ii = (int)s_pw1[i];
            s_pw1[i] = ii >= 0 ? pow_u32(a_n2, (uint32_t)ii) : 0u;
        }
        for (int j = tid; j < n2; j += PGR_THREADS) {
            int ii = (int)s_pw2[j];
            s_pw2[j] = ii >= 0 ? pow_u32(LCG_A, (uint32_t)ii + 1u) : 0u;
        }
    }
    __syncthreads();

    const int total = n1 * n2;
    const int nch = (total + PGR_CH - 1) / PGR_CH;
    const int lane = tid & 31;
    if (tid >= 32) {
        // ---- producer warps: chunk c of the flattened (i, j) sequence -> s_val[c & 1] ----
        const int pt = tid - 32;                  // 0..63
        constexpr int TPT = PGR_CH / PGR_PROD;    // terms per producer thread and chunk
        unsigned cnt = 0;
        int ic = 0, jc = 0;                       // (i, j) of the first term of the chunk
        for (int c = 0; c < nch; c++) {
            if (c >= 2) {                                         // buffer c&1 was consumed
                if (c & 1) asm volatile("bar.sync 5, " PGR_NT ";\n" ::: "memory");
                else asm volatile("bar.sync 4, " PGR_NT ";\n" ::: "memory");
            }
            float* dst = s_val[c & 1];
            int64_t addr[TPT];
#pragma unroll
            for (int u = 0; u < TPT; u++) {
                int i = ic, j = jc + pt + PGR_PROD * u;
                while (j >= n2 && i < n1) { j -= n2; i++; }
                addr[u] = -1;
                if (i < n1) {
                    bool skip = false;
                    if (LCG) {
                        const uint32_t w1 = s_pw1[i], w2 = s_pw2[j];
                        if (w1 != 0u && w2 != 0u) {               // computePatchGraph.cu:75-86
                            uint32_t rnd = rnd0 * w1 * w2;
                            float rndT = (float)rnd / 4294967296.0f;
                            skip = (double)rndT > 0.2;
                        }
                    }
                    const int a = s_q1[i], b = s_q2[j];
                    int dz = (b >> 20) - (a >> 20), dy = ((b >> 10) & 1023) - ((a >> 10) & 1023),
                        dx = (b & 1023) - (a & 1023);
                    int rowb = s_row1[i];
                    // raster index g1 <= g2  <=>  (dz,dy,dx) >= 0 lexicographically (:89-124)
                    if (dz < 0 || (dz == 0 && (dy < 0 || (dy == 0 && dx < 0)))) {
                        dz = -dz; dy = -dy; dx = -dx; rowb = s_row2[j];
                    }
                    // :98-101 / :116-119 (index = offset + ps - 1 in [0, 2ps))
                    if (!skip && !(dz > g.psz || dy < -(g.psy - 1) || dy > g.psy ||
                                   dx < -(g.psx - 1) || dx > g.psx)) {
                        cnt++;
                        const int k = k_of_offset(g, dz, dy, dx);
                        if (k >= 0 && rowb >= 0) addr[u] = (int64_t)rowb * g.K + k;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < TPT; u++)
                dst[pt + PGR_PROD * u] = addr[u] >= 0 ? cons[addr[u]] : 0.0f;
            jc += PGR_CH;
            while (jc >= n2 && ic < n1) { jc -= n2; ic++; }
            __threadfence_block();
            if (c & 1) asm volatile("bar.arrive 3, " PGR_NT ";\n" ::: "memory");     // chunk c is ready
            else asm volatile("bar.arrive 2, " PGR_NT ";\n" ::: "memory");
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) s_cnt[(tid >> 5) - 1] = cnt;
    } else {
        // ---- adder warp: lane 0 adds the chunks in order ---------------------------------
        float acc = 0.0f;
        for (int c = 0; c < nch; c++) {
            if (c & 1) asm volatile("bar.sync 3, " PGR_NT ";\n" ::: "memory");       // wait for chunk c
            else asm volatile("bar.sync 2, " PGR_NT ";\n" ::: "memory");
            if (lane == 0) {
                const float4* v = (const float4*)s_val[c & 1];
#pragma unroll 8
                for (int q = 0; q < PGR_CH / 4; q++) {
                    float4 t = v[q];
                    acc += t.x; acc += t.y; acc += t.z; acc += t.w;
                }
            }
            if (c + 2 < nch) {                                    // buffer free
                __syncwarp();
                if (c & 1) asm volatile("bar.arrive 5, " PGR_NT ";\n" ::: "memory");
                else asm volatile("bar.arrive 4, " PGR_NT ";\n" ::: "memory");
            }
        }
        __syncwarp();
        if (lane == 0) s_val[0][0] = acc;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned c = 0;
        for (int q = 0; q < PGR_PROD / 32; q++) c += s_cnt[q];
        const float acc = s_val[0][0];
        aff[id] = (cfg.graph_flags & 1) ? acc / (float)(c > 1 ? c : 1) : acc;
    }
}

extern "C" int64_t ppp_patch_graph_scratch_bytes(const ppp_cfg* cfg, int64_t n)
{
    Geo g = make_geo(*cfg);
    return (n > 0 ? n : 0) * (int64_t)g.P * 24 + 256;     // two pixel lists per pair
}

template <class Src>
static int patch_graph_launch(Src src, const uint8_t* flags, const int32_t* fgidx,
                              const float* cons, const uint32_t* pairs, int64_t n,
                              const ppp_cfg* cfg, float* aff, void* scratch, void* stream,
                              const int32_t* pair_org = nullptr)
{
    if (n <= 0) return 0;
    Geo g = make_geo(*cfg);
    if (g.P > 32767) return ppp_fail(-1, "ppp_patch_graph: patch too large");
    if (!(cfg->graph_flags & 4)) {                // default: the reference's summation order
        if (g.psz > 128 || g.psy > 128 || g.psx > 128)
            return ppp_fail(-1, "ppp_patch_graph: patch axis larger than 128");
        // rnd0 = z*z2*y*y2*x*x2 (computePatchGraph.cu:24-27) is 0 for every pair of a
        // single-slice volume: no sub-sampling, no factor tables
        const bool lcg = g.Z > 1 || pair_org != nullptr;
        if (scratch == nullptr) return ppp_fail(-1, "ppp_patch_graph: scratch required");
        if (lcg)
            patch_graph_ref_kernel<true, Src><<<(unsigned)n, PGR_THREADS, 0, (cudaStream_t)stream>>>(
                src, flags, fgidx, cons, pairs, *cfg, aff, (int32_t*)scratch, pair_org);
        else
            patch_graph_ref_kernel<false, Src><<<(unsigned)n, PGR_THREADS, 0, (cudaStream_t)stream>>>(
                src, flags, fgidx, cons, pairs, *cfg, aff, (int32_t*)scratch, pair_org);
        return ppp_check("ppp_patch_graph(reference order)");
    }
    size_t smem = (size_t)g.P * 12 + 16;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(patch_graph_kernel<Src>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return ppp_fail((int)e, "ppp_patch_graph: smem attribute");
    }
    patch_graph_kernel<Src><<<(unsigned)n, PG_THREADS, smem, (cudaStream_t)stream>>>(
        src, flags, fgidx, cons, pairs, *cfg, aff, pair_org);
    return ppp_check("ppp_patch_graph");
}

extern "C" int ppp_patch_graph(const float* pred, const uint8_t* flags,
                               const int32_t* fgidx, const float* cons,
                               const uint32_t* pairs, int64_t n, const ppp_cfg* cfg,
                               float* aff, void* scratch, void* stream)
{
    Geo g = make_geo(*cfg);
    return patch_graph_launch(SrcDense{pred, g.V}, flags, fgidx, cons, pairs, n, cfg, aff,
                              scratch, stream);
}

extern "C" int ppp_patch_graph_rows(const uint16_t* patches, const int32_t* vox2row,
                                    const uint8_t* flags, const int32_t* fgidx,
                                    const float* cons, const uint32_t* pairs,
                                    const int32_t* pair_org, int64_t n,
                                    const ppp_cfg* cfg, float* aff, void* scratch, void* stream)
{
    Geo g = make_geo(*cfg);
    return patch_graph_launch(SrcRows{(const __half*)patches, vox2row, g.P}, flags, fgidx, cons,
                              pairs, n, cfg, aff, scratch, stream, pair_org);
}

// ---------------------------------------------------------------------------
// connected components over edges with aff > 0, numbered in the reference's
// order.  networkx yields components in node-insertion order of the positive
// graph, which is built by iterating the edges of the full graph
// (graph_to_labeling.py:50-54); a component therefore ranks by the smallest
// first-appearance position (2*pair + endpoint, over pairs with aff != 0,
// aff_patch_graph.py:35-39) among its nodes.  Lock-free union-find on the
// voxel index of the patch centres.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(int32_t* parent, int v)
{
    int p = parent[v];
    while (p != v) {
        int gp = parent[p];
        if (gp != p) parent[v] = gp;      // path halving (benign race)
        v = p; p = gp;
    }
    return v;
}

__device__ __forceinline__ void uf_union(int32_t* parent, int a, int b)
{
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }      // hook larger root under smaller
        int old = atomicCAS(&parent[a], a, b);
        if (old == a) return;
    }
}

__device__ __forceinline__ int pair_vox(const Geo& g, const uint32_t* pairs, int64_t i, int e)
{
    return (int)(((int64_t)pairs[i * 6 + 3 * e] * g.Y + pairs[i * 6 + 3 * e + 1]) * g.X +
                 pairs[i * 6 + 3 * e + 2]);
}

__global__ void cc_init_kernel(const uint32_t* __restrict__ pairs, int64_t n, ppp_cfg cfg,
                               int32_t* parent, int32_t* first, int32_t* key, int32_t* comp)
{
    Geo g = make_geo(cfg);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int e = 0; e < 2; e++) {
        int v = pair_vox(g, pairs, i, e);
        parent[v] = v; first[v] = 0x7fffffff; key[v] = 0x7fffffff; comp[v] = 0;
    }
}

__global__ void cc_union_kernel(const uint32_t* __restrict__ pairs, const float* __restrict__ aff,
                                int64_t n, ppp_cfg cfg, int32_t* parent, int32_t* first)
{
    Geo g = make_geo(cfg);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = aff[i];
    if (a == 0.0f) return;                           // aff_patch_graph.py:36
    int v0 = pair_vox(g, pairs, i, 0), v1 = pair_vox(g, pairs, i, 1);
    atomicMin(&first[v0], (int)(2 * i));
    atomicMin(&first[v1], (int)(2 * i + 1));
    if (a > 0.0f) uf_union(parent, v0, v1);          // graph_to_labeling.py:52
}

// key[root] = min first-appearance over the members that have a positive edge
__global__ void cc_key_kernel(const uint32_t* __restrict__ pairs, const float* __restrict__ aff,
                              int64_t n, ppp_cfg cfg, int32_t* parent, const int32_t* first,
                              int32_t* key)
{
    Geo g = make_geo(cfg);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!(aff[i] > 0.0f)) return;
    for (int e = 0; e < 2; e++) {
        int v = pair_vox(g, pairs, i, e);
        int r = uf_find(parent, v);
        atomicMin(&key[r], first[v]);
    }
}

// collect each root once: (key, root)
__global__ void cc_roots_kernel(const uint32_t* __restrict__ pairs, const float* __restrict__ aff,
                                int64_t n, ppp_cfg cfg, int32_t* parent, int32_t* key,
                                int32_t* comp, uint64_t* roots, int32_t* n_comp)
{
    Geo g = make_geo(cfg);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!(aff[i] > 0.0f)) return;
    int v = pair_vox(g, pairs, i, 0);
    int r = uf_find(parent, v);
    if (atomicExch(&comp[r], -1) == 0) {            // first visitor claims the root
        int slot = atomicAdd(n_comp, 1);
        roots[slot] = ((uint64_t)(uint32_t)key[r] << 32) | (uint32_t)r;
    }
}

__global__ void cc_number_kernel(const uint64_t* __restrict__ roots_sorted, int n_roots,
                                 int32_t* comp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_roots) return;
    comp[(uint32_t)(roots_sorted[i] & 0xffffffffu)] = -(i + 1) - 1;   // tagged: -(k+1)
}

__global__ void cc_assign_kernel(const uint32_t* __restrict__ pairs, const float* __restrict__ aff,
                                 int64_t n, ppp_cfg cfg, int32_t* parent, int32_t* comp)
{
    Geo g = make_geo(cfg);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!(aff[i] > 0.0f)) return;
    for (int e = 0; e < 2; e++) {
        int v = pair_vox(g, pairs, i, e);
        int r = uf_find(parent, v);
        if (r != v) comp[v] = -comp[r] - 1;         // roots keep the tagged value for now
    }
}

__global__ void cc_untag_kernel(const uint64_t* __restrict__ roots_sorted, int n_roots,
                                int32_t* comp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_roots) return;
    comp[(uint32_t)(roots_sorted[i] & 0xffffffffu)] = i + 1;
}

static size_t roots_sort_bytes(int64_t n)
{
    size_t tb = 0;
    cub::DeviceRadixSort::SortKeys((void*)nullptr, tb, (const uint64_t*)nullptr,
                                   (uint64_t*)nullptr, (int)(n > 0 ? n : 1));
    return tb;
}

extern "C" int64_t ppp_label_scratch_bytes(int64_t V, int64_t n)
{
    size_t tb = roots_sort_bytes(n);
    return (int64_t)(3 * ((V * 4 + 255) / 256) * 256 + 2 * ((n * 8 + 255) / 256 + 1) * 256 +
                     ((tb + 255) / 256) * 256 + 256);
}

extern "C" int ppp_label_cc(const uint32_t* pairs, const float* aff, int64_t n,
                            const ppp_cfg* cfg, int32_t* comp, int32_t* n_comp,
                            void* scratch, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(n_comp, 0, sizeof(int32_t), s);
    if (n <= 0) return ppp_check("ppp_label_cc");
    Geo g = make_geo(*cfg);
    size_t vb = ((g.V * 4 + 255) / 256) * 256;
    size_t rb = ((n * 8 + 255) / 256 + 1) * 256;
    char* base = (char*)scratch;
    int32_t* parent = (int32_t*)base;
    int32_t* first = (int32_t*)(base + vb);
    int32_t* key = (int32_t*)(base + 2 * vb);
    uint64_t* roots = (uint64_t*)(base + 3 * vb);
    uint64_t* roots_sorted = (uint64_t*)(base + 3 * vb + rb);
    void* sort_tmp = base + 3 * vb + 2 * rb;
    size_t tb = roots_sort_bytes(n);
    unsigned nb = (unsigned)((n + 255) / 256);
    cc_init_kernel<<<nb, 256, 0, s>>>(pairs, n, *cfg, parent, first, key, comp);
    cc_union_kernel<<<nb, 256, 0, s>>>(pairs, aff, n, *cfg, parent, first);
    cc_key_kernel<<<nb, 256, 0, s>>>(pairs, aff, n, *cfg, parent, first, key);
    cc_roots_kernel<<<nb, 256, 0, s>>>(pairs, aff, n, *cfg, parent, key, comp, roots, n_comp);
    // the number of roots is only known on the device; sort the full buffer
    // after padding it with +inf keys
    // (n is small: pairs of selected patches)
    cudaMemsetAsync(roots_sorted, 0xff, n * 8, s);
    int32_t h_ncomp = 0;
    cudaMemcpyAsync(&h_ncomp, n_comp, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    if (h_ncomp > 0) {
        cub::DeviceRadixSort::SortKeys(sort_tmp, tb, roots, roots_sorted, h_ncomp, 0, 64, s);
        unsigned rbk = (unsigned)((h_ncomp + 255) / 256);
        cc_number_kernel<<<rbk, 256, 0, s>>>(roots_sorted, h_ncomp, comp);
        cc_assign_kernel<<<nb, 256, 0, s>>>(pairs, aff, n, *cfg, parent, comp);
        cc_untag_kernel<<<rbk, 256, 0, s>>>(roots_sorted, h_ncomp, comp);
    }
    return ppp_check("ppp_label_cc");
}

// ---------------------------------------------------------------------------
// painting (graph_to_labeling.py:68-84): every pixel of a member patch with
// pred > patch_threshold takes the component number; later components
// overwrite earlier ones == the maximum wins.
// ---------------------------------------------------------------------------
__global__ void paint_kernel(const float* __restrict__ pred, const int32_t* __restrict__ nodes,
                             const int32_t* __restrict__ comp, ppp_cfg cfg,
                             int32_t* __restrict__ instances)
{
    Geo g = make_geo(cfg);
    const int vc = nodes[blockIdx.x];
    const int c = comp[vc];
    if (c <= 0) return;
    int cz, cy, cx;
    vox_decode(g, vc, cz, cy, cx);
    for (int po = threadIdx.x; po < g.P; po += blockDim.x) {
        int qz, qy, qx;
        po_decode(g, po, qz, qy, qx);
        int z = cz + qz - g.rz, y = cy + qy - g.ry, x = cx + qx - g.rx;
        if (z < 0 || z >= g.Z || y < 0 || y >= g.Y || x < 0 || x >= g.X) continue;
        if (pred[(int64_t)po * g.V + vc] > cfg.pt_gt)
            atomicMax(&instances[(z * g.Y + y) * g.X + x], c);
    }
}

extern "C" int ppp_paint(const float* pred, const int32_t* nodes, int64_t m,
                         const int32_t* comp, const ppp_cfg* cfg, int32_t* instances,
                         void* stream)
{
    if (m <= 0) return 0;
    paint_kernel<<<(unsigned)m, 128, 0, (cudaStream_t)stream>>>(pred, nodes, comp, *cfg,
                                                                 instances);
    return ppp_check("ppp_paint");
}

// one channel per component (`one_instance_per_channel`, graph_to_labeling.py:57-95):
// channel c-1 of instances [n_comp][V] holds component c painted alone
__global__ void paint_channels_kernel(const float* __restrict__ pred,
                                      const int32_t* __restrict__ nodes,
                                      const int32_t* __restrict__ comp, ppp_cfg cfg,
                                      int32_t* __restrict__ instances)
{
    Geo g = make_geo(cfg);
    const int vc = nodes[blockIdx.x];
    const int c = comp[vc];
    if (c <= 0) return;
    int cz, cy, cx;
    vox_decode(g, vc, cz, cy, cx);
    int32_t* chan = instances + (int64_t)(c - 1) * g.V;
    for (int po = threadIdx.x; po < g.P; po += blockDim.x) {
        int qz, qy, qx;
        po_decode(g, po, qz, qy, qx);
        int z = cz + qz - g.rz, y = cy + qy - g.ry, x = cx + qx - g.rx;
        if (z < 0 || z >= g.Z || y < 0 || y >= g.Y || x < 0 || x >= g.X) continue;
        if (pred[(int64_t)po * g.V + vc] > cfg.pt_gt) chan[(z * g.Y + y) * g.X + x] = c;
    }
}

extern "C" int ppp_paint_channels(const float* pred, const int32_t* nodes, int64_t m,
                                  const int32_t* comp, const ppp_cfg* cfg, int32_t* instances,
                                  void* stream)
{
    if (m <= 0) return 0;
    paint_channels_kernel<<<(unsigned)m, 128, 0, (cudaStream_t)stream>>>(pred, nodes, comp, *cfg,
                                                                          instances);
    return ppp_check("ppp_paint_channels");
}

// same, with the member patches handed over as a compact [m][P] array (the
// blockwise path reads only the selected patches of a volume that does not fit
// the device, stitch_patch_graph.py:380-385)
__global__ void paint_patches_kernel(const float* __restrict__ patches,
                                     const int32_t* __restrict__ nodes,
                                     const int32_t* __restrict__ comp, ppp_cfg cfg,
                                     int32_t* __restrict__ instances)
{
    Geo g = make_geo(cfg);
    const int vc = nodes[blockIdx.x];
    const int c = comp[vc];
    if (c <= 0) return;
    int cz, cy, cx;
    vox_decode(g, vc, cz, cy, cx);
    for (int po = threadIdx.x; po < g.P; po += blockDim.x) {
        int qz, qy, qx;
        po_decode(g, po, qz, qy, qx);
        int z = cz + qz - g.rz, y = cy + qy - g.ry, x = cx + qx - g.rx;
        if (z < 0 || z >= g.Z || y < 0 || y >= g.Y || x < 0 || x >= g.X) continue;
        if (patches[(int64_t)blockIdx.x * g.P + po] > cfg.pt_gt)
            atomicMax(&instances[(z * g.Y + y) * g.X + x], c);
    }
}

extern "C" int ppp_paint_patches(const float* patches, const int32_t* nodes, int64_t m,
                                 const int32_t* comp, const ppp_cfg* cfg, int32_t* instances,
                                 void* stream)
{
    if (m <= 0) return 0;
    paint_patches_kernel<<<(unsigned)m, 128, 0, (cudaStream_t)stream>>>(patches, nodes, comp,
                                                                         *cfg, instances);
    return ppp_check("ppp_paint_patches");
}

// ---------------------------------------------------------------------------
// painting from float16 patch ROWS into a sub-volume (sharded blockwise path):
// node i sits at (cz,cy,cx) in the coordinates of `instances` (shape cfg Z,Y,X;
// the centre itself may lie outside, its window is clipped), its patch is row
// node_row[i] of `patches`, its component node_label[i] (<= 0: not painted).
// ---------------------------------------------------------------------------
__global__ void paint_rows_kernel(const __half* __restrict__ patches,
                                  const int32_t* __restrict__ node_row,
                                  const int32_t* __restrict__ node_zyx,
                                  const int32_t* __restrict__ node_label, ppp_cfg cfg,
                                  int32_t* __restrict__ instances)
{
    Geo g = make_geo(cfg);
    const int64_t i = blockIdx.x;
    const int c = node_label[i];
    const int r = node_row[i];
    if (c <= 0 || r < 0) return;
    const int cz = node_zyx[3 * i], cy = node_zyx[3 * i + 1], cx = node_zyx[3 * i + 2];
    for (int po = threadIdx.x; po < g.P; po += blockDim.x) {
        int qz, qy, qx;
        po_decode(g, po, qz, qy, qx);
        int z = cz + qz - g.rz, y = cy + qy - g.ry, x = cx + qx - g.rx;
        if (z < 0 || z >= g.Z || y < 0 || y >= g.Y || x < 0 || x >= g.X) continue;
        if (__half2float(patches[(int64_t)r * g.P + po]) > cfg.pt_gt)
            atomicMax(&instances[((int64_t)z * g.Y + y) * g.X + x], c);
    }
}

extern "C" int ppp_paint_rows(const uint16_t* patches, const int32_t* node_row,
                              const int32_t* node_zyx, const int32_t* node_label, int64_t m,
                              const ppp_cfg* cfg, int32_t* instances, void* stream)
{
    if (m <= 0) return 0;
    paint_rows_kernel<<<(unsigned)m, 128, 0, (cudaStream_t)stream>>>(
        (const __half*)patches, node_row, node_zyx, node_label, *cfg, instances);
    return ppp_check("ppp_paint_rows");
}
