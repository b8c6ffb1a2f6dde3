// Steps 3+4: greedy foreground cover and greedy set-cover thinning on the
// device (foreground_cover.py:15-256 are serial python loops over numpy
// windows).  Both are restated so that the serial walk disappears where the
// result allows it: the threshold-0 cover is "first coverer of every voxel"
// (cover_first_*), the set cover runs in rounds of local maxima
// (thin_rounds_kernel); the one-decision-per-step kernels (cover_kernel for the
// dense threshold schedule, thin_kernel as cross-check) stay in one CTA.
// What makes the counting fast is the data layout: the mask to cover is a bit volume
// (one 32-bit word per 32 x-voxels, resident in shared memory when it fits)
// and every candidate patch is a P-bit string (`fcmask`, patch > fc_threshold),
// so "how many uncovered voxels would this patch cover" is a handful of
// funnel-shift / AND / POPC per patch row instead of a P-element numpy op.
#include <cub/cub.cuh>
#include "ppp_common.cuh"
#include "ppp_api.cuh"

struct BitVol {
    uint32_t* w;      // [Z*Y][WX]
    int WX;           // words per x-row
};

__host__ __device__ inline int bitvol_wx(int X) { return (X + 31) / 32; }

// 32 mask bits starting at bit position `bit` of x-row `r`
__device__ __forceinline__ uint32_t bv_get32(const BitVol& b, int r, int bit)
{
    const int wi = bit >> 5;
    const uint32_t* p = b.w + (int64_t)r * b.WX + wi;
    return __funnelshift_r(p[0], (wi + 1 < b.WX) ? p[1] : 0u, bit & 31);
}

__device__ __forceinline__ void bv_clear32(const BitVol& b, int r, int bit, uint32_t m)
{
    const int wi = bit >> 5;
    uint32_t* p = b.w + (int64_t)r * b.WX + wi;
    int s = bit & 31;
    p[0] &= ~(m << s);
    if (s && wi + 1 < b.WX) p[1] &= ~(m >> (32 - s));
}

// 32 bits of a patch bit string starting at bit `bit` (zero past the end)
__device__ __forceinline__ uint32_t pm_get32(const uint32_t* __restrict__ pm, int W, int bit)
{
    int a = bit >> 5;
    uint32_t lo = a < W ? pm[a] : 0u;
    uint32_t hi = (a + 1) < W ? pm[a + 1] : 0u;
    return __funnelshift_r(lo, hi, bit & 31);
}

// bits of [x0, x0+nb) that lie inside the radslice x-range [rx, X-rx)
__device__ __forceinline__ uint32_t rad_xmask(const Geo& g, int x0, int nb)
{
    int lo = max(g.rx - x0, 0), hi = min(g.X - g.rx - x0, nb);
    if (hi <= lo) return 0u;
    uint32_t m = (hi - lo) >= 32 ? 0xffffffffu : ((1u << (hi - lo)) - 1u);
    return m << lo;
}

// One warp: count (and optionally clear) the still-uncovered voxels of the
// window of centre (cz,cy,cx) that the patch bit string `pm` marks.
// Returns the warp-wide count; *rad_cleared (lane-local partial, only when
// clearing) counts cleared voxels inside the radslice.
template <bool CLEAR>
__device__ __forceinline__ int patch_window_count(const Geo& g, const BitVol& bv,
                                                  const uint32_t* __restrict__ pm,
                                                  int cz, int cy, int cx, int lane,
                                                  int* rad_cleared)
{
    int cnt = 0;
    const int nrows = g.psz * g.psy;
    for (int rr = lane; rr < nrows; rr += 32) {
        int qz = rr / g.psy, qy = rr - qz * g.psy;
        int z = cz - g.rz + qz, y = cy - g.ry + qy;
        int r = z * g.Y + y;
        bool row_in_rad = z >= g.rz && z < g.Z - g.rz && y >= g.ry && y < g.Y - g.ry;
        for (int j = 0; j < g.psx; j += 32) {
            int nb = min(32, g.psx - j);
            uint32_t keep = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
            int x0 = cx - g.rx + j;
            uint32_t m = bv_get32(bv, r, x0) & pm_get32(pm, g.W, rr * g.psx + j) & keep;
            cnt += __popc(m);
            if (CLEAR && m) {
                bv_clear32(bv, r, x0, m);
                if (row_in_rad) *rad_cleared += __popc(m & rad_xmask(g, x0, nb));
            }
        }
    }
    return CLEAR ? cnt : warp_sum_i(cnt);
}

// pack the u8 mask volume into the bit volume; returns (via *remaining) the
// number of set voxels inside the radslice
__device__ void bitvol_init(const Geo& g, const BitVol& bv, const uint8_t* __restrict__ mask,
                            int* remaining)
{
    const int rows = g.Z * g.Y;
    int local = 0;
    for (int64_t i = threadIdx.x; i < (int64_t)rows * bv.WX; i += blockDim.x) {
        int r = (int)(i / bv.WX), wi = (int)(i - (int64_t)r * bv.WX);
        uint32_t word = 0;
        int x0 = wi * 32;
        for (int t = 0; t < 32; t++) {
            int x = x0 + t;
            if (x < g.X && mask[(int64_t)r * g.X + x]) word |= 1u << t;
        }
        bv.w[i] = word;
        int z = r / g.Y, y = r - z * g.Y;
        if (z >= g.rz && z < g.Z - g.rz && y >= g.ry && y < g.Y - g.ry)
            local += __popc(word & rad_xmask(g, x0, 32));
    }
    atomicAdd(remaining, local);
}

extern "C" int64_t ppp_cover_scratch_bytes(const ppp_cfg* cfg)
{
    Geo g = make_geo(*cfg);
    int64_t bitvol = (int64_t)g.Z * g.Y * bitvol_wx(g.X) * 4;
    int64_t first = g.V * 4 + 16;                 // minpos[V] + r* (threshold-0 path)
    return (bitvol > first ? bitvol : first) + 256;
}

// One THREAD: still-uncovered voxels of the window of centre (cz,cy,cx) that the patch
// bit string `pm` marks (no clearing).  Used where many candidates are counted side by
// side: one candidate per thread keeps 1024 independent load chains in flight.
__device__ __forceinline__ int patch_window_count_thread(const Geo& g, const BitVol& bv,
                                                         const uint32_t* __restrict__ pm,
                                                         int cz, int cy, int cx)
{
    int cnt = 0;
    int bit = 0;
    for (int qz = 0; qz < g.psz; qz++)
        for (int qy = 0; qy < g.psy; qy++) {
            const int r = (cz - g.rz + qz) * g.Y + (cy - g.ry + qy);
            for (int j = 0; j < g.psx; j += 32) {
                int nb = min(32, g.psx - j);
                uint32_t keep = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
                uint32_t m = pm_get32(pm, g.W, bit + j) & keep;
                if (m) cnt += __popc(m & bv_get32(bv, r, cx - g.rx + j));
            }
            bit += g.psx;
        }
    return cnt;
}

#ifdef COVER_PROFILE
__device__ unsigned long long cover_prof[8];
extern "C" int ppp_debug_cover_prof(unsigned long long* out, int reset)
{
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(cover_prof, z, sizeof(z)); return 0; }
    return (int)cudaMemcpyFromSymbol(out, cover_prof, 8 * sizeof(unsigned long long));
}
#define COVER_TICK_INIT long long t_prof = clock64();
#define COVER_TICK(i) if (threadIdx.x == 0) { long long now_ = clock64(); cover_prof[i] += (unsigned long long)(now_ - t_prof); t_prof = now_; }
#define COVER_COUNT(i, v) if (threadIdx.x == 0) cover_prof[i] += (unsigned long long)(v);
#else
#define COVER_TICK_INIT
#define COVER_TICK(i)
#define COVER_COUNT(i, v)
#endif

#define COVER_THREADS 1024
#define THIN_THREADS 1024
#define SMEM_BITVOL_MAX (192 * 1024)

// ---------------------------------------------------------------------------
// greedy cover (foreground_cover.py:111-180).  A patch is selected iff it
// covers more than pixTh still-uncovered voxels; the walk over the ranked list
// stops when nothing inside the radslice is left (:128).
// The serial walk is kept exact but shortened: coverage counts can only fall
// while the mask shrinks, so a candidate whose count is already <= pixTh at
// the start of its chunk is rejected for good.  All 32 warps count a chunk of
// 1024 candidates in parallel (phase A), then warp 0 replays only the
// survivors in rank order against the live mask (phase B).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(COVER_THREADS)
cover_kernel(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ overlap,
             const int32_t* __restrict__ order, int64_t n,
             const int32_t* __restrict__ fgidx, const uint32_t* __restrict__ fcmask,
             ppp_cfg cfg, const int32_t* __restrict__ pix_ths, int n_pix,
             uint8_t* __restrict__ selected, uint32_t* gbits, int use_smem)
{
    Geo g = make_geo(cfg);
    extern __shared__ uint32_t s_bits[];
    __shared__ int s_remaining;
    __shared__ int s_cnt[COVER_THREADS];
    __shared__ int16_t s_surv[COVER_THREADS];
    __shared__ int s_nsurv, s_cnt2[32], s_vc[32];
    uint32_t* s_fc = s_bits + (use_smem ? (size_t)g.Z * g.Y * bitvol_wx(g.X) : 0);   // [32][W]
    BitVol bv;
    bv.WX = bitvol_wx(g.X);
    bv.w = use_smem ? s_bits : gbits;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    COVER_TICK_INIT
    if (threadIdx.x == 0) s_remaining = 0;
    __syncthreads();
    bitvol_init(g, bv, mask, &s_remaining);
    __syncthreads();
    COVER_TICK(0)
    for (int pi = 0; pi < n_pix; pi++) {
        const int pix_th = pix_ths[pi];
        for (int64_t r0 = 0; r0 < n; r0 += COVER_THREADS) {
            if (s_remaining <= 0) break;
            COVER_COUNT(4, 1)
            // phase A: parallel upper bounds, one candidate per thread
            {
                int64_t r = r0 + threadIdx.x;
                int cnt = -1;
                if (r < n && !selected[r]) {
                    int vc = order[r];
                    int row = fgidx[vc];
                    if (!(overlap != nullptr && overlap[vc]) && row >= 0) {   // :144
                        int cz, cy, cx;
                        vox_decode(g, vc, cz, cy, cx);
                        cnt = patch_window_count_thread(g, bv, fcmask + (int64_t)row * g.W,
                                                        cz, cy, cx);
                    }
                }
                s_cnt[threadIdx.x] = cnt;
            }
            __syncthreads();
            COVER_TICK(1)
            // phase B: replay of the survivors in rank order.  Sub-batches of up to 32
            // survivors: every warp stages the bit string of one survivor in shared
            // memory and re-counts it against the live mask in parallel; warp 0 then
            // walks the sub-batch in order, and only a survivor whose window overlaps
            // a patch selected earlier in the same sub-batch is counted again.
            if (threadIdx.x == 0) s_nsurv = 0;
            __syncthreads();
            if (w == 0) {
                int nsv = 0;
                for (int base = 0; base < COVER_THREADS; base += 32) {
                    bool sv = s_cnt[base + lane] > pix_th;
                    unsigned bal = __ballot_sync(0xffffffffu, sv);
                    if (sv) s_surv[nsv + __popc(bal & ((1u << lane) - 1u))] = (int16_t)(base + lane);
                    nsv += __popc(bal);
                }
                if (lane == 0) s_nsurv = nsv;
            }
            __syncthreads();
            const int nsurv = s_nsurv;
            COVER_TICK(2)
            COVER_COUNT(5, nsurv)
            for (int sb = 0; sb < nsurv; sb += 32) {
                if (s_remaining <= 0) break;
                const int nb = min(32, nsurv - sb);
                if (w < nb) {
                    const int64_t r = r0 + s_surv[sb + w];
                    const int vc = order[r];
                    const uint32_t* pm = fcmask + (int64_t)fgidx[vc] * g.W;
                    for (int q = lane; q < g.W; q += 32) s_fc[w * g.W + q] = pm[q];
                    __syncwarp();
                    int cz, cy, cx;
                    vox_decode(g, vc, cz, cy, cx);
                    int cnt = patch_window_count<false>(g, bv, s_fc + w * g.W, cz, cy, cx, lane,
                                                        nullptr);
                    if (lane == 0) { s_cnt2[w] = cnt; s_vc[w] = vc; }
                }
                __syncthreads();
                if (w == 0) {
                    int remaining = s_remaining;
                    unsigned stale = 0;                          // survivors to count again
                    for (int k = 0; k < nb && remaining > 0; k++) {
                        const int vc = s_vc[k];
                        int cz, cy, cx;
                        vox_decode(g, vc, cz, cy, cx);
                        const uint32_t* pm = s_fc + k * g.W;
                        int cnt = s_cnt2[k];
                        if ((stale >> k) & 1u)
                            cnt = patch_window_count<false>(g, bv, pm, cz, cy, cx, lane, nullptr);
                        if (cnt > pix_th) {
                            int rc = 0;
                            patch_window_count<true>(g, bv, pm, cz, cy, cx, lane, &rc);
                            __syncwarp();
                            remaining -= warp_sum_i(rc);
                            if (lane == 0) selected[r0 + s_surv[sb + k]] = 1;
                            // later survivors whose window intersects this one are stale
                            bool hit = false;
                            if (lane > k && lane < nb) {
                                int oz, oy, ox;
                                vox_decode(g, s_vc[lane], oz, oy, ox);
                                hit = abs(oz - cz) < g.psz && abs(oy - cy) < g.psy &&
                                      abs(ox - cx) < g.psx;
                            }
                            stale |= __ballot_sync(0xffffffffu, hit);
                        }
                    }
                    if (lane == 0) s_remaining = remaining;
                }
                __syncthreads();
                COVER_COUNT(6, 1)
            }
            COVER_TICK(3)
        }
        if (s_remaining < 1) break;                                  // :50-51
    }
}

// ---------------------------------------------------------------------------
// greedy cover with pixel threshold 0 (`select_patches_for_sparse_data`, the
// flylight default, foreground_cover.py:35-36) without the serial walk.
// A candidate is selected iff at its turn some mask voxel under its patch is
// still uncovered.  Let first(p) be the best-ranked (non-skipped) candidate whose
// patch marks voxel p.  first(p) is always selected (nobody before it touches p),
// and a selected candidate must be first(p) for the voxel that made it count
// (an earlier coverer of p would itself have been selected and cleared p).  So
//     selected = { first(p) : p in mask }  up to the position r* where the walk
// stops: the last first(p) over the voxels inside the radslice (:128-129), or the
// end of the list if one of them cannot be covered.  Four data-parallel passes.
// ---------------------------------------------------------------------------
__global__ void cover_first_init_kernel(int32_t* __restrict__ minpos, int64_t V, int32_t* rstar)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V) minpos[i] = 0x7fffffff;
    if (i == 0) *rstar = -1;
}

__global__ void __launch_bounds__(256)
cover_first_mark_kernel(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ overlap,
                        const int32_t* __restrict__ order, int64_t n,
                        const int32_t* __restrict__ fgidx, const uint32_t* __restrict__ fcmask,
                        ppp_cfg cfg, int32_t* __restrict__ minpos)
{
    Geo g = make_geo(cfg);
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const int vc = order[r];
    const int row = fgidx[vc];
    if ((overlap != nullptr && overlap[vc]) || row < 0) return;             // :141-145
    const uint32_t* pm = fcmask + (int64_t)row * g.W;
    int cz, cy, cx;
    vox_decode(g, vc, cz, cy, cx);
    const int nrows = g.psz * g.psy;
    for (int rr = lane; rr < nrows; rr += 32) {
        int qz = rr / g.psy, qy = rr - qz * g.psy;
        int z = cz - g.rz + qz, y = cy - g.ry + qy;
        if (z < 0 || z >= g.Z || y < 0 || y >= g.Y) continue;
        const int64_t line = ((int64_t)z * g.Y + y) * g.X;
        for (int j = 0; j < g.psx; j += 32) {
            int nb = min(32, g.psx - j);
            uint32_t keep = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
            uint32_t m = pm_get32(pm, g.W, rr * g.psx + j) & keep;
            while (m) {
                int b = __ffs(m) - 1;
                m &= m - 1;
                int x = cx - g.rx + j + b;
                if (x >= 0 && x < g.X && mask[line + x]) atomicMin(&minpos[line + x], (int)r);
            }
        }
    }
}

// r* = where the reference's walk stops
__global__ void cover_first_stop_kernel(const uint8_t* __restrict__ mask,
                                        const int32_t* __restrict__ minpos, ppp_cfg cfg,
                                        int64_t n, int32_t* rstar)
{
    Geo g = make_geo(cfg);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int v = -1;
    if (i < g.V && mask[i]) {
        int z, y, x;
        vox_decode(g, (int)i, z, y, x);
        if (z >= g.rz && z < g.Z - g.rz && y >= g.ry && y < g.Y - g.ry && x >= g.rx &&
            x < g.X - g.rx) {
            int mp = minpos[i];
            v = mp == 0x7fffffff ? (int)(n - 1) : mp;
        }
    }
    v = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v >= 0) atomicMax(rstar, v);
}

__global__ void cover_first_select_kernel(const uint8_t* __restrict__ mask,
                                          const int32_t* __restrict__ minpos, int64_t V,
                                          const int32_t* __restrict__ rstar,
                                          uint8_t* __restrict__ selected)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V || !mask[i]) return;
    int mp = minpos[i];
    if (mp != 0x7fffffff && mp <= *rstar) selected[mp] = 1;
}

extern "C" int ppp_cover(const uint8_t* mask, const uint8_t* overlap, const int32_t* order,
                         int64_t n, const int32_t* fgidx, const uint32_t* fcmask,
                         const ppp_cfg* cfg, const int32_t* pix_ths, int32_t n_pix,
                         uint8_t* selected, void* scratch, void* stream)
{
    if (n <= 0) return 0;
    Geo g = make_geo(*cfg);
    if (pix_ths == nullptr) {
        // pixel threshold 0 only (select_patches_for_sparse_data): data-parallel form
        cudaStream_t st = (cudaStream_t)stream;
        int32_t* minpos = (int32_t*)scratch;
        int32_t* rstar = minpos + g.V;
        const unsigned nb = (unsigned)((g.V + 255) / 256);
        cover_first_init_kernel<<<nb, 256, 0, st>>>(minpos, g.V, rstar);
        cover_first_mark_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(
            mask, overlap, order, n, fgidx, fcmask, *cfg, minpos);
        cover_first_stop_kernel<<<nb, 256, 0, st>>>(mask, minpos, *cfg, n, rstar);
        cover_first_select_kernel<<<nb, 256, 0, st>>>(mask, minpos, g.V, rstar, selected);
        return ppp_check("ppp_cover(first coverer)");
    }
    size_t bytes = (size_t)g.Z * g.Y * bitvol_wx(g.X) * 4;
    int use_smem = bytes <= SMEM_BITVOL_MAX;
    size_t smem = (use_smem ? bytes : 0) + (size_t)32 * g.W * 4;
    {
        cudaError_t e = cudaFuncSetAttribute(cover_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             SMEM_BITVOL_MAX + 12 * 1024);
        if (e != cudaSuccess) return ppp_fail((int)e, "ppp_cover: smem attribute");
    }
    cover_kernel<<<1, COVER_THREADS, smem, (cudaStream_t)stream>>>(
        mask, overlap, order, n, fgidx, fcmask, *cfg, pix_ths, n_pix, selected,
        (uint32_t*)scratch, use_smem);
    return ppp_check("ppp_cover");
}

// ---------------------------------------------------------------------------
// thinning = greedy set cover (foreground_cover.py:183-256, thin_cover_use_kd
// False): repeatedly keep the patch that still covers the most uncovered
// voxels (first maximum in list order, np.argmax), remove its voxels, until
// the radslice is covered.  |S_i| after the reference's set subtraction equals
// popc(patch bits & running mask), so only the patches whose window overlaps
// the chosen one need re-counting.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(THIN_THREADS)
thin_kernel(const uint8_t* __restrict__ mask, const int32_t* __restrict__ sel, int64_t m,
            const int32_t* __restrict__ fgidx, const uint32_t* __restrict__ fcmask,
            ppp_cfg cfg, uint8_t* __restrict__ keep, uint32_t* gbits, int32_t* gcounts,
            int use_smem)
{
    Geo g = make_geo(cfg);
    extern __shared__ uint32_t s_bits[];
    __shared__ int s_remaining;
    __shared__ int s_best_cnt[THIN_THREADS / 32];
    __shared__ int s_best_idx[THIN_THREADS / 32];
    __shared__ int s_best;
    BitVol bv;
    bv.WX = bitvol_wx(g.X);
    bv.w = use_smem ? s_bits : gbits;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = THIN_THREADS / 32;
    if (threadIdx.x == 0) s_remaining = 0;
    __syncthreads();
    bitvol_init(g, bv, mask, &s_remaining);
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) keep[i] = 0;
    __syncthreads();
    // initial |S_i|
    for (int64_t i = w; i < m; i += nw) {
        int vc = sel[i], row = fgidx[vc];
        int cnt = 0;
        if (row >= 0) {
            int cz, cy, cx;
            vox_decode(g, vc, cz, cy, cx);
            cnt = patch_window_count<false>(g, bv, fcmask + (int64_t)row * g.W, cz, cy, cx,
                                            lane, nullptr);
        }
        if (lane == 0) gcounts[i] = cnt;
    }
    __syncthreads();
    while (s_remaining > 0) {
        // first maximum
        int bc = -1, bi = 0x7fffffff;
        for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
            int c = gcounts[i];
            if (c > bc) { bc = c; bi = (int)i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            int oc = __shfl_xor_sync(0xffffffffu, bc, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oc > bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
        }
        if (lane == 0) { s_best_cnt[w] = bc; s_best_idx[w] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int c = s_best_cnt[0], ix = s_best_idx[0];
            for (int i = 1; i < nw; i++)
                if (s_best_cnt[i] > c || (s_best_cnt[i] == c && s_best_idx[i] < ix)) {
                    c = s_best_cnt[i]; ix = s_best_idx[i];
                }
            s_best = c > 0 ? ix : -1;   // nothing left that covers anything: stop
        }
        __syncthreads();
        const int best = s_best;
        if (best < 0) break;            // the reference would spin here
        int bz, by, bx;
        vox_decode(g, sel[best], bz, by, bx);
        if (w == 0) {
            int rc = 0;
            patch_window_count<true>(g, bv, fcmask + (int64_t)fgidx[sel[best]] * g.W,
                                     bz, by, bx, lane, &rc);
            rc = warp_sum_i(rc);
            if (lane == 0) { keep[best] = 1; s_remaining -= rc; gcounts[best] = 0; }
        }
        __syncthreads();
        // re-count the patches whose window intersects the chosen one
        for (int64_t i = w; i < m; i += nw) {
            if (gcounts[i] == 0) continue;
            int vc = sel[i];
            int cz, cy, cx;
            vox_decode(g, vc, cz, cy, cx);
            if (abs(cz - bz) >= g.psz || abs(cy - by) >= g.psy || abs(cx - bx) >= g.psx) continue;
            int row = fgidx[vc];
            int cnt = patch_window_count<false>(g, bv, fcmask + (int64_t)row * g.W, cz, cy, cx,
                                                lane, nullptr);
            if (lane == 0) gcounts[i] = cnt;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// greedy set cover in ROUNDS instead of one selection per step.
//
// The reference picks the patch with the largest remaining set (first maximum),
// removes its voxels from every other set, and repeats until nothing inside the
// radslice is left (foreground_cover.py:196-254).  Observations:
//  * counts only fall, and only for patches whose window intersects the chosen
//    one ("neighbours");
//  * hence the counts at selection time are non-increasing along the serial
//    run, i.e. the serial order of the selected patches is exactly their order
//    by (count at selection, descending; index, ascending);
//  * a patch that currently beats all its live neighbours in that order (a
//    LOCAL maximum) is selected by the serial run before any of its neighbours,
//    with exactly its current count, provided the run gets that far: nobody who
//    could lower its count can become the global maximum before it.
// So all local maxima of a round are selected at once (their windows are
// pairwise disjoint), their neighbours are re-counted, and the only thing that
// needs the serial order is WHERE the run stops: the first position of the
// (count, index)-sorted selection at which the voxels newly removed from the
// radslice add up to its initial content.  Rounds continue until no live patch
// could still sort before that position; later selections are dropped.
// One CTA; per-patch state in global scratch (L1/L2), bit volume as before.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool thin_better(int ca, int ia, int cb, int ib)
{
    return ca > cb || (ca == cb && ia < ib);       // a is chosen before b
}

// like patch_window_count<true>, but safe when several warps clear disjoint
// windows that may share 32-bit words of the bit volume
__device__ __forceinline__ int patch_window_clear_atomic(const Geo& g, const BitVol& bv,
                                                         const uint32_t* __restrict__ pm,
                                                         int cz, int cy, int cx, int lane)
{
    int rad = 0;
    const int nrows = g.psz * g.psy;
    for (int rr = lane; rr < nrows; rr += 32) {
        int qz = rr / g.psy, qy = rr - qz * g.psy;
        int z = cz - g.rz + qz, y = cy - g.ry + qy;
        int r = z * g.Y + y;
        bool row_in_rad = z >= g.rz && z < g.Z - g.rz && y >= g.ry && y < g.Y - g.ry;
        for (int j = 0; j < g.psx; j += 32) {
            int nb = min(32, g.psx - j);
            uint32_t keep = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
            int x0 = cx - g.rx + j;
            uint32_t m = bv_get32(bv, r, x0) & pm_get32(pm, g.W, rr * g.psx + j) & keep;
            if (m) {
                const int wi = x0 >> 5, sft = x0 & 31;
                uint32_t* p = bv.w + (int64_t)r * bv.WX + wi;
                atomicAnd(p, ~(m << sft));
                if (sft && wi + 1 < bv.WX) atomicAnd(p + 1, ~(m >> (32 - sft)));
                if (row_in_rad) rad += __popc(m & rad_xmask(g, x0, nb));
            }
        }
    }
    return warp_sum_i(rad);
}

// neighbour lists of the thinning (patches whose windows intersect), built by the whole
// GPU before the one-CTA rounds kernel: packed centres + fcmask rows, degrees, and (after an
// exclusive scan of the degrees) the entries.  O(m^2) window tests spread over all SMs.
__device__ __forceinline__ bool thin_overlap(const Geo& g, int a, int b)
{
    return abs((a >> 22) - (b >> 22)) < g.psz &&
           abs(((a >> 11) & 2047) - ((b >> 11) & 2047)) < g.psy &&
           abs((a & 2047) - (b & 2047)) < g.psx;
}

__global__ void thin_centres_kernel(const int32_t* __restrict__ sel, int m,
                                    const int32_t* __restrict__ fgidx, ppp_cfg cfg,
                                    int32_t* __restrict__ ctr, int32_t* __restrict__ row)
{
    Geo g = make_geo(cfg);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int vc = sel[i], z, y, x;
    vox_decode(g, vc, z, y, x);
    ctr[i] = (z << 22) | (y << 11) | x;
    row[i] = fgidx[vc];
}

template <bool FILL>
__global__ void __launch_bounds__(128)
thin_neighbours_kernel(const int32_t* __restrict__ ctr, int m, ppp_cfg cfg,
                       int32_t* __restrict__ deg, const int32_t* __restrict__ off,
                       int32_t* __restrict__ nbr, int64_t cap)
{
    Geo g = make_geo(cfg);
    __shared__ int s_c[128];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int pi = i < m ? ctr[i] : 0;
    int d = 0;
    int64_t o = 0;
    if (FILL) { if (off[m] > cap) return; o = i < m ? off[i] : 0; }      // lists disabled
    for (int j0 = 0; j0 < m; j0 += 128) {
        __syncthreads();
        s_c[threadIdx.x] = j0 + threadIdx.x < m ? ctr[j0 + threadIdx.x] : 0;
        __syncthreads();
        const int nj = min(128, m - j0);
        if (i < m)
            for (int q = 0; q < nj; q++) {
                const int j = j0 + q;
                if (j != i && thin_overlap(g, pi, s_c[q])) {
                    if (FILL) nbr[o + d] = j;
                    d++;
                }
            }
    }
    if (!FILL && i < m) deg[i] = d;
}

#define THIN_LIST 2048      // per-round lists kept in shared memory
#define THIN_DEG 64         // neighbour-list budget per patch (average); beyond: brute force

__global__ void __launch_bounds__(THIN_THREADS)
thin_rounds_kernel(const uint8_t* __restrict__ mask, const int32_t* __restrict__ sel, int64_t m64,
                   const int32_t* __restrict__ fgidx, const uint32_t* __restrict__ fcmask,
                   ppp_cfg cfg, uint8_t* __restrict__ keep, uint32_t* gbits, int32_t* gstate,
                   int use_smem)
{
    Geo g = make_geo(cfg);
    extern __shared__ uint32_t s_bits[];
    __shared__ int s_remaining, s_total, s_nnew, s_nrec, s_go, s_cutc, s_cuti, s_nsel, s_lists;
    __shared__ int s_new[THIN_LIST];               // patches selected in this round
    __shared__ int s_rec[THIN_LIST];               // patches to count again / selected so far
    const int m = (int)m64;
    // per-patch state (global scratch): count, packed centre, fcmask row, count at
    // selection (-1: not selected), voxels newly removed from the radslice, round of
    // selection, neighbour list (patches whose window intersects) as offsets + entries
    int32_t* cnt = gstate;
    int32_t* ctr = cnt + m;
    int32_t* row = ctr + m;
    int32_t* csel = row + m;
    int32_t* rcl = csel + m;
    int32_t* rnd = rcl + m;
    int32_t* off = rnd + m;                        // [m + 1]
    int32_t* hist = off + m + 1;                   // [P + 1] removed voxels per selection count
    int32_t* nbr = hist + g.P + 1;                 // [<= THIN_DEG * m]
    BitVol bv;
    bv.WX = bitvol_wx(g.X);
    bv.w = use_smem ? s_bits : gbits;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = THIN_THREADS / 32;
    if (tid == 0) { s_remaining = 0; s_cutc = -1; s_cuti = -1; }
    __syncthreads();
    bitvol_init(g, bv, mask, &s_remaining);
    for (int i = tid; i < m; i += THIN_THREADS) {        // ctr / row / off / nbr: pre-kernels
        csel[i] = -1;
        rcl[i] = 0;
        rnd[i] = -1;
        keep[i] = 0;
    }
    __syncthreads();
    if (tid == 0) s_total = s_remaining;
    // initial |S_i|
    for (int i = w; i < m; i += nw) {
        int c = ctr[i], rw = row[i], n = 0;
        if (rw >= 0)
            n = patch_window_count<false>(g, bv, fcmask + (int64_t)rw * g.W, c >> 22,
                                          (c >> 11) & 2047, c & 2047, lane, nullptr);
        if (lane == 0) cnt[i] = n;
    }
    auto overlap = [&](int a, int b) { return thin_overlap(g, a, b); };
    if (tid == 0) s_lists = off[m] <= (int64_t)THIN_DEG * m;
    __syncthreads();
    const bool lists = s_lists != 0;
    int round = 0;
    while (true) {
        // ---- 1. local maxima among the live patches ----------------------------------
        if (tid == 0) { s_nnew = 0; s_nrec = 0; s_go = 0; }
        __syncthreads();
        for (int i = tid; i < m; i += THIN_THREADS) {
            const int ci = cnt[i];
            if (ci <= 0 || csel[i] >= 0) continue;
            bool lm = true;
            if (lists) {
                for (int q = off[i], e = off[i + 1]; q < e && lm; q++) {
                    const int j = nbr[q], cj = cnt[j];
                    if (cj > 0 && csel[j] < 0 && thin_better(cj, j, ci, i)) lm = false;
                }
            } else {
                const int pi = ctr[i];
                for (int j = 0; j < m && lm; j++) {
                    const int cj = cnt[j];
                    if (cj > 0 && j != i && csel[j] < 0 && thin_better(cj, j, ci, i) &&
                        overlap(pi, ctr[j]))
                        lm = false;
                }
            }
            if (lm) {
                int k = atomicAdd(&s_nnew, 1);
                if (k < THIN_LIST) s_new[k] = i;
            }
        }
        __syncthreads();
        const int nnew = min(s_nnew, THIN_LIST);   // overflow: the rest is taken next round
        if (nnew == 0) break;                      // nothing left that covers anything
        // ---- 2. select them: windows are pairwise disjoint ----------------------------
        for (int k = w; k < nnew; k += nw) {
            const int i = s_new[k], c = ctr[i];
            int rc = patch_window_clear_atomic(g, bv, fcmask + (int64_t)row[i] * g.W, c >> 22,
                                               (c >> 11) & 2047, c & 2047, lane);
            if (lane == 0) {
                csel[i] = cnt[i]; rcl[i] = rc; rnd[i] = round;
                atomicSub(&s_remaining, rc);
            }
        }
        __syncthreads();
        // ---- 3. count their live neighbours again --------------------------------------
        for (int i = tid; i < m; i += THIN_THREADS) {
            if (cnt[i] <= 0 || csel[i] >= 0) continue;
            bool hit = false;
            if (lists) {
                for (int q = off[i], e = off[i + 1]; q < e && !hit; q++) hit = rnd[nbr[q]] == round;
            } else {
                const int pi = ctr[i];
                for (int k = 0; k < nnew && !hit; k++) hit = overlap(pi, ctr[s_new[k]]);
            }
            if (hit) {
                int k = atomicAdd(&s_nrec, 1);
                if (k < THIN_LIST) s_rec[k] = i;
                else s_go = 2;                     // list overflow: count everything (below)
            }
        }
        __syncthreads();
        if (s_go == 2) {
            for (int i = w; i < m; i += nw) {
                if (cnt[i] <= 0 || csel[i] >= 0) continue;
                int c = ctr[i];
                int n = patch_window_count<false>(g, bv, fcmask + (int64_t)row[i] * g.W, c >> 22,
                                                  (c >> 11) & 2047, c & 2047, lane, nullptr);
                if (lane == 0) cnt[i] = n;
            }
        } else {
            const int nrec = s_nrec;
            for (int k = w; k < nrec; k += nw) {
                const int i = s_rec[k], c = ctr[i];
                int n = patch_window_count<false>(g, bv, fcmask + (int64_t)row[i] * g.W, c >> 22,
                                                  (c >> 11) & 2047, c & 2047, lane, nullptr);
                if (lane == 0) cnt[i] = n;
            }
        }
        round++;
        __syncthreads();
        // ---- 4. can the run already be cut? ---------------------------------------------
        if (s_remaining > 0) continue;             // the radslice is not empty yet
        // position where the serial run stops: first patch (in selection order = count at
        // selection descending, index ascending) at which the removed radslice voxels add up
        // to the initial content.  Histogram of the removed voxels over the counts, a scan
        // from the largest count down, then the order inside the one count where it crosses.
        if (tid == 0) { s_cutc = -1; s_cuti = 0x7fffffff; s_go = 0; s_nsel = 0; }
        for (int c = tid; c <= g.P; c += THIN_THREADS) hist[c] = 0;
        __syncthreads();
        for (int i = tid; i < m; i += THIN_THREADS)
            if (csel[i] >= 0) atomicAdd(&hist[csel[i]], rcl[i]);
        __syncthreads();
        const int total = s_total;
        if (tid == 0) {
            int acc = 0, c = g.P;
            for (; c >= 0; c--) { if (acc + hist[c] >= total) break; acc += hist[c]; }
            s_cutc = c;                            // -1: not reached (cannot happen here)
            s_nsel = acc;                          // voxels removed by all larger counts
        }
        __syncthreads();
        const int cutc = s_cutc, before = s_nsel;
        for (int i = tid; i < m; i += THIN_THREADS) {
            if (csel[i] != cutc) continue;
            int acc = before;
            for (int j = 0; j <= i; j++)
                if (csel[j] == cutc) acc += rcl[j];
            if (acc >= total) atomicMin(&s_cuti, i);
        }
        __syncthreads();
        const int cuti = s_cuti;
        // a live patch that could still be selected before the cut keeps the rounds going
        for (int i = tid; i < m; i += THIN_THREADS)
            if (cnt[i] > 0 && csel[i] < 0 && thin_better(cnt[i], i, cutc, cuti)) s_go = 1;
        __syncthreads();
        if (s_go == 0) break;
    }
    __syncthreads();
    // ---- result: the selection up to the cut (everything if the radslice never emptied) --
    const bool cut = s_remaining <= 0 && s_cutc >= 0;
    const int cutc = s_cutc, cuti = s_cuti;
    for (int i = tid; i < m; i += THIN_THREADS) {
        const int ci = csel[i];
        keep[i] = (ci >= 0 && (!cut || (i == cuti) || thin_better(ci, i, cutc, cuti))) ? 1 : 0;
    }
}

static size_t thin_scan_bytes(int64_t m)
{
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum((void*)nullptr, tb, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (int)(m > 0 ? m + 1 : 1));
    return ((tb + 255) / 256) * 256 + 256;
}

extern "C" int64_t ppp_thin_scratch_bytes(const ppp_cfg* cfg, int64_t m)
{
    // bit volume + per-patch state (6 ints + offset) + histogram + degrees + scan temp +
    // neighbour lists (THIN_DEG per patch)
    Geo g = make_geo(*cfg);
    return ppp_cover_scratch_bytes(cfg) + 4096 + 4 * ((int64_t)g.P + 1) +
           (36 + 4 * THIN_DEG) * (m > 0 ? m : 0) + (int64_t)thin_scan_bytes(m);
}

extern "C" int ppp_thin(const uint8_t* mask, const int32_t* sel, int64_t m,
                        const int32_t* fgidx, const uint32_t* fcmask, const ppp_cfg* cfg,
                        uint8_t* keep, void* scratch, void* stream)
{
    if (m <= 0) return 0;
    Geo g = make_geo(*cfg);
    size_t bytes = (size_t)g.Z * g.Y * bitvol_wx(g.X) * 4;
    int use_smem = bytes <= SMEM_BITVOL_MAX;
    size_t smem = use_smem ? bytes : 0;
    if (use_smem) {
        cudaError_t e = cudaFuncSetAttribute(thin_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             SMEM_BITVOL_MAX);
        if (e != cudaSuccess) return ppp_fail((int)e, "ppp_thin: smem attribute");
    }
    // scratch: [bit volume][per-patch state 5 x m]
    uint32_t* gbits = (uint32_t*)scratch;
    int32_t* gcounts = (int32_t*)((char*)scratch + ((bytes + 255) / 256) * 256);
    if (!(cfg->reserved & 0x10000) && g.Z < 512 && g.Y < 2048 && g.X < 2048 &&
        m < 0x7fffffff / 8) {
        if (use_smem) {
            cudaError_t e = cudaFuncSetAttribute(thin_rounds_kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 SMEM_BITVOL_MAX);
            if (e != cudaSuccess) return ppp_fail((int)e, "ppp_thin: smem attribute");
        }
        // per-patch state, then histogram, neighbour lists, degrees, scan temp (layout shared
        // with thin_rounds_kernel)
        cudaStream_t st = (cudaStream_t)stream;
        const int mi = (int)m;
        int32_t* ctr = gcounts + m;
        int32_t* row = gcounts + 2 * m;
        int32_t* off = gcounts + 6 * m;
        int32_t* hist = off + m + 1;
        int32_t* nbr = hist + g.P + 1;
        const int64_t cap = (int64_t)THIN_DEG * m;
        int32_t* deg = nbr + cap;
        void* scan_tmp = (void*)((((uintptr_t)(deg + m + 1)) + 255) / 256 * 256);
        size_t stb = thin_scan_bytes(m);
        thin_centres_kernel<<<(mi + 255) / 256, 256, 0, st>>>(sel, mi, fgidx, *cfg, ctr, row);
        const int64_t maxdeg = (int64_t)(2 * g.psz - 1) * (2 * g.psy - 1) * (2 * g.psx - 1);
        if ((maxdeg < m ? maxdeg : m) * m < 0x7fffffffLL) {
            thin_neighbours_kernel<false><<<(mi + 127) / 128, 128, 0, st>>>(ctr, mi, *cfg, deg,
                                                                             nullptr, nullptr, 0);
            cudaMemsetAsync(deg + m, 0, 4, st);
            cub::DeviceScan::ExclusiveSum(scan_tmp, stb, deg, off, mi + 1, st);
            thin_neighbours_kernel<true><<<(mi + 127) / 128, 128, 0, st>>>(ctr, mi, *cfg, nullptr,
                                                                            off, nbr, cap);
        } else {
            const int32_t big = 0x7fffffff;                       // lists off: brute force
            cudaMemcpyAsync(off + m, &big, 4, cudaMemcpyHostToDevice, st);
        }
        thin_rounds_kernel<<<1, THIN_THREADS, smem, st>>>(
            mask, sel, m, fgidx, fcmask, *cfg, keep, gbits, gcounts, use_smem);
        return ppp_check("ppp_thin(rounds)");
    }
    thin_kernel<<<1, THIN_THREADS, smem, (cudaStream_t)stream>>>(
        mask, sel, m, fgidx, fcmask, *cfg, keep, gbits, gcounts, use_smem);
    return ppp_check("ppp_thin");
}
