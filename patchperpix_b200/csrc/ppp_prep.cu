// Step 0 of instance assembly on the device: foreground gate, row compaction
// and the centre-major class-folded patch array.  Replaces the host-side
// np.where / list-comprehension filtering of vote_instances.py:276-287 and the
// per-thread re-reading of thresholds in every reference kernel.
#include "ppp_common.cuh"
#include "ppp_api.cuh"

// ---------------------------------------------------------------------------
// gate: flags[v] = FG | GATED | CENTRE  (fillConsensusArray.cu:25-33, 53-60)
// ---------------------------------------------------------------------------
template <class Src>
__global__ void gate_kernel(Src src, int mid,
                            const uint8_t* __restrict__ overlap,
                            const uint8_t* __restrict__ cand,
                            ppp_cfg cfg, uint8_t* __restrict__ flags)
{
    Geo g = make_geo(cfg);
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.V) return;
    int z, y, x;
    vox_decode(g, (int)v, z, y, x);
    bool fg = src.at(mid, v) > cfg.th_gt;
    bool ov = cfg.use_overlap && overlap != nullptr && overlap[v] != 0;
    bool interior = x >= g.rx && x < g.X - g.rx && y >= g.ry && y < g.Y - g.ry &&
                    z >= g.rz && z < g.Z - g.rz;
    uint8_t f = interior ? PPP_FLAG_INTERIOR : 0;
    if (cand != nullptr && cand[v] != 0) f |= PPP_FLAG_CAND;
    if (fg) {
        f |= PPP_FLAG_FG;
        if (!ov) f |= PPP_FLAG_GATED;
        if (interior) f |= PPP_FLAG_CENTRE;
    }
    flags[v] = f;
}

extern "C" int ppp_gate(const float* pred, const uint8_t* overlap, const uint8_t* cand,
                        const ppp_cfg* cfg, uint8_t* flags, void* stream)
{
    Geo g = make_geo(*cfg);
    if (g.V <= 0 || g.V > 0x7fffffffLL) return ppp_fail(-1, "ppp_gate: bad volume size");
    int mid = g.P / 2;
    int threads = 256;
    int64_t blocks = (g.V + threads - 1) / threads;
    gate_kernel<SrcDense><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        SrcDense{pred, g.V}, mid, overlap, cand, *cfg, flags);
    return ppp_check("ppp_gate");
}

extern "C" int ppp_gate_rows(const uint16_t* patches, const int32_t* vox2row,
                             const uint8_t* overlap, const uint8_t* cand,
                             const ppp_cfg* cfg, uint8_t* flags, void* stream)
{
    Geo g = make_geo(*cfg);
    if (g.V <= 0 || g.V > 0x7fffffffLL) return ppp_fail(-1, "ppp_gate_rows: bad volume size");
    int threads = 256;
    int64_t blocks = (g.V + threads - 1) / threads;
    gate_kernel<SrcRows><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        SrcRows{(const __half*)patches, vox2row, g.P}, g.P / 2, overlap, cand, *cfg, flags);
    return ppp_check("ppp_gate_rows");
}

// ---------------------------------------------------------------------------
// compaction: three-phase exclusive scan over (flags & FG)
// ---------------------------------------------------------------------------
#define SCAN_ITEMS 2048     // voxels per block (256 threads x 8)

__global__ void scan_count_kernel(const uint8_t* __restrict__ flags, int64_t V,
                                  int32_t* __restrict__ block_counts)
{
    int64_t base = (int64_t)blockIdx.x * SCAN_ITEMS;
    int cnt = 0;
    for (int i = threadIdx.x; i < SCAN_ITEMS; i += blockDim.x) {
        int64_t v = base + i;
        if (v < V) cnt += (flags[v] & PPP_FLAG_ROW) ? 1 : 0;
    }
    __shared__ int ws[8];
    cnt = warp_sum_i(cnt);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += ws[i];
        block_counts[blockIdx.x] = s;
    }
}

// single block: exclusive scan of block_counts in place, total -> n_rows
__global__ void scan_blocks_kernel(int32_t* __restrict__ block_counts, int nb,
                                   int64_t* __restrict__ n_rows)
{
    __shared__ int ws[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += blockDim.x) {
        int i = base + threadIdx.x;
        int val = i < nb ? block_counts[i] : 0;
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        int inc = val;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) ws[w] = inc;
        __syncthreads();
        if (w == 0) {
            int t = lane < (int)(blockDim.x >> 5) ? ws[lane] : 0;
            int ti = t;
            for (int o = 1; o < 32; o <<= 1) {
                int u = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += u;
            }
            ws[lane] = ti - t;       // exclusive warp offsets
        }
        __syncthreads();
        int carry = carry_s;
        int excl = carry + ws[w] + inc - val;
        if (i < nb) block_counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = excl + val;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_rows = carry_s;
}

__global__ void scan_write_kernel(const uint8_t* __restrict__ flags, int64_t V,
                                  const int32_t* __restrict__ block_offsets,
                                  int32_t* __restrict__ fgidx,
                                  int32_t* __restrict__ rowvox)
{
    // each thread owns 8 consecutive voxels so that rows stay in raster order
    __shared__ int ws[8];
    int64_t base = (int64_t)blockIdx.x * SCAN_ITEMS + (int64_t)threadIdx.x * 8;
    int f[8];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int64_t v = base + i;
        f[i] = (v < V && (flags[v] & PPP_FLAG_ROW)) ? 1 : 0;
        cnt += f[i];
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    int woff = 0;
    for (int i = 0; i < w; i++) woff += ws[i];
    int row = block_offsets[blockIdx.x] + woff + inc - cnt;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int64_t v = base + i;
        if (v < V) {
            if (f[i]) { fgidx[v] = row; rowvox[row] = (int32_t)v; row++; }
            else fgidx[v] = -1 - row;     // encodes the number of rows before v
        }
    }
}

extern "C" int64_t ppp_compact_scratch_bytes(int64_t V)
{
    int64_t nb = (V + SCAN_ITEMS - 1) / SCAN_ITEMS;
    return (nb + 1) * (int64_t)sizeof(int32_t);
}

extern "C" int ppp_compact(const uint8_t* flags, int64_t V, int32_t* fgidx,
                           int32_t* rowvox, int64_t* n_rows, void* scratch,
                           void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    int64_t nb = (V + SCAN_ITEMS - 1) / SCAN_ITEMS;
    int32_t* bc = (int32_t*)scratch;
    scan_count_kernel<<<(unsigned)nb, 256, 0, s>>>(flags, V, bc);
    scan_blocks_kernel<<<1, 1024, 0, s>>>(bc, (int)nb, n_rows);
    scan_write_kernel<<<(unsigned)nb, 256, 0, s>>>(flags, V, bc, fgidx, rowvox);
    return ppp_check("ppp_compact");
}

// ---------------------------------------------------------------------------
// prepare: dense [P][V] planes -> centre-major rows through a 32x32 transpose
// tile, so that the plane reads (fixed po, consecutive centres) and the row
// writes (fixed centre, consecutive po) are both coalesced.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prepare_kernel(const float* __restrict__ pred, const uint8_t* __restrict__ flags,
               const int32_t* __restrict__ rowvox, int64_t F, ppp_cfg cfg,
               float* __restrict__ dp, uint32_t* __restrict__ fcmask,
               uint32_t* __restrict__ ptmask)
{
    Geo g = make_geo(cfg);
    __shared__ float tileD[32][33];
    __shared__ uint8_t tileB[32][36];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * 32;
    const int po0 = blockIdx.y * 32;

    // phase 1: lane <-> centre row, warp <-> po
    int64_t row = row0 + lane;
    int v = -1, z = 0, y = 0, x = 0;
    bool centre = false, interior = false;
    if (row < F) {
        v = rowvox[row];
        vox_decode(g, v, z, y, x);
        centre = (flags[v] & PPP_FLAG_CENTRE) != 0;
        interior = (flags[v] & PPP_FLAG_INTERIOR) != 0;
    }
    for (int pl = w; pl < 32; pl += 8) {
        int po = po0 + pl;
        float d = 0.0f;
        uint8_t b = 0;
        if (interior && po < g.P) {
            float val = pred[(int64_t)po * g.V + v];
            if (centre) {
                int qz, qy, qx;
                po_decode(g, po, qz, qy, qx);
                int pv = ((z + qz - g.rz) * g.Y + (y + qy - g.ry)) * g.X + (x + qx - g.rx);
                if (flags[pv] & PPP_FLAG_GATED) d = fold_class(val, cfg.th_gt, cfg.bg_lt);
            }
            b = (val > cfg.fc_gt ? 1 : 0) | (val > cfg.pt_gt ? 2 : 0);
        }
        tileD[pl][lane] = d;
        tileB[pl][lane] = b;
    }
    __syncthreads();
    // phase 2: lane <-> po, warp <-> 4 rows
    for (int rl = w; rl < 32; rl += 8) {
        int64_t r = row0 + rl;
        if (r >= F) break;
        int po = po0 + lane;
        float d = tileD[lane][rl];
        uint8_t b = tileB[lane][rl];
        if (dp != nullptr && po < g.P) dp[dp_index(g, F, r, po)] = d;
        unsigned m1 = __ballot_sync(0xffffffffu, b & 1);
        unsigned m2 = __ballot_sync(0xffffffffu, b & 2);
        if (lane == 0) {
            if (fcmask != nullptr) fcmask[r * g.W + blockIdx.y] = m1;
            if (ptmask != nullptr) ptmask[r * g.W + blockIdx.y] = m2;
        }
        // slot 0 of every patch row: x coordinate of the centre (see ppp_common.cuh)
        if (dp != nullptr && blockIdx.y == 0) {
            int xr = rowvox[r] % g.X;
            for (int pr = lane; pr < g.psz * g.psy; pr += 32)
                dp[((int64_t)pr * F + r) * g.rsg] = __int_as_float(xr);
        }
    }
}

// ---------------------------------------------------------------------------
// "received" class bits: for every gated voxel b and every centre line offset
// (dz,dy), one 64-bit word whose bit (dx + rx) says that the centre c = b + d
// calls b "high" (H word) or "background" (L word).  The vote COUNTERS of the
// consensus are popcounts over these words (ppp_consensus.cu).  Lanes walk
// consecutive rows, so the dense plane reads are coalesced.
// rbits u64 [psz*psy][F][2] (line major: the consensus kernels read one word pair
// per centre line from CONSECUTIVE partner rows, which are then contiguous).
// ---------------------------------------------------------------------------
template <class Src>
__global__ void __launch_bounds__(128)
received_bits_kernel(Src src, const uint8_t* __restrict__ flags,
                     const int32_t* __restrict__ rowvox, int64_t F, ppp_cfg cfg,
                     unsigned long long* __restrict__ rbits)
{
    Geo g = make_geo(cfg);
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= F) return;
    const int w = blockIdx.y;                   // (dz + rz) * psy + (dy + ry)
    const int dz = w / g.psy - g.rz, dy = w % g.psy - g.ry;
    const int v = rowvox[row];
    unsigned long long hb = 0ull, lb = 0ull;
    if (flags[v] & PPP_FLAG_GATED) {
        int bz, by, bx;
        vox_decode(g, v, bz, by, bx);
        const int cz = bz + dz, cy = by + dy;
        if (cz >= g.rz && cz < g.Z - g.rz && cy >= g.ry && cy < g.Y - g.ry) {
            const int64_t line = ((int64_t)cz * g.Y + cy) * g.X;
            const int porow = ((g.rz - dz) * g.psy + (g.ry - dy)) * g.psx;
            for (int t = 0; t < g.psx; t++) {
                int cx = bx - g.rx + t;
                if (cx < g.rx || cx >= g.X - g.rx) continue;
                if (!(flags[line + cx] & PPP_FLAG_CENTRE)) continue;
                // pixel b seen from centre c sits at patch index r - d
                float val = src.at(porow + (g.psx - 1 - t), line + cx);
                if (val > cfg.th_gt) hb |= 1ull << t;
                else if (val < cfg.bg_lt) lb |= 1ull << t;
            }
        }
    }
    const int64_t o = ((int64_t)w * F + row) * 2;
    rbits[o] = hb;
    rbits[o + 1] = lb;
}

extern "C" int ppp_prepare_patches(const float* pred, const uint8_t* flags,
                                   const int32_t* rowvox, int64_t F,
                                   const ppp_cfg* cfg, float* dp,
                                   uint32_t* fcmask, uint32_t* ptmask,
                                   uint64_t* rbits, void* stream)
{
    if (F <= 0) return 0;
    Geo g = make_geo(*cfg);
    dim3 grid((unsigned)((F + 31) / 32), (unsigned)g.W);
    prepare_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        pred, flags, rowvox, F, *cfg, dp, fcmask, ptmask);
    if (rbits != nullptr) {
        if (g.psx > 64) return ppp_fail(-1, "ppp_prepare_patches: psx > 64 has no bit path");
        dim3 grid2((unsigned)((F + 127) / 128), (unsigned)(g.psz * g.psy));
        received_bits_kernel<SrcDense><<<grid2, 128, 0, (cudaStream_t)stream>>>(
            SrcDense{pred, g.V}, flags, rowvox, F, *cfg, (unsigned long long*)rbits);
    }
    return ppp_check("ppp_prepare_patches");
}

// ---------------------------------------------------------------------------
// the same from float16 patch ROWS (compact ppp+dec form): one warp per row,
// lanes walk the patch, so the source reads (consecutive po of one row) and the
// dp / mask writes are coalesced without a transpose tile.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prepare_rows_kernel(const __half* __restrict__ patches, const int32_t* __restrict__ vox2row,
                    const uint8_t* __restrict__ flags, const int32_t* __restrict__ rowvox,
                    int64_t F, ppp_cfg cfg, float* __restrict__ dp,
                    uint32_t* __restrict__ fcmask, uint32_t* __restrict__ ptmask)
{
    Geo g = make_geo(cfg);
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= F) return;
    const int v = rowvox[row];
    int z, y, x;
    vox_decode(g, v, z, y, x);
    const uint8_t fl = flags[v];
    const bool centre = (fl & PPP_FLAG_CENTRE) != 0, interior = (fl & PPP_FLAG_INTERIOR) != 0;
    const int srow = vox2row[v];
    const __half* prow = patches + (int64_t)(srow < 0 ? 0 : srow) * g.P;
    for (int base = 0; base < g.W * 32; base += 32) {
        const int po = base + lane;
        float d = 0.0f;
        bool b1 = false, b2 = false;
        if (interior && srow >= 0 && po < g.P) {
            const float val = __half2float(prow[po]);
            if (centre) {
                int qz, qy, qx;
                po_decode(g, po, qz, qy, qx);
                int pv = ((z + qz - g.rz) * g.Y + (y + qy - g.ry)) * g.X + (x + qx - g.rx);
                if (flags[pv] & PPP_FLAG_GATED) d = fold_class(val, cfg.th_gt, cfg.bg_lt);
            }
            b1 = val > cfg.fc_gt;
            b2 = val > cfg.pt_gt;
        }
        if (dp != nullptr && po < g.P) dp[dp_index(g, F, row, po)] = d;
        unsigned m1 = __ballot_sync(0xffffffffu, b1);
        unsigned m2 = __ballot_sync(0xffffffffu, b2);
        if (lane == 0) {
            if (fcmask != nullptr) fcmask[row * g.W + (base >> 5)] = m1;
            if (ptmask != nullptr) ptmask[row * g.W + (base >> 5)] = m2;
        }
    }
    if (dp != nullptr)
        for (int pr = lane; pr < g.psz * g.psy; pr += 32)
            dp[((int64_t)pr * F + row) * g.rsg] = __int_as_float(x);
}

extern "C" int ppp_prepare_rows(const uint16_t* patches, const int32_t* vox2row,
                                const uint8_t* flags, const int32_t* rowvox, int64_t F,
                                const ppp_cfg* cfg, float* dp, uint32_t* fcmask,
                                uint32_t* ptmask, uint64_t* rbits, void* stream)
{
    if (F <= 0) return 0;
    Geo g = make_geo(*cfg);
    prepare_rows_kernel<<<(unsigned)((F + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)patches, vox2row, flags, rowvox, F, *cfg, dp, fcmask, ptmask);
    if (rbits != nullptr) {
        if (g.psx > 64) return ppp_fail(-1, "ppp_prepare_rows: psx > 64 has no bit path");
        dim3 grid2((unsigned)((F + 127) / 128), (unsigned)(g.psz * g.psy));
        received_bits_kernel<SrcRows><<<grid2, 128, 0, (cudaStream_t)stream>>>(
            SrcRows{(const __half*)patches, vox2row, g.P}, flags, rowvox, F, *cfg,
            (unsigned long long*)rbits);
    }
    return ppp_check("ppp_prepare_rows");
}

// ---------------------------------------------------------------------------
// "received" patches for small windows (psx <= 8, e.g. the 7^3 flylight patches):
// for every gated voxel b the class-folded value EVERY surrounding centre assigns
// to it, rv[row(b)][d] = D_{b+d}(b) with d in window raster order (0 where b+d is
// no patch centre), and the class bits of the same table packed per centre line:
// rb16[row][w] = (L bits << 8) | H bits, bit t <-> centre b + (dz, dy, t - rx).
// The consensus of a voxel pair is then a masked dot product of two such rows
// (ppp_consensus_small): partner data is one contiguous row instead of gathers
// through fgidx into the centre-major array.  One warp per row.
// ---------------------------------------------------------------------------
template <class Src>
__global__ void __launch_bounds__(256)
received_values_kernel(Src src, const uint8_t* __restrict__ flags,
                       const int32_t* __restrict__ rowvox, const uint8_t* __restrict__ need,
                       int64_t F, ppp_cfg cfg,
                       float* __restrict__ rv, uint16_t* __restrict__ rb16, int rbw)
{
    Geo g = make_geo(cfg);
    __shared__ unsigned s_bits[8][64];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * 8 + w;
    if (row >= F) return;
    if (need != nullptr && !need[row]) return;
    const int nrw = g.psz * g.psy;
    const int v = rowvox[row];
    for (int i = lane; i < 64; i += 32) s_bits[w][i] = 0u;
    __syncwarp();
    if (flags[v] & PPP_FLAG_GATED) {
        int bz, by, bx;
        vox_decode(g, v, bz, by, bx);
        for (int base = 0; base < g.P; base += 32) {
            const int d = base + lane;
            if (d < g.P) {
                int qz, qy, qx;
                po_decode(g, d, qz, qy, qx);
                const int cz = bz + qz - g.rz, cy = by + qy - g.ry, cx = bx + qx - g.rx;
                float val = 0.0f;
                if (cz >= g.rz && cz < g.Z - g.rz && cy >= g.ry && cy < g.Y - g.ry &&
                    cx >= g.rx && cx < g.X - g.rx) {
                    const int64_t vc = ((int64_t)cz * g.Y + cy) * g.X + cx;
                    if (flags[vc] & PPP_FLAG_CENTRE)    // b seen from c sits at patch index r - d
                        val = fold_class(src.at(g.P - 1 - d, vc), cfg.th_gt, cfg.bg_lt);
                }
                rv[row * g.P + d] = val;
                if (val > 0.0f) atomicOr(&s_bits[w][qz * g.psy + qy], 1u << qx);
                else if (val < 0.0f) atomicOr(&s_bits[w][qz * g.psy + qy], 256u << qx);
            }
        }
        __syncwarp();
    }
    for (int i = lane; i < rbw; i += 32)
        rb16[row * rbw + i] = (uint16_t)(i < nrw ? s_bits[w][i] : 0u);
}

static int received_check(const Geo& g, int64_t F)
{
    if (g.psx > 8) return ppp_fail(-1, "ppp_received: psx > 8 (use the rbits path)");
    if (g.psz * g.psy > 64) return ppp_fail(-1, "ppp_received: more than 64 centre lines");
    if (F * (int64_t)g.P > 0x7fffffffffLL) return ppp_fail(-1, "ppp_received: too many rows");
    return 0;
}

extern "C" int64_t ppp_received_row_words(const ppp_cfg* cfg)
{
    Geo g = make_geo(*cfg);
    return ((g.psz * g.psy + 7) / 8) * 8;
}

extern "C" int ppp_received(const float* pred, const uint8_t* flags, const int32_t* rowvox,
                            const uint8_t* need, int64_t F, const ppp_cfg* cfg, float* rv,
                            uint16_t* rb16, void* stream)
{
    if (F <= 0) return 0;
    Geo g = make_geo(*cfg);
    if (int rc = received_check(g, F)) return rc;
    received_values_kernel<SrcDense><<<(unsigned)((F + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        SrcDense{pred, g.V}, flags, rowvox, need, F, *cfg, rv, rb16,
        (int)ppp_received_row_words(cfg));
    return ppp_check("ppp_received");
}

extern "C" int ppp_received_rows(const uint16_t* patches, const int32_t* vox2row,
                                 const uint8_t* flags, const int32_t* rowvox,
                                 const uint8_t* need, int64_t F,
                                 const ppp_cfg* cfg, float* rv, uint16_t* rb16, void* stream)
{
    if (F <= 0) return 0;
    Geo g = make_geo(*cfg);
    if (int rc = received_check(g, F)) return rc;
    received_values_kernel<SrcRows><<<(unsigned)((F + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        SrcRows{(const __half*)patches, vox2row, g.P}, flags, rowvox, need, F, *cfg, rv, rb16,
        (int)ppp_received_row_words(cfg));
    return ppp_check("ppp_received_rows");
}

// ---------------------------------------------------------------------------
// row masks for partial consensus runs (face jobs of the blockwise stitcher only
// read the rows inside the windows of their candidate patches): need[row] = 1 for
// every row whose voxel lies within (hz,hy,hx) of one of the m centres.  One CTA
// per centre; need must be zeroed by the caller.
// ---------------------------------------------------------------------------
__global__ void mark_windows_kernel(const int32_t* __restrict__ centres, ppp_cfg cfg,
                                    int hz, int hy, int hx,
                                    const int32_t* __restrict__ fgidx, uint8_t* __restrict__ need)
{
    Geo g = make_geo(cfg);
    const int cz = centres[3 * blockIdx.x], cy = centres[3 * blockIdx.x + 1],
              cx = centres[3 * blockIdx.x + 2];
    const int ny = 2 * hy + 1, nx = 2 * hx + 1, n = (2 * hz + 1) * ny * nx;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int z = cz - hz + i / (ny * nx), y = cy - hy + (i / nx) % ny, x = cx - hx + i % nx;
        if (z < 0 || z >= g.Z || y < 0 || y >= g.Y || x < 0 || x >= g.X) continue;
        const int r = fgidx[((int64_t)z * g.Y + y) * g.X + x];
        if (r >= 0) need[r] = 1;
    }
}

extern "C" int ppp_mark_windows(const int32_t* centres, int64_t m, const ppp_cfg* cfg,
                                int32_t hz, int32_t hy, int32_t hx, const int32_t* fgidx,
                                uint8_t* need, void* stream)
{
    if (m <= 0) return 0;
    mark_windows_kernel<<<(unsigned)m, 128, 0, (cudaStream_t)stream>>>(centres, *cfg, hz, hy, hx,
                                                                      fgidx, need);
    return ppp_check("ppp_mark_windows");
}
