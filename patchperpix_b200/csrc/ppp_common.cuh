// Shared device helpers for the PatchPerPix B200 assembly kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/ppp_b200.h"

struct Geo {
    int Z, Y, X;
    int psz, psy, psx;
    int rz, ry, rx;
    int nz, ny, nx;        // 2*ps-1
    int P;                 // psz*psy*psx
    int K;                 // (nz*ny*nx-1)/2
    int64_t V;
    int W;                 // (P+31)/32
    int rsg;               // padded length of one patch x-row in `dp` (guards + psx, x4)
    int rp;                // floats per `dp` row: psz*psy*rsg
};

// `dp` layout: [patch row (qz,qy)][row][rsg] — patch-row major, so that one
// patch x-row of CONSECUTIVE centres is one contiguous, 16-byte aligned block
// that a single bulk copy (cp.async.bulk, TMA engine) brings into shared memory.
// The psx values of a patch x-row sit at [DP_GUARD, DP_GUARD+psx), the guards
// are zero so that a T-wide tile can be read without range checks; slot 0 of
// every patch row holds the x coordinate of the centre (int bits).
#define DP_GUARD 8

__host__ __device__ inline Geo make_geo(const ppp_cfg& c)
{
    Geo g;
    g.Z = c.Z; g.Y = c.Y; g.X = c.X;
    g.psz = c.psz; g.psy = c.psy; g.psx = c.psx;
    g.rz = c.psz / 2; g.ry = c.psy / 2; g.rx = c.psx / 2;
    g.nz = 2 * c.psz - 1; g.ny = 2 * c.psy - 1; g.nx = 2 * c.psx - 1;
    g.P = c.psz * c.psy * c.psx;
    g.K = (g.nz * g.ny * g.nx - 1) / 2;
    g.V = (int64_t)c.Z * c.Y * c.X;
    g.W = (g.P + 31) / 32;
    g.rsg = ((c.psx + 2 * DP_GUARD + 3) / 4) * 4;
    g.rp = c.psz * c.psy * g.rsg;
    return g;
}

// Where the patch predictions come from.  SrcDense = the reference's dense
// float32 [P][Z][Y][X] block (vote_instances.py:193-200).  SrcRows = the compact
// form a ppp+dec run produces (decode.py:39-65 decodes the foreground voxels only,
// everything else stays zero): float16 [G][P] patch rows + vox2row[V] = row of
// block voxel v in `patches`, or -1 (an all-zero patch).
struct SrcDense {
    const float* __restrict__ pred;
    int64_t V;
    __device__ __forceinline__ float at(int po, int64_t v) const { return pred[(int64_t)po * V + v]; }
};
struct SrcRows {
    const __half* __restrict__ patches;
    const int32_t* __restrict__ vox2row;
    int P;
    __device__ __forceinline__ float at(int po, int64_t v) const
    {
        const int r = vox2row[v];
        return r < 0 ? 0.0f : __half2float(patches[(int64_t)r * P + po]);
    }
};

__device__ __forceinline__ void vox_decode(const Geo& g, int v, int& z, int& y, int& x)
{
    x = v % g.X;
    int t = v / g.X;
    y = t % g.Y;
    z = t / g.Y;
}

__device__ __forceinline__ void po_decode(const Geo& g, int po, int& qz, int& qy, int& qx)
{
    qx = po % g.psx;
    int t = po / g.psx;
    qy = t % g.psy;
    qz = t / g.psy;
}

// index of patch channel po of row `row` in `dp` (F rows in total)
__host__ __device__ __forceinline__ int64_t dp_index(const Geo& g, int64_t F, int64_t row, int po)
{
    return ((int64_t)(po / g.psx) * F + row) * g.rsg + DP_GUARD + po % g.psx;
}

// position of patch pixel po in the (2ps-1)^3 offset raster: k(o) for
// o = off(po2)-off(po1), po2 > po1, is lin(po2) - lin(po1) - 1.
__device__ __forceinline__ int po_lin(const Geo& g, int qz, int qy, int qx)
{
    return (qz * g.ny + qy) * g.nx + qx;
}

// slot of offset (dz,dy,dx); -1 if outside the cube, centre or negative half
__device__ __forceinline__ int k_of_offset(const Geo& g, int dz, int dy, int dx)
{
    if (abs(dz) >= g.psz || abs(dy) >= g.psy || abs(dx) >= g.psx) return -1;
    int lin = ((dz + g.psz - 1) * g.ny + (dy + g.psy - 1)) * g.nx + (dx + g.psx - 1);
    return lin - g.K - 1;
}

// class-folded patch value (see include/ppp_b200.h)
__device__ __forceinline__ float fold_class(float v, float th_gt, float bg_lt)
{
    if (v > th_gt) return v;
    if (v < bg_lt) return -(1.0f - v);
    return 0.0f;
}

// number of rows whose voxel index is < v (fgidx of a non-row voxel stores
// -1 - that count, see scan_write_kernel); v == V gives F.
__device__ __forceinline__ int rows_before(const int32_t* __restrict__ fgidx, int64_t v,
                                           int64_t V, int F)
{
    if (v >= V) return F;
    int f = fgidx[v];
    return f >= 0 ? f : -1 - f;
}

__device__ __forceinline__ double warp_sum_d(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int warp_sum_i(int v)
{
    return __reduce_add_sync(0xffffffffu, v);
}
