/* TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the arithmetic of
 * the reference's CUDA path for instance assembly.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product (patchperpix_b200/) never does.
 *
 * Each function restates one reference kernel with run-time shapes, the same
 * float/double mixed arithmetic and the same SERIAL visiting order as the
 * reference kernel run through oracle/ref_shim (grid of 8x8x8 thread blocks,
 * utilVoteInstances.py:452-462, blocks z,y,x then threads z,y,x), so that its
 * outputs are bit-identical to the golden vectors in tests/golden/ (generated
 * from the unmodified reference by tools/gen_golden.py).
 *
 *   ppp_oracle_consensus   <- cuda/fillConsensusArray.cu:5-175
 *   ppp_oracle_norm        <- cuda/normConsensusArray.cu:5-28
 *   ppp_oracle_rank        <- cuda/rankPatches.cu:1-161
 *   ppp_oracle_patch_graph <- cuda/computePatchGraph.cu:3-136
 *
 * Output layout is the compact one of patchperpix_b200/layout.py: row =
 * gated fg voxel (raster order), column k = lexicographically positive offset.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (no -march: no FMA fusion).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct {
    int Z, Y, X;
    int psz, psy, psx;
    double th;        /* TH literal (utilVoteInstances.py:361) */
    double thi;       /* THI literal (:362-365) */
    int bg_mode;      /* 0 USE_INV_TH, 1 USE_HALF_TH, 2 USE_LESS_THAN_TH */
    int prod_mode;    /* 0 counter, 1 PROB_PRODUCT, 2 NORM_PROB_PRODUCT */
} ppp_oracle_cfg;

static inline int64_t vox(const ppp_oracle_cfg* c, int z, int y, int x)
{ return ((int64_t)z * c->Y + y) * c->X + x; }

static inline int is_bg(const ppp_oracle_cfg* c, float v2)
{
    switch (c->bg_mode) {
    case 0: return v2 < c->thi;
    case 1: return v2 < c->th / 2;
    default: return v2 < c->th;
    }
}

/* slot of offset (dz,dy,dx) in the compact row, -1 if not lex-positive /
 * outside the (2ps-1)^3 cube */
static inline int64_t kslot(const ppp_oracle_cfg* c, int dz, int dy, int dx)
{
    int nz = 2 * c->psz - 1, ny = 2 * c->psy - 1, nx = 2 * c->psx - 1;
    if (abs(dz) >= c->psz || abs(dy) >= c->psy || abs(dx) >= c->psx) return -1;
    int64_t N = (int64_t)nz * ny * nx;
    int64_t lin = ((int64_t)(dz + c->psz - 1) * ny + (dy + c->psy - 1)) * nx
                  + (dx + c->psx - 1);
    return lin - (N - 1) / 2 - 1;     /* < 0 for centre and negative half */
}

/* gate (fillConsensusArray.cu:53-60): pred[mid] > TH and not overlap.
 * fgidx[v] = compact row or -1; returns the number of rows. */
int64_t ppp_oracle_gate(const ppp_oracle_cfg* c, const float* pred,
                        const uint8_t* overlap, int32_t* fgidx)
{
    int64_t V = (int64_t)c->Z * c->Y * c->X, n = 0;
    int mid = (c->psz * c->psy * c->psx) / 2;
    const float* pm = pred + (int64_t)mid * V;
    for (int64_t v = 0; v < V; v++) {
        int g = !(pm[v] <= c->th) && !(overlap && overlap[v] != 0);
        fgidx[v] = g ? (int32_t)n++ : -1;
    }
    return n;
}

#define FOR_CENTRES_BLOCKED(c)                                                 \
    for (int bz = 0; bz < ((c)->Z + 7) / 8; bz++)                              \
    for (int by = 0; by < ((c)->Y + 7) / 8; by++)                              \
    for (int bx = 0; bx < ((c)->X + 7) / 8; bx++)                              \
    for (int tz = 0; tz < 8; tz++)                                             \
    for (int ty = 0; ty < 8; ty++)                                             \
    for (int tx = 0; tx < 8; tx++)

/* cons_raw f32 [F][K] (may be NULL), cnt_pos/cnt_neg u16 [F][K] (may be NULL);
 * all zero-initialised by the caller. */
void ppp_oracle_consensus(const ppp_oracle_cfg* c, const float* pred,
                          const uint8_t* overlap, const int32_t* fgidx,
                          float* cons_raw, uint16_t* cnt_pos, uint16_t* cnt_neg)
{
    const int PSZ = c->psz, PSY = c->psy, PSX = c->psx;
    const int PSZH = PSZ / 2, PSYH = PSY / 2, PSXH = PSX / 2;
    const int mid = (PSX * PSY * PSZ) / 2;
    const int64_t V = (int64_t)c->Z * c->Y * c->X;
    const int64_t K = ((int64_t)(2 * PSZ - 1) * (2 * PSY - 1) * (2 * PSX - 1) - 1) / 2;
    const double TH = c->th;
    const float* pm = pred + (int64_t)mid * V;
    FOR_CENTRES_BLOCKED(c) {
        int idz = bz * 8 + tz, idy = by * 8 + ty, idx = bx * 8 + tx;
        if (!(idx < c->X - PSXH && idy < c->Y - PSYH && idz < c->Z - PSZH &&
              idx >= PSXH && idy >= PSYH && idz >= PSZH)) continue;
        int64_t vc = vox(c, idz, idy, idx);
        if (pm[vc] <= TH) continue;
        for (int pz1 = 0; pz1 < PSZ; pz1++)
        for (int py1 = 0; py1 < PSY; py1++)
        for (int px1 = 0; px1 < PSX; px1++) {
            int po1 = px1 + PSX * py1 + PSX * PSY * pz1;
            float v1 = pred[(int64_t)po1 * V + vc];
            if (v1 <= TH) continue;
            int z1 = idz + pz1 - PSZH, y1 = idy + py1 - PSYH, x1 = idx + px1 - PSXH;
            int64_t v1i = vox(c, z1, y1, x1);
            if (pm[v1i] <= TH) continue;
            if (overlap && overlap[v1i] != 0) continue;
            for (int pz2 = 0; pz2 < PSZ; pz2++)
            for (int py2 = 0; py2 < PSY; py2++)
            for (int px2 = 0; px2 < PSX; px2++) {
                int po2 = px2 + PSX * py2 + PSX * PSY * pz2;
                if (po1 == po2) continue;
                int z2 = idz + pz2 - PSZH, y2 = idy + py2 - PSYH, x2 = idx + px2 - PSXH;
                int64_t v2i = vox(c, z2, y2, x2);
                if (pm[v2i] <= TH) continue;
                if (overlap && overlap[v2i] != 0) continue;
                float v2 = pred[(int64_t)po2 * V + vc];
                if (v2 > TH) {
                    if (po2 <= po1) continue;
                    int64_t row = fgidx[v1i];
                    int64_t k = kslot(c, pz2 - pz1, py2 - py1, px2 - px1);
                    if (cnt_pos) cnt_pos[row * K + k] += 1;
                    if (cons_raw) {
                        float v3;
                        if (c->prod_mode == 2)
                            v3 = (v1 * v2 - TH * TH) / (1.0 - TH * TH);
                        else if (c->prod_mode == 1)
                            v3 = v1 * v2;
                        else
                            v3 = 1;
                        cons_raw[row * K + k] += v3;
                    }
                } else if (is_bg(c, v2)) {
                    float v3 = 1;
                    if (c->prod_mode == 2)
                        v3 = (v1 * (1 - v2) - TH * TH) / (1.0 - TH * TH);
                    else if (c->prod_mode == 1)
                        v3 = v1 * (1 - v2);
                    int64_t row, k;
                    if (po2 <= po1) {   /* reversed: base = pixel 2 */
                        row = fgidx[v2i];
                        k = kslot(c, pz1 - pz2, py1 - py2, px1 - px2);
                    } else {
                        row = fgidx[v1i];
                        k = kslot(c, pz2 - pz1, py2 - py1, px2 - px1);
                    }
                    if (cnt_neg) cnt_neg[row * K + k] += 1;
                    if (cons_raw) cons_raw[row * K + k] += -v3;
                }
            }
        }
    }
}

/* normConsensusArray.cu:19-26: cons /= cnt where cnt != 0 (every compact row
 * is a pred[mid] > TH voxel, the kernel's only gate). */
void ppp_oracle_norm(int64_t n, float* cons, const uint16_t* cnt_pos,
                     const uint16_t* cnt_neg)
{
    for (int64_t i = 0; i < n; i++) {
        float cnt = (float)((int)cnt_pos[i] + (int)cnt_neg[i]);
        if (cnt != 0) cons[i] = cons[i] / cnt;
    }
}

static inline float cons_at(const ppp_oracle_cfg* c, const float* cons,
                            const int32_t* fgidx, int64_t K, int64_t v,
                            int dz, int dy, int dx)
{
    int64_t row = fgidx[v];
    if (row < 0) return 0.0f;          /* never written in the dense layout */
    int64_t k = kslot(c, dz, dy, dx);
    if (k < 0) return 0.0f;
    return cons[row * K + k];
}

/* flags: bit0 NORM_PATCH_RANK, bit1 COUNT_POS_NEG */
void ppp_oracle_rank(const ppp_oracle_cfg* c, const float* pred,
                     const uint8_t* overlap, const int32_t* fgidx,
                     const float* cons, int flags, float* score)
{
    const int PSZ = c->psz, PSY = c->psy, PSX = c->psx;
    const int PSZH = PSZ / 2, PSYH = PSY / 2, PSXH = PSX / 2;
    const int mid = (PSX * PSY * PSZ) / 2;
    const int64_t V = (int64_t)c->Z * c->Y * c->X;
    const int64_t K = ((int64_t)(2 * PSZ - 1) * (2 * PSY - 1) * (2 * PSX - 1) - 1) / 2;
    const double TH = c->th;
    const float* pm = pred + (int64_t)mid * V;
    for (int idz = 0; idz < c->Z; idz++)
    for (int idy = 0; idy < c->Y; idy++)
    for (int idx = 0; idx < c->X; idx++) {
        int64_t vc = vox(c, idz, idy, idx);
        if (!(idx < c->X - PSXH && idy < c->Y - PSYH && idz < c->Z - PSZH &&
              idx >= PSXH && idy >= PSYH && idz >= PSZH)) {
            score[vc] = (flags & 1) ? -1.0 : -9999999.0;
            continue;
        }
        if (pm[vc] <= TH) continue;          /* keeps the zero-init value */
        float acc = 0.0f;
        double dacc = 0.0;      /* flags bit2: exact-ish sum, diagnostic only */
        unsigned fgCnt = 0;
        for (int pz1 = 0; pz1 < PSZ; pz1++)
        for (int py1 = 0; py1 < PSY; py1++)
        for (int px1 = 0; px1 < PSX; px1++) {
            int po1 = px1 + PSX * py1 + PSX * PSY * pz1;
            float v1 = pred[(int64_t)po1 * V + vc];
            if (v1 <= TH) continue;
            int z1 = idz + pz1 - PSZH, y1 = idy + py1 - PSYH, x1 = idx + px1 - PSXH;
            int64_t v1i = vox(c, z1, y1, x1);
            if (pm[v1i] <= TH) continue;
            if (overlap && overlap[v1i] != 0) continue;
            for (int pz2 = 0; pz2 < PSZ; pz2++)
            for (int py2 = 0; py2 < PSY; py2++)
            for (int px2 = 0; px2 < PSX; px2++) {
                int po2 = px2 + PSX * py2 + PSX * PSY * pz2;
                if (po1 == po2) continue;
                int z2 = idz + pz2 - PSZH, y2 = idy + py2 - PSYH, x2 = idx + px2 - PSXH;
                int64_t v2i = vox(c, z2, y2, x2);
                if (pm[v2i] <= TH) continue;
                if (overlap && overlap[v2i] != 0) continue;
                float v2 = pred[(int64_t)po2 * V + vc];
                if (v2 > TH) {
                    if (po2 <= po1) continue;
                    float v3 = cons_at(c, cons, fgidx, K, v1i,
                                       pz2 - pz1, py2 - py1, px2 - px1);
                    dacc += v3;
                    if (flags & 2) {
                        if (v3 != 0) acc += copysignf(1, v3);
                        else acc -= 1;
                    } else acc += v3;
                } else if (is_bg(c, v2)) {
                    float v3;
                    if (po2 <= po1)
                        v3 = cons_at(c, cons, fgidx, K, v2i,
                                     pz1 - pz2, py1 - py2, px1 - px2);
                    else
                        v3 = cons_at(c, cons, fgidx, K, v1i,
                                     pz2 - pz1, py2 - py1, px2 - px1);
                    dacc -= v3;
                    if (flags & 2) {
                        if (v3 != 0) acc -= copysignf(1, v3);
                        else acc -= 1;
                    } else acc -= v3;
                }
                fgCnt += 1;
            }
        }
        if (flags & 4) acc = (float)dacc;
        if (flags & 1) score[vc] = acc / (float)(fgCnt > 1 ? fgCnt : 1);
        else score[vc] = acc;
    }
}

/* flags: bit0 NORM_PATCH_AFFINITY.  pairs u32 [n][6] (z,y,x,z2,y2,x2). */
void ppp_oracle_patch_graph(const ppp_oracle_cfg* c, const float* pred,
                            const int32_t* fgidx, const float* cons,
                            const uint32_t* pairs, int64_t n, int flags,
                            float* aff)
{
    const int PSZ = c->psz, PSY = c->psy, PSX = c->psx;
    const int PSZH = PSZ / 2, PSYH = PSY / 2, PSXH = PSX / 2;
    const int mid = (PSX * PSY * PSZ) / 2;
    const int64_t V = (int64_t)c->Z * c->Y * c->X;
    const int64_t K = ((int64_t)(2 * PSZ - 1) * (2 * PSY - 1) * (2 * PSX - 1) - 1) / 2;
    const double TH = c->th;
    const float* pm = pred + (int64_t)mid * V;
    for (int64_t id1 = 0; id1 < n; id1++) {
        int idz = pairs[id1 * 6], idy = pairs[id1 * 6 + 1], idx = pairs[id1 * 6 + 2];
        int idz2 = pairs[id1 * 6 + 3], idy2 = pairs[id1 * 6 + 4], idx2 = pairs[id1 * 6 + 5];
        uint32_t rnd = (uint32_t)idz * (uint32_t)idz2 * (uint32_t)idy *
                       (uint32_t)idy2 * (uint32_t)idx * (uint32_t)idx2;
        int64_t vc1 = vox(c, idz, idy, idx), vc2 = vox(c, idz2, idy2, idx2);
        float acc = 0.0f;
        double dacc = 0.0;      /* flags bit2: exact-ish sum, diagnostic only */
        unsigned fgCnt = 0;
        for (int pz1 = 0; pz1 < PSZ; pz1++)
        for (int py1 = 0; py1 < PSY; py1++)
        for (int px1 = 0; px1 < PSX; px1++) {
            int z1 = idz + pz1 - PSZH, y1 = idy + py1 - PSYH, x1 = idx + px1 - PSXH;
            int64_t v1i = vox(c, z1, y1, x1);
            if (pm[v1i] <= TH) continue;
            int po1 = px1 + PSX * py1 + PSX * PSY * pz1;
            if (pred[(int64_t)po1 * V + vc1] <= TH) continue;
            for (int pz2 = 0; pz2 < PSZ; pz2++)
            for (int py2 = 0; py2 < PSY; py2++)
            for (int px2 = 0; px2 < PSX; px2++) {
                int z2 = idz2 + pz2 - PSZH, y2 = idy2 + py2 - PSYH, x2 = idx2 + px2 - PSXH;
                int64_t v2i = vox(c, z2, y2, x2);
                if (pm[v2i] <= TH) continue;
                int po2 = px2 + PSX * py2 + PSX * PSY * pz2;
                if (pred[(int64_t)po2 * V + vc2] <= TH) continue;
                int gz1 = x1 + c->X * y1 + c->X * c->Y * z1;
                int gz2 = x2 + c->X * y2 + c->X * c->Y * z2;
                if (abs(x1 - idx2) <= PSXH && abs(y1 - idy2) <= PSYH &&
                    abs(z1 - idz2) <= PSZH && abs(x2 - idx) <= PSXH &&
                    abs(y2 - idy) <= PSYH && abs(z2 - idz) <= PSZH) {
                    rnd = rnd * 1103515245U;
                    float rndT = rnd / 4294967296.0f;
                    if (rndT > 0.2) continue;
                }
                if (gz1 <= gz2) {
                    int zo = z2 - z1 + PSZ - 1, yo = y2 - y1 + PSY - 1, xo = x2 - x1 + PSX - 1;
                    if (zo < 0 || zo >= 2 * PSZ || yo < 0 || yo >= 2 * PSY ||
                        xo < 0 || xo >= 2 * PSX) continue;
                    { float v3 = cons_at(c, cons, fgidx, K, v1i, z2 - z1, y2 - y1, x2 - x1);
                      acc += v3; dacc += v3; }
                    fgCnt += 1;
                } else {
                    int zo = z1 - z2 + PSZ - 1, yo = y1 - y2 + PSY - 1, xo = x1 - x2 + PSX - 1;
                    if (zo < 0 || zo >= 2 * PSZ || yo < 0 || yo >= 2 * PSY ||
                        xo < 0 || xo >= 2 * PSX) continue;
                    { float v3 = cons_at(c, cons, fgidx, K, v2i, z1 - z2, y1 - y2, x1 - x2);
                      acc += v3; dacc += v3; }
                    fgCnt += 1;
                }
            }
        }
        if (flags & 4) acc = (float)dacc;
        if (flags & 1) aff[id1] = acc / (float)(fgCnt > 1 ? fgCnt : 1);
        else aff[id1] = acc;
    }
}
