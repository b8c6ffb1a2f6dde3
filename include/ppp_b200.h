/* ppp_b200.h — C ABI of the B200-native PatchPerPix instance-assembly path.
 *
 * This is the device boundary that replaces the reference's pycuda layer
 * (PatchPerPix/vote_instances/cuda_code.py:5-59) and the three operator
 * functions that sit on it:
 *   create_consensus_array_cuda   consensus_array.py:71-206
 *   rank_patches_cuda             ranked_patches.py:33-74
 *   computePatchGraph_cuda        aff_patch_graph.py:113-187
 * plus device versions of the serial host steps that separate them
 * (foreground_cover.py:15-256, graph_to_labeling.py:34-155).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *    the name ends in _h; the caller owns every buffer, nothing is allocated
 *    or freed behind its back (the reference leaks managed memory instead,
 *    SURVEY.md §8b "Ownership");
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it;
 *  - return value 0 = success, otherwise a cudaError_t (or -1 for argument
 *    errors); ppp_last_error() gives the text;
 *  - volumes are [Z][Y][X] row-major, predictions [P][Z][Y][X] float32 with
 *    P = psz*psy*psx exactly as the reference feeds them
 *    (vote_instances.py:193-200).
 *
 * Data layout produced by this library
 *  - "row" = one foreground voxel (pred[mid] > TH, or a host-side candidate),
 *    rows in raster order;
 *    fgidx[v] = row or -1; rowvox[row] = v.
 *  - flags[v]: bit0 fg, bit1 gated (fg and not overlap,
 *    fillConsensusArray.cu:53-60), bit2 valid patch centre (fg and interior,
 *    fillConsensusArray.cu:25-33).
 *  - dp[row][P]: the patch of that centre, class-folded: v if "high"
 *    (v > TH), -(1-v) if "background" (the USE_*_TH test), 0 otherwise or if
 *    the pixel it talks about is not gated.
 *  - cons[row][K], cnt[row][K]: consensus of base voxel `row` with the voxel
 *    at the k-th lexicographically positive offset; cnt packs
 *    (negative votes << 16) | positive votes.
 */
#ifndef PPP_B200_H
#define PPP_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPP_FLAG_FG      1
#define PPP_FLAG_GATED   2
#define PPP_FLAG_CENTRE  4
#define PPP_FLAG_CAND    8    /* host-side foreground (candidate patch centre) */
#define PPP_FLAG_INTERIOR 16  /* at least patchshape//2 away from every face */
#define PPP_FLAG_ROW     (PPP_FLAG_FG | PPP_FLAG_CAND)

/* kernel variant switches: the reference compiles one kernel per combination
 * of -D flags (utilVoteInstances.py:389-449); here they are run-time fields. */
typedef struct ppp_cfg {
    int32_t Z, Y, X;            /* block shape */
    int32_t psz, psy, psx;      /* patchshape */
    float th_gt;                /* "v > TH"  <=> v > th_gt (TH is a double literal) */
    float bg_lt;                /* background test <=> v < bg_lt (THI, TH/2 or TH) */
    float fc_gt;                /* host-side patch > fc_threshold (f32-rounded)   */
    float pt_gt;                /* host-side patch > patch_threshold (f32-rounded) */
    double th2;                 /* TH*TH         (fillConsensusArray.cu:105) */
    double one_m_th2;           /* 1.0 - TH*TH */
    int32_t prod_mode;          /* 0 vote counter, 1 PROB_PRODUCT, 2 NORM_PROB_PRODUCT */
    int32_t norm_aff;           /* consensus_norm_aff: divide sums by vote count */
    int32_t use_overlap;        /* -DOVERLAP */
    int32_t rank_flags;         /* bit0 NORM_PATCH_RANK, bit1 COUNT_POS_NEG, bit2 fast
                                   parallel double sum instead of the reference's serial
                                   float order (scores then differ by its rounding) */
    int32_t graph_flags;        /* bit0 NORM_PATCH_AFFINITY, bit2 parallel double sum instead of
                                 * the reference's serial float order */
    int32_t reserved;
} ppp_cfg;

const char* ppp_last_error(void);
int ppp_version(void);
/* kernels launched by this library since it was loaded (own kernels + the CUB kernels
 * compiled into it), counted at the runtime's launch entry. */
int64_t ppp_launch_count(void);

/* ---- step 0: gate + compaction (vote_instances.py:276-287 does this on the
 * host with np.where + a python list comprehension) ------------------------ */
/* flags[V] from pred[mid], the optional overlap volume (u8, may be NULL) and
 * the optional host-side foreground `cand` (u8, may be NULL): voxels the
 * reference's python side treats as patch candidates even if pred[mid] <= TH
 * (loadFg / returnFg, utilVoteInstances.py:275-322). */
int ppp_gate(const float* pred, const uint8_t* overlap, const uint8_t* cand,
             const ppp_cfg* cfg, uint8_t* flags, void* stream);
/* exclusive scan of (flags & ROW): fgidx[V], rowvox[<=V], *n_rows (device int64).
 * scratch: at least ppp_compact_scratch_bytes(V) bytes. */
int64_t ppp_compact_scratch_bytes(int64_t V);
int ppp_compact(const uint8_t* flags, int64_t V, int32_t* fgidx, int32_t* rowvox,
                int64_t* n_rows, void* scratch, void* stream);

/* ---- step 0b: centre-major class-folded patches + bit masks --------------
 * dp f32 [F][P]; fcmask/ptmask u32 [F][W], W = (P+31)/32: bit po set iff
 * pred[po][c] > fc_gt / > pt_gt (foreground_cover.py:158, graph_to_labeling.py:84).
 * rbits u64 [psz*psy][F][2] (centre-line major): "received" class bits of every gated voxel — bit
 * (dx + rx) of word (dz,dy) says the centre at b + d calls b high (word 0) /
 * background (word 1); the vote counters are popcounts over them.
 * Any output pointer may be NULL. */
int ppp_prepare_patches(const float* pred, const uint8_t* flags,
                        const int32_t* rowvox, int64_t F, const ppp_cfg* cfg,
                        float* dp, uint32_t* fcmask, uint32_t* ptmask,
                        uint64_t* rbits, void* stream);

/* ---- step 1: consensus (fillConsensusArray.cu + normConsensusArray.cu) ----
 * cons f32 [F][K] (normalised iff cfg->norm_aff), cnt u32 [F][K].
 * impl: 0 = automatic: small patches (psx < 16) use the bit-guided gather (one
 * kernel, visits only the centres that vote), larger ones the tiled pair of
 * kernels (vote counters from `rbits`, sums from TMA-staged shared-memory tiles);
 * 1 = the simple one-CTA-per-voxel gather kept as a cross-check (rbits, scratch
 * and cnt may be NULL); 2 = force the bit-guided gather; 3 = force tiled. */
int64_t ppp_consensus_scratch_bytes(const ppp_cfg* cfg);
int ppp_consensus(const float* dp, const uint64_t* rbits, const uint8_t* flags,
                  const int32_t* fgidx, const int32_t* rowvox, int64_t F,
                  const ppp_cfg* cfg, float* cons, uint32_t* cnt, int32_t impl,
                  void* scratch, void* stream);

/* small windows (psx <= 8, psz*psy <= 64: the 7^3 flylight patches): "received"
 * tables, one contiguous row per voxel instead of gathers into the centre-major
 * array.  rv f32 [F][P]: rv[row(b)][d] = class-folded value the centre b + d
 * (d in window raster order) assigns to b, 0 if that voxel is no patch centre or
 * b is not gated; rb16 u16 [F][W16], W16 = ppp_received_row_words(cfg): word
 * (dz,dy) = (background bits << 8) | high bits of the same table.
 * ppp_consensus_small gives bit-identical cons / cnt to ppp_consensus (same
 * centre order).  need u8 [F] or NULL: rows with 0 are skipped and left
 * unwritten (callers that only read the rows of known patches; the tables must
 * then cover those rows AND their partners, i.e. the mask dilated by 2*ps-1). */
int64_t ppp_received_row_words(const ppp_cfg* cfg);
int ppp_received(const float* pred, const uint8_t* flags, const int32_t* rowvox,
                 const uint8_t* need, int64_t F, const ppp_cfg* cfg, float* rv,
                 uint16_t* rb16, void* stream);
int ppp_received_rows(const uint16_t* patches, const int32_t* vox2row,
                      const uint8_t* flags, const int32_t* rowvox,
                      const uint8_t* need, int64_t F,
                      const ppp_cfg* cfg, float* rv, uint16_t* rb16, void* stream);
int ppp_consensus_small(const float* rv, const uint16_t* rb16, const uint8_t* flags,
                        const int32_t* fgidx, const int32_t* rowvox,
                        const uint8_t* need, int64_t F, const ppp_cfg* cfg,
                        float* cons, uint32_t* cnt, void* stream);

/* need[row] = 1 for every row whose voxel lies within (hz,hy,hx) of one of the m
 * centres (i32 [m][3], block coordinates); need u8 [F] zeroed by the caller.  Builds the
 * `need` masks above from the candidate patches of face jobs. */
int ppp_mark_windows(const int32_t* centres, int64_t m, const ppp_cfg* cfg,
                     int32_t hz, int32_t hy, int32_t hx, const int32_t* fgidx,
                     uint8_t* need, void* stream);

/* ---- step 2: rank (rankPatches.cu) ----------------------------------------
 * score f32 [Z][Y][X]: border voxels -1 / -9999999, non-fg interior 0. */
int64_t ppp_rank_scratch_bytes(const ppp_cfg* cfg, int64_t F);
int ppp_rank(const float* dp, const uint8_t* flags, const int32_t* fgidx,
             const int32_t* rowvox, int64_t F, const float* cons,
             const ppp_cfg* cfg, float* score, void* scratch, void* stream);
/* stable descending sort of the candidate voxels by score
 * (ranked_patches.py:21-30): cand[n] voxel indices in raster order ->
 * order[n] = candidate voxel indices, best first.  scratch from
 * ppp_rank_sort_scratch_bytes(n). */
int64_t ppp_rank_sort_scratch_bytes(int64_t n);
int ppp_rank_sort(const float* score, const int32_t* cand, int64_t n,
                  int32_t* order, void* scratch, void* stream);

/* ---- steps 3+4: greedy foreground cover and thinning ----------------------
 * mask u8 [V] (mask_to_cover, overlap already removed), overlap u8 [V] or NULL
 * (centres to skip, foreground_cover.py:144).  order[n] from ppp_rank_sort.
 * selected u8 [n] in/out (zero it before the first call).
 * pix_ths[n_pix] (device): the pixTh schedule (foreground_cover.py:35-39);
 * NULL = the single threshold 0 of `select_patches_for_sparse_data`, computed
 * without the serial walk (selected = first coverer of every mask voxel).
 * scratch: ppp_cover_scratch_bytes(cfg) bytes. */
int64_t ppp_cover_scratch_bytes(const ppp_cfg* cfg);
int ppp_cover(const uint8_t* mask, const uint8_t* overlap, const int32_t* order,
              int64_t n, const int32_t* fgidx, const uint32_t* fcmask,
              const ppp_cfg* cfg, const int32_t* pix_ths, int32_t n_pix,
              uint8_t* selected, void* scratch, void* stream);
/* greedy set cover over the selected patches (foreground_cover.py:183-256):
 * sel[m] voxel indices in ranked order; keep u8 [m] out.
 * scratch: ppp_thin_scratch_bytes(cfg, m). */
int64_t ppp_thin_scratch_bytes(const ppp_cfg* cfg, int64_t m);
int ppp_thin(const uint8_t* mask, const int32_t* sel, int64_t m,
             const int32_t* fgidx, const uint32_t* fcmask, const ppp_cfg* cfg,
             uint8_t* keep, void* scratch, void* stream);

/* ---- step 5: patch graph (computePatchGraph.cu) ---------------------------
 * pairs u32 [n][6] = (z,y,x,z2,y2,x2); aff f32 [n].  Default: the terms are
 * added into one float in the reference's loop order (the mutex watershed
 * orders edges by |aff|); graph_flags bit2: parallel sum in double.
 * scratch: ppp_patch_graph_scratch_bytes(cfg, n) (the voting-pixel lists of the
 * pairs; may be NULL with graph_flags bit2). */
int64_t ppp_patch_graph_scratch_bytes(const ppp_cfg* cfg, int64_t n);
int ppp_patch_graph(const float* pred, const uint8_t* flags,
                    const int32_t* fgidx, const float* cons,
                    const uint32_t* pairs, int64_t n, const ppp_cfg* cfg,
                    float* aff, void* scratch, void* stream);

/* ---- step 6: connected components over aff > 0 and painting ---------------
 * (aff_patch_graph.py:31-40, graph_to_labeling.py:50-84).  Node = voxel index
 * of a patch centre.  comp i32 [V] scratch/out: component number (1-based, in
 * the reference's order) of every node that has a positive edge, else 0.
 * n_comp device int32.  scratch: ppp_label_scratch_bytes(V, n). */
int64_t ppp_label_scratch_bytes(int64_t V, int64_t n);
int ppp_label_cc(const uint32_t* pairs, const float* aff, int64_t n,
                 const ppp_cfg* cfg, int32_t* comp, int32_t* n_comp,
                 void* scratch, void* stream);
/* mutex watershed instead of thresholded components -- the flylight default
 * `mws = true` (graph_mws.py:7-85 on the graph of aff_patch_graph.py:31-40).
 * HOST function, HOST pointers: one serial greedy pass over the edges sorted by
 * |aff| (every decision depends on the earlier ones; a few thousand edges).
 * node_vox / node_label i32 [<= 2n] out: voxel index and component id of every
 * graph node in insertion order; id 0 = the node joined nothing and is not
 * painted.  Ids equal the reference's instance values (merged-away ids leave
 * gaps).  *n_nodes, *n_labels out; n_labels = ids ever created = length of the
 * reference's component list (the largest surviving label may be smaller).  Scatter node_label into
 * comp[node_vox] and paint with ppp_paint / ppp_paint_patches. */
int ppp_mws_host(const uint32_t* pairs, const float* aff, int64_t n,
                 const ppp_cfg* cfg, int32_t* node_vox, int32_t* node_label,
                 int64_t* n_nodes, int32_t* n_labels);

/* HOST helper for the pair enumeration (aff_patch_graph.py:57-110): order[k] =
 * index of the k-th pair that CPython yields when iterating the set built by
 * inserting the tuples (pairs[i][0], pairs[i][1]) one by one -- what the reference
 * does with cKDTree.query_pairs' result.  pairs i64 [n][2], non-negative, distinct. */
int ppp_pyset_order(const int64_t* pairs, int64_t n, int64_t* order);
/* same, followed by the reference's distance filter (aff_patch_graph.py:61-69: drop a pair if
 * |pts[i][d] - pts[j][d]| > thr[d] on any axis); out i64 [<= n][2] in set order, *n_out kept. */
int ppp_pyset_pairs(const int64_t* pairs, int64_t n, const uint32_t* pts, const double* thr,
                    int64_t* out, int64_t* n_out);

/* instances i32 [V] (zeroed by the caller): for every node with comp > 0,
 * window pixels with pred > pt_gt take max(comp) ("later components overwrite
 * earlier ones", graph_to_labeling.py:84).  nodes[m] voxel indices. */
int ppp_paint(const float* pred, const int32_t* nodes, int64_t m,
              const int32_t* comp, const ppp_cfg* cfg, int32_t* instances,
              void* stream);

/* `one_instance_per_channel` (graph_to_labeling.py:57-95): instances i32
 * [n_comp][V] (zeroed by the caller), channel c-1 = component c painted alone. */
int ppp_paint_channels(const float* pred, const int32_t* nodes, int64_t m,
                       const int32_t* comp, const ppp_cfg* cfg, int32_t* instances,
                       void* stream);

/* same with the member patches as a compact f32 [m][P] array (blockwise path:
 * only the selected patches of a volume that does not fit the device are read,
 * stitch_patch_graph.py:380-385). */
int ppp_paint_patches(const float* patches, const int32_t* nodes, int64_t m,
                      const int32_t* comp, const ppp_cfg* cfg, int32_t* instances,
                      void* stream);

/* ---- compact patch rows: the form a ppp+dec run produces -------------------
 * decode.py:39-65 decodes the foreground voxels only (everything else of the
 * prediction volume stays zero) and stores float16 (decode.py:102-109).  Instead
 * of scattering into a dense [P][Z][Y][X] array the patches stay ROWS:
 *   patches  f16 [G][P]   (bit pattern as uint16_t; one row per stored voxel)
 *   vox2row  i32 [V]      row of block voxel v in `patches`, or -1 = all-zero patch
 * The *_rows entry points equal their dense counterparts on the dense array that
 * scattering the rows would give (tests/test_rows_path.py). */
int ppp_gate_rows(const uint16_t* patches, const int32_t* vox2row,
                  const uint8_t* overlap, const uint8_t* cand,
                  const ppp_cfg* cfg, uint8_t* flags, void* stream);
int ppp_prepare_rows(const uint16_t* patches, const int32_t* vox2row,
                     const uint8_t* flags, const int32_t* rowvox, int64_t F,
                     const ppp_cfg* cfg, float* dp, uint32_t* fcmask,
                     uint32_t* ptmask, uint64_t* rbits, void* stream);
/* pair_org i32 [n][3] or NULL: origin of the region the reference handed pair i to its
 * kernel in (the sub-sampling seed is the product of the REGION-relative centre
 * coordinates, computePatchGraph.cu:24-27); lets the cross pairs of many face regions
 * (stitch_patch_graph.py:252-336) run as one launch over one shared region. */
int ppp_patch_graph_rows(const uint16_t* patches, const int32_t* vox2row,
                         const uint8_t* flags, const int32_t* fgidx,
                         const float* cons, const uint32_t* pairs,
                         const int32_t* pair_org, int64_t n,
                         const ppp_cfg* cfg, float* aff, void* scratch, void* stream);
/* painting of graph nodes into a sub-volume `instances` i32 [cfg Z][Y][X] (zeroed
 * by the caller): node i has its centre at node_zyx[i] (coordinates of that
 * sub-volume; may lie outside, the window is clipped), its patch is row
 * node_row[i], its component node_label[i] (<= 0: skipped); maximum wins
 * (graph_to_labeling.py:84, "later components overwrite earlier ones"). */
int ppp_paint_rows(const uint16_t* patches, const int32_t* node_row,
                   const int32_t* node_zyx, const int32_t* node_label, int64_t m,
                   const ppp_cfg* cfg, int32_t* instances, void* stream);

/* ---- ppp+dec: code -> patch decoder on the tensor cores --------------------
 * (experiments/flylight/setups/setup01/decode.py:16-65 + Autoencoder.forward,
 * torch_model.py:523-544; flylight sizes: code 176 = 22 x 2^3, patch 7^3).
 * codes f32 [B][176]; patches f32 [B][343] (logits, or sigmoid if requested).
 * Weights (device): w_fc f32 [128][22]; w_up0 fp16 [27][2][64][64], w_c0a,
 * w_c0b fp16 [27][1][64][64] = [tap][cin chunk][cout][cin]; biases f32 [64];
 * w_up1 fp16 [27][1][16][64]: the last up-sampling layer (nearest x2 + 3^3 conv
 * 64->1) folded into one conv per output parity (row p < 8, rows 8..15 zero);
 * w_c1a, w_c1b f32 [27]; scalar biases f32 [1].
 * scratch: ppp_decode_scratch_bytes(B). */
int64_t ppp_decode_scratch_bytes(int64_t B);
int ppp_decode(const float* codes, int64_t B, const float* w_fc, const float* b_fc,
               const void* w_up0, const float* b_up0, const void* w_c0a,
               const float* b_c0a, const void* w_c0b, const float* b_c0b,
               const void* w_up1, const float* b_up1, const float* w_c1a,
               const float* b_c1a, const float* w_c1b, const float* b_c1b,
               int32_t apply_sigmoid, float* patches, void* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif
