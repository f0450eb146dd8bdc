/* wesup_b200 -- C ABI of the B200-native WESUP superpixel stage.
 *
 * The reference (mrcfps/WESUP) is pure Python and has no FFI of its own; its
 * boundary for this path is the Python surface of models/wesup.py.  This header
 * is the native boundary *underneath* that surface: every entry point replaces
 * the torch op sequence of one reference call site (cited per function).  The
 * Python mirror in wesup_b200/models/wesup.py binds these with ctypes
 * (wesup_b200/_lib.py); INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer on the current CUDA device unless the
 *    comment says "host"; the caller owns all buffers including workspaces
 *    (sizes from the *_workspace_bytes functions); nothing is allocated here;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *    never synchronises and is safe to capture in a CUDA graph;
 *  - return value: 0 = OK, negative = bad argument (WESUP_E_*), positive = the
 *    cudaError_t of a failed launch.  wesup_last_error() returns a thread-local
 *    message for the most recent non-zero return on the calling thread;
 *  - there is no CPU fallback: without a CUDA device every compute entry point
 *    returns a cudaError_t.
 */
#ifndef WESUP_B200_H
#define WESUP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WESUP_ABI_VERSION 7
#define WESUP_MAX_LEVELS 16

/* element type of the hypercolumn tensor */
#define WESUP_F32 0
#define WESUP_BF16 1
/* layout of the hypercolumn / side tensors: channel-major (C,H,W) as in the
 * reference (models/wesup.py:281) or pixel-major (H*W,C) (what the pooling
 * gather and the pixel-inference GEMM, models/wesup.py:398, consume). */
#define WESUP_CHW 0
#define WESUP_HWC 1

#define WESUP_E_ARG (-1)      /* null pointer / non-positive size */
#define WESUP_E_UNSUPPORTED (-2) /* valid request outside what the kernels cover */
#define WESUP_E_ALIGN (-3)    /* pointer or channel count breaks the 16-byte vector contract */

int wesup_abi_version(void);
const char *wesup_last_error(void);
/* diagnostic: number of CUDA kernels this library has launched in this process
 * (all threads); bench.py reports the delta over its timed region. */
unsigned long long wesup_kernel_launches(void);

/* ---- (a) hypercolumn: bilinear(align_corners=True) upsample + concat --------
 * Replaces WESUP._hook_fn's F.interpolate + torch.cat chain
 * (models/wesup.py:254-261) for all 13 side outputs in ONE launch.
 * side[l]: fp32, (C[l],h[l],w[l]) for WESUP_CHW or (h[l],w[l],C[l]) for
 * WESUP_HWC.  out: (sum C, H, W) or (H*W, sum C) in out_dtype.  `side`, `C`,
 * `h`, `w` are HOST arrays of n_levels entries.  HWC needs C[l] % 4 == 0. */
int wesup_hypercolumn_fwd(const void *const *side, const int *C, const int *h, const int *w,
                          int n_levels, int H, int W, void *out, int out_dtype, int layout,
                          void *stream);
/* adjoint of the above (what autograd derives for models/wesup.py:254-261):
 * grad_side[l] (fp32, same layout as side[l]) = bilinear^T of the level's
 * channel slice of grad_out.  Deterministic (no atomics).  `ws` (size from
 * wesup_hypercolumn_bwd_workspace_bytes) enables the separable two-pass kernels
 * that stream grad_out once (WESUP_HWC); with ws == NULL, or for WESUP_CHW, the
 * single-pass gather kernels run. */
size_t wesup_hypercolumn_bwd_workspace_bytes(const int *C, const int *h, const int *w, int n_levels,
                                             int H, int W);
int wesup_hypercolumn_bwd(const void *grad_out, int grad_dtype, int layout, const int *C,
                          const int *h, const int *w, int n_levels, int H, int W,
                          void *const *grad_side, void *ws, void *stream);

/* ---- superpixel statistics: replaces _preprocess_superpixels ---------------
 * (models/wesup.py:18-63) without the dense (N,H,W) maps.
 * labels: (H*W) int32 ids in [0,n_sp).  mask: (n_cls,H,W) int64 one-hot-or-zero
 * planes (utils/data.py:140-142,501-508) or NULL (= no supervision, :53-54).
 * Outputs, all in the reference's row order "labeled ids ascending, then
 * unlabeled ids ascending" (:45-47):
 *   order[k]      original id of row k                           (n_sp)
 *   row_labels[p] row index of pixel p (the relabelled map)      (H*W)
 *   counts[k]     |S_k|                                          (n_sp)
 *   seg_offsets   CSR offsets into seg_pixels                    (n_sp+1)
 *   seg_pixels    pixel ids grouped by row, ascending inside     (H*W)
 *   sp_labels     quantised multi-hot labels (:50-52), rows >= n_labeled zero (n_sp*n_cls)
 *   n_labeled     device scalar */
size_t wesup_sp_stats_workspace_bytes(int H, int W, int n_sp, int n_cls);
int wesup_sp_stats(const int32_t *labels, const int64_t *mask, int H, int W, int n_cls, int n_sp,
                   int32_t *order, int32_t *row_labels, int32_t *counts, int32_t *seg_offsets,
                   int32_t *seg_pixels, float *sp_labels, int32_t *n_labeled, void *ws,
                   void *stream);

/* ---- (b) superpixel mean pooling -------------------------------------------
 * fwd replaces torch.mm(sp_maps, x.t()) (models/wesup.py:284-285):
 *   pooled[k,c] = mean_{p in S_k} feat[p,c]             pooled: (N,C) fp32
 * bwd is the adjoint autograd derives for that mm:
 *   grad_feat[p,c] = grad_pooled[row_labels[p],c] / counts[row_labels[p]] */
int wesup_sp_pool_fwd(const void *feat, int dtype, int layout, const int32_t *seg_offsets,
                      const int32_t *seg_pixels, int HW, int C, int N, float *pooled,
                      void *stream);
int wesup_sp_pool_bwd(const float *grad_pooled, const int32_t *row_labels, const int32_t *counts,
                      int HW, int C, int N, void *grad_feat, int dtype, int layout, void *stream);

/* ---- fused (b) o (a): superpixel means straight from the feature levels -------
 * pooled = wesup_sp_pool_fwd(wesup_hypercolumn_fwd(level)) (models/wesup.py:254-261
 * then :284-285) WITHOUT the (H*W, sum C) tensor, in the footprint formulation: the
 * bilinear tap weights of a superpixel's pixels are aggregated once per low-resolution
 * cell (scalar work shared by all channels and all levels of that resolution), then
 *   pooled[k, coff_l + c] = 1/|S_k| * sum_cells G_k(cell) * level_l[cell, c].
 * level[l]: fp32 pixel-major (h[l], w[l], C[l]); `level`, `C`, `h`, `w` are HOST arrays;
 * pooled: (N, sum C) fp32.  Callers: the 13 side outputs (sum C = 2112), or the 13
 * backbone conv outputs ("pool first", sum C = 4224) followed by the 1x1 side
 * convolutions on N rows -- mean and 1x1 conv commute.  C[l] % 4 == 0, C[l] <= 1536. */
int wesup_levels_pool_fwd(const void *const *level, const int *C, const int *h, const int *w,
                          int n_levels, int H, int W, const int32_t *seg_offsets,
                          const int32_t *seg_pixels, int N, float *pooled, void *stream);
/* adjoint of the above (what autograd derives for the mm at models/wesup.py:284-285
 * followed by the cat + interpolate chain at :254-261), evaluated from the pooled
 * gradient (N, sum C): grad_level[l] (fp32 (h[l], w[l], C[l]), fully overwritten).
 * One warp per low-resolution cell: reads the cell's footprint of the label map once, folds
 * the tap weights per superpixel (hash table, fixed-point integer adds), then gathers the
 * grad_pooled rows of the superpixels it met in ascending id order.  Deterministic.
 * `ws` (wesup_levels_pool_bwd_workspace_bytes) holds small per-axis footprint tables. */
size_t wesup_levels_pool_bwd_workspace_bytes(const int *C, const int *h, const int *w, int n_levels,
                                             int H, int W);
int wesup_levels_pool_bwd(const float *grad_pooled, const int32_t *row_labels, const int32_t *counts,
                          const int *C, const int *h, const int *w, int n_levels, int H, int W,
                          int N, void *const *grad_level, void *ws, void *stream);

/* ---- the same two operators over PRECOMPUTED footprints ------------------------
 * The weights G_k(cell) depend only on the label map and the level resolutions: they are
 * the sparse form of the reference's dense `sp_maps` (models/wesup.py:57-61) composed with
 * the interpolation of :254-255, and like `sp_maps` they are built once per image in
 * preprocessing.  wesup_footprint_build fills an opaque device blob `fp`
 * (wesup_footprint_bytes; depends on h, w, H, W, N only) with, per distinct non-identity
 * resolution, the forward lists (per superpixel: cells + weights) and -- with_bwd != 0 --
 * their transpose (per cell: superpixel rows + weights / |S_k|, ascending rows).  The pooling
 * kernels are then prologue-free streaming gathers; results equal wesup_levels_pool_fwd/bwd
 * up to fp32 summation order.  `h`, `w`, H, W, N must be the ones the blob was built with;
 * rows k with an empty CSR segment pool to zero (fixed-capacity callers, CUDA graphs).
 * Any C[l] % 4 == 0.  Levels of 32/64/128/256/512 channels (forward) and 128/256/512 (backward)
 * take the whole-cell kernels -- a warp reads / writes complete C[l]*4-byte cell rows; other
 * channel counts take the 128-channel-chunk kernels, which the environment variables
 * WESUP_FP_FWD=chunks / WESUP_FP_BWD=chunks also select (cross-check hook of the tests). */
size_t wesup_footprint_bytes(const int *h, const int *w, int n_levels, int H, int W, int N);
int wesup_footprint_build(const int *h, const int *w, int n_levels, int H, int W, int N,
                          const int32_t *seg_offsets, const int32_t *seg_pixels,
                          const int32_t *row_labels, const int32_t *counts, int with_bwd,
                          void *fp, void *stream);
int wesup_levels_pool_fwd_fp(const void *const *level, const int *C, const int *h, const int *w,
                             int n_levels, int H, int W, const int32_t *seg_offsets,
                             const int32_t *seg_pixels, int N, const void *fp, float *pooled,
                             void *stream);
int wesup_levels_pool_bwd_fp(const float *grad_pooled, const int32_t *row_labels,
                             const int32_t *counts, const int *C, const int *h, const int *w,
                             int n_levels, int H, int W, int N, const void *fp,
                             void *const *grad_level, void *stream);

/* Historical names of the fused path (same signatures as ABI 3): they now run the two
 * footprint kernels above (`ws` from wesup_sp_pool_hypercolumn_bwd_workspace_bytes serves both). */
int wesup_hypercolumn_pool_fwd(const void *const *side, const int *C, const int *h, const int *w,
                               int n_levels, int H, int W, const int32_t *seg_offsets,
                               const int32_t *seg_pixels, int N, float *pooled, void *stream);
size_t wesup_sp_pool_hypercolumn_bwd_workspace_bytes(const int *C, const int *h, const int *w,
                                                     int n_levels, int H, int W, int N);
int wesup_sp_pool_hypercolumn_bwd(const float *grad_pooled, const int32_t *row_labels,
                                  const int32_t *counts, const int *C, const int *h, const int *w,
                                  int n_levels, int H, int W, int N, void *const *grad_side,
                                  void *ws, void *stream);
/* The per-pixel walk formulation of the same two operators (4 taps per pixel per channel;
 * kept as an independent cross-check of the footprint kernels and as their measured
 * predecessor).  The bwd needs `ws` of wesup_sp_pool_hypercolumn_bwd_workspace_bytes. */
int wesup_hypercolumn_pool_fwd_walk(const void *const *side, const int *C, const int *h, const int *w,
                                    int n_levels, int H, int W, const int32_t *seg_offsets,
                                    const int32_t *seg_pixels, int N, float *pooled, void *stream);
int wesup_sp_pool_hypercolumn_bwd_walk(const float *grad_pooled, const int32_t *row_labels,
                                       const int32_t *counts, const int *C, const int *h, const int *w,
                                       int n_levels, int H, int W, int N, void *const *grad_side,
                                       void *ws, void *stream);

/* ---- bias gradient of a channels_last convolution ----------------------------
 * out[c] = sum_p x[p, c] over a pixel-major (rows, C) fp32 matrix: what autograd computes as
 * grad_out.sum((0,2,3)) for every nn.Conv2d the reference builds (models/wesup.py:190-210).
 * Deterministic two-stage reduction; `ws` from wesup_colsum_workspace_bytes.  C % 4 == 0, C <= 1024. */
size_t wesup_colsum_workspace_bytes(long rows, int C);
int wesup_colsum(const float *x, long rows, int C, float *out, void *ws, void *stream);

/* ---- paint: replaces argmax + per-superpixel index_put loop -----------------
 * (models/wesup.py:295-304): out[p] = sp_pred[row_labels[p], cls]. */
int wesup_sp_paint(const int32_t *row_labels, const float *sp_pred, int HW, int n_cls, int cls,
                   float *out, void *stream);

/* ---- (c) label propagation: replaces _label_propagate ----------------------
 * (models/wesup.py:99-139).  feats (N,D) fp32, rows [0,n_l) labeled;
 * y_l (n_l,n_cls).  For every unlabeled row u: sim = exp(-min_j ||f_u-f_j||^2),
 * src = first arg-max, y_u[u] = y_l[src] iff sim > thr (strict) else 0.
 * y_u (N-n_l,n_cls); src_idx, max_sim (N-n_l) may be NULL. */
size_t wesup_label_propagate_workspace_bytes(int N, int D, int n_l);
/* dispatcher: the tcgen05 path when D == 32 and there are at least
 * WESUP_LP_TC_MIN_LABELED (default 64, the measured crossover) labeled rows,
 * else the CUDA-core path; both give bit-identical src_idx / max_sim / y_u. */
int wesup_label_propagate(const float *feats, int N, int D, int n_l, const float *y_l, int n_cls,
                          float thr, float *y_u, int32_t *src_idx, float *max_sim, void *ws,
                          void *stream);
/* CUDA-core path: direct-difference fp32 distances, any D. */
int wesup_label_propagate_exact(const float *feats, int N, int D, int n_l, const float *y_l,
                                int n_cls, float thr, float *y_u, int32_t *src_idx, float *max_sim,
                                void *ws, void *stream);
/* tensor-core path (D == 32): split-TF32 tcgen05.mma cross term with TMEM
 * accumulators as a candidate filter + exact fp32 re-evaluation of the
 * survivors in the epilogue (see csrc/label_propagate_tc.cu).  Two launches:
 * a prep kernel that splits every feature row once into the operand layout
 * (held in `ws`: 288 bytes per 128-row-padded row + the merge buffers) and the
 * warp-specialised pipeline (cp.async.bulk ring, four TMEM accumulators, eight
 * epilogue warps), chained by programmatic dependent launch. */
size_t wesup_label_propagate_tc_workspace_bytes(int N, int D, int n_l);
int wesup_label_propagate_tc(const float *feats, int N, int D, int n_l, const float *y_l, int n_cls,
                             float thr, float *y_u, int32_t *src_idx, float *max_sim, void *ws,
                             void *stream);
/* diagnostics of the last tensor-core call on `ws` (synchronise the stream first;
 * HOST out[2]): out[0] = exact re-evaluations, out[1] = float bits of
 * max |approx d2 - exact d2| / (|a|^2 + |b|^2) over the re-evaluated pairs. */
int wesup_label_propagate_tc_stats(const void *ws, int N, int n_l, unsigned long long *out_host);

/* Same operator with the row counts in DEVICE memory (counts_dev = {n_rows, n_labeled}, both
 * <= n_max): the launch depends only on the capacity n_max, so one CUDA graph serves images
 * with different superpixel counts.  feats (n_max, D); y_l (n_max, n_cls), rows >= n_labeled
 * ignored; y_full (n_max, n_cls) is written completely: rows [n_labeled, n_rows) as y_u above,
 * all other rows zero (the rows _cross_entropy skips, models/wesup.py:84-90).  Same fp32
 * arithmetic as wesup_label_propagate_exact. */
int wesup_label_propagate_dev(const float *feats, int n_max, int D, const int32_t *counts_dev,
                              const float *y_l, int n_cls, float thr, float *y_full, void *stream);

/* ---- pixel-wise inference: first MLP layer without the hypercolumn ----------------
 * out[p, :] = act(bias + sum_g bilinear_align_corners(z[g])[p, :]) for n_terms <= 5 pixel-major terms
 * z[g] (h[g], w[g], C) of one dtype (WESUP_F32 / WESUP_BF16), out (H*W, C) in the same dtype, bias (C) fp32 or
 * NULL, relu != 0 applies max(., 0).  With z[g] = cat_{levels of resolution g}(backbone level) . W'_g^T this is
 * ReLU(Linear(2112,1024)(hypercolumn)) of WESUPPixelInference.forward (models/wesup.py:392-400, :246-261) by
 * linearity of the 1x1 side convolutions, the upsampling and the Linear layer -- the GEMMs run at the levels' own
 * resolution and the (H*W, 2112) tensor is never formed.  `z`, `h`, `w` are HOST arrays.  C % 4 == 0, C <= 1024
 * (bf16 with C % 8 == 0 takes the 8-channels-per-thread kernel).  At most one full-resolution and four
 * low-resolution terms. */
int wesup_upsample_sum(const void *const *z, const int *h, const int *w, int n_terms, int H, int W, int C,
                       int dtype, const float *bias, int relu, void *out, void *stream);

/* ---- (d) SLIC: replaces skimage.segmentation.slic at models/wesup.py:471-476
 * rgb: fp32 in [0,1], (3,H,W) for WESUP_CHW (what the trainer holds) or (H,W,3).
 * labels: (H*W) int32, 0-based, contiguous, numbered in raster order of first
 * pixel; n_labels: device scalar.  Connectivity enforcement applies skimage's
 * min_size = int(0.5*H*W/n_segments) merge and max_size = int(3*H*W/n_segments) cut.
 * Two persistent cooperative launches per call (k-means sweeps; connectivity).
 * The _batch form runs B same-sized images (rgb (B,3,H,W) or (B,H,W,3), labels
 * (B,H*W), n_labels (B)) in the same two launches; results equal B single calls
 * bit for bit (integer / fixed-point cluster sums). */
size_t wesup_slic_workspace_bytes(int H, int W, int n_segments);
int wesup_slic(const float *rgb, int rgb_layout, int H, int W, int n_segments, double compactness,
               int max_iter, int enforce_connectivity, int32_t *labels, int32_t *n_labels,
               void *ws, void *stream);
size_t wesup_slic_batch_workspace_bytes(int B, int H, int W, int n_segments);
int wesup_slic_batch(const float *rgb, int rgb_layout, int B, int H, int W, int n_segments,
                     double compactness, int max_iter, int enforce_connectivity, int32_t *labels,
                     int32_t *n_labels, void *ws, void *stream);
/* Measurement tooling: %globaltimer stamps (ns) that block 0 wrote after every grid barrier of the last
 * wesup_slic[_batch] call on `ws`; out_host[64].  Synchronises the device. */
int wesup_slic_debug_times(const void *ws, int B, int H, int W, int n_segments, unsigned long long *out_host);
/* The connectivity pass alone on arbitrary int32 label maps seg (B,H*W)
 * (skimage's _enforce_label_connectivity_cython: 4-connected pieces in raster
 * order, breadth-first search capped at max_size, pieces below min_size merged
 * into the last labelled neighbour seen).  Requires min_size <= max_size. */
size_t wesup_enforce_connectivity_workspace_bytes(int B, int H, int W);
int wesup_enforce_connectivity(const int32_t *seg, int B, int H, int W, int min_size, int max_size,
                               int32_t *labels, int32_t *n_labels, void *ws, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* WESUP_B200_H */
