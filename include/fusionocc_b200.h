/*
 * fusionocc_b200 — C ABI of the B200-native camera->voxel view transformation
 * (bev_pool_v2 forward / backward splat + the rank-precompute stage).
 *
 * This header is the drop-in boundary.  Every entry point takes plain device
 * pointers, sizes and a CUDA stream; no torch / pybind types appear.  The
 * library allocates nothing and frees nothing: all buffers (inputs, outputs and
 * scratch) are owned by the caller, exactly like the reference extension, whose
 * Python side allocates `out` / the gradients and lends raw pointers for the
 * duration of one launch (reference: mmdet3d/ops/bev_pool_v2/bev_pool.py:27,67-68;
 * src/bev_pool.cpp:40-56,86-103).
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp:7-9    void bev_pool_v2(int c, int n_intervals, ...)
 *   mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp:11-14  void bev_pool_v2_grad(int c, int n_intervals, ...)
 *   mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp:30-57  bev_pool_v2_forward  (pybind, at::Tensor)
 *   mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp:74-104 bev_pool_v2_backward (pybind, at::Tensor)
 *   mmdet3d/ops/bev_pool_v2/bev_pool.py:47-57       backward re-sort + interval rebuild (torch ops)
 *   mmdet3d/ops/bev_pool_v2/bev_pool.py:91          permute(0,4,1,2,3).contiguous()
 *   projects/FusionOcc/fusionocc/necks/view_transformer.py:223-281  voxel_pooling_prepare_v2 (torch ops)
 *
 * Conventions
 *   - All functions return FO_OK (0) or an FO_ERR_* code; fo_last_error() returns a
 *     thread-local human-readable message for the last failure on this thread.
 *   - All launches go to `stream` (a cudaStream_t passed as void*); nothing uses
 *     the legacy default stream implicitly and nothing synchronises the host.
 *   - Re-entrant: no global mutable state besides the thread-local error string.
 *   - Index arrays are int32, values are fp32, like the reference extension.
 *   - "ranks" arrays are sorted by ranks_bev (ties in ascending ranks_depth);
 *     interval k covers positions [interval_starts[k], +interval_lengths[k]).
 *
 * Built for sm_100a only.  There is no CPU fallback.
 */
#ifndef FUSIONOCC_B200_H_
#define FUSIONOCC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FO_ABI_VERSION 2

#define FO_OK                  0
#define FO_ERR_INVALID_ARG     1   /* null pointer, negative size, bad enum, misaligned buffer   */
#define FO_ERR_CUDA            2   /* a CUDA runtime call or launch failed (message has details) */
#define FO_ERR_UNSUPPORTED     3   /* shape outside what the kernels implement                  */
#define FO_ERR_SCRATCH         4   /* scratch / plan buffer smaller than fo_*_bytes() asked for */

/* fo_bev_pool_v2_forward flags */
#define FO_FWD_ASSUME_SORTED 1   /* plan is known clean (built by fo_rank_prepare, or its flags were read back
                                    as 0): skip launching the two device-guarded order-agnostic kernels */

/* Memory layout of the dense voxel tensor (the forward output / the backward's out_grad). */
#define FO_LAYOUT_BCZYX  0   /* contiguous (B,C,Z,Y,X): what bev_pool_v2() returns (bev_pool.py:91)      */
#define FO_LAYOUT_BZYXC  1   /* contiguous (B,Z,Y,X,C): what the reference extension itself writes/reads */

/* Element type of a tensor passed as void* (fo_lift_prepare_*). */
#define FO_DTYPE_F32   0
#define FO_DTYPE_F16   1
#define FO_DTYPE_BF16  2

typedef void *fo_stream_t;   /* cudaStream_t */

int         fo_abi_version(void);
const char *fo_last_error(void);
/* Compile-time facts, for logs and tests: "sm_100a", tile size, etc. */
const char *fo_build_info(void);

/* ------------------------------------------------------------------------------------------------
 * Forward plan.  A small device-side index over the interval list that lets the forward kernel own
 * dense output tiles: tile_first_interval[t] for every tile of FO_TILE_VOXELS consecutive voxels,
 * a voxel-sortedness flag, and pos->interval ids for the backward.  Depends only on
 * (ranks_bev, interval_starts, interval_lengths, B, Z*Y*X): build once and reuse while those are
 * unchanged (the reference's `accelerate=True` caching, view_transformer.py:175-194), or rebuild per
 * call (a few microseconds).
 *
 * n_intervals_dev: optional device pointer to the live interval count (int32).  When non-NULL the
 * kernels read the count from there (sync-free pipelines) and `n_intervals` is only the capacity of
 * the interval arrays; when NULL `n_intervals` is the count.
 * ------------------------------------------------------------------------------------------------ */
size_t fo_fwd_plan_bytes(int64_t n_voxels_total /* B*Z*Y*X */, int64_t n_points_capacity);

int fo_fwd_plan_build(fo_stream_t stream,
                      const int32_t *ranks_bev, const int32_t *interval_starts,
                      const int32_t *interval_lengths,
                      int64_t n_points, int64_t n_intervals, const int32_t *n_intervals_dev,
                      int32_t B, int64_t n_voxels_per_sample /* Z*Y*X */,
                      void *plan, size_t plan_bytes);

/* ------------------------------------------------------------------------------------------------
 * bev_pool_v2 forward.  Replaces bev_pool.cpp:7-9 + the 82 MB new_zeros (bev_pool.py:27) + the
 * permute copy (bev_pool.py:91): every element of `out` is written exactly once, zeros included, in
 * `out_layout` order, with per-interval sequential FMA accumulation in interval order starting from
 * +0.0f (bit-identical to bev_pool_cuda.cu:39-47).
 *
 *   depth        fp32 [n_depth]           flat (B,N,D,H,W), indexed by ranks_depth
 *   feat         fp32 [n_feat_rows, c]    flat (B,N,H,W,C), rows indexed by ranks_feat
 *   out          fp32 B*c*n_voxels_per_sample elements, layout `out_layout`
 *   plan         from fo_fwd_plan_build for the same index arrays
 * Interval voxels that are not strictly increasing, or ranks outside the grid, are detected by the
 * plan; the call then takes the order-agnostic scatter path (same results for valid input).  The
 * choice is made ON THE DEVICE from the plan's flag word (no host sync): both paths are enqueued and
 * the wrong one exits at once, unless FO_FWD_ASSUME_SORTED is passed.
 * ------------------------------------------------------------------------------------------------ */
int fo_bev_pool_v2_forward(fo_stream_t stream, int32_t c,
                           const float *depth, const float *feat,
                           const int32_t *ranks_depth, const int32_t *ranks_feat,
                           const int32_t *ranks_bev,
                           const int32_t *interval_starts, const int32_t *interval_lengths,
                           int64_t n_points, int64_t n_intervals, const int32_t *n_intervals_dev,
                           int32_t B, int64_t n_voxels_per_sample,
                           float *out, int32_t out_layout, int32_t flags,
                           const void *plan, size_t plan_bytes);

/* Same, writing a CHANNEL SLICE [c_offset, c_offset + c) of a wider voxel tensor with c_total channels
 * ((B,c_total,Z,Y,X) or (B,Z,Y,X,c_total), `out` = base of the wide tensor): what the consumer builds
 * with torch.cat over temporal frames (projects/FusionOcc/fusionocc/detectors/fusion_occ.py:316-326) is
 * written in place, one call per frame, no concatenation copy.  Only the slice is
 * written. */
int fo_bev_pool_v2_forward_slice(fo_stream_t stream, int32_t c,
                                 const float *depth, const float *feat,
                                 const int32_t *ranks_depth, const int32_t *ranks_feat,
                                 const int32_t *ranks_bev,
                                 const int32_t *interval_starts, const int32_t *interval_lengths,
                                 int64_t n_points, int64_t n_intervals, const int32_t *n_intervals_dev,
                                 int32_t B, int64_t n_voxels_per_sample,
                                 float *out, int32_t out_layout, int32_t c_total, int32_t c_offset,
                                 int32_t flags, const void *plan, size_t plan_bytes);

/* ------------------------------------------------------------------------------------------------
 * Backward plan = the inverse interval ordering: forward positions regrouped by ranks_feat (stable),
 * i.e. the arrays the reference rebuilds with argsort / where on every backward
 * (bev_pool.py:47-57), stored as (depth index, forward interval) entries per image pixel.  Depends only
 * on the index arrays; cache it with the forward plan.  Pass the SAME plan_bytes to every call that
 * takes a plan buffer: the buffer layout is derived from it.
 * ------------------------------------------------------------------------------------------------ */
size_t fo_bwd_plan_bytes(int64_t n_points_capacity, int64_t n_feat_rows);

/* flags of fo_bwd_plan_build */
#define FO_BWD_PLAN_STRUCTURED 1  /* the forward plan was produced by fo_rank_prepare: every point p =
                                     ((b*N+n)*D+d)*hw_size+hw belongs to feature row (b*N+n)*hw_size+hw and the
                                     plan holds the point->position map, so the inverse ordering is built
                                     per pixel in registers (no sort).  n_points_capacity must be >= n_depth. */

int fo_bwd_plan_build(fo_stream_t stream, const int32_t *ranks_depth, const int32_t *ranks_feat,
                      int64_t n_points, const int32_t *n_points_dev,
                      int64_t n_depth, int64_t n_feat_rows, int32_t hw_size /* H*W, structured only */,
                      int32_t flags,
                      const void *fwd_plan, size_t fwd_plan_bytes, int32_t B, int64_t n_voxels_per_sample,
                      void *plan, size_t plan_bytes);

/* ------------------------------------------------------------------------------------------------
 * bev_pool_v2 backward.  Replaces bev_pool.py:44-83 + bev_pool.cpp:11-14: no re-sort, no
 * out_grad.contiguous() copy, no atomics.  depth_grad / feat_grad are fully written (zeros where no
 * point contributes).  depth_grad[p] = sum_c out_grad[v,c]*feat[q,c] sequentially over c;
 * feat_grad[q,c] = sum_i out_grad[v_i,c]*depth[p_i] sequentially in (ranks_bev, position) order —
 * the reference's orders, so both are bit-identical to bev_pool_cuda.cu:91-120.
 *
 *   out_grad     fp32, layout `og_layout` (FO_LAYOUT_BCZYX is the gradient of bev_pool_v2()'s output)
 *   n_points / n_intervals   capacities the plans were built with (live counts are in the plans)
 *   fwd_plan / bwd_plan      the index arrays themselves are not needed: the plans carry everything
 *   scratch      fo_bwd_scratch_bytes() bytes (compact gathered out_grad rows when og is BCZYX)
 * ------------------------------------------------------------------------------------------------ */
size_t fo_bwd_scratch_bytes(int64_t n_intervals_capacity, int32_t c, int32_t og_layout);

int fo_bev_pool_v2_backward(fo_stream_t stream, int32_t c,
                            const float *out_grad, int32_t og_layout,
                            const float *depth, const float *feat,
                            int64_t n_points, int64_t n_intervals,
                            int32_t B, int64_t n_voxels_per_sample,
                            int64_t n_depth, int64_t n_feat_rows,
                            float *depth_grad, float *feat_grad,
                            const void *fwd_plan, size_t fwd_plan_bytes,
                            const void *bwd_plan, size_t bwd_plan_bytes,
                            void *scratch, size_t scratch_bytes);

/* Backward INCLUDING the construction of the structured backward plan (what fo_bwd_plan_build with
 * FO_BWD_PLAN_STRUCTURED followed by fo_bev_pool_v2_backward does; `bwd_plan` is written by this call and can be
 * passed to later fo_bev_pool_v2_backward calls while the rank arrays are unchanged).  When out_grad is contiguous
 * (B,C,Z,Y,X), D <= 128 and B*Z*Y*X < 2^24 the plan is built by extra warps of the gather kernel — the gather is
 * DRAM-bound, the plan issue-bound, one kernel makes them share every SM — otherwise by a launch of its own.
 * Requires a forward plan produced by fo_rank_prepare[_calib].  n_points_dev: the live point count (counts_dev). */
int fo_bev_pool_v2_backward_with_plan(fo_stream_t stream, int32_t c,
                                      const float *out_grad, int32_t og_layout,
                                      const float *depth, const float *feat,
                                      int64_t n_points, const int32_t *n_points_dev, int64_t n_intervals,
                                      int32_t B, int64_t n_voxels_per_sample,
                                      int64_t n_depth, int64_t n_feat_rows, int32_t hw_size,
                                      float *depth_grad, float *feat_grad,
                                      const void *fwd_plan, size_t fwd_plan_bytes,
                                      void *bwd_plan, size_t bwd_plan_bytes,
                                      void *scratch, size_t scratch_bytes);

/* Same, reading the gradient of a channel slice [c_offset, c_offset + c) out of the gradient of the wide
 * tensor (`out_grad` = base of the (B,c_total,Z,Y,X) / (B,Z,Y,X,c_total) gradient): no slicing copy. */
int fo_bev_pool_v2_backward_slice(fo_stream_t stream, int32_t c,
                                  const float *out_grad, int32_t og_layout, int32_t c_total, int32_t c_offset,
                                  const float *depth, const float *feat,
                                  int64_t n_points, int64_t n_intervals,
                                  int32_t B, int64_t n_voxels_per_sample,
                                  int64_t n_depth, int64_t n_feat_rows,
                                  float *depth_grad, float *feat_grad,
                                  const void *fwd_plan, size_t fwd_plan_bytes,
                                  const void *bwd_plan, size_t bwd_plan_bytes,
                                  void *scratch, size_t scratch_bytes);

/* ------------------------------------------------------------------------------------------------
 * Rank precompute.  Replaces view_transformer.py:223-281 (about 50 eager torch launches, >= 4 host
 * syncs, a library sort) by: voxelise + count, one scan, placement, in-interval ordering.
 *
 *   coor         fp32 [B*N*D*H*W, 3] frustum points in ego space (get_lidar_coor output)
 *   lower/interval  grid_lower_bound[3], grid_interval[3] (host floats; index = trunc((coor-lb)/itv),
 *                IEEE sub then IEEE div, truncation toward zero, view_transformer.py:246-248)
 *   grid X,Y,Z   integer grid size
 * Outputs (capacity n_points_total each for ranks_*, min(n_points_total, B*Z*Y*X) for intervals):
 *   ranks_bev / ranks_depth / ranks_feat / interval_starts / interval_lengths  int32
 *   counts_dev   int32[4] on device: {n_kept, n_intervals, 0, 0}
 *   fwd_plan     required: filled as by fo_fwd_plan_build (the scan produces the tile table for free)
 * Integer arithmetic is exact (the reference's fp32 rank arithmetic is exact only for
 * B*Z*Y*X < 2^24; above that the reference itself is wrong).
 * ------------------------------------------------------------------------------------------------ */
size_t fo_rank_prepare_scratch_bytes(int64_t n_points_total, int64_t n_voxels_total);

int fo_rank_prepare(fo_stream_t stream, const float *coor,
                    int32_t B, int32_t N, int32_t D, int32_t H, int32_t W,
                    const float lower_bound[3], const float interval[3],
                    int32_t X, int32_t Y, int32_t Z,
                    int32_t *ranks_bev, int32_t *ranks_depth, int32_t *ranks_feat,
                    int32_t *interval_starts, int32_t *interval_lengths,
                    int32_t *counts_dev,
                    void *fwd_plan, size_t fwd_plan_bytes,
                    void *scratch, size_t scratch_bytes);

/* ------------------------------------------------------------------------------------------------
 * Rank precompute with the geometry fused in.  Replaces view_transformer.py:135-173
 * (get_lidar_coor: ~15 eager launches and 4-6 full-size (B,N,D,H,W,3) temporaries) AND :223-281: the
 * frustum point is computed per thread from the calibration and never stored (unless coor_out != NULL).
 *
 *   frustum      fp32 [D*H*W, 3]   the (x_px, y_px, depth) template of create_frustum (:105-133)
 *   cam_mats     fp32 [B*N, 24]    per camera: inv(post_rots) 3x3 row-major | post_trans 3 |
 *                                  combine = sensor2ego[:3,:3] @ inv(cam2img) 3x3 | sensor2ego[:3,3]
 *                                  (the caller forms the two small products with the reference's own ops)
 *   bda          fp32 [B, 12]      bda 3x3 row-major | translation 3 (STCOcc's 4x4 bda; zeros otherwise)
 *   matvec_mode  fp32 summation order of the three per-point 3x3 products: 0 = FMA chain k = 0,1,2;
 *                1 = separate multiplies and adds; 2 = FMA chain k = 2,1,0; 3 = fma(m1,y, m0*x) + m2*z.
 *                The reference runs them through a batched library GEMM; mode 3 reproduces its bits on a
 *                B200 with torch 2.11 / CUDA 12.8 (tests/test_gpu_fused_geometry.py checks that, and
 *                measures the voxel-index mismatch rate of the other orders).
 *   coor_out     optional fp32 [B*N*D*H*W, 3]: also materialise the points (debug / parity tests)
 * Everything else as fo_rank_prepare.
 * ------------------------------------------------------------------------------------------------ */
int fo_rank_prepare_calib(fo_stream_t stream, const float *frustum, const float *cam_mats, const float *bda,
                          int32_t bda_has_translation, int32_t matvec_mode, float *coor_out,
                          int32_t B, int32_t N, int32_t D, int32_t H, int32_t W,
                          const float lower_bound[3], const float interval[3],
                          int32_t X, int32_t Y, int32_t Z,
                          int32_t *ranks_bev, int32_t *ranks_depth, int32_t *ranks_feat,
                          int32_t *interval_starts, int32_t *interval_lengths,
                          int32_t *counts_dev,
                          void *fwd_plan, size_t fwd_plan_bytes,
                          void *scratch, size_t scratch_bytes);

/* ------------------------------------------------------------------------------------------------
 * Rank pipeline from integer bucket ids — the sibling ops with the same sort -> interval -> reduce skeleton
 * (SURVEY.md §8f-4): bev_pool v1 (projects/BEVFusion/bevfusion/ops/bev_pool/bev_pool.py:85-99: ranks from
 * integer coords, argsort, kept/where interval rebuild) and occ_pool (projects/CONet/.../occ_pooling/
 * OCC_Pool.py:39-71).  Stable: ties keep ascending original index.
 *
 *   keys              int32 [n_points]  bucket (voxel) id in [0, n_buckets); anything else drops the point
 *   sorted_keys       int32 [n_points]  out: keys in sorted order                    (= ranks_bev)
 *   order             int32 [n_points]  out: original index of every sorted position (= ranks_depth/_feat)
 *   interval_starts / interval_lengths  out, capacity min(n_points, n_buckets)
 *   counts_dev        int32[4] on device: {n_kept, n_intervals, 0, 0}
 * ------------------------------------------------------------------------------------------------ */
size_t fo_rank_from_keys_scratch_bytes(int64_t n_points, int64_t n_buckets);

int fo_rank_from_keys(fo_stream_t stream, const int32_t *keys, int64_t n_points, int64_t n_buckets,
                      int32_t *sorted_keys, int32_t *order,
                      int32_t *interval_starts, int32_t *interval_lengths, int32_t *counts_dev,
                      void *scratch, size_t scratch_bytes);

/* ------------------------------------------------------------------------------------------------
 * The step before the splat: depth softmax + channel split + NCHW -> NHWC transpose +
 * fp16/bf16 -> fp32 conversion of the depth-net output in one pass.  Replaces view_transformer.py:329-336
 * (x[:, :D].softmax(dim=1), x[:, D:D+C]) and the feat.contiguous().float() transpose copy of bev_pool.py:20-21.
 *
 *   x            (BN, c_in, H*W) contiguous, element type x_dtype; channels [0,D) depth logits,
 *                [D, D+C) context features, [D+C, c_in) ignored
 *   depth        fp32 (BN, D, H*W): softmax over D, computed in fp32 as exp(x - max) / sum
 *   feat_nhwc    fp32 (BN, H*W, C): the rows bev_pool_v2 gathers (ranks_feat index them)
 * Backward: x_grad[:, :D] = (depth_grad - sum_d(depth_grad * depth)) * depth (softmax Jacobian),
 * x_grad[:, D:D+C] = feat_nhwc_grad transposed back, x_grad[:, D+C:] = 0; x_grad has x's shape and dtype.
 * ------------------------------------------------------------------------------------------------ */
int fo_lift_prepare_forward(fo_stream_t stream, const void *x, int32_t x_dtype,
                            int64_t BN, int32_t c_in, int32_t D, int32_t C, int32_t HW,
                            float *depth, float *feat_nhwc);

int fo_lift_prepare_backward(fo_stream_t stream, const float *depth, const float *depth_grad,
                             const float *feat_nhwc_grad,
                             int64_t BN, int32_t c_in, int32_t D, int32_t C, int32_t HW,
                             void *x_grad, int32_t x_dtype);

/* ------------------------------------------------------------------------------------------------
 * Source-compatible L0 symbols.  Same C signatures, semantics (assign into a caller-zeroed
 * (B,Z,Y,X,C) `out`; backward arrays already re-sorted by ranks_feat) and stream behaviour (legacy
 * default stream) as the two launchers bev_pool.cpp declares at :7-14, so the reference's own
 * bev_pool.cpp can be linked against this library unchanged.
 * ------------------------------------------------------------------------------------------------ */
void fo_compat_bev_pool_v2(int c, int n_intervals, const float *depth, const float *feat,
                           const int *ranks_depth, const int *ranks_feat, const int *ranks_bev,
                           const int *interval_starts, const int *interval_lengths, float *out);

void fo_compat_bev_pool_v2_grad(int c, int n_intervals, const float *out_grad,
                                const float *depth, const float *feat,
                                const int *ranks_depth, const int *ranks_feat, const int *ranks_bev,
                                const int *interval_starts, const int *interval_lengths,
                                float *depth_grad, float *feat_grad);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer convenience entry (end-to-end measurement and non-torch hosts): H2D of the inputs,
 * rank precompute + forward (+ backward when out_grad_host != NULL), D2H of the results, all on
 * `stream`, using a caller-provided device workspace.  Host buffers should be pinned.  `upload_stream`
 * (may be NULL = `stream`) carries the out_grad upload, so that it overlaps the forward and the download
 * of the voxel tensor; the call orders the two streams with events and returns without synchronising.
 * ------------------------------------------------------------------------------------------------ */
size_t fo_view_transform_host_workspace_bytes(int32_t B, int32_t N, int32_t D, int32_t H, int32_t W,
                                              int32_t c, int32_t X, int32_t Y, int32_t Z,
                                              int32_t with_backward);

int fo_view_transform_host(fo_stream_t stream,
                           const float *coor_host, const float *depth_host, const float *feat_host,
                           const float *out_grad_host /* may be NULL: forward only */,
                           int32_t B, int32_t N, int32_t D, int32_t H, int32_t W, int32_t c,
                           const float lower_bound[3], const float interval[3],
                           int32_t X, int32_t Y, int32_t Z,
                           float *out_host /* (B,c,Z,Y,X) */,
                           float *depth_grad_host, float *feat_grad_host,
                           int32_t counts_host[4],
                           void *workspace_dev, size_t workspace_bytes,
                           fo_stream_t upload_stream);

/* Same, from the CALIBRATION instead of a materialised (B,N,D,H,W,3) point tensor: the rank precompute computes
 * every frustum point from the per-camera matrices (fo_rank_prepare_calib), so the host uploads 0.74 MB of frustum
 * template + 96 bytes per camera instead of 4.46 MB of points per sample.  Argument meaning as fo_rank_prepare_calib
 * (frustum [D*H*W,3], cam_mats [B*N,24], bda [B,12]) and fo_view_transform_host. */
size_t fo_view_transform_host_calib_workspace_bytes(int32_t B, int32_t N, int32_t D, int32_t H, int32_t W,
                                                    int32_t c, int32_t X, int32_t Y, int32_t Z,
                                                    int32_t with_backward);

int fo_view_transform_host_calib(fo_stream_t stream,
                                 const float *frustum_host, const float *cam_mats_host, const float *bda_host,
                                 int32_t bda_has_translation, int32_t matvec_mode,
                                 const float *depth_host, const float *feat_host,
                                 const float *out_grad_host /* may be NULL: forward only */,
                                 int32_t B, int32_t N, int32_t D, int32_t H, int32_t W, int32_t c,
                                 const float lower_bound[3], const float interval[3],
                                 int32_t X, int32_t Y, int32_t Z,
                                 float *out_host /* (B,c,Z,Y,X) */,
                                 float *depth_grad_host, float *feat_grad_host,
                                 int32_t counts_host[4],
                                 void *workspace_dev, size_t workspace_bytes,
                                 fo_stream_t upload_stream);

#ifdef __cplusplus
}
#endif
#endif /* FUSIONOCC_B200_H_ */
