/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's two bev_pool_v2 CUDA kernels, with the
 * exact arithmetic contract the sm_100a build of the reference has (every
 * multiply-add contracted to a single-rounding FFMA):
 *
 *   forward : /root/reference/mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-48
 *   backward: /root/reference/mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:67-121
 *
 * Parity status: PINNED by the reference's only known-answer test
 * (mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176, tests/test_oracle_kat.py) and, on
 * the GPU box, by the reference CUDA extension itself built from the reference
 * sources into oracle/_ref (tests/test_gpu_vs_reference_ext.py).
 *
 * Built by oracle/Makefile into oracle/liboracle_bevpool.so with
 * -ffp-contract=off so that only the explicit fmaf() calls fuse.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* bev_pool_cuda.cu:21-48 — one (interval, channel) per CUDA thread there; a
 * double loop here.  psum starts at +0.0f and accumulates in interval order;
 * the result is ASSIGNED to out[ranks_bev[start]*c + ch] (caller pre-zeroes). */
void oracle_bev_pool_v2_fwd(int c, int n_intervals,
                            const float *depth, const float *feat,
                            const int32_t *ranks_depth, const int32_t *ranks_feat,
                            const int32_t *ranks_bev,
                            const int32_t *interval_starts, const int32_t *interval_lengths,
                            float *out)
{
    for (int k = 0; k < n_intervals; ++k) {
        const int s = interval_starts[k];
        const int len = interval_lengths[k];
        float *o = out + (size_t)ranks_bev[s] * c;
        for (int ch = 0; ch < c; ++ch) {
            float psum = 0.0f;
            for (int i = 0; i < len; ++i) {
                const float d = depth[ranks_depth[s + i]];
                const float f = feat[(size_t)ranks_feat[s + i] * c + ch];
                psum = fmaf(f, d, psum);                 /* :42  psum += *cur_feat * *cur_depth */
            }
            o[ch] = psum;                                 /* :47 */
        }
    }
}

/* bev_pool_cuda.cu:67-121 — one backward interval (a run of equal ranks_feat in
 * the ranks_feat-sorted arrays, bev_pool.py:47-57) per CUDA thread there. */
void oracle_bev_pool_v2_bwd(int c, int n_intervals,
                            const float *out_grad, const float *depth, const float *feat,
                            const int32_t *ranks_depth, const int32_t *ranks_feat,
                            const int32_t *ranks_bev,
                            const int32_t *interval_starts, const int32_t *interval_lengths,
                            float *depth_grad, float *feat_grad)
{
    for (int k = 0; k < n_intervals; ++k) {
        const int s = interval_starts[k];
        const int len = interval_lengths[k];
        for (int i = 0; i < len; ++i) {                   /* :91-105 depth grad, sequential over c */
            const float *g = out_grad + (size_t)ranks_bev[s + i] * c;
            const float *f = feat + (size_t)ranks_feat[s + i] * c;
            float sum = 0.0f;
            for (int ch = 0; ch < c; ++ch)
                sum = fmaf(g[ch], f[ch], sum);            /* :100 */
            depth_grad[ranks_depth[s + i]] = sum;         /* :104 */
        }
        float *fg = feat_grad + (size_t)ranks_feat[s] * c;
        for (int ch = 0; ch < c; ++ch) {                  /* :109-120 feat grad, sequential over the run */
            float sum = 0.0f;
            for (int i = 0; i < len; ++i) {
                const float g = out_grad[(size_t)ranks_bev[s + i] * c + ch];
                const float d = depth[ranks_depth[s + i]];
                sum = fmaf(g, d, sum);                    /* :116 */
            }
            fg[ch] = sum;                                 /* :119 */
        }
    }
}

/* (B,Z,Y,X,C) -> (B,C,Z,Y,X) contiguous, the wrapper's final permute
 * (/root/reference/mmdet3d/ops/bev_pool_v2/bev_pool.py:91). n_vox = Z*Y*X. */
void oracle_permute_to_bczyx(int b, int c, size_t n_vox, const float *in, float *out)
{
    for (int ib = 0; ib < b; ++ib)
        for (size_t v = 0; v < n_vox; ++v)
            for (int ch = 0; ch < c; ++ch)
                out[((size_t)ib * c + ch) * n_vox + v] = in[((size_t)ib * n_vox + v) * c + ch];
}

/* Inverse of the above: out_grad.contiguous() at bev_pool.py:69. */
void oracle_permute_to_bzyxc(int b, int c, size_t n_vox, const float *in, float *out)
{
    for (int ib = 0; ib < b; ++ib)
        for (int ch = 0; ch < c; ++ch)
            for (size_t v = 0; v < n_vox; ++v)
                out[((size_t)ib * n_vox + v) * c + ch] = in[((size_t)ib * c + ch) * n_vox + v];
}
