// L0 drop-in proof: the two launchers that the reference's own, UNMODIFIED
// mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp declares at :7-14 and calls at :46-56 / :92-103, implemented by
// forwarding to the source-compatible symbols of libfusionocc_b200.so.  This file replaces
// src/bev_pool_cuda.cu in the reference's setup.py:235-244; nothing else of the reference changes.
#include "fusionocc_b200.h"

void bev_pool_v2(int c, int n_intervals, const float* depth, const float* feat, const int* ranks_depth,
                 const int* ranks_feat, const int* ranks_bev, const int* interval_starts,
                 const int* interval_lengths, float* out) {
  fo_compat_bev_pool_v2(c, n_intervals, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts,
                        interval_lengths, out);
}

void bev_pool_v2_grad(int c, int n_intervals, const float* out_grad, const float* depth, const float* feat,
                      const int* ranks_depth, const int* ranks_feat, const int* ranks_bev,
                      const int* interval_starts, const int* interval_lengths, float* depth_grad,
                      float* feat_grad) {
  fo_compat_bev_pool_v2_grad(c, n_intervals, out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                             interval_starts, interval_lengths, depth_grad, feat_grad);
}
