// fusionocc_b200 — C-ABI plumbing: error state, version, and the host-buffer convenience entry.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace fo {

static thread_local char g_err[512] = "";

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace fo

using namespace fo;

#define FO_STR2(x) #x
#define FO_STR(x) FO_STR2(x)
#define FO_STR_CUDA_VERSION FO_STR(__CUDACC_VER_MAJOR__) "." FO_STR(__CUDACC_VER_MINOR__)

extern "C" int fo_abi_version(void) { return FO_ABI_VERSION; }
extern "C" const char *fo_last_error(void) { return g_err; }
extern "C" const char *fo_build_info(void) {
    return "fusionocc_b200 abi=2 arch=sm_100a subtile=32vox fwd=1warp/subtile bwd=tma-gather+multi-pixel cuda=" FO_STR_CUDA_VERSION;
}

// ------------------------------------------------------------------------------------------------
// Host-buffer entry.  Workspace carving (all 256-byte aligned), in this order:
//   coor (or frustum | cam_mats | bda) | depth | feat | out | [out_grad | depth_grad | feat_grad] | ranks_bev | ranks_depth |
//   ranks_feat | interval_starts | interval_lengths | counts | fwd plan | rank scratch |
//   [bwd plan | bwd scratch]
// ------------------------------------------------------------------------------------------------
namespace {
struct HostWs {
    float *coor, *frustum, *cam, *bda, *depth, *feat, *out, *og, *dg, *fg;
    int32_t *rb, *rd, *rf, *st, *ln, *counts;
    void *fwd_plan; size_t fwd_plan_bytes;
    void *rank_scratch; size_t rank_scratch_bytes;
    void *bwd_plan; size_t bwd_plan_bytes;
    void *bwd_scratch; size_t bwd_scratch_bytes;
    size_t total;
};
HostWs carve(void *base, int64_t B, int64_t N, int64_t D, int64_t H, int64_t W, int64_t c, int64_t X, int64_t Y,
             int64_t Z, bool bwd, bool calib = false) {
    HostWs w;
    char *p = (char *)base;
    auto take = [&](int64_t bytes) { char *r = p; p += align_up(bytes, 256); return (void *)r; };
    const int64_t P = B * N * D * H * W, V = X * Y * Z, NV = B * V, rows = B * N * H * W;
    const int64_t cap_iv = P < NV ? P : NV;
    w.coor = calib ? nullptr : (float *)take(P * 12);
    w.frustum = calib ? (float *)take(D * H * W * 12) : nullptr;
    w.cam = calib ? (float *)take(B * N * 24 * 4) : nullptr;
    w.bda = calib ? (float *)take(B * 12 * 4) : nullptr;
    w.depth = (float *)take(P * 4);
    w.feat = (float *)take(rows * c * 4);
    w.out = (float *)take(NV * c * 4);
    w.og = bwd ? (float *)take(NV * c * 4) : nullptr;
    w.dg = bwd ? (float *)take(P * 4) : nullptr;
    w.fg = bwd ? (float *)take(rows * c * 4) : nullptr;
    w.rb = (int32_t *)take(P * 4);
    w.rd = (int32_t *)take(P * 4);
    w.rf = (int32_t *)take(P * 4);
    w.st = (int32_t *)take(cap_iv * 4);
    w.ln = (int32_t *)take(cap_iv * 4);
    w.counts = (int32_t *)take(256);
    w.fwd_plan_bytes = fo_fwd_plan_bytes(NV, P);
    w.fwd_plan = take((int64_t)w.fwd_plan_bytes);
    w.rank_scratch_bytes = fo_rank_prepare_scratch_bytes(P, NV);
    w.rank_scratch = take((int64_t)w.rank_scratch_bytes);
    w.bwd_plan_bytes = bwd ? fo_bwd_plan_bytes(P, rows) : 0;
    w.bwd_plan = bwd ? take((int64_t)w.bwd_plan_bytes) : nullptr;
    w.bwd_scratch_bytes = bwd ? fo_bwd_scratch_bytes(cap_iv, (int32_t)c, FO_LAYOUT_BCZYX) : 0;
    w.bwd_scratch = bwd ? take((int64_t)w.bwd_scratch_bytes) : nullptr;
    w.total = (size_t)(p - (char *)base);
    return w;
}
}  // namespace

extern "C" size_t fo_view_transform_host_workspace_bytes(int32_t B, int32_t N, int32_t D, int32_t H, int32_t W,
                                                         int32_t c, int32_t X, int32_t Y, int32_t Z,
                                                         int32_t with_backward) {
    if (B < 1 || N < 1 || D < 1 || H < 1 || W < 1 || c < 1 || X < 1 || Y < 1 || Z < 1) return 0;
    return carve(nullptr, B, N, D, H, W, c, X, Y, Z, with_backward != 0).total;
}

namespace {
struct HostCalib {
    const float *frustum, *cam, *bda;
    int32_t bda_has_t, mode;
};
int view_transform_host_impl(fo_stream_t stream_, const float *coor_host, const HostCalib *cal, const float *depth_host,
                             const float *feat_host, const float *out_grad_host, int32_t B, int32_t N, int32_t D,
                             int32_t H, int32_t W, int32_t c, const float lower_bound[3], const float interval[3],
                             int32_t X, int32_t Y, int32_t Z, float *out_host, float *depth_grad_host,
                             float *feat_grad_host, int32_t counts_host[4], void *workspace_dev, size_t workspace_bytes,
                             fo_stream_t upload_stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaStream_t up = upload_stream_ ? (cudaStream_t)upload_stream_ : stream;
    FO_CHECK_ARG(B >= 1 && N >= 1 && D >= 1 && H >= 1 && W >= 1 && c >= 1 && X >= 1 && Y >= 1 && Z >= 1,
                 "non-positive dimension");
    FO_CHECK_ARG((coor_host || cal) && depth_host && feat_host && out_host && workspace_dev, "NULL buffer");
    FO_CHECK_ARG(!cal || (cal->frustum && cal->cam && cal->bda), "NULL calibration buffer");
    const bool bwd = out_grad_host != nullptr;
    FO_CHECK_ARG(!bwd || (depth_grad_host && feat_grad_host), "backward requested but gradient outputs are NULL");
    FO_CHECK_ARG(((uintptr_t)workspace_dev & 255) == 0, "workspace must be 256-byte aligned");
    HostWs w = carve(workspace_dev, B, N, D, H, W, c, X, Y, Z, bwd, cal != nullptr);
    if (workspace_bytes < w.total)
        return set_error(FO_ERR_SCRATCH, "workspace is %zu bytes, need %zu", workspace_bytes, w.total);
    const int64_t P = (int64_t)B * N * D * H * W, V = (int64_t)X * Y * Z, NV = B * V, rows = (int64_t)B * N * H * W;
    const int64_t cap_iv = P < NV ? P : NV;

    if (cal) {
        FO_CUDA(cudaMemcpyAsync(w.frustum, cal->frustum, (size_t)D * H * W * 12, cudaMemcpyHostToDevice, stream));
        FO_CUDA(cudaMemcpyAsync(w.cam, cal->cam, (size_t)B * N * 96, cudaMemcpyHostToDevice, stream));
        FO_CUDA(cudaMemcpyAsync(w.bda, cal->bda, (size_t)B * 48, cudaMemcpyHostToDevice, stream));
    } else {
        FO_CUDA(cudaMemcpyAsync(w.coor, coor_host, (size_t)P * 12, cudaMemcpyHostToDevice, stream));
    }
    FO_CUDA(cudaMemcpyAsync(w.depth, depth_host, (size_t)P * 4, cudaMemcpyHostToDevice, stream));
    FO_CUDA(cudaMemcpyAsync(w.feat, feat_host, (size_t)rows * c * 4, cudaMemcpyHostToDevice, stream));
    // The 82 MB/sample out_grad upload is not needed before the backward: on a second stream it overlaps the
    // forward and, PCIe being full duplex, the 82 MB/sample download of the voxel tensor.
    cudaEvent_t og_ready = nullptr;
    if (bwd) {
        if (up != stream) {
            cudaEvent_t ws_free;
            FO_CUDA(cudaEventCreateWithFlags(&ws_free, cudaEventDisableTiming));
            FO_CUDA(cudaEventRecord(ws_free, stream));             // earlier users of the workspace are done
            FO_CUDA(cudaStreamWaitEvent(up, ws_free, 0));
            FO_CUDA(cudaEventDestroy(ws_free));
        }
        FO_CUDA(cudaMemcpyAsync(w.og, out_grad_host, (size_t)NV * c * 4, cudaMemcpyHostToDevice, up));
        if (up != stream) {
            FO_CUDA(cudaEventCreateWithFlags(&og_ready, cudaEventDisableTiming));
            FO_CUDA(cudaEventRecord(og_ready, up));
        }
    }

    int rc;
    if (cal)
        rc = fo_rank_prepare_calib(stream_, w.frustum, w.cam, w.bda, cal->bda_has_t, cal->mode, nullptr, B, N, D, H, W,
                                   lower_bound, interval, X, Y, Z, w.rb, w.rd, w.rf, w.st, w.ln, w.counts, w.fwd_plan,
                                   w.fwd_plan_bytes, w.rank_scratch, w.rank_scratch_bytes);
    else
        rc = fo_rank_prepare(stream_, w.coor, B, N, D, H, W, lower_bound, interval, X, Y, Z, w.rb, w.rd, w.rf, w.st,
                             w.ln, w.counts, w.fwd_plan, w.fwd_plan_bytes, w.rank_scratch, w.rank_scratch_bytes);
    if (rc) return rc;
    if (int rc2 = fo_bev_pool_v2_forward(stream_, c, w.depth, w.feat, w.rd, w.rf, w.rb, w.st, w.ln, P, cap_iv,
                                         w.counts + 1, B, V, w.out, FO_LAYOUT_BCZYX, FO_FWD_ASSUME_SORTED, w.fwd_plan,
                                         w.fwd_plan_bytes))
        return rc2;
    FO_CUDA(cudaMemcpyAsync(out_host, w.out, (size_t)NV * c * 4, cudaMemcpyDeviceToHost, stream));
    if (bwd) {
        if (og_ready) {
            FO_CUDA(cudaStreamWaitEvent(stream, og_ready, 0));
            FO_CUDA(cudaEventDestroy(og_ready));
        }
        if (int rc2 = fo_bwd_plan_build(stream_, w.rd, w.rf, P, w.counts, P, rows, H * W, FO_BWD_PLAN_STRUCTURED,
                                        w.fwd_plan, w.fwd_plan_bytes, B, V, w.bwd_plan, w.bwd_plan_bytes))
            return rc2;
        if (int rc2 = fo_bev_pool_v2_backward(stream_, c, w.og, FO_LAYOUT_BCZYX, w.depth, w.feat, P, cap_iv, B, V, P,
                                              rows, w.dg, w.fg, w.fwd_plan, w.fwd_plan_bytes, w.bwd_plan,
                                              w.bwd_plan_bytes, w.bwd_scratch, w.bwd_scratch_bytes))
            return rc2;
        FO_CUDA(cudaMemcpyAsync(depth_grad_host, w.dg, (size_t)P * 4, cudaMemcpyDeviceToHost, stream));
        FO_CUDA(cudaMemcpyAsync(feat_grad_host, w.fg, (size_t)rows * c * 4, cudaMemcpyDeviceToHost, stream));
    }
    if (counts_host)
        FO_CUDA(cudaMemcpyAsync(counts_host, w.counts, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    return FO_OK;
}
}  // namespace

extern "C" int fo_view_transform_host(fo_stream_t stream_, const float *coor_host, const float *depth_host,
                                      const float *feat_host, const float *out_grad_host, int32_t B, int32_t N,
                                      int32_t D, int32_t H, int32_t W, int32_t c, const float lower_bound[3],
                                      const float interval[3], int32_t X, int32_t Y, int32_t Z, float *out_host,
                                      float *depth_grad_host, float *feat_grad_host, int32_t counts_host[4],
                                      void *workspace_dev, size_t workspace_bytes, fo_stream_t upload_stream_) {
    FO_CHECK_ARG(coor_host != nullptr, "coor_host is NULL");
    return view_transform_host_impl(stream_, coor_host, nullptr, depth_host, feat_host, out_grad_host, B, N, D, H, W, c,
                                    lower_bound, interval, X, Y, Z, out_host, depth_grad_host, feat_grad_host,
                                    counts_host, workspace_dev, workspace_bytes, upload_stream_);
}

extern "C" size_t fo_view_transform_host_calib_workspace_bytes(int32_t B, int32_t N, int32_t D, int32_t H, int32_t W,
                                                               int32_t c, int32_t X, int32_t Y, int32_t Z,
                                                               int32_t with_backward) {
    if (B < 1 || N < 1 || D < 1 || H < 1 || W < 1 || c < 1 || X < 1 || Y < 1 || Z < 1) return 0;
    return carve(nullptr, B, N, D, H, W, c, X, Y, Z, with_backward != 0, true).total;
}

extern "C" int fo_view_transform_host_calib(fo_stream_t stream_, const float *frustum_host, const float *cam_mats_host,
                                            const float *bda_host, int32_t bda_has_translation, int32_t matvec_mode,
                                            const float *depth_host, const float *feat_host,
                                            const float *out_grad_host, int32_t B, int32_t N, int32_t D, int32_t H,
                                            int32_t W, int32_t c, const float lower_bound[3], const float interval[3],
                                            int32_t X, int32_t Y, int32_t Z, float *out_host, float *depth_grad_host,
                                            float *feat_grad_host, int32_t counts_host[4], void *workspace_dev,
                                            size_t workspace_bytes, fo_stream_t upload_stream_) {
    HostCalib cal{frustum_host, cam_mats_host, bda_host, bda_has_translation ? 1 : 0, matvec_mode};
    return view_transform_host_impl(stream_, nullptr, &cal, depth_host, feat_host, out_grad_host, B, N, D, H, W, c,
                                    lower_bound, interval, X, Y, Z, out_host, depth_grad_host, feat_grad_host,
                                    counts_host, workspace_dev, workspace_bytes, upload_stream_);
}
