// fusionocc_b200 — the step BEFORE the splat: depth softmax + channel split + NCHW -> NHWC
// transpose + fp16/bf16 -> fp32 conversion in ONE pass over the depth-net output.
//
// Reference (projects/FusionOcc/fusionocc/necks/view_transformer.py:329-336 and the same lines of the
// mmdet3d / TEOcc / STCOcc copies; fusion_view_transformer.py:247-250):
//     x = depth_net(x)                                   (B*N, D + C [+ extra], H, W)
//     depth    = x[:, :D].softmax(dim=1)                 slice view + softmax kernel (fp32 under autocast)
//     tran_feat = x[:, D:D+C]                            slice view
//     ... feat.permute(0,1,3,4,2) -> bev_pool.py:20-21   .contiguous().float(): transpose copy (+ cast copies)
// i.e. 3-5 launches that each stream the tensor again.  Here: one CTA per 32 consecutive pixels of one image
// reads every channel row once (128-byte coalesced), keeps the tile in shared memory, writes the fp32 softmax in
// (B*N, D, H, W) and the context features channels-last in (B*N, H, W, C), the layout the splat gathers rows from.
// The backward applies the softmax Jacobian per pixel, grad_x = (grad - sum_d(grad * y)) * y, and transposes the
// feature gradient back.  HBM-bound elementwise work; fp32 arithmetic like the reference (exp(x - max) / sum).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace fo {

constexpr int kLiftPix = 32;        // pixels per CTA (one 128-byte row segment per channel)
constexpr int kLiftThreads = 128;   // 4 warps: warp w takes channels w, w+4, ...

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// smem: tile[(D + C)][33] floats (+1 padding: the transpose reads columns), red[4][32]
template <typename T>
__global__ void __launch_bounds__(kLiftThreads) lift_prepare_fwd_kernel(const T *__restrict__ x, int c_in, int D, int C,
                                                                        int HW, float *__restrict__ depth,
                                                                        float *__restrict__ feat) {
    extern __shared__ float sm[];
    float *tile = sm;                                  // [(D + C)][33]
    float *red = sm + (size_t)(D + C) * 33;            // [4][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t bn = blockIdx.y;
    const int hw0 = blockIdx.x * kLiftPix, hw = hw0 + lane;
    const bool ok = hw < HW;
    const T *xb = x + bn * (int64_t)c_in * HW;
    // one pass over the rows of this tile; running max of the depth logits per pixel
    float mx = -INFINITY;
    for (int c = warp; c < D + C; c += 4) {
        const float v = ok ? to_f32<T>(xb[(int64_t)c * HW + hw]) : 0.f;
        tile[c * 33 + lane] = v;
        if (c < D) mx = fmaxf(mx, v);
    }
    red[warp * 32 + lane] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[lane], red[32 + lane]), fmaxf(red[64 + lane], red[96 + lane]));
    __syncthreads();
    float sum = 0.f;
    for (int d = warp; d < D; d += 4) {
        const float e = expf(tile[d * 33 + lane] - mx);
        tile[d * 33 + lane] = e;
        sum += e;
    }
    red[warp * 32 + lane] = sum;
    __syncthreads();
    sum = (red[lane] + red[32 + lane]) + (red[64 + lane] + red[96 + lane]);
    if (ok) {
        float *db = depth + bn * (int64_t)D * HW + hw;
        for (int d = warp; d < D; d += 4) db[(int64_t)d * HW] = tile[d * 33 + lane] / sum;
    }
    // context features: rows of C floats per pixel, the tile's pixels are consecutive rows -> one contiguous run
    const int npix = min(kLiftPix, HW - hw0);
    float *fb = feat + (bn * HW + hw0) * (int64_t)C;
    for (int e = threadIdx.x; e < npix * C; e += kLiftThreads) {
        const int p = e / C, c = e - p * C;
        fb[e] = tile[(D + c) * 33 + p];
    }
}

template <typename T>
__global__ void __launch_bounds__(kLiftThreads) lift_prepare_bwd_kernel(const float *__restrict__ y,
                                                                        const float *__restrict__ dy,
                                                                        const float *__restrict__ dfeat, int c_in,
                                                                        int D, int C, int HW, T *__restrict__ dx) {
    extern __shared__ float sm[];
    float *tile = sm;                                  // [C][33]: the feature gradient, transposed
    float *red = sm + (size_t)C * 33;                  // [4][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t bn = blockIdx.y;
    const int hw0 = blockIdx.x * kLiftPix, hw = hw0 + lane;
    const bool ok = hw < HW;
    const int npix = min(kLiftPix, HW - hw0);
    const float *fb = dfeat + (bn * HW + hw0) * (int64_t)C;
    for (int e = threadIdx.x; e < npix * C; e += kLiftThreads) {
        const int p = e / C, c = e - p * C;
        tile[c * 33 + p] = fb[e];
    }
    // softmax Jacobian: s = sum_d dy*y ; dx = (dy - s) * y
    const float *yb = y + bn * (int64_t)D * HW + hw, *gb = dy + bn * (int64_t)D * HW + hw;
    float s = 0.f;
    if (ok)
        for (int d = warp; d < D; d += 4) s = fmaf(gb[(int64_t)d * HW], yb[(int64_t)d * HW], s);
    red[warp * 32 + lane] = s;
    __syncthreads();
    s = (red[lane] + red[32 + lane]) + (red[64 + lane] + red[96 + lane]);
    T *xb = dx + bn * (int64_t)c_in * HW + hw;
    if (ok) {
        for (int d = warp; d < D; d += 4)
            xb[(int64_t)d * HW] = from_f32<T>((gb[(int64_t)d * HW] - s) * yb[(int64_t)d * HW]);
        for (int c = warp; c < C; c += 4) xb[(int64_t)(D + c) * HW] = from_f32<T>(tile[c * 33 + lane]);
        for (int c = D + C + warp; c < c_in; c += 4) xb[(int64_t)c * HW] = from_f32<T>(0.f);
    }
}

}  // namespace fo

using namespace fo;

namespace {
int check_lift(const void *x, int32_t dtype, int64_t BN, int32_t c_in, int32_t D, int32_t C, int32_t HW) {
    FO_CHECK_ARG(x != nullptr, "x is NULL");
    FO_CHECK_ARG(dtype >= FO_DTYPE_F32 && dtype <= FO_DTYPE_BF16, "unknown dtype %d", dtype);
    FO_CHECK_ARG(BN >= 1 && BN <= 65535 && D >= 1 && C >= 1 && HW >= 1 && c_in >= D + C,
                 "bad sizes BN=%lld c_in=%d D=%d C=%d HW=%d", (long long)BN, c_in, D, C, HW);
    FO_CHECK_ARG((size_t)(D + C) * 33 * 4 + 512 <= 200 * 1024, "D + C = %d channels do not fit the shared-memory tile", D + C);
    return FO_OK;
}
}  // namespace

extern "C" int fo_lift_prepare_forward(fo_stream_t stream_, const void *x, int32_t x_dtype, int64_t BN, int32_t c_in,
                                       int32_t D, int32_t C, int32_t HW, float *depth, float *feat_nhwc) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc = check_lift(x, x_dtype, BN, c_in, D, C, HW)) return rc;
    FO_CHECK_ARG(depth && feat_nhwc, "NULL output");
    const size_t smem = ((size_t)(D + C) * 33 + 128) * sizeof(float);
    const dim3 grid((HW + kLiftPix - 1) / kLiftPix, (unsigned)BN);
#define FO_LIFT_FWD(T)                                                                                              \
    do {                                                                                                            \
        if (smem > 48 * 1024)                                                                                       \
            FO_CUDA(cudaFuncSetAttribute(lift_prepare_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                         (int)smem));                                                               \
        lift_prepare_fwd_kernel<T><<<grid, kLiftThreads, smem, stream>>>((const T *)x, c_in, D, C, HW, depth,       \
                                                                         feat_nhwc);                                \
    } while (0)
    if (x_dtype == FO_DTYPE_F32) FO_LIFT_FWD(float);
    else if (x_dtype == FO_DTYPE_F16) FO_LIFT_FWD(__half);
    else FO_LIFT_FWD(__nv_bfloat16);
#undef FO_LIFT_FWD
    FO_LAUNCH_CHECK("lift_prepare_fwd_kernel");
    return FO_OK;
}

extern "C" int fo_lift_prepare_backward(fo_stream_t stream_, const float *depth, const float *depth_grad,
                                        const float *feat_nhwc_grad, int64_t BN, int32_t c_in, int32_t D, int32_t C,
                                        int32_t HW, void *x_grad, int32_t x_dtype) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc = check_lift(x_grad, x_dtype, BN, c_in, D, C, HW)) return rc;
    FO_CHECK_ARG(depth && depth_grad && feat_nhwc_grad, "NULL input");
    const size_t smem = ((size_t)C * 33 + 128) * sizeof(float);
    const dim3 grid((HW + kLiftPix - 1) / kLiftPix, (unsigned)BN);
#define FO_LIFT_BWD(T)                                                                                              \
    do {                                                                                                            \
        if (smem > 48 * 1024)                                                                                       \
            FO_CUDA(cudaFuncSetAttribute(lift_prepare_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                         (int)smem));                                                               \
        lift_prepare_bwd_kernel<T><<<grid, kLiftThreads, smem, stream>>>(depth, depth_grad, feat_nhwc_grad, c_in, D, \
                                                                         C, HW, (T *)x_grad);                       \
    } while (0)
    if (x_dtype == FO_DTYPE_F32) FO_LIFT_BWD(float);
    else if (x_dtype == FO_DTYPE_F16) FO_LIFT_BWD(__half);
    else FO_LIFT_BWD(__nv_bfloat16);
#undef FO_LIFT_BWD
    FO_LAUNCH_CHECK("lift_prepare_bwd_kernel");
    return FO_OK;
}
