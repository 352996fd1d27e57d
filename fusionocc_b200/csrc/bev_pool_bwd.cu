// fusionocc_b200 — bev_pool_v2 backward (sm_100a).
//
// Reference: mmdet3d/ops/bev_pool_v2/bev_pool.py:44-83 (argsort by ranks_feat, interval rebuild with a
// host sync, 164 MB out_grad.contiguous() un-permute) + src/bev_pool_cuda.cu:67-121 (ONE THREAD per
// image pixel: 4 224 threads at the headline shape).  Here:
//
//   gather   (only when out_grad is (B,C,Z,Y,X)): a CTA per forward tile reads the C x 128 block with
//            128-byte coalesced loads — predicated per lane on the tile's occupancy mask, so only the
//            32-byte sectors that contain an occupied voxel are fetched — transposes it through shared
//            memory and emits one compact channels-last row per interval: G[k, 0..C).  No 164 MB copy.
//   pixel    an 8-lane group per backward interval (= image pixel) walks the pixel's points in the
//            inverse interval ordering (backward plan).  depth_grad[p] is the sequential FMA chain over
//            c = 0..C-1 (passed lane to lane), feat_grad[q, :] the sequential FMA over the pixel's points:
//            the reference's exact orders, no atomics, every output written once.
#include <limits.h>

#include "common.cuh"

namespace fo {

struct GatherArgs {
    const float *og;                // (B,C,Z,Y,X)
    const int32_t *rb, *starts;
    int32_t C;
    int64_t V;
    const FwdPlanHeader *hdr;
    const int32_t *tile_off;
    float *G;                       // [n_intervals, C]
};

template <int NCHUNK>
__global__ void __launch_bounds__(kThreads) bwd_gather_kernel(GatherArgs a) {
    extern __shared__ __align__(16) float stage[];      // [kTile][C+1]
    __shared__ unsigned s_mask[kTile / 32];
    __shared__ int s_vl[kTile];
    const int tile = blockIdx.x;
    const int k0 = a.tile_off[tile], k1 = a.tile_off[tile + 1];
    if (k1 <= k0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = a.C, S = C + 1;
    const int tps = a.hdr->tiles_per_sample;
    const int b = tile / tps;
    const int64_t v0 = (int64_t)(tile - b * tps) * kTile;
    const int nv = (int)min((int64_t)kTile, a.V - v0);
    const int64_t vbase = (int64_t)b * a.V + v0;
    const int nk = min(k1 - k0, kTile);

    if (tid < kTile / 32) s_mask[tid] = 0u;
    __syncthreads();
    for (int kk = tid; kk < nk; kk += kThreads) {
        int vl = (int)(__ldg(a.rb + __ldg(a.starts + k0 + kk)) - vbase);
        if ((unsigned)vl >= (unsigned)nv) vl = -1;
        s_vl[kk] = vl;
        if (vl >= 0) atomicOr(&s_mask[vl >> 5], 1u << (vl & 31));
    }
    __syncthreads();
    {
        const float *plane0 = a.og + ((int64_t)b * C) * a.V + v0;
        unsigned m[kTile / 32];
#pragma unroll
        for (int r = 0; r < kTile / 32; ++r) m[r] = s_mask[r];
        for (int c = warp; c < C; c += kThreads / 32) {
            const float *src = plane0 + (int64_t)c * a.V;
#pragma unroll
            for (int r = 0; r < kTile / 32; ++r) {
                const int vl = lane + 32 * r;
                if ((m[r] >> lane) & 1u) stage[vl * S + c] = __ldcs(src + vl);
            }
        }
    }
    __syncthreads();
    if constexpr (NCHUNK > 0) {
        const int g = tid / kGroupLanes, gl = tid % kGroupLanes, c4 = C >> 2;
        for (int kk = g; kk < nk; kk += kGroupsPerCta) {
            const int vl = s_vl[kk];
            if (vl < 0) continue;
            const float *row = stage + vl * S;
            float *dst = a.G + (int64_t)(k0 + kk) * C;
#pragma unroll
            for (int ch = 0; ch < NCHUNK; ++ch) {
                const int idx = gl + kGroupLanes * ch;
                if (idx < c4)
                    *reinterpret_cast<float4 *>(dst + 4 * idx) =
                        make_float4(row[4 * idx], row[4 * idx + 1], row[4 * idx + 2], row[4 * idx + 3]);
            }
        }
    } else {
        for (int kk = warp; kk < nk; kk += kThreads / 32) {
            const int vl = s_vl[kk];
            if (vl < 0) continue;
            for (int c = lane; c < C; c += 32) a.G[(int64_t)(k0 + kk) * C + c] = stage[vl * S + c];
        }
    }
}

struct PixelArgs {
    const float *G;                 // gathered rows [n_intervals, C]  or  out_grad in (B,Z,Y,X,C)
    const int32_t *row_of_pos;      // pos2iv (G rows by interval) or ranks_bev (channels-last out_grad)
    const float *depth, *feat;
    const int32_t *rd, *rf;
    const int32_t *bwd_pos;         // nullptr: arrays are already in backward order (compat entry)
    const int32_t *bwd_starts, *bwd_lengths;
    const int32_t *n_bwd_dev;       // live number of backward intervals (nullptr: use n_bwd)
    int64_t n_bwd;
    int64_t n_rows_G;               // bound for row indices
    int64_t n_depth, n_feat_rows, n_points;
    int32_t C;
    float *depth_grad, *feat_grad;
};

// One 8-lane group per backward interval.  NCHUNK >= 1: float4 lanes (C % 4 == 0).
template <int NCHUNK>
__global__ void __launch_bounds__(kThreads) bwd_pixel_kernel(PixelArgs a) {
    const int lane = threadIdx.x & 31;
    const int gl = lane & (kGroupLanes - 1);
    const unsigned gmask = 0xffu << (lane & 24);
    const int gbase = lane & 24;
    const int C = a.C, c4 = C >> 2;
    const int64_t n = a.n_bwd_dev ? min((int64_t)max(*a.n_bwd_dev, 0), a.n_bwd) : a.n_bwd;
    const int64_t group0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kGroupLanes;
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / kGroupLanes;
    const int last_owner = (c4 - 1) & (kGroupLanes - 1);

    for (int64_t m = group0; m < n; m += ngroups) {
        const int s = __ldg(a.bwd_starts + m), len = __ldg(a.bwd_lengths + m);
        if (len <= 0 || s < 0 || (int64_t)s + len > a.n_points) continue;
        const int i0 = a.bwd_pos ? __ldg(a.bwd_pos + s) : s;
        const int q = __ldg(a.rf + i0);
        if (q < 0 || q >= a.n_feat_rows) continue;
        float4 f[NCHUNK], fg[NCHUNK];
#pragma unroll
        for (int ch = 0; ch < NCHUNK; ++ch) {
            const int idx = gl + kGroupLanes * ch;
            f[ch] = (idx < c4) ? ldg4(a.feat + ((int64_t)q * c4 + idx) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            fg[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // software pipeline: indices and rows of point j+1 are in flight while point j is reduced
        int p_n = 0, row_n = 0;
        float d_n = 0.f;
        float4 g_n[NCHUNK];
        auto fetch = [&](int j) {
            const int i = a.bwd_pos ? __ldg(a.bwd_pos + s + j) : s + j;
            p_n = __ldg(a.rd + i);
            row_n = __ldg(a.row_of_pos + i);
            const bool ok = p_n >= 0 && p_n < a.n_depth && row_n >= 0 && row_n < a.n_rows_G;
            d_n = ok ? __ldg(a.depth + p_n) : 0.f;
            if (!ok) p_n = -1;
#pragma unroll
            for (int ch = 0; ch < NCHUNK; ++ch) {
                const int idx = gl + kGroupLanes * ch;
                g_n[ch] = (ok && idx < c4) ? ldg4(a.G + ((int64_t)row_n * c4 + idx) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        fetch(0);
        for (int j = 0; j < len; ++j) {
            const int p = p_n;
            const float d = d_n;
            float4 g[NCHUNK];
#pragma unroll
            for (int ch = 0; ch < NCHUNK; ++ch) g[ch] = g_n[ch];
            if (j + 1 < len) fetch(j + 1);
            // feat grad: sequential over the pixel's points (bev_pool_cuda.cu:109-120)
#pragma unroll
            for (int ch = 0; ch < NCHUNK; ++ch) fma4(fg[ch], g[ch], d);
            // depth grad: sequential FMA chain over c = 0..C-1 (bev_pool_cuda.cu:96-101), handed from
            // lane to lane inside the group
            float sum = 0.f;
#pragma unroll
            for (int ch = 0; ch < NCHUNK; ++ch) {
#pragma unroll
                for (int o = 0; o < kGroupLanes; ++o) {
                    const int idx = ch * kGroupLanes + o;
                    if (idx < c4) {
                        if (idx > 0) sum = __shfl_sync(gmask, sum, gbase + ((o + kGroupLanes - 1) & (kGroupLanes - 1)));
                        if (gl == o) {
                            sum = fmaf(g[ch].x, f[ch].x, sum);
                            sum = fmaf(g[ch].y, f[ch].y, sum);
                            sum = fmaf(g[ch].z, f[ch].z, sum);
                            sum = fmaf(g[ch].w, f[ch].w, sum);
                        }
                    }
                }
            }
            if (gl == last_owner && p >= 0) a.depth_grad[p] = sum;
        }
#pragma unroll
        for (int ch = 0; ch < NCHUNK; ++ch) {
            const int idx = gl + kGroupLanes * ch;
            if (idx < c4) *reinterpret_cast<float4 *>(a.feat_grad + ((int64_t)q * c4 + idx) * 4) = fg[ch];
        }
    }
}

// Scalar path (any C): one warp per backward interval; lane 0 runs the depth chain from shared rows.
__global__ void __launch_bounds__(kThreads) bwd_pixel_scalar_kernel(PixelArgs a) {
    const int lane = threadIdx.x & 31;
    const int C = a.C;
    const int64_t n = a.n_bwd_dev ? min((int64_t)max(*a.n_bwd_dev, 0), a.n_bwd) : a.n_bwd;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m = warp0; m < n; m += nwarps) {
        const int s = a.bwd_starts[m], len = a.bwd_lengths[m];
        if (len <= 0 || s < 0 || (int64_t)s + len > a.n_points) continue;
        const int i0 = a.bwd_pos ? a.bwd_pos[s] : s;
        const int q = a.rf[i0];
        if (q < 0 || q >= a.n_feat_rows) continue;
        // feat grad: lane per channel, sequential over points
        for (int c = lane; c < C; c += 32) {
            float sum = 0.f;
            for (int j = 0; j < len; ++j) {
                const int i = a.bwd_pos ? a.bwd_pos[s + j] : s + j;
                const int p = a.rd[i], row = a.row_of_pos[i];
                if (p < 0 || p >= a.n_depth || row < 0 || row >= a.n_rows_G) continue;
                sum = fmaf(a.G[(int64_t)row * C + c], a.depth[p], sum);
            }
            a.feat_grad[(int64_t)q * C + c] = sum;
        }
        // depth grad: lane per point, sequential over channels
        for (int j = lane; j < len; j += 32) {
            const int i = a.bwd_pos ? a.bwd_pos[s + j] : s + j;
            const int p = a.rd[i], row = a.row_of_pos[i];
            if (p < 0 || p >= a.n_depth || row < 0 || row >= a.n_rows_G) continue;
            float sum = 0.f;
            for (int c = 0; c < C; ++c) sum = fmaf(a.G[(int64_t)row * C + c], a.feat[(int64_t)q * C + c], sum);
            a.depth_grad[p] = sum;
        }
    }
}

}  // namespace fo

using namespace fo;

extern "C" size_t fo_bwd_scratch_bytes(int64_t n_intervals_capacity, int32_t c, int32_t og_layout) {
    if (n_intervals_capacity < 0 || c < 1) return 0;
    if (og_layout == FO_LAYOUT_BZYXC) return 256;
    return (size_t)align_up(n_intervals_capacity * (int64_t)c * 4, 256) + 256;
}

namespace {
int launch_pixel(const PixelArgs &pa, bool vec, cudaStream_t stream) {
    const int64_t groups = pa.n_bwd;
    if (groups <= 0) return FO_OK;
    if (vec) {
        int64_t blocks = (groups * kGroupLanes + kThreads - 1) / kThreads;
        if (blocks > 148 * 32) blocks = 148 * 32;
        const int chunks = (pa.C / 4 + kGroupLanes - 1) / kGroupLanes;
        switch (chunks) {
            case 1: bwd_pixel_kernel<1><<<(int)blocks, kThreads, 0, stream>>>(pa); break;
            case 2: bwd_pixel_kernel<2><<<(int)blocks, kThreads, 0, stream>>>(pa); break;
            case 3: bwd_pixel_kernel<3><<<(int)blocks, kThreads, 0, stream>>>(pa); break;
            default: bwd_pixel_kernel<4><<<(int)blocks, kThreads, 0, stream>>>(pa); break;
        }
    } else {
        int64_t blocks = (groups * 32 + kThreads - 1) / kThreads;
        if (blocks > 148 * 32) blocks = 148 * 32;
        bwd_pixel_scalar_kernel<<<(int)blocks, kThreads, 0, stream>>>(pa);
    }
    FO_LAUNCH_CHECK("bwd_pixel_kernel");
    return FO_OK;
}
}  // namespace

// layout of the backward plan (must match rank_prepare.cu)
namespace {
struct BwdPlanPtrs { const BwdPlanHeader *hdr; const int32_t *pos, *starts, *lengths; };
BwdPlanPtrs bwd_plan_ptrs(const void *plan, int64_t n_pts, int64_t n_rows) {
    const char *p = (const char *)plan;
    BwdPlanPtrs r;
    r.hdr = (const BwdPlanHeader *)p;   p += 256;
    r.pos = (const int32_t *)p;         p += align_up(n_pts * 4, 256);
    r.starts = (const int32_t *)p;      p += align_up(n_rows * 4, 256);
    r.lengths = (const int32_t *)p;
    return r;
}
}  // namespace

extern "C" int fo_bev_pool_v2_backward(fo_stream_t stream_, int32_t c, const float *out_grad, int32_t og_layout,
                                       const float *depth, const float *feat, const int32_t *ranks_depth,
                                       const int32_t *ranks_feat, const int32_t *ranks_bev,
                                       const int32_t *interval_starts, const int32_t *interval_lengths,
                                       int64_t n_points, int64_t n_intervals, const int32_t *n_counts_dev, int32_t B,
                                       int64_t n_vox, int64_t n_depth, int64_t n_feat_rows, float *depth_grad,
                                       float *feat_grad, const void *fwd_plan, size_t fwd_plan_bytes,
                                       const void *bwd_plan, size_t bwd_plan_bytes, void *scratch,
                                       size_t scratch_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    (void)n_counts_dev; (void)interval_lengths;
    FO_CHECK_ARG(c >= 1 && B >= 1 && n_vox >= 1, "c, B and voxels per sample must be positive");
    FO_CHECK_ARG(og_layout == FO_LAYOUT_BCZYX || og_layout == FO_LAYOUT_BZYXC, "unknown og_layout %d", og_layout);
    FO_CHECK_ARG(depth_grad && feat_grad, "NULL gradient output");
    FO_CHECK_ARG(n_depth >= 0 && n_feat_rows >= 0 && n_points >= 0 && n_intervals >= 0, "negative size");
    FO_CHECK_ARG(n_points < INT_MAX && (int64_t)B * n_vox < INT_MAX, "sizes exceed int32 ranks");
    FO_CUDA(cudaMemsetAsync(depth_grad, 0, (size_t)n_depth * 4, stream));
    FO_CUDA(cudaMemsetAsync(feat_grad, 0, (size_t)n_feat_rows * c * 4, stream));
    if (n_points == 0 || n_intervals == 0) return FO_OK;
    FO_CHECK_ARG(out_grad && depth && feat && ranks_depth && ranks_feat && ranks_bev && interval_starts,
                 "NULL input array");
    FO_CHECK_ARG(fwd_plan && bwd_plan, "forward and backward plans are required");
    const int64_t tps = tiles_per_sample(n_vox);
    const int64_t n_tiles = tps * B;
    if (fwd_plan_bytes < 256 + fwd_plan_tile_bytes(n_tiles) + (size_t)align_up(n_points * 4, 256))
        return set_error(FO_ERR_SCRATCH, "forward plan buffer too small for backward (%zu bytes)", fwd_plan_bytes);
    if (bwd_plan_bytes < fo_bwd_plan_bytes(n_points, n_feat_rows))
        return set_error(FO_ERR_SCRATCH, "backward plan buffer too small (%zu bytes)", bwd_plan_bytes);
    FwdPlanView pv = fwd_plan_view(const_cast<void *>(fwd_plan), n_tiles);
    BwdPlanPtrs bp = bwd_plan_ptrs(bwd_plan, n_points, n_feat_rows);
    const bool vec = (c % 4 == 0) && (c <= 4 * kGroupLanes * kMaxChunks) && (((uintptr_t)feat & 15) == 0) &&
                     (((uintptr_t)feat_grad & 15) == 0);

    PixelArgs pa;
    pa.depth = depth; pa.feat = feat; pa.rd = ranks_depth; pa.rf = ranks_feat;
    pa.bwd_pos = bp.pos; pa.bwd_starts = bp.starts; pa.bwd_lengths = bp.lengths;
    pa.n_bwd_dev = &bp.hdr->n_bwd_intervals; pa.n_bwd = n_feat_rows;
    pa.n_depth = n_depth; pa.n_feat_rows = n_feat_rows; pa.n_points = n_points;
    pa.C = c; pa.depth_grad = depth_grad; pa.feat_grad = feat_grad;

    if (og_layout == FO_LAYOUT_BCZYX) {
        const size_t need = fo_bwd_scratch_bytes(n_intervals, c, og_layout);
        if (!scratch || scratch_bytes < need)
            return set_error(FO_ERR_SCRATCH, "backward scratch is %zu bytes, need %zu", scratch_bytes, need);
        const size_t smem = (size_t)kTile * (c + 1) * sizeof(float);
        if (smem > 200 * 1024) return set_error(FO_ERR_UNSUPPORTED, "C=%d too large for the gather tile", c);
        GatherArgs ga;
        ga.og = out_grad; ga.rb = ranks_bev; ga.starts = interval_starts; ga.C = c; ga.V = n_vox;
        ga.hdr = pv.hdr; ga.tile_off = pv.tile_off; ga.G = (float *)scratch;
        const bool gvec = (c % 4 == 0) && (c <= 4 * kGroupLanes * kMaxChunks);
        const int chunks = gvec ? (c / 4 + kGroupLanes - 1) / kGroupLanes : 0;
#define FO_GATHER(NC)                                                                                          \
    do {                                                                                                       \
        if (smem > 48 * 1024)                                                                                  \
            FO_CUDA(cudaFuncSetAttribute(bwd_gather_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                         (int)smem));                                                          \
        bwd_gather_kernel<NC><<<(int)n_tiles, kThreads, smem, stream>>>(ga);                                   \
    } while (0)
        switch (chunks) {
            case 1: FO_GATHER(1); break;
            case 2: FO_GATHER(2); break;
            case 3: FO_GATHER(3); break;
            case 4: FO_GATHER(4); break;
            default: FO_GATHER(0); break;
        }
#undef FO_GATHER
        FO_LAUNCH_CHECK("bwd_gather_kernel");
        pa.G = (const float *)scratch; pa.row_of_pos = pv.pos2iv; pa.n_rows_G = n_intervals;
    } else {
        pa.G = out_grad; pa.row_of_pos = ranks_bev; pa.n_rows_G = (int64_t)B * n_vox;
    }
    const bool pvec = vec && (((uintptr_t)pa.G & 15) == 0);
    return launch_pixel(pa, pvec, stream);
}

// Source-compatible launcher: bev_pool.cpp:11-14 / bev_pool_cuda.cu:133-140 semantics — arrays are
// already re-sorted by ranks_feat with backward intervals, out_grad is (B,Z,Y,X,C), outputs are
// caller-zeroed, legacy default stream, no status.
extern "C" void fo_compat_bev_pool_v2_grad(int c, int n_intervals, const float *out_grad, const float *depth,
                                           const float *feat, const int *ranks_depth, const int *ranks_feat,
                                           const int *ranks_bev, const int *interval_starts,
                                           const int *interval_lengths, float *depth_grad, float *feat_grad) {
    if (n_intervals <= 0 || c <= 0) return;
    PixelArgs pa;
    pa.G = out_grad; pa.row_of_pos = ranks_bev; pa.depth = depth; pa.feat = feat;
    pa.rd = ranks_depth; pa.rf = ranks_feat; pa.bwd_pos = nullptr;
    pa.bwd_starts = interval_starts; pa.bwd_lengths = interval_lengths;
    pa.n_bwd_dev = nullptr; pa.n_bwd = n_intervals;
    pa.n_rows_G = INT_MAX; pa.n_depth = INT_MAX; pa.n_feat_rows = INT_MAX; pa.n_points = INT_MAX;
    pa.C = c; pa.depth_grad = depth_grad; pa.feat_grad = feat_grad;
    const bool vec = (c % 4 == 0) && (c <= 4 * kGroupLanes * kMaxChunks) && (((uintptr_t)feat & 15) == 0) &&
                     (((uintptr_t)feat_grad & 15) == 0) && (((uintptr_t)out_grad & 15) == 0);
    launch_pixel(pa, vec, (cudaStream_t)0);
}
