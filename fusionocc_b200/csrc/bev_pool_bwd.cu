// fusionocc_b200 — bev_pool_v2 backward (sm_100a).
//
// Reference: mmdet3d/ops/bev_pool_v2/bev_pool.py:44-83 (argsort by ranks_feat, interval rebuild with a
// host sync, 164 MB out_grad.contiguous() un-permute) + src/bev_pool_cuda.cu:67-121 (ONE THREAD per
// image pixel: 4 224 threads at the headline shape).  Here:
//
//   gather   (only when out_grad is (B,C,Z,Y,X)): a CTA per forward tile reads the C x 128 block with
//            128-byte coalesced loads — predicated per lane on the tile's occupancy mask, so only the
//            sectors that contain an occupied voxel are fetched — transposes it through shared memory
//            and emits one compact channels-last row per interval: G[k, 0..C).  No 164 MB copy.
//   pixel    an 8-lane group per backward interval (= image pixel) walks the pixel's points in the
//            inverse interval ordering (backward plan entries: depth index + forward interval), eight
//            points per batch so that each batch costs two memory round trips.  depth_grad[p] is the
//            sequential FMA chain over c = 0..C-1 (handed lane to lane), feat_grad[q, :] the sequential
//            FMA over the pixel's points: the reference's exact orders, no atomics, every output
//            written once.
#include "common.cuh"

namespace fo {

struct GatherArgs {
    const float *og;                // (B,C,Z,Y,X)
    int32_t C;
    int64_t V;
    const FwdPlanHeader *hdr;
    const int32_t *tile_off;
    const int32_t *iv_vox;
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
        int vl = (int)(__ldg(a.iv_vox + k0 + kk) - vbase);
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
            float v[kTile / 32];
#pragma unroll
            for (int r = 0; r < kTile / 32; ++r)
                v[r] = ((m[r] >> lane) & 1u) ? __ldcs(src + lane + 32 * r) : 0.f;
#pragma unroll
            for (int r = 0; r < kTile / 32; ++r)
                if ((m[r] >> lane) & 1u) stage[(lane + 32 * r) * S + c] = v[r];
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
    const int32_t *row_map;         // nullptr: G row = entry's interval id; else G row = row_map[interval id]
    const float *depth, *feat;
    const int32_t *ent_p, *ent_iv;  // backward plan entries
    const int32_t *bwd_starts, *bwd_lengths, *bwd_ids;
    const int32_t *n_bwd_dev;       // live number of backward intervals (nullptr: use n_bwd)
    int64_t n_bwd;
    int64_t n_rows_G;               // bound for G row indices
    int64_t n_iv;                   // bound for interval ids (row_map length)
    int64_t n_depth, n_feat_rows, n_entries;
    int32_t C;
    float *depth_grad, *feat_grad;
};

constexpr int kPixThreads = 256;

// One 8-lane group per backward interval.  NCHUNK >= 1: float4 lanes (C % 4 == 0).
template <int NCHUNK>
__global__ void __launch_bounds__(kPixThreads) bwd_pixel_kernel(PixelArgs a) {
    constexpr int kPixBatch = NCHUNK <= 2 ? 8 : 4;      // points per batch (register budget)
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
        const int q = __ldg(a.bwd_ids + m);
        if (len <= 0 || s < 0 || (int64_t)s + len > a.n_entries || q < 0 || q >= a.n_feat_rows) continue;
        float4 f[NCHUNK], fg[NCHUNK];
#pragma unroll
        for (int ch = 0; ch < NCHUNK; ++ch) {
            const int idx = gl + kGroupLanes * ch;
            f[ch] = (idx < c4) ? ldg4(a.feat + ((int64_t)q * c4 + idx) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            fg[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int j0 = 0; j0 < len; j0 += kPixBatch) {
            // lane u of the group fetches the indices of point j0+u, then they are broadcast
            int my_p = -1, my_row = -1;
            if (gl < kPixBatch && j0 + gl < len) {
                my_p = __ldg(a.ent_p + s + j0 + gl);
                const int iv = __ldg(a.ent_iv + s + j0 + gl);
                my_row = iv;
                if (a.row_map) my_row = (iv >= 0 && iv < a.n_iv) ? __ldg(a.row_map + iv) : -1;
                if (my_p < 0 || my_p >= a.n_depth || my_row < 0 || my_row >= a.n_rows_G) { my_p = -1; my_row = -1; }
            }
            float my_d = (my_p >= 0) ? __ldg(a.depth + my_p) : 0.f;
            float4 g[kPixBatch][NCHUNK];
#pragma unroll
            for (int u = 0; u < kPixBatch; ++u) {
                const int row = __shfl_sync(gmask, my_row, gbase + u);
#pragma unroll
                for (int ch = 0; ch < NCHUNK; ++ch) {
                    const int idx = gl + kGroupLanes * ch;
                    g[u][ch] = (row >= 0 && idx < c4) ? ldg4(a.G + ((int64_t)row * c4 + idx) * 4)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < kPixBatch; ++u) {
                const int p = __shfl_sync(gmask, my_p, gbase + u);
                const float d = __shfl_sync(gmask, my_d, gbase + u);
                if (j0 + u < len) {        // uniform across the group
                    // feat grad: sequential over the pixel's points (bev_pool_cuda.cu:109-120)
#pragma unroll
                    for (int ch = 0; ch < NCHUNK; ++ch) fma4(fg[ch], g[u][ch], d);
                    // depth grad: sequential FMA chain over c = 0..C-1 (bev_pool_cuda.cu:96-101)
                    float sum = 0.f;
#pragma unroll
                    for (int ch = 0; ch < NCHUNK; ++ch) {
#pragma unroll
                        for (int o = 0; o < kGroupLanes; ++o) {
                            const int idx = ch * kGroupLanes + o;
                            if (idx < c4) {
                                if (idx > 0)
                                    sum = __shfl_sync(gmask, sum, gbase + ((o + kGroupLanes - 1) & (kGroupLanes - 1)));
                                if (gl == o) {
                                    sum = fmaf(g[u][ch].x, f[ch].x, sum);
                                    sum = fmaf(g[u][ch].y, f[ch].y, sum);
                                    sum = fmaf(g[u][ch].z, f[ch].z, sum);
                                    sum = fmaf(g[u][ch].w, f[ch].w, sum);
                                }
                            }
                        }
                    }
                    if (gl == last_owner && p >= 0) a.depth_grad[p] = sum;
                }
            }
        }
#pragma unroll
        for (int ch = 0; ch < NCHUNK; ++ch) {
            const int idx = gl + kGroupLanes * ch;
            if (idx < c4) *reinterpret_cast<float4 *>(a.feat_grad + ((int64_t)q * c4 + idx) * 4) = fg[ch];
        }
    }
}

// Scalar path (any C): one warp per backward interval.
__global__ void __launch_bounds__(kPixThreads) bwd_pixel_scalar_kernel(PixelArgs a) {
    const int lane = threadIdx.x & 31;
    const int C = a.C;
    const int64_t n = a.n_bwd_dev ? min((int64_t)max(*a.n_bwd_dev, 0), a.n_bwd) : a.n_bwd;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m = warp0; m < n; m += nwarps) {
        const int s = a.bwd_starts[m], len = a.bwd_lengths[m], q = a.bwd_ids[m];
        if (len <= 0 || s < 0 || (int64_t)s + len > a.n_entries || q < 0 || q >= a.n_feat_rows) continue;
        auto row_of = [&](int j) -> int {
            int iv = a.ent_iv[s + j];
            if (a.row_map) iv = (iv >= 0 && iv < a.n_iv) ? a.row_map[iv] : -1;
            return (iv >= 0 && iv < a.n_rows_G) ? iv : -1;
        };
        // feat grad: lane per channel, sequential over points
        for (int c = lane; c < C; c += 32) {
            float sum = 0.f;
            for (int j = 0; j < len; ++j) {
                const int p = a.ent_p[s + j], row = row_of(j);
                if (p < 0 || p >= a.n_depth || row < 0) continue;
                sum = fmaf(a.G[(int64_t)row * C + c], a.depth[p], sum);
            }
            a.feat_grad[(int64_t)q * C + c] = sum;
        }
        // depth grad: lane per point, sequential over channels
        for (int j = lane; j < len; j += 32) {
            const int p = a.ent_p[s + j], row = row_of(j);
            if (p < 0 || p >= a.n_depth || row < 0) continue;
            float sum = 0.f;
            for (int c = 0; c < C; ++c) sum = fmaf(a.G[(int64_t)row * C + c], a.feat[(int64_t)q * C + c], sum);
            a.depth_grad[p] = sum;
        }
    }
}

// Source-compatible launcher's kernel: arrays already in backward order (bev_pool_cuda.cu:67-121
// contract).  One warp per backward interval, same summation orders.
__global__ void __launch_bounds__(256) compat_grad_kernel(int c, int n_intervals, const float *__restrict__ out_grad,
                                                          const float *__restrict__ depth,
                                                          const float *__restrict__ feat,
                                                          const int *__restrict__ rd, const int *__restrict__ rf,
                                                          const int *__restrict__ rb,
                                                          const int *__restrict__ starts,
                                                          const int *__restrict__ lengths, float *depth_grad,
                                                          float *feat_grad) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m = warp0; m < n_intervals; m += nwarps) {
        const int s = starts[m], len = lengths[m];
        if (len <= 0) continue;
        const int64_t q = rf[s];
        for (int ch = lane; ch < c; ch += 32) {
            float sum = 0.f;
            for (int i = 0; i < len; ++i) sum = fmaf(out_grad[(int64_t)rb[s + i] * c + ch], depth[rd[s + i]], sum);
            feat_grad[q * c + ch] = sum;
        }
        for (int i = lane; i < len; i += 32) {
            const float *g = out_grad + (int64_t)rb[s + i] * c;
            const float *f = feat + (int64_t)rf[s + i] * c;
            float sum = 0.f;
            for (int ch = 0; ch < c; ++ch) sum = fmaf(g[ch], f[ch], sum);
            depth_grad[rd[s + i]] = sum;
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
        const int blocks = grid_for(groups * kGroupLanes, kPixThreads, 16);
        const int chunks = (pa.C / 4 + kGroupLanes - 1) / kGroupLanes;
        switch (chunks) {
            case 1: bwd_pixel_kernel<1><<<blocks, kPixThreads, 0, stream>>>(pa); break;
            case 2: bwd_pixel_kernel<2><<<blocks, kPixThreads, 0, stream>>>(pa); break;
            case 3: bwd_pixel_kernel<3><<<blocks, kPixThreads, 0, stream>>>(pa); break;
            default: bwd_pixel_kernel<4><<<blocks, kPixThreads, 0, stream>>>(pa); break;
        }
    } else {
        bwd_pixel_scalar_kernel<<<grid_for(groups * 32, kPixThreads, 16), kPixThreads, 0, stream>>>(pa);
    }
    FO_LAUNCH_CHECK("bwd_pixel_kernel");
    return FO_OK;
}
}  // namespace

extern "C" int fo_bev_pool_v2_backward(fo_stream_t stream_, int32_t c, const float *out_grad, int32_t og_layout,
                                       const float *depth, const float *feat, int64_t n_points,
                                       int64_t n_intervals, int32_t B, int64_t n_vox, int64_t n_depth,
                                       int64_t n_feat_rows, float *depth_grad, float *feat_grad,
                                       const void *fwd_plan, size_t fwd_plan_bytes, const void *bwd_plan,
                                       size_t bwd_plan_bytes, void *scratch, size_t scratch_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FO_CHECK_ARG(c >= 1 && B >= 1 && n_vox >= 1, "c, B and voxels per sample must be positive");
    FO_CHECK_ARG(og_layout == FO_LAYOUT_BCZYX || og_layout == FO_LAYOUT_BZYXC, "unknown og_layout %d", og_layout);
    FO_CHECK_ARG(depth_grad && feat_grad, "NULL gradient output");
    FO_CHECK_ARG(n_depth >= 0 && n_feat_rows >= 1 && n_points >= 0 && n_intervals >= 0, "negative size");
    FO_CHECK_ARG(n_points < INT_MAX && (int64_t)B * n_vox < INT_MAX && n_depth < INT_MAX, "sizes exceed int32 ranks");
    FO_CUDA(cudaMemsetAsync(depth_grad, 0, (size_t)n_depth * 4, stream));
    FO_CUDA(cudaMemsetAsync(feat_grad, 0, (size_t)n_feat_rows * c * 4, stream));
    if (n_points == 0 || n_intervals == 0) return FO_OK;
    FO_CHECK_ARG(out_grad && depth && feat, "NULL input array");
    FO_CHECK_ARG(bwd_plan != nullptr, "backward plan is required");
    FwdPlanView pv; int64_t n_tiles; int tps;
    if (int rc = open_fwd_plan_const(fwd_plan, fwd_plan_bytes, B, n_vox, n_points, &pv, &n_tiles, &tps)) return rc;
    BwdPlanView bv;
    if (!bwd_plan_view(const_cast<void *>(bwd_plan), n_feat_rows, bwd_plan_bytes, &bv))
        return set_error(FO_ERR_SCRATCH, "backward plan buffer too small (%zu bytes)", bwd_plan_bytes);
    const bool vec = (c % 4 == 0) && (c <= 4 * kGroupLanes * kMaxChunks) && (((uintptr_t)feat & 15) == 0) &&
                     (((uintptr_t)feat_grad & 15) == 0);

    PixelArgs pa;
    pa.depth = depth; pa.feat = feat;
    pa.ent_p = bv.ent_p; pa.ent_iv = bv.ent_iv;
    pa.bwd_starts = bv.starts; pa.bwd_lengths = bv.lengths; pa.bwd_ids = bv.ids;
    pa.n_bwd_dev = &bv.hdr->n_bwd_intervals; pa.n_bwd = n_feat_rows;
    pa.n_depth = n_depth; pa.n_feat_rows = n_feat_rows; pa.n_entries = bv.cap; pa.n_iv = n_intervals;
    pa.C = c; pa.depth_grad = depth_grad; pa.feat_grad = feat_grad;

    if (og_layout == FO_LAYOUT_BCZYX) {
        const size_t need = fo_bwd_scratch_bytes(n_intervals, c, og_layout);
        if (!scratch || scratch_bytes < need)
            return set_error(FO_ERR_SCRATCH, "backward scratch is %zu bytes, need %zu", scratch_bytes, need);
        FO_CHECK_ARG(((uintptr_t)scratch & 15) == 0, "scratch must be 16-byte aligned");
        const size_t smem = (size_t)kTile * (c + 1) * sizeof(float);
        if (smem > 200 * 1024) return set_error(FO_ERR_UNSUPPORTED, "C=%d too large for the gather tile", c);
        GatherArgs ga;
        ga.og = out_grad; ga.C = c; ga.V = n_vox;
        ga.hdr = pv.hdr; ga.tile_off = pv.tile_off; ga.iv_vox = pv.iv_vox; ga.G = (float *)scratch;
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
        pa.G = (const float *)scratch; pa.row_map = nullptr; pa.n_rows_G = n_intervals;
    } else {
        pa.G = out_grad; pa.row_map = pv.iv_vox; pa.n_rows_G = (int64_t)B * n_vox;
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
    compat_grad_kernel<<<grid_for((int64_t)n_intervals * 32, 256, 16), 256, 0, 0>>>(
        c, n_intervals, out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths,
        depth_grad, feat_grad);
}
