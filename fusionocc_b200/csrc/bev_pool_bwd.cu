// fusionocc_b200 — bev_pool_v2 backward (sm_100a).
//
// Reference: mmdet3d/ops/bev_pool_v2/bev_pool.py:44-83 (argsort by ranks_feat, interval rebuild with a
// host sync, 164 MB out_grad.contiguous() un-permute) + src/bev_pool_cuda.cu:67-121 (ONE THREAD per
// image pixel: 4 224 threads at the headline shape).  Here:
//
//   gather   (only when out_grad is (B,C,Z,Y,X)): a warp per 32-voxel sub-tile reads the C x 32 block
//            with 128-bit loads — predicated on the sub-tile's occupancy mask, so only the sectors that
//            contain an occupied voxel are fetched — transposes it through shared memory and emits one
//            compact channels-last row per interval: G[k, 0..C).  No 164 MB copy.
//   pixel    one warp per backward interval (= image pixel) walks the pixel's points in the inverse
//            interval ordering (backward plan entries: depth index + forward interval), 32 points per
//            pass staged through shared memory.  depth_grad[p] is the sequential FMA chain over
//            c = 0..C-1 run by one thread, feat_grad[q, :] the sequential FMA over the pixel's points:
//            the reference's exact orders, no atomics, every output written once.
#include "common.cuh"

namespace fo {

struct GatherArgs {
    const float *og;                // (B,C,Z,Y,X), possibly a channel slice of a wider tensor
    int64_t og_bstride;             // elements between samples
    int32_t C;
    int32_t sps;                    // sub-tiles per sample
    int64_t V;
    const FwdPlanHeader *hdr;
    const int32_t *sub_iv;
    const int32_t *iv_vox;
    float *G;                       // [n_intervals, C]
};

__device__ __forceinline__ void sts_f4(unsigned addr, const float4 &v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds_f4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// One warp per 32-voxel sub-tile (same tiling and grid as the forward).  The C x 32 block of out_grad
// is read with 128-bit loads, four full 128-byte lines per instruction, a 16-byte chunk being fetched
// only if one of its four voxels is occupied; it is staged in the swizzled channel-major layout and one
// compact channels-last row G[k, 0..C) is emitted per interval.
template <int NACC, bool EXACT>
__global__ void __launch_bounds__(kThreads) bwd_gather_kernel(GatherArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = EXACT ? 32 * NACC : a.C;
    const int sps = a.sps;                     // kernel argument: one dependent load less per CTA
    const int b = blockIdx.x;
    const int su = blockIdx.y * kWarpsPerCta + warp;
    if (su >= sps) return;
    const int u = b * sps + su;
    const int ia = __ldg(a.sub_iv + u), ib = __ldg(a.sub_iv + u + 1);
    if (ib <= ia) return;
    const int v0 = su << kSubShift;
    const int nv = (int)min((int64_t)kSub, a.V - v0);
    const int vbase = (int)((int64_t)b * a.V) + v0;
    const int ni = min(ib - ia, kSub);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem + warp * C * kSub);

    int my_v = -1;
    if (lane < ni) {
        my_v = __ldg(a.iv_vox + ia + lane) - vbase;
        if ((unsigned)my_v >= (unsigned)nv) my_v = -1;
    }
    const unsigned occ = __reduce_or_sync(0xffffffffu, my_v >= 0 ? (1u << my_v) : 0u);
    const float *plane0 = a.og + (int64_t)b * a.og_bstride + v0;
    if ((a.V & 3) == 0) {
        const int riq = lane >> 3, chunk = lane & 7;
        if (((occ >> (4 * chunk)) & 0xFu) != 0u) {
            const float *src = plane0 + (int64_t)riq * a.V + 4 * chunk;
            const int64_t step = 4 * a.V;
            // smem position of logical chunk `chunk` in row r is (chunk + r) & 7: two alternating values
            const unsigned s0 = sbase + ((unsigned)riq << 7) + ((unsigned)((chunk + riq) & 7) << 4);
            const unsigned s1 = sbase + ((unsigned)(riq + 4) << 7) + ((unsigned)((chunk + riq + 4) & 7) << 4);
            int r = riq;
            for (; r + 28 < C; r += 32, src += 8 * step) {       // eight 128-bit loads in flight per lane
                float4 x[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) x[t] = __ldcs(reinterpret_cast<const float4 *>(src + t * step));
                const unsigned o = (unsigned)(r - riq) << 7;
#pragma unroll
                for (int t = 0; t < 8; ++t) sts_f4(((t & 1) ? s1 : s0) + o + 1024u * (t >> 1), x[t]);
            }
            for (; r + 12 < C; r += 16, src += 4 * step) {
                const float4 x0 = __ldcs(reinterpret_cast<const float4 *>(src));
                const float4 x1 = __ldcs(reinterpret_cast<const float4 *>(src + step));
                const float4 x2 = __ldcs(reinterpret_cast<const float4 *>(src + 2 * step));
                const float4 x3 = __ldcs(reinterpret_cast<const float4 *>(src + 3 * step));
                const unsigned o = (unsigned)(r - riq) << 7;
                sts_f4(s0 + o, x0); sts_f4(s1 + o, x1); sts_f4(s0 + o + 1024u, x2); sts_f4(s1 + o + 1024u, x3);
            }
            for (; r < C; r += 4, src += step) {
                const float4 x = __ldcs(reinterpret_cast<const float4 *>(src));
                sts_f4(sbase + ((unsigned)r << 7) + ((unsigned)((chunk + r) & 7) << 4), x);
            }
        }
    } else {
        if ((occ >> lane) & 1u)
            for (int c = 0; c < C; ++c)
                sts_f32(sbase + 4u * stage_index(c, lane), __ldcs(plane0 + (int64_t)c * a.V + lane));
    }
    __syncwarp();
    // one compact row per interval: lane = channel, conflict-light swizzled read, 128-byte coalesced store
    float *dst = a.G + (int64_t)ia * C + lane;
    const unsigned lane_row = sbase + ((unsigned)lane << 7);
    const unsigned lane_rot = ((unsigned)lane & 7u) << 4;
    for (int l = 0; l < ni; ++l, dst += C) {
        const unsigned v = (unsigned)__shfl_sync(0xffffffffu, my_v, l) & 31u;   // invalid voxels were masked to -1:
        const bool ok = ((occ >> v) & 1u) != 0u;                                 // they read an unoccupied column
        const unsigned addr = lane_row + (((v << 2) + lane_rot) & 127u);
#pragma unroll
        for (int k = 0; k < NACC; ++k)
            if (ok && (EXACT || lane + 32 * k < C)) dst[32 * k] = lds_f32(addr + 4096u * k);
    }
}

struct PixelArgs {
    const float *G;                 // gathered rows [n_intervals, C]  or  out_grad in (B,Z,Y,X,C)
    int32_t g_rowstride;            // elements between rows of G (C; C_total for a channel slice of (B,Z,Y,X,C_total))
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

#ifndef FO_PIX_THREADS
#define FO_PIX_THREADS 256
#endif
constexpr int kPixThreads = FO_PIX_THREADS;
constexpr int kPixWarps   = kPixThreads / 32;
constexpr int kPixChunk   = 32;     // points staged per pass: one per lane

// One WARP per backward interval (= image pixel), points taken 32 at a time in the plan's order.
//   records lane j fetches entry j (depth index, G row) and its depth value; they are broadcast through
//           shared memory (LDS.128 = two records); the next pass's records are fetched during this pass.
//   rows    lanes = channels: each point's gathered out_grad row is one coalesced 128-byte load, eight
//           rows in flight; the row goes (a) straight into feat_grad[q][c] += row[c] * depth[p] — the
//           sequential FMA over the pixel's points in plan order (bev_pool_cuda.cu:109-120), accumulators
//           in registers across passes — and (b) into the per-warp tile rows[32][C+4].
//   depth   lane j owns point j: depth_grad[p_j] = sum_c rows[j][c] * feat[q][c] as ONE sequential FMA
//           chain over c = 0..C-1 in one thread (bev_pool_cuda.cu:96-101): 2 LDS.128 + 4 FFMA per four
//           channels for 32 points at once (row stride C+4 keeps the LDS.128 conflict-free).
// NACC = ceil(C / 32) feat accumulators per lane; EXACT: C == 32 * NACC.  Vector path: C % 4 == 0.
// All indices are 32-bit (the host checks rows * C < 2^31).
template <int NACC, bool EXACT>
__global__ void __launch_bounds__(kPixThreads) bwd_pixel_kernel(PixelArgs a) {
#ifndef FO_PIX_U
#define FO_PIX_U 16
#endif
    constexpr int U = NACC == 1 ? FO_PIX_U : (NACC == 2 ? 8 : 4);   // gathered rows in flight per lane
    extern __shared__ __align__(16) float psm[];
    __shared__ __align__(16) int2 s_rec[kPixWarps][kPixChunk + 16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = EXACT ? 32 * NACC : a.C;
    const int c4 = C >> 2, S = C + 4;
    float *rows = psm + warp * (kPixChunk * S + C);                          // [32][S]
    float *fs = rows + kPixChunk * S;                                        // feat row of the pixel [C]
    const unsigned rows_s = (unsigned)__cvta_generic_to_shared(rows);
    const unsigned fs_s = (unsigned)__cvta_generic_to_shared(fs);
    const unsigned my_col = rows_s + 4u * lane;                              // column `lane` of row 0
    int2 *rec = s_rec[warp];
    const int n = a.n_bwd_dev ? min(max(*a.n_bwd_dev, 0), (int)a.n_bwd) : (int)a.n_bwd;
    const int cta_warps = blockDim.x >> 5;                                   // <= kPixWarps
    const int warp0 = blockIdx.x * cta_warps + warp;
    const int nwarps = gridDim.x * cta_warps;
    const int n_depth = (int)a.n_depth, n_rows_G = (int)a.n_rows_G, n_iv = (int)a.n_iv;

    for (int m = warp0; m < n; m += nwarps) {
        const int s = __ldg(a.bwd_starts + m), len = __ldg(a.bwd_lengths + m);
        const int q = __ldg(a.bwd_ids + m);
        if (len <= 0 || s < 0 || (int64_t)s + len > a.n_entries || q < 0 || q >= a.n_feat_rows) continue;
        const int32_t *ep = a.ent_p + s, *ei = a.ent_iv + s;
        auto load_rec = [&](int j0, int &p_out) -> int2 {
            int2 r = make_int2(-1, 0);                    // (G row or -1, depth bits)
            p_out = -1;
            if (j0 + lane < len) {
                const int p = __ldg(ep + j0 + lane);
                int row = __ldg(ei + j0 + lane);
                if (a.row_map) row = ((unsigned)row < (unsigned)n_iv) ? __ldg(a.row_map + row) : -1;
                if ((unsigned)p < (unsigned)n_depth && (unsigned)row < (unsigned)n_rows_G) {
                    r.x = row;
                    r.y = __float_as_int(__ldg(a.depth + p));
                    p_out = p;
                }
            }
            return r;
        };
        int my_p;
        int2 mine = load_rec(0, my_p);
        __syncwarp();
        for (int i = lane; i < c4; i += 32) reinterpret_cast<float4 *>(fs)[i] = ldg4(a.feat + (q * c4 + i) * 4);
        float fg[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) fg[k] = 0.f;

        for (int j0 = 0; j0 < len; j0 += kPixChunk) {
            const int np = min(kPixChunk, len - j0);
            const int cur_p = my_p;
            __syncwarp();                               // previous pass finished reading rows / rec
            rec[lane] = mine;
            if (lane < 16) rec[kPixChunk + lane] = make_int2(-1, 0);
            __syncwarp();
            if (j0 + kPixChunk < len) mine = load_rec(j0 + kPixChunk, my_p);
            for (int j = 0; j < np; j += U) {
                float g[U][NACC];
                int2 r[U];
#pragma unroll
                for (int t = 0; t < U; ++t) r[t] = rec[j + t];
#pragma unroll
                for (int t = 0; t < U; ++t) {
                    const int base = max(r[t].x, 0) * a.g_rowstride + lane;      // invalid rows read row 0 and are discarded
#pragma unroll
                    for (int k = 0; k < NACC; ++k) g[t][k] = (EXACT || lane + 32 * k < C) ? __ldg(a.G + base + 32 * k) : 0.f;
                }
                const unsigned dst = my_col + 4u * (j * S);
#pragma unroll
                for (int t = 0; t < U; ++t) {
                    const bool ok = r[t].x >= 0;          // warp-uniform; padding records are invalid
                    const float d = __int_as_float(r[t].y);
#pragma unroll
                    for (int k = 0; k < NACC; ++k) {
                        if (EXACT || lane + 32 * k < C) sts_f32(dst + 4u * (t * S + 32 * k), g[t][k]);
                        const float x = fmaf(g[t][k], d, fg[k]);
                        fg[k] = ok ? x : fg[k];
                    }
                }
            }
            __syncwarp();
            // depth grad: one thread, one sequential chain over all C channels
            if (lane < np && cur_p >= 0) {
                const unsigned ra = rows_s + 4u * (lane * S);
                float sum = 0.f;
#pragma unroll 2
                for (int i = 0; i < c4; ++i) {
                    const float4 g = lds_f4(ra + 16u * i), f = lds_f4(fs_s + 16u * i);
                    sum = fmaf(g.x, f.x, sum);
                    sum = fmaf(g.y, f.y, sum);
                    sum = fmaf(g.z, f.z, sum);
                    sum = fmaf(g.w, f.w, sum);
                }
                a.depth_grad[cur_p] = sum;
            }
        }
        float *fgp = a.feat_grad + q * C + lane;
#pragma unroll
        for (int k = 0; k < NACC; ++k)
            if (EXACT || lane + 32 * k < C) fgp[32 * k] = fg[k];
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
                sum = fmaf(a.G[(int64_t)row * a.g_rowstride + c], a.depth[p], sum);
            }
            a.feat_grad[(int64_t)q * C + c] = sum;
        }
        // depth grad: lane per point, sequential over channels
        for (int j = lane; j < len; j += 32) {
            const int p = a.ent_p[s + j], row = row_of(j);
            if (p < 0 || p >= a.n_depth || row < 0) continue;
            float sum = 0.f;
            for (int c = 0; c < C; ++c) sum = fmaf(a.G[(int64_t)row * a.g_rowstride + c], a.feat[(int64_t)q * C + c], sum);
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
    const int64_t pixels = pa.n_bwd;
    if (pixels <= 0) return FO_OK;
    const int C = pa.C;
#ifndef FO_PIX_WARPS_BIGC
#define FO_PIX_WARPS_BIGC 8
#endif
    const int nacc = (C + 31) / 32;
    // the per-warp tile grows with C: smaller CTAs keep more warps resident for wide channel counts
    const int cta_warps = nacc >= 3 ? FO_PIX_WARPS_BIGC : kPixWarps;
    const size_t smem = (size_t)cta_warps * (kPixChunk * (C + 4) + C) * sizeof(float);
    const bool idx32 = pa.n_rows_G * pa.g_rowstride < INT_MAX && pa.n_feat_rows * C < INT_MAX && pixels < INT_MAX;
    if (vec && idx32 && nacc <= 4 && smem <= 200 * 1024) {
#ifndef FO_PIX_CTAS_PER_SM
#define FO_PIX_CTAS_PER_SM 8
#endif
        const int blocks = grid_for(pixels, cta_warps, FO_PIX_CTAS_PER_SM * kPixWarps / cta_warps);
#define FO_PIX(NA, EX)                                                                                          \
    do {                                                                                                        \
        if (smem > 48 * 1024)                                                                                   \
            FO_CUDA(cudaFuncSetAttribute(bwd_pixel_kernel<NA, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                           \
        bwd_pixel_kernel<NA, EX><<<blocks, 32 * cta_warps, smem, stream>>>(pa);                                 \
    } while (0)
        if (C % 32 == 0) {
            switch (nacc) {
                case 1: FO_PIX(1, true); break;
                case 2: FO_PIX(2, true); break;
                case 3: FO_PIX(3, true); break;
                default: FO_PIX(4, true); break;
            }
        } else {
            switch (nacc) {
                case 1: FO_PIX(1, false); break;
                case 2: FO_PIX(2, false); break;
                case 3: FO_PIX(3, false); break;
                default: FO_PIX(4, false); break;
            }
        }
#undef FO_PIX
    } else {
        bwd_pixel_scalar_kernel<<<grid_for(pixels * 32, kPixThreads, 16), kPixThreads, 0, stream>>>(pa);
    }
    FO_LAUNCH_CHECK("bwd_pixel_kernel");
    return FO_OK;
}
}  // namespace

namespace {
int backward_impl(cudaStream_t stream, int32_t c, const float *out_grad, int32_t og_layout, int32_t c_total,
                  int32_t c_offset, const float *depth, const float *feat, int64_t n_points, int64_t n_intervals,
                  int32_t B, int64_t n_vox, int64_t n_depth, int64_t n_feat_rows, float *depth_grad,
                  float *feat_grad, const void *fwd_plan, size_t fwd_plan_bytes, const void *bwd_plan,
                  size_t bwd_plan_bytes, void *scratch, size_t scratch_bytes) {
    FO_CHECK_ARG(c >= 1 && c_offset >= 0 && c_total >= c && c_offset + c <= c_total,
                 "channel slice [%d, %d) does not fit %d channels", c_offset, c_offset + c, c_total);
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
    FwdPlanView pv; int64_t n_subs; int sps;
    if (int rc = open_fwd_plan_const(fwd_plan, fwd_plan_bytes, B, n_vox, n_points, &pv, &n_subs, &sps)) return rc;
    BwdPlanView bv;
    if (!bwd_plan_view(const_cast<void *>(bwd_plan), n_feat_rows, bwd_plan_bytes, &bv))
        return set_error(FO_ERR_SCRATCH, "backward plan buffer too small (%zu bytes)", bwd_plan_bytes);
    const bool vec = (c % 4 == 0) && (((uintptr_t)feat & 15) == 0);

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
        const size_t smem = (size_t)kWarpsPerCta * kSub * c * sizeof(float);
        if (smem > 200 * 1024) return set_error(FO_ERR_UNSUPPORTED, "C=%d too large for the gather tile", c);
        GatherArgs ga;
        ga.og = out_grad + (int64_t)c_offset * n_vox; ga.og_bstride = (int64_t)c_total * n_vox; ga.C = c; ga.V = n_vox;
        ga.sps = sps;
        ga.hdr = pv.hdr; ga.sub_iv = pv.sub_iv; ga.iv_vox = pv.iv_vox; ga.G = (float *)scratch;
        const int n_ctas = (sps + kWarpsPerCta - 1) / kWarpsPerCta;
        if (n_ctas > 65535 || B > 65535 || c > 256)
            return set_error(FO_ERR_UNSUPPORTED, "grid or channel count too large for the gather kernel");
#define FO_GATHER(NA, EX)                                                                                     \
    do {                                                                                                      \
        if (smem > 48 * 1024)                                                                                 \
            FO_CUDA(cudaFuncSetAttribute(bwd_gather_kernel<NA, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                         \
        bwd_gather_kernel<NA, EX><<<dim3(B, n_ctas), kThreads, smem, stream>>>(ga);                           \
    } while (0)
        const int nacc = (c + 31) / 32;
        if (c % 32 == 0 && nacc <= 4) {
            switch (nacc) {
                case 1: FO_GATHER(1, true); break;
                case 2: FO_GATHER(2, true); break;
                case 3: FO_GATHER(3, true); break;
                default: FO_GATHER(4, true); break;
            }
        } else {
            switch (nacc) {
                case 1: FO_GATHER(1, false); break;
                case 2: FO_GATHER(2, false); break;
                case 3: FO_GATHER(3, false); break;
                case 4: FO_GATHER(4, false); break;
                case 5: FO_GATHER(5, false); break;
                case 6: FO_GATHER(6, false); break;
                case 7: FO_GATHER(7, false); break;
                default: FO_GATHER(8, false); break;
            }
        }
#undef FO_GATHER
        FO_LAUNCH_CHECK("bwd_gather_kernel");
        pa.G = (const float *)scratch; pa.row_map = nullptr; pa.n_rows_G = n_intervals; pa.g_rowstride = c;
    } else {
        pa.G = out_grad + c_offset; pa.row_map = pv.iv_vox; pa.n_rows_G = (int64_t)B * n_vox; pa.g_rowstride = c_total;
    }
    const bool pvec = vec && (((uintptr_t)pa.G & 15) == 0) && (pa.g_rowstride % 4 == 0);
    return launch_pixel(pa, pvec, stream);
}
}  // namespace

extern "C" int fo_bev_pool_v2_backward(fo_stream_t stream_, int32_t c, const float *out_grad, int32_t og_layout,
                                       const float *depth, const float *feat, int64_t n_points,
                                       int64_t n_intervals, int32_t B, int64_t n_vox, int64_t n_depth,
                                       int64_t n_feat_rows, float *depth_grad, float *feat_grad,
                                       const void *fwd_plan, size_t fwd_plan_bytes, const void *bwd_plan,
                                       size_t bwd_plan_bytes, void *scratch, size_t scratch_bytes) {
    return backward_impl((cudaStream_t)stream_, c, out_grad, og_layout, c, 0, depth, feat, n_points, n_intervals, B,
                         n_vox, n_depth, n_feat_rows, depth_grad, feat_grad, fwd_plan, fwd_plan_bytes, bwd_plan,
                         bwd_plan_bytes, scratch, scratch_bytes);
}

extern "C" int fo_bev_pool_v2_backward_slice(fo_stream_t stream_, int32_t c, const float *out_grad,
                                             int32_t og_layout, int32_t c_total, int32_t c_offset,
                                             const float *depth, const float *feat, int64_t n_points,
                                             int64_t n_intervals, int32_t B, int64_t n_vox, int64_t n_depth,
                                             int64_t n_feat_rows, float *depth_grad, float *feat_grad,
                                             const void *fwd_plan, size_t fwd_plan_bytes, const void *bwd_plan,
                                             size_t bwd_plan_bytes, void *scratch, size_t scratch_bytes) {
    return backward_impl((cudaStream_t)stream_, c, out_grad, og_layout, c_total, c_offset, depth, feat, n_points,
                         n_intervals, B, n_vox, n_depth, n_feat_rows, depth_grad, feat_grad, fwd_plan, fwd_plan_bytes,
                         bwd_plan, bwd_plan_bytes, scratch, scratch_bytes);
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
