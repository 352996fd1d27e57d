// fusionocc_b200 — rank precompute and plan construction (sm_100a).
//
// Replaces projects/FusionOcc/fusionocc/necks/view_transformer.py:223-281 (voxel_pooling_prepare_v2)
// and the backward re-sort of mmdet3d/ops/bev_pool_v2/bev_pool.py:47-57.  See bucket_sort.cuh for
// the sort itself.  Every kernel here is HBM/L2-bound integer work; nothing is reshaped into GEMMs.
#include <stdlib.h>

#include "bucket_sort.cuh"
#include "bwd_plan.cuh"
#include "rank_fast.cuh"

namespace fo {

// ------------------------------------------------------------------------------------------------
// K1 (forward flavour): voxelise + count.
//   idx = trunc_toward_zero((coor - lb) / itv) per axis — IEEE fp32 subtract, IEEE fp32 divide,
//   cvt.rzi.s64.f32 (what torch's .long() does on device), view_transformer.py:246-248; kept iff inside
//   the grid (:254-256); key = ((b*Z + z)*Y + y)*X + x in exact integers (:262-265).
// Four points (three float4) per thread so that global loads are 128-bit and fully used.
// ------------------------------------------------------------------------------------------------
struct VoxArgs {
    const float *coor;
    int64_t n_points;            // B*N*D*H*W
    int64_t points_per_sample;   // N*D*H*W
    float lbx, lby, lbz, ivx, ivy, ivz;
    int32_t X, Y, Z;
    int32_t *cnt;                // [B*Z*Y*X] zero-initialised
    int32_t *key;                // [n_points] voxel id or -1 (lives in the plan: pt2vox)
    int32_t *slot;               // [n_points]
    FwdPlanHeader *hdr;          // static fields initialised by thread 0
    int32_t n_subs, subs_per_sample;
};

__device__ __forceinline__ int voxel_key(float x, float y, float z, int64_t b, const VoxArgs &a) {
    const long long ix = (long long)__fdiv_rn(__fsub_rn(x, a.lbx), a.ivx);
    const long long iy = (long long)__fdiv_rn(__fsub_rn(y, a.lby), a.ivy);
    const long long iz = (long long)__fdiv_rn(__fsub_rn(z, a.lbz), a.ivz);
    const bool kept = ix >= 0 && ix < a.X && iy >= 0 && iy < a.Y && iz >= 0 && iz < a.Z;
    return kept ? (int)(((b * a.Z + iz) * a.Y + iy) * a.X + ix) : -1;
}

// Fused geometry: the frustum point is computed from the calibration instead of being
// read from a (B,N,D,H,W,3) tensor — view_transformer.py:161-172 per point:
//     p = frustum - post_trans;  p = inv(post_rots) p;  p = (x z, y z, z);  p = (R_s2e inv(K)) p + t_s2e;  p = bda p
// The four small matrices per camera are inputs (computed by the caller with the reference's own torch ops, so
// their bits are the reference's); the per-point arithmetic is three 3x3 matrix-vector products whose fp32
// summation order is chosen by `mode` (the reference runs them through a batched library GEMM):
//   0: k-ascending FMA chain  fma(m2,z, fma(m1,y, m0*x))     1: separate multiplies and adds, k-ascending
//   2: k-descending FMA chain fma(m0,x, fma(m1,y, m2*z))     3: fma(m1,y, m0*x) + m2*z
// Mode 3 is what torch 2.11 / cuBLAS 12.8 executes for these broadcast 3x3 @ 3x1 products on a B200 (measured
// by profiles/matmul_order_probe.py against an fp64 emulation of eleven candidate orders: 0 of 2 230 272
// floats differ for mode 3, 15-42 % for every other order).
struct CalibArgs {
    const float *frustum;        // [D*H*W, 3]  (x_px, y_px, depth)                 view_transformer.py:105-133
    const float *cam;            // [B*N, 24]   inv(post_rots) 9 | post_trans 3 | combine 9 | t_s2e 3
    const float *bda;            // [B, 12]     bda 3x3 | translation 3 (zeros for a 3x3 bda)
    int32_t n_cams, dhw, mode, bda_has_t;
    float *coor_out;             // optional [n_points, 3]
};

__device__ __forceinline__ float dot3(int mode, float m0, float m1, float m2, float x, float y, float z) {
    if (mode == 0) return fmaf(m2, z, fmaf(m1, y, __fmul_rn(m0, x)));
    if (mode == 1) return __fadd_rn(__fadd_rn(__fmul_rn(m0, x), __fmul_rn(m1, y)), __fmul_rn(m2, z));
    if (mode == 2) return fmaf(m0, x, fmaf(m1, y, __fmul_rn(m2, z)));
    return __fadd_rn(fmaf(m1, y, __fmul_rn(m0, x)), __fmul_rn(m2, z));
}

__device__ __forceinline__ void calib_point(const CalibArgs &g, int p, float &ox, float &oy, float &oz) {
    const int bn = p / g.dhw, r = p - bn * g.dhw;
    const float *f = g.frustum + 3 * r;
    const float *m = g.cam + 24 * bn;
    const float *bd = g.bda + 12 * (bn / g.n_cams);
    const float x0 = __fsub_rn(__ldg(f), __ldg(m + 9)), y0 = __fsub_rn(__ldg(f + 1), __ldg(m + 10)),
                z0 = __fsub_rn(__ldg(f + 2), __ldg(m + 11));
    const float x1 = dot3(g.mode, __ldg(m), __ldg(m + 1), __ldg(m + 2), x0, y0, z0);
    const float y1 = dot3(g.mode, __ldg(m + 3), __ldg(m + 4), __ldg(m + 5), x0, y0, z0);
    const float z1 = dot3(g.mode, __ldg(m + 6), __ldg(m + 7), __ldg(m + 8), x0, y0, z0);
    const float x2 = __fmul_rn(x1, z1), y2 = __fmul_rn(y1, z1);
    const float x3 = __fadd_rn(dot3(g.mode, __ldg(m + 12), __ldg(m + 13), __ldg(m + 14), x2, y2, z1), __ldg(m + 21));
    const float y3 = __fadd_rn(dot3(g.mode, __ldg(m + 15), __ldg(m + 16), __ldg(m + 17), x2, y2, z1), __ldg(m + 22));
    const float z3 = __fadd_rn(dot3(g.mode, __ldg(m + 18), __ldg(m + 19), __ldg(m + 20), x2, y2, z1), __ldg(m + 23));
    ox = dot3(g.mode, __ldg(bd), __ldg(bd + 1), __ldg(bd + 2), x3, y3, z3);
    oy = dot3(g.mode, __ldg(bd + 3), __ldg(bd + 4), __ldg(bd + 5), x3, y3, z3);
    oz = dot3(g.mode, __ldg(bd + 6), __ldg(bd + 7), __ldg(bd + 8), x3, y3, z3);
    if (g.bda_has_t) {
        ox = __fadd_rn(ox, __ldg(bd + 9)); oy = __fadd_rn(oy, __ldg(bd + 10)); oz = __fadd_rn(oz, __ldg(bd + 11));
    }
}

template <bool CALIB>
__global__ void __launch_bounds__(256) voxelize_count_kernel(VoxArgs a, CalibArgs g) {
    pdl_wait();
    pdl_launch();
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid == 0) {
        a.hdr->flags = 0;
        a.hdr->n_subs = a.n_subs;
        a.hdr->subs_per_sample = a.subs_per_sample;
        a.hdr->structured = 1;
        a.hdr->fwd_heavy[0] = a.hdr->fwd_heavy[1] = a.hdr->fwd_heavy[2] = 0;
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_quads = a.n_points >> 2;
    for (int64_t qd = gtid; qd < n_quads; qd += stride) {
        float xs[4], ys[4], zs[4];
        // n_points < 2^31 is enforced by the host wrapper: 32-bit index arithmetic
        const int p0 = (int)(qd << 2), pps = (int)a.points_per_sample;
        if (CALIB) {
#pragma unroll
            for (int j = 0; j < 4; ++j) calib_point(g, p0 + j, xs[j], ys[j], zs[j]);
            if (g.coor_out) {
                float4 *dst = reinterpret_cast<float4 *>(g.coor_out) + qd * 3;
                dst[0] = make_float4(xs[0], ys[0], zs[0], xs[1]);
                dst[1] = make_float4(ys[1], zs[1], xs[2], ys[2]);
                dst[2] = make_float4(zs[2], xs[3], ys[3], zs[3]);
            }
        } else {
            const float4 *src = reinterpret_cast<const float4 *>(a.coor) + qd * 3;
            const float4 v0 = __ldcs(src), v1 = __ldcs(src + 1), v2 = __ldcs(src + 2);
            xs[0] = v0.x; xs[1] = v0.w; xs[2] = v1.z; xs[3] = v2.y;
            ys[0] = v0.y; ys[1] = v1.x; ys[2] = v1.w; ys[3] = v2.z;
            zs[0] = v0.z; zs[1] = v1.y; zs[2] = v2.x; zs[3] = v2.w;
        }
        int keys[4], slots[4];
        const int b0 = p0 / pps, rem = p0 - b0 * pps;
#pragma unroll
        for (int j = 0; j < 4; ++j) keys[j] = voxel_key(xs[j], ys[j], zs[j], b0 + (rem + j) / pps, a);
#pragma unroll
        for (int j = 0; j < 4; ++j) slots[j] = keys[j] >= 0 ? atomicAdd(a.cnt + keys[j], 1) : 0;
        reinterpret_cast<int4 *>(a.key)[qd] = make_int4(keys[0], keys[1], keys[2], keys[3]);
        reinterpret_cast<int4 *>(a.slot)[qd] = make_int4(slots[0], slots[1], slots[2], slots[3]);
    }
    // tail (n_points % 4)
    for (int64_t p = (n_quads << 2) + gtid; p < a.n_points; p += stride) {
        float x, y, z;
        if (CALIB) {
            calib_point(g, (int)p, x, y, z);
            if (g.coor_out) { g.coor_out[3 * p] = x; g.coor_out[3 * p + 1] = y; g.coor_out[3 * p + 2] = z; }
        } else {
            x = a.coor[3 * p]; y = a.coor[3 * p + 1]; z = a.coor[3 * p + 2];
        }
        const int k = voxel_key(x, y, z, p / a.points_per_sample, a);
        a.key[p] = k;
        a.slot[p] = k >= 0 ? atomicAdd(a.cnt + k, 1) : 0;
    }
}

}  // namespace fo
#include "rank_chunk.cuh"
namespace fo {

// K1 (backward flavour): keys are given (ranks_feat of each forward position).
__global__ void __launch_bounds__(256) count_keys_kernel(const int32_t *__restrict__ keys, int64_t n_cap,
                                                         const int32_t *__restrict__ n_dev, int64_t n_buckets,
                                                         int32_t *cnt, int32_t *slot, BwdPlanHeader *hdr) {
    const int64_t n = n_dev ? min((int64_t)max(*n_dev, 0), n_cap) : n_cap;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid == 0 && hdr) { hdr->n_bwd_intervals = 0; hdr->n_points = (int)n; }
    for (int64_t i = gtid; i < n; i += stride) {
        const int k = keys[i];
        slot[i] = (k >= 0 && k < n_buckets) ? atomicAdd(cnt + k, 1) : -1;
    }
}

// K3: placement.  sorted[offset[key] + slot] = original index.
__global__ void __launch_bounds__(256) place_kernel(const int32_t *__restrict__ key, const int32_t *__restrict__ slot,
                                                    const int32_t *__restrict__ offs, int64_t n_cap,
                                                    const int32_t *__restrict__ n_dev, int64_t n_buckets,
                                                    int32_t *__restrict__ sorted) {
    pdl_wait();
    pdl_launch();
    const int64_t n = n_dev ? min((int64_t)max(*n_dev, 0), n_cap) : n_cap;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int k = key[i];
        const int s = slot[i];
        if (k >= 0 && k < n_buckets && s >= 0) sorted[offs[k] + s] = (int)i;
    }
}

// ------------------------------------------------------------------------------------------------
// Forward plan from caller-supplied interval arrays (fo_fwd_plan_build).
// ------------------------------------------------------------------------------------------------
__global__ void init_fwd_header_kernel(FwdPlanHeader *hdr, int n_subs, int sps, int n_intervals) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        hdr->flags = 0;
        hdr->n_subs = n_subs;
        hdr->subs_per_sample = sps;
        hdr->n_intervals = n_intervals;
        hdr->structured = 0;
        hdr->fwd_heavy[0] = hdr->fwd_heavy[1] = hdr->fwd_heavy[2] = 0;
    }
}

__device__ __forceinline__ int interval_voxel(const int32_t *rb, const int32_t *starts, const int32_t *lengths,
                                              int k, int64_t n_points, int64_t n_vox_total) {
    const int s = starts[k], len = lengths[k];
    if (s < 0 || len < 0 || (int64_t)s + len > n_points || s >= n_points) return -1;
    const int v = rb[s];
    return (v >= 0 && v < n_vox_total) ? v : -1;
}

__global__ void __launch_bounds__(256) plan_from_intervals_kernel(
    const int32_t *__restrict__ rb, const int32_t *__restrict__ starts, const int32_t *__restrict__ lengths,
    int64_t n_points, int64_t n_cap, const int32_t *__restrict__ n_dev, int64_t vox_per_sample,
    int64_t n_vox_total, int sps, int n_subs, FwdPlanHeader *hdr, int32_t *sub_iv, int32_t *sub_pt,
    int32_t *pos2iv, int32_t *iv_vox) {
    const int64_t n = n_dev ? min((int64_t)max(*n_dev, 0), n_cap) : n_cap;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid == 0) hdr->n_intervals = (int)n;
    if (n == 0) {
        for (int64_t t = gtid; t <= n_subs; t += stride) { sub_iv[t] = 0; sub_pt[t] = 0; }
        return;
    }
    for (int64_t k = gtid; k < n; k += stride) {
        const int v = interval_voxel(rb, starts, lengths, (int)k, n_points, n_vox_total);
        const int vp = k > 0 ? interval_voxel(rb, starts, lengths, (int)k - 1, n_points, n_vox_total) : -1;
        iv_vox[k] = v;
        if (v < 0) {
            atomicOr(&hdr->flags, kFlagOutOfRange | kFlagUnsorted);
            continue;
        }
        const int s = starts[k], len = lengths[k];
        for (int j = 0; j < len; ++j) pos2iv[s + j] = (int)k;
        // fast path needs: strictly increasing voxels, intervals back to back, every point of an interval
        // carrying the interval's voxel (then runs of equal ranks_bev == intervals)
        bool clean = len >= 1 && !(k > 0 && (vp < 0 || v <= vp || starts[k - 1] + lengths[k - 1] != s));
        for (int j = 1; j < len && clean; ++j) clean = rb[s + j] == v;
        if (!clean) {
            atomicOr(&hdr->flags, kFlagUnsorted);
            continue;
        }
        const int64_t u = (v / vox_per_sample) * sps + ((v % vox_per_sample) >> kSubShift);
        const int64_t up = (k > 0) ? ((vp / vox_per_sample) * sps + ((vp % vox_per_sample) >> kSubShift)) : -1;
        for (int64_t t = up + 1; t <= u; ++t) { sub_iv[t] = (int)k; sub_pt[t] = s; }
        if (k == n - 1)
            for (int64_t t = u + 1; t <= n_subs; ++t) { sub_iv[t] = (int)n; sub_pt[t] = s + len; }
    }
}

// lists the dense sub-tiles of a plan built from caller-supplied intervals (the rank precompute does it inside
// its order pass)
__global__ void __launch_bounds__(256) heavy_queue_kernel(const int32_t *__restrict__ sub_pt, int n_subs,
                                                          FwdPlanHeader *hdr, int32_t *heavy_list, int heavy_pts) {
    if (hdr->flags & kFlagUnsorted) return;               // sub_pt is not meaningful then
    const int stride = gridDim.x * blockDim.x;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n_subs; u += stride)
        if (sub_pt[u + 1] - sub_pt[u] > heavy_pts) heavy_list[atomicAdd(hdr->fwd_heavy, 1)] = u;
}

// ------------------------------------------------------------------------------------------------
// Backward plan, structured build (plans produced by fo_rank_prepare).
// One warp per image pixel q = (b*N+n)*HW + hw: its <= D candidate points are p = ((b*N+n)*D + d)*HW + hw;
// pt2vox gives each one's voxel (or -1).  The forward order is (voxel id, point index) — a stable sort by
// voxel — so the pixel's entries in ascending forward position (the order of bev_pool.py:47-49) are its
// points sorted by (voxel id, depth bin d): no point -> position map is needed, and the gathered out_grad
// row of a point is vox2iv[voxel].  No sort passes over memory, no atomics, fixed-stride rows of D entries.
// ------------------------------------------------------------------------------------------------
template <int R>   // R = ceil(D / 32) registers per lane; rank-by-counting on 64-bit (voxel, d) keys, D <= 256
__global__ void __launch_bounds__(256) bwd_plan_structured_kernel(const FwdPlanHeader *fhdr,
                                                                  const int32_t *__restrict__ pt2vox,
                                                                  const int32_t *__restrict__ vox2iv_, int D, int HW,
                                                                  int n_rows, BwdPlanHeader *hdr, int32_t *ent_p,
                                                                  int32_t *ent_iv, int32_t *starts, int32_t *lengths,
                                                                  int32_t *ids, const int32_t *n_points_dev) {
    const int lane = threadIdx.x & 31;
    const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int32_t *vox2iv = plan_vox2iv(fhdr, vox2iv_);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        hdr->n_bwd_intervals = n_rows;
        hdr->n_points = n_points_dev ? *n_points_dev : 0;
        hdr->structured = 1;
    }
    for (int q = warp0; q < n_rows; q += nwarps) {
        const int bn = q / HW, hw = q - bn * HW;
        long long key[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int d = lane + 32 * r;
            key[r] = LLONG_MAX;
            if (d < D) {
                const int v = __ldg(pt2vox + ((int64_t)bn * D + d) * HW + hw);
                if (v >= 0) key[r] = ((long long)v << 8) | d;
            }
        }
        int rank[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rank[r] = 0;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
#pragma unroll 8
            for (int l = 0; l < 32; ++l) {
                const long long other = __shfl_sync(0xffffffffu, key[rr], l);
#pragma unroll
                for (int r = 0; r < R; ++r) rank[r] += (other < key[r]) ? 1 : 0;
            }
        }
        int cnt = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (key[r] != LLONG_MAX) {
                const int64_t e = (int64_t)q * D + rank[r];
                ent_p[e] = (bn * D + lane + 32 * r) * HW + hw;
                ent_iv[e] = vox2iv ? __ldg(vox2iv + (int)(key[r] >> 8)) : (int)(key[r] >> 8);
                ++cnt;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) { starts[q] = q * D; lengths[q] = cnt; ids[q] = q; }
    }
}

// Same job with a bitonic network over packed 32-bit keys (bwd_plan.cuh: plan_pixel_bitonic).
template <int R>   // R in {1, 2, 4}: 32 * R >= D
__global__ void __launch_bounds__(256) bwd_plan_structured_bitonic_kernel(
    const FwdPlanHeader *fhdr, const int32_t *__restrict__ pt2vox, const int32_t *__restrict__ vox2iv_, int D, int HW, int n_rows,
    BwdPlanHeader *hdr, int32_t *ent_p, int32_t *ent_iv, int32_t *starts, int32_t *lengths, int32_t *ids,
    const int32_t *n_points_dev) {
    __shared__ int s_cmp[8][32 * R];
    pdl_wait();
    pdl_launch();
    const int lane = threadIdx.x & 31;
    int *cmp = s_cmp[threadIdx.x >> 5];
    const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int32_t *vox2iv = plan_vox2iv(fhdr, vox2iv_);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        hdr->n_bwd_intervals = n_rows;
        hdr->n_points = n_points_dev ? *n_points_dev : 0;
        hdr->structured = 1;
    }
    for (int q = warp0; q < n_rows; q += nwarps)
        plan_pixel_bitonic<R>(pt2vox, vox2iv, D, HW, q, cmp, lane, ent_p, ent_iv, starts, lengths, ids);
}

// Generic build, last step: sorted forward positions -> (depth index, forward interval) entries.
__global__ void __launch_bounds__(256) bwd_plan_fill_entries_kernel(const int32_t *__restrict__ pos,
                                                                    const int32_t *__restrict__ rd,
                                                                    const int32_t *__restrict__ pos2iv,
                                                                    const BwdPlanHeader *hdr, int64_t cap,
                                                                    int32_t *ent_p, int32_t *ent_iv) {
    const int64_t n = min((int64_t)max(hdr->totals[0], 0), cap);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int i = pos[j];
        ent_p[j] = rd[i];
        ent_iv[j] = pos2iv[i];
    }
}

}  // namespace fo

using namespace fo;

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" size_t fo_fwd_plan_bytes(int64_t n_voxels_total, int64_t n_points_capacity) {
    if (n_voxels_total < 1 || n_points_capacity < 0) return 0;
    return fwd_plan_bytes_for(n_voxels_total, n_points_capacity);
}

namespace {
// shared checks for entry points that take a forward plan
int open_fwd_plan(void *plan, size_t plan_bytes, int32_t B, int64_t n_vox, int64_t n_points, FwdPlanView *pv,
                  int64_t *n_subs, int *sps) {
    FO_CHECK_ARG(plan != nullptr, "forward plan is NULL");
    FO_CHECK_ARG(B >= 1 && n_vox >= 1, "B=%d and n_voxels_per_sample=%lld must be positive", B, (long long)n_vox);
    FO_CHECK_ARG((int64_t)B * n_vox < INT_MAX, "B*Z*Y*X = %lld does not fit int32 ranks", (long long)B * n_vox);
    FO_CHECK_ARG(B <= 4096, "B=%d exceeds the plan's sample bound (4096)", B);
    FO_CHECK_ARG(((uintptr_t)plan & 255) == 0, "forward plan must be 256-byte aligned");
    if (!fwd_plan_view(plan, (int64_t)B * n_vox, plan_bytes, pv) || pv->p_cap < n_points)
        return set_error(FO_ERR_SCRATCH, "forward plan buffer is %zu bytes, need %zu for %lld points", plan_bytes,
                         fwd_plan_bytes_for((int64_t)B * n_vox, n_points), (long long)n_points);
    *sps = (int)subs_per_sample(n_vox);
    *n_subs = (int64_t)(*sps) * B;
    return FO_OK;
}

struct SortScratch {
    int32_t *cnt;
    unsigned long long *agg;     // per-scan-tile (points << 32 | non-empty buckets)
    unsigned long long *agg_group;   // the same summed over groups of kScanGroup tiles (zero-initialised)
    int32_t *counter;            // long-interval queue length
    size_t zero_bytes;
};
SortScratch sort_scratch_view(void *base, int64_t n_buckets) {
    const int64_t n_scan_tiles = (n_buckets + kScanTile - 1) / kScanTile;
    SortScratch s;
    char *p = (char *)base;
    s.cnt = (int32_t *)p;                       p += align_up(n_buckets * 4, 256);
    s.agg = (unsigned long long *)p;            p += align_up(n_scan_tiles * 8, 256);
    s.agg_group = (unsigned long long *)p;      p += align_up((n_scan_tiles / kScanGroup + 1) * 8, 256);
    s.counter = (int32_t *)p;                   p += 256;
    s.zero_bytes = (size_t)(p - (char *)base);
    return s;
}
static_assert(kScanTile >= 2048 && kScanItems % 4 == 0, "bucket_zero_bytes() in common.cuh sizes the aggregates for 2048-bucket scan tiles");
}  // namespace

// exported for the other translation units
namespace fo {
int open_fwd_plan_const(const void *plan, size_t plan_bytes, int32_t B, int64_t n_vox, int64_t n_points,
                        FwdPlanView *pv, int64_t *n_subs, int *sps) {
    return open_fwd_plan(const_cast<void *>(plan), plan_bytes, B, n_vox, n_points, pv, n_subs, sps);
}
}  // namespace fo

extern "C" int fo_fwd_plan_build(fo_stream_t stream_, const int32_t *ranks_bev, const int32_t *interval_starts,
                                 const int32_t *interval_lengths, int64_t n_points, int64_t n_intervals,
                                 const int32_t *n_intervals_dev, int32_t B, int64_t n_vox, void *plan,
                                 size_t plan_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FO_CHECK_ARG(n_points >= 0 && n_intervals >= 0 && n_points < INT_MAX, "negative or oversized counts");
    FO_CHECK_ARG(n_intervals == 0 || (ranks_bev && interval_starts && interval_lengths), "NULL index array");
    FwdPlanView pv; int64_t n_subs; int sps;
    if (int rc = open_fwd_plan(plan, plan_bytes, B, n_vox, n_points, &pv, &n_subs, &sps)) return rc;
    FO_CHECK_ARG(n_intervals <= pv.iv_cap, "n_intervals=%lld exceeds the plan's interval capacity %lld",
                 (long long)n_intervals, (long long)pv.iv_cap);
    // a non-canonical interval list leaves (parts of) these tables unwritten: give them defined contents, so that
    // nothing downstream can index with uninitialised values (the backward reads pos2iv for every point)
    FO_CUDA(cudaMemsetAsync(pv.sub_iv, 0, (size_t)(n_subs + 1) * 4, stream));
    FO_CUDA(cudaMemsetAsync(pv.sub_pt, 0, (size_t)(n_subs + 1) * 4, stream));
    FO_CUDA(cudaMemsetAsync(pv.sub_mask, 0, ((size_t)n_subs + 1) * 4, stream));    // occupancy masks unknown
    if (n_points > 0) FO_CUDA(cudaMemsetAsync(pv.pos2iv, 0xFF, (size_t)n_points * 4, stream));
    init_fwd_header_kernel<<<1, 32, 0, stream>>>(pv.hdr, (int)n_subs, sps, (int)n_intervals);
    FO_LAUNCH_CHECK("init_fwd_header_kernel");
    const int64_t work = n_intervals > n_subs + 1 ? n_intervals : n_subs + 1;
    plan_from_intervals_kernel<<<grid_for(work, 256, 16), 256, 0, stream>>>(
        ranks_bev, interval_starts, interval_lengths, n_points, n_intervals, n_intervals_dev, n_vox,
        (int64_t)B * n_vox, sps, (int)n_subs, pv.hdr, pv.sub_iv, pv.sub_pt, pv.pos2iv, pv.iv_vox);
    FO_LAUNCH_CHECK("plan_from_intervals_kernel");
    heavy_queue_kernel<<<grid_for(n_subs, 256, 4), 256, 0, stream>>>(pv.sub_pt, (int)n_subs, pv.hdr, pv.heavy_list,
                                                                     heavy_threshold(B, n_vox));
    FO_LAUNCH_CHECK("heavy_queue_kernel");
    return FO_OK;
}

extern "C" size_t fo_rank_prepare_scratch_bytes(int64_t n_points_total, int64_t n_voxels_total) {
    if (n_points_total < 0 || n_voxels_total < 0) return 0;
    // counters + aggregates | slot[P] | two interval queues of the order pass [P/8 + 1 each] | (offset, interval) pairs [NV]
    const size_t legacy = bucket_zero_bytes(n_voxels_total) + (size_t)align_up(n_points_total * 4, 256) +
                          2 * (size_t)align_up((n_points_total / 8 + 1) * 4, 256) + (size_t)align_up(n_voxels_total * 8, 256);
    const size_t chunked = chunk_scratch_view(nullptr, chunk_bound(n_voxels_total), n_points_total).total_bytes;
    return legacy > chunked ? legacy : chunked;
}

namespace {
int rank_prepare_impl(cudaStream_t stream, const float *coor, const CalibArgs *calib, int32_t B, int32_t N, int32_t D,
                      int32_t H, int32_t W, const float lower_bound[3], const float interval[3], int32_t X,
                      int32_t Y, int32_t Z, int32_t *ranks_bev, int32_t *ranks_depth, int32_t *ranks_feat,
                      int32_t *interval_starts, int32_t *interval_lengths, int32_t *counts_dev, void *fwd_plan,
                      size_t fwd_plan_bytes, void *scratch, size_t scratch_bytes) {
    FO_CHECK_ARG(B >= 1 && N >= 1 && D >= 1 && H >= 1 && W >= 1, "non-positive frustum dims");
    FO_CHECK_ARG(X >= 1 && Y >= 1 && Z >= 1, "non-positive grid dims");
    FO_CHECK_ARG(lower_bound && interval, "NULL grid description");
    FO_CHECK_ARG(ranks_bev && ranks_depth && ranks_feat && interval_starts && interval_lengths && counts_dev,
                 "NULL output array");
    FO_CHECK_ARG(scratch != nullptr && ((uintptr_t)scratch & 255) == 0, "scratch must be non-NULL, 256-byte aligned");
    const int64_t pps = (int64_t)N * D * H * W;
    const int64_t P = pps * B;
    const int64_t n_vox = (int64_t)X * Y * Z;
    const int64_t NV = n_vox * B;
    FO_CHECK_ARG(P < INT_MAX && NV < INT_MAX, "point count %lld or voxel count %lld does not fit int32 ranks",
                 (long long)P, (long long)NV);
    const size_t need = fo_rank_prepare_scratch_bytes(P, NV);
    if (scratch_bytes < need)
        return set_error(FO_ERR_SCRATCH, "rank scratch is %zu bytes, need %zu", scratch_bytes, need);
    FwdPlanView pv; int64_t n_subs; int sps;
    if (int rc = open_fwd_plan(fwd_plan, fwd_plan_bytes, B, n_vox, P, &pv, &n_subs, &sps)) return rc;

    // ---- FO_RANK_IMPL=1: the two-level sort with chunks of 1024 voxels ordered in shared memory (rank_chunk.cuh);
    //      same results, measured slower than the global bucket sort below on a B200 (see the header of that file),
    //      so it is opt-in.
    {
        const char *re = getenv("FO_RANK_IMPL");
        const int impl = (re && *re) ? atoi(re) : 0;
        const ChunkGeom g = chunk_geom(n_vox, B);
        if (impl >= 1 && g.cps <= kChunkMaxPerSample && g.n_chunks <= chunk_bound(NV) && B <= 65535) {
            ChunkScratch sc = chunk_scratch_view(scratch, chunk_bound(NV), P);
            FO_CUDA(cudaMemsetAsync(scratch, 0, sc.zero_bytes, stream));
            FO_CUDA(cudaMemsetAsync(counts_dev, 0, 4 * sizeof(int32_t), stream));
            FO_CUDA(cudaMemsetAsync(pv.sub_mask, 0, ((size_t)n_subs + 1) * 4, stream));  // occupancy masks unknown on this path
            ChunkArgs ca;
            VoxArgs &va = ca.vox;
            va.coor = coor; va.n_points = P; va.points_per_sample = pps;
            va.lbx = lower_bound[0]; va.lby = lower_bound[1]; va.lbz = lower_bound[2];
            va.ivx = interval[0]; va.ivy = interval[1]; va.ivz = interval[2];
            va.X = X; va.Y = Y; va.Z = Z;
            va.cnt = nullptr; va.key = pv.pt2vox; va.slot = nullptr;
            va.hdr = pv.hdr; va.n_subs = (int)n_subs; va.subs_per_sample = sps;
            ca.calib = calib ? *calib : CalibArgs{};
            ca.g = g; ca.B = B; ca.pps = pps; ca.n_cams = N; ca.dhw_pts = D * H * W;
            ca.hist = sc.hist; ca.cursor = sc.cursor; ca.chunk_off = sc.chunk_off; ca.iv_off = sc.iv_off; ca.ctrl = sc.ctrl;
            ca.list = sc.list; ca.list_m = sc.list_m; ca.list_l = sc.list_l; ca.counts = counts_dev;
            ca.rb = ranks_bev; ca.rd = ranks_depth; ca.rf = ranks_feat;
            ca.iv_starts = interval_starts; ca.iv_lengths = interval_lengths;
            ca.sub_iv = pv.sub_iv; ca.sub_pt = pv.sub_pt; ca.heavy_list = pv.heavy_list; ca.vox2iv = pv.vox2iv;
            ca.iv_vox = pv.iv_vox; ca.hdr = pv.hdr; ca.heavy_pts = heavy_threshold(B, n_vox);
            ca.dhw = make_fastdiv((uint32_t)(D * H * W)); ca.hw = make_fastdiv((uint32_t)(H * W));
            ca.n_subs = (int)n_subs;
            FO_CHECK_ARG((int64_t)B * N <= 65535, "B*N=%lld exceeds the grid bound", (long long)B * N);
            const dim3 grid_a((unsigned)(((int64_t)D * H * W + kChunkBlockPts - 1) / kChunkBlockPts), (unsigned)(B * N));
            const size_t smem_a = (size_t)g.cps * sizeof(int);
            if (calib) chunk_voxelize_kernel<true><<<grid_a, kChunkThreads, smem_a, stream>>>(ca);
            else chunk_voxelize_kernel<false><<<grid_a, kChunkThreads, smem_a, stream>>>(ca);
            FO_LAUNCH_CHECK("chunk_voxelize_kernel");
            chunk_scatter_kernel<<<grid_a, kChunkThreads, smem_a, stream>>>(ca);
            FO_LAUNCH_CHECK("chunk_scatter_kernel");
            chunk_distinct_kernel<<<(g.n_chunks + 7) / 8, kChunkThreads, 0, stream>>>(ca);
            FO_LAUNCH_CHECK("chunk_distinct_kernel");
            const size_t smem_s = 4 * ((chunk_sort_smem(kClassS) + 15) / 16 * 16);
            const size_t smem_m = chunk_sort_smem(kClassM), smem_l = chunk_sort_smem(kStageL);
            // only the dense class needs more than the default 48 KB of dynamic shared memory
            FO_CUDA(cudaFuncSetAttribute(chunk_sort_cta_kernel<16, kStageL, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
            FO_CUDA(cudaFuncSetAttribute(chunk_sort_cta_kernel<8, kClassM, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m));
            // the dense classes first: their CTAs are the long ones
            const int sms = sm_count();
            chunk_sort_cta_kernel<16, kStageL, true><<<sms < g.n_chunks ? sms : g.n_chunks, 512, smem_l, stream>>>(ca);
            FO_LAUNCH_CHECK("chunk_sort_cta_kernel<large>");
            chunk_sort_cta_kernel<8, kClassM, false><<<4 * sms < g.n_chunks ? 4 * sms : g.n_chunks, 256, smem_m, stream>>>(ca);
            FO_LAUNCH_CHECK("chunk_sort_cta_kernel<medium>");
            chunk_sort_small_kernel<<<(g.n_chunks + 3) / 4, 128, smem_s, stream>>>(ca);
            FO_LAUNCH_CHECK("chunk_sort_small_kernel");
            return FO_OK;
        }
    }
    SortScratch ss = sort_scratch_view(scratch, NV);
    int32_t *slot = (int32_t *)((char *)scratch + ss.zero_bytes);
    int32_t *key = pv.pt2vox;                    // voxel id of every frustum point, -1 if outside the grid

    // zero the counters / aggregates and the four count words (counts_dev: 16-byte aligned like every tensor base)
    if (((uintptr_t)counts_dev & 15) == 0) {
        FO_CUDA(launch_pdl(kPdlZero, zero2_kernel, dim3(grid_for((int64_t)(ss.zero_bytes / 16), 256, 8)), dim3(256), 0, stream,
                           (uint4 *)scratch, (int64_t)(ss.zero_bytes / 16), (uint4 *)counts_dev, (int64_t)1));
    } else {
        FO_CUDA(cudaMemsetAsync(scratch, 0, ss.zero_bytes, stream));
        FO_CUDA(cudaMemsetAsync(counts_dev, 0, 4 * sizeof(int32_t), stream));
    }

    VoxArgs va;
    va.coor = coor; va.n_points = P; va.points_per_sample = pps;
    va.lbx = lower_bound[0]; va.lby = lower_bound[1]; va.lbz = lower_bound[2];
    va.ivx = interval[0]; va.ivy = interval[1]; va.ivz = interval[2];
    va.X = X; va.Y = Y; va.Z = Z;
    va.cnt = ss.cnt; va.key = key; va.slot = slot;
    va.hdr = pv.hdr; va.n_subs = (int)n_subs; va.subs_per_sample = sps;
    // FO_VOX_IMPL (A/B): 0 = round-1 point loop (per-point calibration loads, point -> camera division, IEEE divisions),
    // 1 = one camera per CTA: calibration in registers, division only near voxel boundaries
    const char *ve = getenv("FO_VOX_IMPL");
    const int vox_impl = (ve && *ve) ? atoi(ve) : (calib ? 1 : 0);
    if (vox_impl >= 1 && (int64_t)B * N <= 65535) {
        const int dhw = D * H * W, per = kChunkThreads * FO_VOXCAM_PPT;
        const dim3 grid_v((unsigned)((dhw + per - 1) / per), (unsigned)(B * N));
        if (calib) {
            switch (calib->mode) {
                case 0: FO_CUDA(launch_pdl(kPdlRank, voxelize_count_cam_kernel<true, 0>, grid_v, dim3(kChunkThreads), 0, stream, va, *calib, N, dhw)); break;
                case 1: FO_CUDA(launch_pdl(kPdlRank, voxelize_count_cam_kernel<true, 1>, grid_v, dim3(kChunkThreads), 0, stream, va, *calib, N, dhw)); break;
                case 2: FO_CUDA(launch_pdl(kPdlRank, voxelize_count_cam_kernel<true, 2>, grid_v, dim3(kChunkThreads), 0, stream, va, *calib, N, dhw)); break;
                default: FO_CUDA(launch_pdl(kPdlRank, voxelize_count_cam_kernel<true, 3>, grid_v, dim3(kChunkThreads), 0, stream, va, *calib, N, dhw)); break;
            }
        } else FO_CUDA(launch_pdl(kPdlRank, voxelize_count_cam_kernel<false, 0>, grid_v, dim3(kChunkThreads), 0, stream, va, CalibArgs{}, N, dhw));
    } else if (calib) FO_CUDA(launch_pdl(kPdlRank, voxelize_count_kernel<true>, dim3(grid_for((P + 3) / 4, 256)), dim3(256), 0, stream, va, *calib));
    else FO_CUDA(launch_pdl(kPdlRank, voxelize_count_kernel<false>, dim3(grid_for((P + 3) / 4, 256)), dim3(256), 0, stream, va, CalibArgs{}));
    FO_LAUNCH_CHECK("voxelize_count_kernel");

    ScanArgs sa;
    sa.cnt = ss.cnt; sa.n_buckets = NV;
    sa.iv_starts = interval_starts; sa.iv_lengths = interval_lengths; sa.iv_bucket = pv.iv_vox;
    sa.bucket2iv = pv.vox2iv;
    sa.totals = counts_dev;
    sa.sub_iv = pv.sub_iv; sa.sub_pt = pv.sub_pt; sa.sub_mask = nullptr; sa.vox_per_sample = n_vox; sa.subs_per_sample = sps;
    sa.n_subs = (int)n_subs;
    sa.fwd_hdr = pv.hdr; sa.bwd_hdr = nullptr;
    sa.agg = ss.agg; sa.agg_group = ss.agg_group;
    const int scan_blocks = (int)((NV + kScanTile - 1) / kScanTile);
    FO_CUDA(launch_pdl(kPdlRank, tile_reduce_kernel, dim3(scan_blocks), dim3(kScanThreads), 0, stream, (const int32_t *)ss.cnt, NV,
                       ss.agg, ss.agg_group));

    // FO_RANK_FAST (A/B): 1 (default) = scan with staged outputs and a queue of the long intervals + ONE order launch;
    // 0 = the round-1 passes (scan, order_short queueing the long intervals, order_long)
    const char *fe = getenv("FO_RANK_FAST");
    const int fast = (fe && *fe) ? atoi(fe) : 1;
    const FastDiv fd_dhw = make_fastdiv((uint32_t)(D * H * W)), fd_hw = make_fastdiv((uint32_t)(H * W));
    if (fast >= 1) {
        // long-interval queue: lives behind the slot array (both are live until the placement has run)
        ListArgs la;
        la.long_list = (int32_t *)((char *)slot + align_up(P * 4, 256)); la.long_cap = (int32_t)(P / 8 + 1);
        la.mid_list = (int32_t *)((char *)la.long_list + align_up(((int64_t)P / 8 + 1) * 4, 256)); la.mid_cap = (int32_t)(P / 8 + 1);
        la.counts = ss.counter;
        {
            // 9..16-point intervals one lane each for dense frusta (more points than voxels: 512x1408 rank precompute
            // 283.4 -> 269.7 us at batch 8; at the headline shape there are too few of them to pay for the registers,
            // 95.8 -> 98.0 us); FO_RANK_MID=0/1 forces it off / on (A/B)
            const char *me = getenv("FO_RANK_MID");
            const bool mid = (me && *me) ? atoi(me) != 0 : P > NV;
            la.mid_max = mid ? kMidSortMax : kLaneSortMax;
        }
        sa.sub_mask = pv.sub_mask;
        // FO_RANK_FAST=2 (opt-in): (offset, interval id) pairs + a placement that turns pt2vox into pt2iv, so that the
        // backward plan needs no voxel -> interval gather.  Measured (profiles/r02_summary.md): backward -7 us, rank
        // precompute +6 us at batch 8 (+19 / -16 us at 512x1408): a wash with a backward, a loss without one.
        la.ofiv = fast >= 2 ? (int2 *)((char *)la.mid_list + align_up(((int64_t)P / 8 + 1) * 4, 256)) : nullptr;
        if (la.ofiv) sa.bucket2iv = nullptr;
        FO_CUDA(launch_pdl(kPdlRank, scan_buckets2_kernel, dim3(scan_blocks), dim3(kScan2Threads), 0, stream, sa, la));
        if (la.ofiv)
            FO_CUDA(launch_pdl(kPdlRank, place_iv_kernel, dim3(grid_for((P + 3) / 4, 256)), dim3(256), 0, stream, key,
                               (const int32_t *)slot, (const int2 *)la.ofiv, P, ranks_depth, pv.hdr));
        else
            FO_CUDA(launch_pdl(kPdlRank, place_kernel, dim3(grid_for(P, 256)), dim3(256), 0, stream, (const int32_t *)key,
                               (const int32_t *)slot, (const int32_t *)ss.cnt, P, (const int32_t *)nullptr, NV, ranks_depth));
        Order2Args o2;
        o2.sorted = ranks_depth; o2.ranks_feat = ranks_feat; o2.ranks_bev = ranks_bev; o2.iv_starts = interval_starts;
        o2.iv_lengths = interval_lengths; o2.iv_bucket = pv.iv_vox; o2.n_intervals = counts_dev + 1;
        o2.l = la; o2.dhw = fd_dhw; o2.hw = fd_hw;
        o2.sub_pt = pv.sub_pt; o2.n_subs = (int32_t)n_subs; o2.heavy_list = pv.heavy_list; o2.heavy_n = pv.hdr->fwd_heavy;
        o2.heavy_pts = heavy_threshold(B, n_vox);
        if (la.mid_max > kLaneSortMax) FO_CUDA(launch_pdl(kPdlRank, order2_kernel<true>, dim3(sm_count() * 16), dim3(kSortThreads), 0, stream, o2));
        else FO_CUDA(launch_pdl(kPdlRank, order2_kernel<false>, dim3(sm_count() * 16), dim3(kSortThreads), 0, stream, o2));
        return FO_OK;
    }
    FO_CUDA(cudaMemsetAsync(pv.sub_mask, 0, ((size_t)n_subs + 1) * 4, stream));   // occupancy masks unknown on this path
    FO_CUDA(launch_pdl(kPdlRank, scan_buckets_kernel, dim3(scan_blocks), dim3(kScanThreads), 0, stream, sa));
    FO_CUDA(launch_pdl(kPdlRank, place_kernel, dim3(grid_for(P, 256)), dim3(256), 0, stream, (const int32_t *)key,
                       (const int32_t *)slot, (const int32_t *)ss.cnt, P, (const int32_t *)nullptr, NV, ranks_depth));

    OrderArgs oa;
    oa.sorted = ranks_depth; oa.iv_starts = interval_starts; oa.iv_lengths = interval_lengths;
    oa.iv_bucket = pv.iv_vox; oa.n_intervals = counts_dev + 1;
    oa.ranks_feat = ranks_feat; oa.ranks_bev = ranks_bev;
    oa.sub_pt = pv.sub_pt; oa.n_subs = (int32_t)n_subs; oa.heavy_list = pv.heavy_list; oa.heavy_n = pv.hdr->fwd_heavy;
    oa.heavy_pts = heavy_threshold(B, n_vox);
    oa.dhw = fd_dhw; oa.hw = fd_hw;
    oa.long_list = slot; oa.long_count = ss.counter;       // the slot array is dead after the placement
    oa.long_cap = (int32_t)P;
    const int64_t cap_iv = P < NV ? P : NV;
    FO_CUDA(launch_pdl(kPdlRank, order_short_kernel<true>, dim3(grid_for(cap_iv, 256)), dim3(256), 0, stream, oa));
    FO_CUDA(launch_pdl(kPdlRank, order_long_kernel<true>, dim3(sm_count() * 16), dim3(kSortThreads), 0, stream, oa));
    return FO_OK;
}
}  // namespace

extern "C" int fo_rank_prepare(fo_stream_t stream_, const float *coor, int32_t B, int32_t N, int32_t D, int32_t H,
                               int32_t W, const float lower_bound[3], const float interval[3], int32_t X, int32_t Y,
                               int32_t Z, int32_t *ranks_bev, int32_t *ranks_depth, int32_t *ranks_feat,
                               int32_t *interval_starts, int32_t *interval_lengths, int32_t *counts_dev,
                               void *fwd_plan, size_t fwd_plan_bytes, void *scratch, size_t scratch_bytes) {
    FO_CHECK_ARG(coor != nullptr && ((uintptr_t)coor & 15) == 0, "coor must be non-NULL and 16-byte aligned");
    return rank_prepare_impl((cudaStream_t)stream_, coor, nullptr, B, N, D, H, W, lower_bound, interval, X, Y, Z,
                             ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths, counts_dev,
                             fwd_plan, fwd_plan_bytes, scratch, scratch_bytes);
}

extern "C" int fo_rank_prepare_calib(fo_stream_t stream_, const float *frustum, const float *cam_mats,
                                     const float *bda, int32_t bda_has_translation, int32_t matvec_mode,
                                     float *coor_out, int32_t B, int32_t N, int32_t D, int32_t H, int32_t W,
                                     const float lower_bound[3], const float interval[3], int32_t X, int32_t Y,
                                     int32_t Z, int32_t *ranks_bev, int32_t *ranks_depth, int32_t *ranks_feat,
                                     int32_t *interval_starts, int32_t *interval_lengths, int32_t *counts_dev,
                                     void *fwd_plan, size_t fwd_plan_bytes, void *scratch, size_t scratch_bytes) {
    FO_CHECK_ARG(frustum && cam_mats && bda, "NULL calibration input");
    FO_CHECK_ARG(matvec_mode >= 0 && matvec_mode <= 3, "unknown matvec_mode %d", matvec_mode);
    FO_CHECK_ARG(coor_out == nullptr || ((uintptr_t)coor_out & 15) == 0, "coor_out must be 16-byte aligned");
    FO_CHECK_ARG(B >= 1 && N >= 1 && D >= 1 && H >= 1 && W >= 1, "non-positive frustum dims");
    CalibArgs g;
    g.frustum = frustum; g.cam = cam_mats; g.bda = bda;
    g.n_cams = N; g.dhw = D * H * W; g.mode = matvec_mode; g.bda_has_t = bda_has_translation ? 1 : 0;
    g.coor_out = coor_out;
    return rank_prepare_impl((cudaStream_t)stream_, nullptr, &g, B, N, D, H, W, lower_bound, interval, X, Y, Z,
                             ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths, counts_dev,
                             fwd_plan, fwd_plan_bytes, scratch, scratch_bytes);
}

// ------------------------------------------------------------------------------------------------
// Rank pipeline from caller-supplied integer bucket ids (sibling ops bev_pool v1 / occ_pool, SURVEY.md §8f-4:
// projects/BEVFusion/bevfusion/ops/bev_pool/bev_pool.py:85-99 computes ranks from integer coords, argsorts and
// re-discovers the runs with torch ops).  Same stable one-digit bucket sort as fo_rank_prepare.
// ------------------------------------------------------------------------------------------------
extern "C" size_t fo_rank_from_keys_scratch_bytes(int64_t n_points, int64_t n_buckets) {
    if (n_points < 0 || n_buckets < 1) return 0;
    const int64_t cap_iv = n_points < n_buckets ? n_points : n_buckets;
    return bucket_zero_bytes(n_buckets) + (size_t)align_up(n_points * 4, 256) + (size_t)align_up(cap_iv * 4, 256);
}

extern "C" int fo_rank_from_keys(fo_stream_t stream_, const int32_t *keys, int64_t n_points, int64_t n_buckets,
                                 int32_t *sorted_keys, int32_t *order, int32_t *interval_starts,
                                 int32_t *interval_lengths, int32_t *counts_dev, void *scratch,
                                 size_t scratch_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FO_CHECK_ARG(n_points >= 0 && n_points < INT_MAX && n_buckets >= 1 && n_buckets < INT_MAX, "bad sizes");
    FO_CHECK_ARG(counts_dev != nullptr, "counts_dev is NULL");
    FO_CUDA(cudaMemsetAsync(counts_dev, 0, 4 * sizeof(int32_t), stream));
    if (n_points == 0) return FO_OK;
    FO_CHECK_ARG(keys && sorted_keys && order && interval_starts && interval_lengths, "NULL array");
    FO_CHECK_ARG(scratch != nullptr && ((uintptr_t)scratch & 255) == 0, "scratch must be non-NULL, 256-byte aligned");
    const size_t need = fo_rank_from_keys_scratch_bytes(n_points, n_buckets);
    if (scratch_bytes < need)
        return set_error(FO_ERR_SCRATCH, "rank scratch is %zu bytes, need %zu", scratch_bytes, need);
    SortScratch ss = sort_scratch_view(scratch, n_buckets);
    int32_t *slot = (int32_t *)((char *)scratch + ss.zero_bytes);
    int32_t *iv_bucket = (int32_t *)((char *)slot + align_up(n_points * 4, 256));
    FO_CUDA(cudaMemsetAsync(scratch, 0, ss.zero_bytes, stream));
    count_keys_kernel<<<grid_for(n_points, 256), 256, 0, stream>>>(keys, n_points, nullptr, n_buckets, ss.cnt, slot,
                                                                   nullptr);
    FO_LAUNCH_CHECK("count_keys_kernel");
    ScanArgs sa;
    sa.cnt = ss.cnt; sa.n_buckets = n_buckets;
    sa.iv_starts = interval_starts; sa.iv_lengths = interval_lengths; sa.iv_bucket = iv_bucket;
    sa.bucket2iv = nullptr;
    sa.totals = counts_dev;
    sa.sub_iv = nullptr; sa.sub_pt = nullptr; sa.sub_mask = nullptr; sa.vox_per_sample = 1; sa.subs_per_sample = 0; sa.n_subs = 0;
    sa.fwd_hdr = nullptr; sa.bwd_hdr = nullptr;
    sa.agg = ss.agg; sa.agg_group = ss.agg_group;
    const int scan_blocks = (int)((n_buckets + kScanTile - 1) / kScanTile);
    tile_reduce_kernel<<<scan_blocks, kScanThreads, 0, stream>>>(ss.cnt, n_buckets, ss.agg, ss.agg_group);
    FO_LAUNCH_CHECK("tile_reduce_kernel");
    scan_buckets_kernel<<<scan_blocks, kScanThreads, 0, stream>>>(sa);
    FO_LAUNCH_CHECK("scan_buckets_kernel");
    place_kernel<<<grid_for(n_points, 256), 256, 0, stream>>>(keys, slot, ss.cnt, n_points, nullptr, n_buckets, order);
    FO_LAUNCH_CHECK("place_kernel");
    OrderArgs oa;
    oa.sorted = order; oa.iv_starts = interval_starts; oa.iv_lengths = interval_lengths;
    oa.iv_bucket = iv_bucket; oa.n_intervals = counts_dev + 1;
    oa.sub_pt = nullptr; oa.n_subs = 0; oa.heavy_list = nullptr; oa.heavy_n = nullptr; oa.heavy_pts = kHeavyPts;
    oa.ranks_feat = nullptr; oa.ranks_bev = sorted_keys;
    oa.dhw = make_fastdiv(1); oa.hw = make_fastdiv(1);
    oa.long_list = slot; oa.long_count = ss.counter; oa.long_cap = (int32_t)n_points;
    const int64_t cap_iv = n_points < n_buckets ? n_points : n_buckets;
    order_short_kernel<true><<<grid_for(cap_iv, 256), 256, 0, stream>>>(oa);
    FO_LAUNCH_CHECK("order_short_kernel<keys>");
    order_long_kernel<true><<<sm_count() * 16, kSortThreads, 0, stream>>>(oa);
    FO_LAUNCH_CHECK("order_long_kernel<keys>");
    return FO_OK;
}

extern "C" size_t fo_bwd_plan_bytes(int64_t n_points_capacity, int64_t n_feat_rows) {
    if (n_points_capacity < 0 || n_feat_rows < 1) return 0;
    return bwd_plan_bytes_for(n_points_capacity, n_feat_rows);
}

extern "C" int fo_bwd_plan_build(fo_stream_t stream_, const int32_t *ranks_depth, const int32_t *ranks_feat,
                                 int64_t n_points, const int32_t *n_points_dev, int64_t n_depth,
                                 int64_t n_feat_rows, int32_t hw, int32_t flags, const void *fwd_plan,
                                 size_t fwd_plan_bytes, int32_t B, int64_t n_vox, void *plan, size_t plan_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FO_CHECK_ARG(plan != nullptr && ((uintptr_t)plan & 255) == 0, "backward plan must be non-NULL, 256-byte aligned");
    FO_CHECK_ARG(n_points >= 0 && n_points < INT_MAX && n_feat_rows >= 1 && n_feat_rows < INT_MAX && n_depth >= 0,
                 "bad sizes n_points=%lld n_feat_rows=%lld", (long long)n_points, (long long)n_feat_rows);
    FwdPlanView fv; int64_t n_subs; int sps;
    if (int rc = open_fwd_plan_const(fwd_plan, fwd_plan_bytes, B, n_vox, n_points, &fv, &n_subs, &sps)) return rc;
    BwdPlanView bv;
    const bool structured = (flags & FO_BWD_PLAN_STRUCTURED) != 0;
    const int64_t need_cap = structured ? n_depth : n_points;
    if (!bwd_plan_view(plan, n_feat_rows, plan_bytes, &bv) || bv.cap < need_cap)
        return set_error(FO_ERR_SCRATCH, "backward plan buffer is %zu bytes, need %zu", plan_bytes,
                         bwd_plan_bytes_for(need_cap, n_feat_rows));
    if (structured) {
        FO_CHECK_ARG(hw >= 1 && n_feat_rows % hw == 0 && n_depth % n_feat_rows == 0,
                     "structured plan needs n_depth = n_feat_rows * D and n_feat_rows = B*N*hw");
        const int D = (int)(n_depth / n_feat_rows);
        FO_CHECK_ARG(n_depth <= fv.p_cap, "forward plan holds %lld points, structured build needs %lld",
                     (long long)fv.p_cap, (long long)n_depth);
        const int R = (D + 31) / 32;
        if (R > 8) return set_error(FO_ERR_UNSUPPORTED, "structured backward plan supports D <= 256 (got %d)", D);
        #ifndef FO_PLAN_CTAS_PER_SM
#define FO_PLAN_CTAS_PER_SM 32     // one pixel per warp, many short CTAs: 125 -> 111 us at 512x1408 (8 -> 32)
#endif
        const int blocks = grid_for(n_feat_rows * 32, 256, FO_PLAN_CTAS_PER_SM);
        if (D <= 128 && (int64_t)B * n_vox < (1 << 24)) {   // packed-key bitonic variant
#define FO_BITONIC(RR)                                                                                          \
    FO_CUDA(launch_pdl(kPdlGather, bwd_plan_structured_bitonic_kernel<RR>, dim3(blocks), dim3(256), 0, stream,   \
                       (const FwdPlanHeader *)fv.hdr, (const int32_t *)fv.pt2vox, (const int32_t *)fv.vox2iv, D, \
                       (int)hw, (int)n_feat_rows, bv.hdr, bv.ent_p, bv.ent_iv, bv.starts, bv.lengths, bv.ids,   \
                       n_points_dev))
            if (R == 1) FO_BITONIC(1);
            else if (R == 2) FO_BITONIC(2);
            else FO_BITONIC(4);
#undef FO_BITONIC
            FO_LAUNCH_CHECK("bwd_plan_structured_bitonic_kernel");
            return FO_OK;
        }
#define FO_STRUCT(RR)                                                                                         \
    bwd_plan_structured_kernel<RR><<<blocks, 256, 0, stream>>>(fv.hdr, fv.pt2vox, fv.vox2iv, D, hw, (int)n_feat_rows, \
                                                              bv.hdr, bv.ent_p, bv.ent_iv, bv.starts,        \
                                                              bv.lengths, bv.ids, n_points_dev)
        switch (R) {
            case 1: FO_STRUCT(1); break;
            case 2: FO_STRUCT(2); break;
            case 3: FO_STRUCT(3); break;
            case 4: FO_STRUCT(4); break;
            case 5: FO_STRUCT(5); break;
            case 6: FO_STRUCT(6); break;
            case 7: FO_STRUCT(7); break;
            default: FO_STRUCT(8); break;
        }
#undef FO_STRUCT
        FO_LAUNCH_CHECK("bwd_plan_structured_kernel");
        return FO_OK;
    }
    // generic build: stable bucket sort of forward positions by ranks_feat
    FO_CHECK_ARG(n_points == 0 || (ranks_feat && ranks_depth), "ranks_depth / ranks_feat is NULL");
    SortScratch ss = sort_scratch_view(bv.counters, n_feat_rows);
    FO_CUDA(cudaMemsetAsync(bv.counters, 0, ss.zero_bytes, stream));
    FO_CUDA(cudaMemsetAsync(bv.hdr, 0, sizeof(BwdPlanHeader), stream));
    if (n_points == 0) return FO_OK;
    count_keys_kernel<<<grid_for(n_points, 256), 256, 0, stream>>>(ranks_feat, n_points, n_points_dev, n_feat_rows,
                                                                   ss.cnt, bv.slot, bv.hdr);
    FO_LAUNCH_CHECK("count_keys_kernel");
    ScanArgs sa;
    sa.cnt = ss.cnt; sa.n_buckets = n_feat_rows;
    sa.iv_starts = bv.starts; sa.iv_lengths = bv.lengths; sa.iv_bucket = bv.ids;
    sa.bucket2iv = nullptr;
    sa.totals = bv.hdr->totals;
    sa.sub_iv = nullptr; sa.sub_pt = nullptr; sa.sub_mask = nullptr; sa.vox_per_sample = 1; sa.subs_per_sample = 0; sa.n_subs = 0;
    sa.fwd_hdr = nullptr; sa.bwd_hdr = bv.hdr;
    sa.agg = ss.agg; sa.agg_group = ss.agg_group;
    const int scan_blocks = (int)((n_feat_rows + kScanTile - 1) / kScanTile);
    tile_reduce_kernel<<<scan_blocks, kScanThreads, 0, stream>>>(ss.cnt, n_feat_rows, ss.agg, ss.agg_group);
    FO_LAUNCH_CHECK("tile_reduce_kernel");
    scan_buckets_kernel<<<scan_blocks, kScanThreads, 0, stream>>>(sa);
    FO_LAUNCH_CHECK("scan_buckets_kernel");
    place_kernel<<<grid_for(n_points, 256), 256, 0, stream>>>(ranks_feat, bv.slot, ss.cnt, n_points, n_points_dev,
                                                              n_feat_rows, bv.pos);
    FO_LAUNCH_CHECK("place_kernel");
    OrderArgs oa;
    oa.sorted = bv.pos; oa.iv_starts = bv.starts; oa.iv_lengths = bv.lengths; oa.iv_bucket = nullptr;
    oa.n_intervals = &bv.hdr->n_bwd_intervals;
    oa.sub_pt = nullptr; oa.n_subs = 0; oa.heavy_list = nullptr; oa.heavy_n = nullptr; oa.heavy_pts = kHeavyPts;
    oa.ranks_feat = nullptr; oa.ranks_bev = nullptr; oa.dhw = make_fastdiv(1); oa.hw = make_fastdiv(1);
    oa.long_list = bv.slot; oa.long_count = ss.counter; oa.long_cap = (int32_t)bv.cap;
    order_short_kernel<false><<<grid_for(n_feat_rows, 256, 8), 256, 0, stream>>>(oa);
    FO_LAUNCH_CHECK("order_short_kernel<bwd>");
    order_long_kernel<false><<<sm_count() * 16, kSortThreads, 0, stream>>>(oa);
    FO_LAUNCH_CHECK("order_long_kernel<bwd>");
    bwd_plan_fill_entries_kernel<<<grid_for(n_points, 256), 256, 0, stream>>>(bv.pos, ranks_depth, fv.pos2iv, bv.hdr,
                                                                              bv.cap, bv.ent_p, bv.ent_iv);
    FO_LAUNCH_CHECK("bwd_plan_fill_entries_kernel");
    return FO_OK;
}
