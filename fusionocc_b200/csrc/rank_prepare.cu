// fusionocc_b200 — rank precompute and plan construction (sm_100a).
//
// Replaces projects/FusionOcc/fusionocc/necks/view_transformer.py:223-281 (voxel_pooling_prepare_v2)
// and the backward re-sort of mmdet3d/ops/bev_pool_v2/bev_pool.py:47-57.  See bucket_sort.cuh for
// the sort itself.  Every kernel here is HBM/L2-bound integer work; nothing is reshaped into GEMMs.
#include <limits.h>

#include "bucket_sort.cuh"

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
    int32_t *key;                // [n_points] voxel id or -1
    int32_t *slot;               // [n_points]
    FwdPlanHeader *hdr;          // optional: static fields initialised by thread 0
    int32_t n_tiles, tiles_per_sample;
};

__device__ __forceinline__ int voxel_key(float x, float y, float z, int64_t b, const VoxArgs &a) {
    const long long ix = (long long)__fdiv_rn(__fsub_rn(x, a.lbx), a.ivx);
    const long long iy = (long long)__fdiv_rn(__fsub_rn(y, a.lby), a.ivy);
    const long long iz = (long long)__fdiv_rn(__fsub_rn(z, a.lbz), a.ivz);
    const bool kept = ix >= 0 && ix < a.X && iy >= 0 && iy < a.Y && iz >= 0 && iz < a.Z;
    return kept ? (int)(((b * a.Z + iz) * a.Y + iy) * a.X + ix) : -1;
}

__global__ void __launch_bounds__(256) voxelize_count_kernel(VoxArgs a) {
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid == 0 && a.hdr) {
        a.hdr->flags = 0;
        a.hdr->n_tiles = a.n_tiles;
        a.hdr->tiles_per_sample = a.tiles_per_sample;
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_quads = a.n_points >> 2;
    for (int64_t qd = gtid; qd < n_quads; qd += stride) {
        const float4 *src = reinterpret_cast<const float4 *>(a.coor) + qd * 3;
        const float4 v0 = __ldcs(src), v1 = __ldcs(src + 1), v2 = __ldcs(src + 2);
        const float xs[4] = {v0.x, v0.w, v1.z, v2.y};
        const float ys[4] = {v0.y, v1.x, v1.w, v2.z};
        const float zs[4] = {v0.z, v1.y, v2.x, v2.w};
        int keys[4], slots[4];
        // n_points < 2^31 is enforced by the host wrapper: 32-bit index arithmetic
        const int p0 = (int)(qd << 2), pps = (int)a.points_per_sample;
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
        const int k = voxel_key(a.coor[3 * p], a.coor[3 * p + 1], a.coor[3 * p + 2], p / a.points_per_sample, a);
        a.key[p] = k;
        a.slot[p] = k >= 0 ? atomicAdd(a.cnt + k, 1) : 0;
    }
}

// K1 (backward flavour): keys are given (ranks_feat of each forward position).
__global__ void __launch_bounds__(256) count_keys_kernel(const int32_t *__restrict__ keys, int64_t n_cap,
                                                         const int32_t *__restrict__ n_dev, int64_t n_buckets,
                                                         int32_t *cnt, int32_t *slot) {
    const int64_t n = n_dev ? min((int64_t)max(*n_dev, 0), n_cap) : n_cap;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int k = keys[i];
        slot[i] = (k >= 0 && k < n_buckets) ? atomicAdd(cnt + k, 1) : -1;
    }
}

// K3: placement.  sorted[offset[key] + slot] = original index.
__global__ void __launch_bounds__(256) place_kernel(const int32_t *__restrict__ key, const int32_t *__restrict__ slot,
                                                    const int32_t *__restrict__ offs, int64_t n_cap,
                                                    const int32_t *__restrict__ n_dev, int64_t n_buckets,
                                                    int32_t *__restrict__ sorted) {
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
__global__ void init_fwd_header_kernel(FwdPlanHeader *hdr, int n_tiles, int tps, int n_intervals) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        hdr->flags = 0;
        hdr->n_tiles = n_tiles;
        hdr->tiles_per_sample = tps;
        hdr->n_intervals = n_intervals;
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
    int64_t n_vox_total, int tps, int n_tiles, FwdPlanHeader *hdr, int32_t *tile_off, int32_t *pos2iv) {
    const int64_t n = n_dev ? min((int64_t)max(*n_dev, 0), n_cap) : n_cap;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid == 0) hdr->n_intervals = (int)n;
    if (n == 0) {
        for (int64_t t = gtid; t <= n_tiles; t += stride) tile_off[t] = 0;
        return;
    }
    for (int64_t k = gtid; k < n; k += stride) {
        const int v = interval_voxel(rb, starts, lengths, (int)k, n_points, n_vox_total);
        const int vp = k > 0 ? interval_voxel(rb, starts, lengths, (int)k - 1, n_points, n_vox_total) : -1;
        if (v < 0) {
            atomicOr(&hdr->flags, kFlagOutOfRange | kFlagUnsorted);
            continue;
        }
        const int s = starts[k], len = lengths[k];
        for (int j = 0; j < len; ++j) pos2iv[s + j] = (int)k;
        if (k > 0 && (vp < 0 || v <= vp)) {   // vp < 0: interval k-1 is invalid, its own thread raised the flags
            atomicOr(&hdr->flags, kFlagUnsorted);
            continue;
        }
        const int64_t t = (v / vox_per_sample) * tps + (v % vox_per_sample) / kTile;
        const int64_t tp = (k > 0) ? ((vp / vox_per_sample) * tps + (vp % vox_per_sample) / kTile) : -1;
        for (int64_t u = tp + 1; u <= t; ++u) tile_off[u] = (int)k;
        if (k == n - 1)
            for (int64_t u = t + 1; u <= n_tiles; ++u) tile_off[u] = (int)n;
    }
}

}  // namespace fo

using namespace fo;

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" size_t fo_fwd_plan_bytes(int64_t n_voxels_total, int64_t n_points_capacity) {
    if (n_voxels_total < 0 || n_points_capacity < 0) return 0;
    // upper bound on tiles: every sample may add one partial tile; callers pass B*Z*Y*X so use a
    // conservative bound of n_vox/kTile + (number of samples <= n_vox) ... bounded by n_vox itself.
    const int64_t n_tiles_max = n_voxels_total / kTile + 4096 + 1;
    return (size_t)(256 + fwd_plan_tile_bytes(n_tiles_max) + align_up(n_points_capacity * 4, 256));
}

static int check_fwd_plan(size_t plan_bytes, int32_t B, int64_t n_vox, int64_t n_points_cap, int64_t *n_tiles_out,
                          int *tps_out) {
    const int64_t tps = tiles_per_sample(n_vox);
    const int64_t n_tiles = tps * B;
    FO_CHECK_ARG(B >= 1 && n_vox >= 1, "B=%d and n_voxels_per_sample=%lld must be positive", B, (long long)n_vox);
    FO_CHECK_ARG((int64_t)B * n_vox < INT_MAX, "B*Z*Y*X = %lld does not fit int32 ranks", (long long)B * n_vox);
    FO_CHECK_ARG(B <= 4096, "B=%d exceeds the plan's sample bound (4096)", B);
    const size_t need = 256 + fwd_plan_tile_bytes(n_tiles) + (size_t)align_up(n_points_cap * 4, 256);
    if (plan_bytes < need)
        return set_error(FO_ERR_SCRATCH, "forward plan buffer is %zu bytes, need %zu", plan_bytes, need);
    *n_tiles_out = n_tiles;
    *tps_out = (int)tps;
    return FO_OK;
}

extern "C" int fo_fwd_plan_build(fo_stream_t stream_, const int32_t *ranks_bev, const int32_t *interval_starts,
                                 const int32_t *interval_lengths, int64_t n_points, int64_t n_intervals,
                                 const int32_t *n_intervals_dev, int32_t B, int64_t n_vox, void *plan,
                                 size_t plan_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FO_CHECK_ARG(plan != nullptr, "plan is NULL");
    FO_CHECK_ARG(n_points >= 0 && n_intervals >= 0 && n_points < INT_MAX, "negative or oversized counts");
    FO_CHECK_ARG(n_intervals == 0 || (ranks_bev && interval_starts && interval_lengths), "NULL index array");
    int64_t n_tiles; int tps;
    if (int rc = check_fwd_plan(plan_bytes, B, n_vox, n_points, &n_tiles, &tps)) return rc;
    FwdPlanView pv = fwd_plan_view(plan, n_tiles);
    init_fwd_header_kernel<<<1, 32, 0, stream>>>(pv.hdr, (int)n_tiles, tps, (int)n_intervals);
    FO_LAUNCH_CHECK("init_fwd_header_kernel");
    const int64_t work = n_intervals > n_tiles + 1 ? n_intervals : n_tiles + 1;
    const int blocks = (int)((work + 255) / 256 > 148 * 16 ? 148 * 16 : (work + 255) / 256);
    plan_from_intervals_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>(
        ranks_bev, interval_starts, interval_lengths, n_points, n_intervals, n_intervals_dev, n_vox,
        (int64_t)B * n_vox, tps, (int)n_tiles, pv.hdr, pv.tile_off, pv.pos2iv);
    FO_LAUNCH_CHECK("plan_from_intervals_kernel");
    return FO_OK;
}

// ---- shared scratch layout of one bucket sort: [cnt | scan_state | counter] zeroed, then the rest
namespace {
struct SortScratch {
    int32_t *cnt;
    unsigned long long *state;
    int32_t *counter;
    size_t zero_bytes;
    char *rest;
};
size_t sort_zero_bytes(int64_t n_buckets) {
    const int64_t n_scan_tiles = (n_buckets + kScanTile - 1) / kScanTile;
    return (size_t)(align_up(n_buckets * 4, 256) + align_up(n_scan_tiles * 8, 256) + 256);
}
SortScratch sort_scratch_view(void *base, int64_t n_buckets) {
    const int64_t n_scan_tiles = (n_buckets + kScanTile - 1) / kScanTile;
    SortScratch s;
    char *p = (char *)base;
    s.cnt = (int32_t *)p;                       p += align_up(n_buckets * 4, 256);
    s.state = (unsigned long long *)p;          p += align_up(n_scan_tiles * 8, 256);
    s.counter = (int32_t *)p;                   p += 256;
    s.zero_bytes = (size_t)(p - (char *)base);
    s.rest = p;
    return s;
}
int grid_for(int64_t work_items, int per_block) {
    int64_t b = (work_items + per_block - 1) / per_block;
    const int64_t cap = 148 * 8;               // persistent-ish: 8 CTAs of 256 threads per SM
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}
}  // namespace

extern "C" size_t fo_rank_prepare_scratch_bytes(int64_t n_points_total, int64_t n_voxels_total) {
    if (n_points_total < 0 || n_voxels_total < 0) return 0;
    const int64_t cap_iv = n_points_total < n_voxels_total ? n_points_total : n_voxels_total;
    return sort_zero_bytes(n_voxels_total) + (size_t)(2 * align_up(n_points_total * 4, 256) +
                                                      align_up(cap_iv * 4, 256));
}

extern "C" int fo_rank_prepare(fo_stream_t stream_, const float *coor, int32_t B, int32_t N, int32_t D, int32_t H,
                               int32_t W, const float lower_bound[3], const float interval[3], int32_t X, int32_t Y,
                               int32_t Z, int32_t *ranks_bev, int32_t *ranks_depth, int32_t *ranks_feat,
                               int32_t *interval_starts, int32_t *interval_lengths, int32_t *counts_dev,
                               void *fwd_plan, size_t fwd_plan_bytes, void *scratch, size_t scratch_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FO_CHECK_ARG(B >= 1 && N >= 1 && D >= 1 && H >= 1 && W >= 1, "non-positive frustum dims");
    FO_CHECK_ARG(X >= 1 && Y >= 1 && Z >= 1, "non-positive grid dims");
    FO_CHECK_ARG(coor && lower_bound && interval, "NULL geometry input");
    FO_CHECK_ARG(ranks_bev && ranks_depth && ranks_feat && interval_starts && interval_lengths && counts_dev,
                 "NULL output array");
    FO_CHECK_ARG(scratch != nullptr, "scratch is NULL");
    FO_CHECK_ARG(((uintptr_t)coor & 15) == 0, "coor must be 16-byte aligned");
    const int64_t pps = (int64_t)N * D * H * W;
    const int64_t P = pps * B;
    const int64_t n_vox = (int64_t)X * Y * Z;
    const int64_t NV = n_vox * B;
    FO_CHECK_ARG(P < INT_MAX && NV < INT_MAX, "point count %lld or voxel count %lld does not fit int32 ranks",
                 (long long)P, (long long)NV);
    const size_t need = fo_rank_prepare_scratch_bytes(P, NV);
    if (scratch_bytes < need)
        return set_error(FO_ERR_SCRATCH, "rank scratch is %zu bytes, need %zu", scratch_bytes, need);

    SortScratch ss = sort_scratch_view(scratch, NV);
    int32_t *key = (int32_t *)ss.rest;
    int32_t *slot = (int32_t *)(ss.rest + align_up(P * 4, 256));
    int32_t *iv_bucket = (int32_t *)(ss.rest + 2 * align_up(P * 4, 256));

    FwdPlanView pv{nullptr, nullptr, nullptr};
    int64_t n_tiles = 0; int tps = 0;
    if (fwd_plan) {
        if (int rc = check_fwd_plan(fwd_plan_bytes, B, n_vox, P, &n_tiles, &tps)) return rc;
        pv = fwd_plan_view(fwd_plan, n_tiles);
    }
    FO_CUDA(cudaMemsetAsync(scratch, 0, ss.zero_bytes, stream));
    FO_CUDA(cudaMemsetAsync(counts_dev, 0, 4 * sizeof(int32_t), stream));

    VoxArgs va;
    va.coor = coor; va.n_points = P; va.points_per_sample = pps;
    va.lbx = lower_bound[0]; va.lby = lower_bound[1]; va.lbz = lower_bound[2];
    va.ivx = interval[0]; va.ivy = interval[1]; va.ivz = interval[2];
    va.X = X; va.Y = Y; va.Z = Z;
    va.cnt = ss.cnt; va.key = key; va.slot = slot;
    va.hdr = pv.hdr; va.n_tiles = (int)n_tiles; va.tiles_per_sample = tps;
    voxelize_count_kernel<<<grid_for((P + 3) / 4, 256), 256, 0, stream>>>(va);
    FO_LAUNCH_CHECK("voxelize_count_kernel");

    ScanArgs sa;
    sa.cnt = ss.cnt; sa.n_buckets = NV;
    sa.iv_starts = interval_starts; sa.iv_lengths = interval_lengths; sa.iv_bucket = iv_bucket;
    sa.totals = counts_dev;
    sa.tile_off = pv.tile_off; sa.vox_per_sample = n_vox; sa.tiles_per_sample = tps; sa.n_tiles = (int)n_tiles;
    sa.fwd_hdr = pv.hdr; sa.bwd_hdr = nullptr;
    sa.state = ss.state; sa.tile_counter = ss.counter;
    const int scan_blocks = (int)((NV + kScanTile - 1) / kScanTile);
    scan_buckets_kernel<<<scan_blocks, kScanThreads, 0, stream>>>(sa);
    FO_LAUNCH_CHECK("scan_buckets_kernel");

    place_kernel<<<grid_for(P, 256), 256, 0, stream>>>(key, slot, ss.cnt, P, nullptr, NV, ranks_depth);
    FO_LAUNCH_CHECK("place_kernel");

    OrderArgs oa;
    oa.sorted = ranks_depth; oa.iv_starts = interval_starts; oa.iv_lengths = interval_lengths;
    oa.iv_bucket = iv_bucket; oa.n_intervals = counts_dev + 1;
    oa.ranks_feat = ranks_feat; oa.ranks_bev = ranks_bev; oa.pos2iv = pv.pos2iv;
    oa.dhw = D * H * W; oa.hw = H * W;
    const int64_t cap_iv = P < NV ? P : NV;
    if (pv.pos2iv)
        order_segments_kernel<true><<<grid_for(cap_iv, kSortThreads), kSortThreads, 0, stream>>>(oa);
    else
        return set_error(FO_ERR_INVALID_ARG, "fo_rank_prepare needs a forward plan buffer (fwd_plan is NULL)");
    FO_LAUNCH_CHECK("order_segments_kernel<fwd>");
    return FO_OK;
}

// ---- backward plan: [header(64) pad to 256 | bwd_pos[n_pts] | starts[n_rows] | lengths[n_rows] |
//                      bucket_ids[n_rows] | sort scratch (cnt/state/counter) | slot[n_pts]]
namespace {
struct BwdLayout {
    BwdPlanHeader *hdr;
    int32_t *pos, *starts, *lengths, *ids;
    void *sort_base;
    int32_t *slot;
    size_t total;
};
BwdLayout bwd_layout(void *base, int64_t n_pts, int64_t n_rows) {
    BwdLayout L;
    char *p = (char *)base;
    L.hdr = (BwdPlanHeader *)p;        p += 256;
    L.pos = (int32_t *)p;              p += align_up(n_pts * 4, 256);
    L.starts = (int32_t *)p;           p += align_up(n_rows * 4, 256);
    L.lengths = (int32_t *)p;          p += align_up(n_rows * 4, 256);
    L.ids = (int32_t *)p;              p += align_up(n_rows * 4, 256);
    L.sort_base = p;                   p += sort_zero_bytes(n_rows);
    L.slot = (int32_t *)p;             p += align_up(n_pts * 4, 256);
    L.total = (size_t)(p - (char *)base);
    return L;
}
}  // namespace

extern "C" size_t fo_bwd_plan_bytes(int64_t n_points_capacity, int64_t n_feat_rows) {
    if (n_points_capacity < 0 || n_feat_rows < 0) return 0;
    return bwd_layout(nullptr, n_points_capacity, n_feat_rows).total;
}

__global__ void init_bwd_header_kernel(BwdPlanHeader *hdr) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { hdr->n_bwd_intervals = 0; hdr->n_points = 0; }
}

extern "C" int fo_bwd_plan_build(fo_stream_t stream_, const int32_t *ranks_feat, int64_t n_points,
                                 const int32_t *n_points_dev, int64_t n_feat_rows, void *plan, size_t plan_bytes) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FO_CHECK_ARG(plan != nullptr, "plan is NULL");
    FO_CHECK_ARG(n_points >= 0 && n_points < INT_MAX && n_feat_rows >= 1 && n_feat_rows < INT_MAX,
                 "bad sizes n_points=%lld n_feat_rows=%lld", (long long)n_points, (long long)n_feat_rows);
    FO_CHECK_ARG(n_points == 0 || ranks_feat, "ranks_feat is NULL");
    BwdLayout L = bwd_layout(plan, n_points, n_feat_rows);
    if (plan_bytes < L.total)
        return set_error(FO_ERR_SCRATCH, "backward plan buffer is %zu bytes, need %zu", plan_bytes, L.total);
    SortScratch ss = sort_scratch_view(L.sort_base, n_feat_rows);
    FO_CUDA(cudaMemsetAsync(L.sort_base, 0, ss.zero_bytes, stream));
    init_bwd_header_kernel<<<1, 32, 0, stream>>>(L.hdr);
    FO_LAUNCH_CHECK("init_bwd_header_kernel");
    if (n_points == 0) return FO_OK;
    count_keys_kernel<<<grid_for(n_points, 256), 256, 0, stream>>>(ranks_feat, n_points, n_points_dev, n_feat_rows,
                                                                   ss.cnt, L.slot);
    FO_LAUNCH_CHECK("count_keys_kernel");
    ScanArgs sa;
    sa.cnt = ss.cnt; sa.n_buckets = n_feat_rows;
    sa.iv_starts = L.starts; sa.iv_lengths = L.lengths; sa.iv_bucket = L.ids;
    sa.totals = &L.hdr->reserved[0];
    sa.tile_off = nullptr; sa.vox_per_sample = 1; sa.tiles_per_sample = 0; sa.n_tiles = 0;
    sa.fwd_hdr = nullptr; sa.bwd_hdr = L.hdr;
    sa.state = ss.state; sa.tile_counter = ss.counter;
    const int scan_blocks = (int)((n_feat_rows + kScanTile - 1) / kScanTile);
    scan_buckets_kernel<<<scan_blocks, kScanThreads, 0, stream>>>(sa);
    FO_LAUNCH_CHECK("scan_buckets_kernel");
    place_kernel<<<grid_for(n_points, 256), 256, 0, stream>>>(ranks_feat, L.slot, ss.cnt, n_points, n_points_dev,
                                                              n_feat_rows, L.pos);
    FO_LAUNCH_CHECK("place_kernel");
    OrderArgs oa;
    oa.sorted = L.pos; oa.iv_starts = L.starts; oa.iv_lengths = L.lengths; oa.iv_bucket = nullptr;
    oa.n_intervals = &L.hdr->n_bwd_intervals;
    oa.ranks_feat = nullptr; oa.ranks_bev = nullptr; oa.pos2iv = nullptr; oa.dhw = 1; oa.hw = 1;
    order_segments_kernel<false><<<grid_for(n_feat_rows, kSortThreads), kSortThreads, 0, stream>>>(oa);
    FO_LAUNCH_CHECK("order_segments_kernel<bwd>");
    return FO_OK;
}
