// fusionocc_b200 — stable one-digit radix ("bucket") sort building blocks for the integer rank
// pipeline (sm_100a).
//
// The reference sorts frustum points by voxel rank with a library sort and then re-discovers the runs
// with eager torch ops (projects/FusionOcc/fusionocc/necks/view_transformer.py:262-278), and does the
// same again by ranks_feat in every backward (mmdet3d/ops/bev_pool_v2/bev_pool.py:47-57).  Keys here
// are dense small integers (voxel id < B*Z*Y*X, feature row < B*N*H*W) with ~1.5 points per occupied
// bucket, so the whole sort is ONE radix digit: count -> scan -> place -> order-within-bucket.  The
// bucket counter array (2.56 MB per sample at 200x200x16) lives in B200's 126 MB L2.
//
//   count    slot[i] = atomicAdd(&cnt[key_i], 1)                  (arrival order, arbitrary)
//   scan     single-pass decoupled look-back exclusive scan of cnt, fused with run extraction:
//            emits interval_starts / interval_lengths / bucket ids for non-empty buckets, the totals,
//            and (forward flavour) the per-tile first-interval table of the forward plan
//   place    sorted[ cnt[key_i] + slot[i] ] = i
//   order    every bucket's segment is sorted ascending => the result is exactly the STABLE sort
//            (ties in ascending original index), independent of the atomics' arrival order
#pragma once

#include "common.cuh"

namespace fo {

constexpr int kScanThreads = 256;
constexpr int kScanItems   = 16;                          // buckets per thread (4 x int4)
constexpr int kScanTile    = kScanThreads * kScanItems;   // 4096 buckets per CTA

// Decoupled look-back descriptor: [63:62] status, [61:31] point sum, [30:0] non-empty-bucket count.
constexpr unsigned long long kStEmpty = 0ull, kStAgg = 1ull, kStPrefix = 2ull;
__device__ __forceinline__ unsigned long long desc_pack(unsigned long long st, unsigned pts, unsigned ne) {
    return (st << 62) | ((unsigned long long)(pts & 0x7fffffffu) << 31) | (unsigned long long)(ne & 0x7fffffffu);
}
__device__ __forceinline__ unsigned desc_status(unsigned long long d) { return (unsigned)(d >> 62); }
__device__ __forceinline__ unsigned desc_pts(unsigned long long d) { return (unsigned)((d >> 31) & 0x7fffffffu); }
__device__ __forceinline__ unsigned desc_ne(unsigned long long d) { return (unsigned)(d & 0x7fffffffu); }

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct ScanArgs {
    int32_t *cnt;            // in: per-bucket counts; out: exclusive point offsets (in place)
    int64_t  n_buckets;
    int32_t *iv_starts;      // compact outputs, one entry per non-empty bucket
    int32_t *iv_lengths;
    int32_t *iv_bucket;
    int32_t *totals;         // totals[0] = number of points, totals[1] = number of non-empty buckets
    // forward flavour: first-interval table per output tile (nullptr to skip)
    int32_t *tile_off;
    int64_t  vox_per_sample;
    int32_t  tiles_per_sample;
    int32_t  n_tiles;
    FwdPlanHeader *fwd_hdr;  // may be nullptr
    BwdPlanHeader *bwd_hdr;  // may be nullptr
    unsigned long long *state;   // [ceil(n_buckets / kScanTile)] zero-initialised
    int32_t *tile_counter;       // zero-initialised
};

// One CTA = kScanTile consecutive buckets.  Tile ids are handed out by an atomic counter so that every
// predecessor of a running tile has already started (forward-progress guarantee of the look-back).
__global__ void __launch_bounds__(kScanThreads) scan_buckets_kernel(ScanArgs a) {
    __shared__ int s_tile;
    __shared__ unsigned long long s_warp[kScanThreads / 32];
    __shared__ unsigned long long s_prefix;     // exclusive prefix of this tile: (pts << 32) | ne
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(a.tile_counter, 1);
    __syncthreads();
    const int tile = s_tile;
    const int64_t base = (int64_t)tile * kScanTile + (int64_t)tid * kScanItems;

    int c[kScanItems];
    if (base + kScanItems <= a.n_buckets) {
        const int4 *src = reinterpret_cast<const int4 *>(a.cnt + base);
#pragma unroll
        for (int j = 0; j < kScanItems / 4; ++j) {
            int4 v = src[j];
            c[4 * j] = v.x; c[4 * j + 1] = v.y; c[4 * j + 2] = v.z; c[4 * j + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) c[j] = (base + j < a.n_buckets) ? a.cnt[base + j] : 0;
    }
    unsigned long long mine = 0;                // (pts << 32) | ne
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
        mine += ((unsigned long long)(unsigned)c[j] << 32) | (c[j] > 0 ? 1ull : 0ull);

    // block-wide exclusive scan of `mine`
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long warp_off = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        unsigned long long t = s_warp[w];
        if (w < warp) warp_off += t;
        block_total += t;
    }
    const unsigned blk_pts = (unsigned)(block_total >> 32), blk_ne = (unsigned)(block_total & 0xffffffffu);

    if (warp == 0) {
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_u64(a.state, desc_pack(kStPrefix, blk_pts, blk_ne));
        } else {
            if (lane == 0) st_relaxed_u64(a.state + tile, desc_pack(kStAgg, blk_pts, blk_ne));
            int look = tile - 1;
            while (true) {
                const int idx = look - lane;
                unsigned long long d = desc_pack(kStPrefix, 0, 0);      // virtual tile -1: empty inclusive prefix
                if (idx >= 0) {
                    do { d = ld_relaxed_u64(a.state + idx); } while (desc_status(d) == kStEmpty);
                }
                const unsigned pmask = __ballot_sync(0xffffffffu, desc_status(d) == kStPrefix);
                const int first = pmask ? (__ffs(pmask) - 1) : 31;
                unsigned long long contrib = (lane <= first) ? (((unsigned long long)desc_pts(d) << 32) | desc_ne(d)) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                excl += contrib;
                if (pmask) break;
                look -= 32;
            }
            if (lane == 0)
                st_relaxed_u64(a.state + tile, desc_pack(kStPrefix, (unsigned)(excl >> 32) + blk_pts,
                                                         (unsigned)(excl & 0xffffffffu) + blk_ne));
        }
        if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    unsigned long long run = s_prefix + warp_off + (incl - mine);
    unsigned pts = (unsigned)(run >> 32), ne = (unsigned)(run & 0xffffffffu);

    // position of this thread's first bucket inside its sample (for the tile table)
    int64_t vin = 0, sample = 0;
    const bool want_tiles = a.tile_off != nullptr;
    if (want_tiles) { sample = base / a.vox_per_sample; vin = base - sample * a.vox_per_sample; }

    int o[kScanItems];
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        const int64_t v = base + j;
        o[j] = (int)pts;
        if (v < a.n_buckets) {
            if (want_tiles) {
                if ((vin & (kTile - 1)) == 0)
                    a.tile_off[sample * a.tiles_per_sample + (vin >> 7)] = (int)ne;
                if (++vin == a.vox_per_sample) { vin = 0; ++sample; }
            }
            if (c[j] > 0) {
                a.iv_starts[ne] = (int)pts;
                a.iv_lengths[ne] = c[j];
                if (a.iv_bucket) a.iv_bucket[ne] = (int)v;
                pts += (unsigned)c[j];
                ++ne;
            }
        }
    }
    static_assert(kTile == 128, "tile shift above assumes 128");
    if (base + kScanItems <= a.n_buckets) {
        int4 *dst = reinterpret_cast<int4 *>(a.cnt + base);
#pragma unroll
        for (int j = 0; j < kScanItems / 4; ++j) dst[j] = make_int4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < kScanItems; ++j)
            if (base + j < a.n_buckets) a.cnt[base + j] = o[j];
    }
    // the thread that owns the last bucket publishes the totals
    if (base <= a.n_buckets - 1 && a.n_buckets - 1 < base + kScanItems) {
        a.totals[0] = (int)pts;
        a.totals[1] = (int)ne;
        if (want_tiles) a.tile_off[a.n_tiles] = (int)ne;
        if (a.fwd_hdr) a.fwd_hdr->n_intervals = (int)ne;
        if (a.bwd_hdr) { a.bwd_hdr->n_bwd_intervals = (int)ne; a.bwd_hdr->n_points = (int)pts; }
    }
}

// ----------------------------------------------------------------------------------------------
// In-segment ordering.  One lane per interval for the short ones (len <= 4: sorting network in
// registers); longer intervals are handled one at a time by the whole warp: bitonic in registers for
// len <= 32, in shared memory for len <= kSortSmem, in global memory (single warp, any length) above.
// ----------------------------------------------------------------------------------------------
constexpr int kSortThreads = 128;
constexpr int kSortSmem    = 1024;   // ints of shared memory per warp

__device__ __forceinline__ void cswap(int &a, int &b) {
    const int lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}

__device__ inline void warp_sort_segment(int32_t *seg, int len, int *smem /* kSortSmem ints, per warp */, int lane) {
    if (len <= 32) {
        int v = (lane < len) ? seg[lane] : INT_MAX;
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const int other = __shfl_xor_sync(0xffffffffu, v, j);
                const bool up = ((lane & k) == 0);
                const bool lower = ((lane & j) == 0);
                v = (lower == up) ? min(v, other) : max(v, other);
            }
        }
        if (lane < len) seg[lane] = v;
        return;
    }
    int n2 = 64;
    while (n2 < len) n2 <<= 1;
    if (n2 <= kSortSmem) {
        for (int i = lane; i < n2; i += 32) smem[i] = (i < len) ? seg[i] : INT_MAX;
        __syncwarp();
        for (int k = 2; k <= n2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = lane; i < n2; i += 32) {
                    const int l = i ^ j;
                    if (l > i) {
                        const int a = smem[i], b = smem[l];
                        const bool up = ((i & k) == 0);
                        if ((a > b) == up) { smem[i] = b; smem[l] = a; }
                    }
                }
                __syncwarp();
            }
        }
        for (int i = lane; i < len; i += 32) seg[i] = smem[i];
        __syncwarp();
        return;
    }
    // very long segment (degenerate geometry): sorting network directly in global memory.  This is the
    // "always ascending" bitonic form — the first stage of every merge pairs i with its mirror
    // i ^ (k-1), later stages with i ^ j — so every compare-exchange moves the minimum to the lower
    // index and the virtual +inf padding beyond `len` never has to move: no out-of-range access.
    // Compare-exchange pairs are disjoint within a stage.
    for (long long k = 2; k <= n2; k <<= 1) {
        for (long long j = k >> 1; j > 0; j >>= 1) {
            const long long x = (j == (k >> 1)) ? (k - 1) : j;
            for (long long i = lane; i < len; i += 32) {
                const long long l = i ^ x;
                if (l > i && l < len) {
                    const int a = seg[i], b = seg[l];
                    if (a > b) { seg[i] = b; seg[l] = a; }
                }
            }
            __threadfence_block();
            __syncwarp();
        }
    }
}

struct OrderArgs {
    int32_t *sorted;            // in/out: per-bucket segments of original indices (becomes ranks_depth)
    const int32_t *iv_starts;
    const int32_t *iv_lengths;
    const int32_t *iv_bucket;   // forward flavour only (= iv_vox of the plan)
    const int32_t *n_intervals; // device count
    // forward flavour outputs (nullptr for the backward plan)
    int32_t *ranks_feat;
    int32_t *ranks_bev;
    int32_t *pos2iv;
    int32_t *pt2pos;            // frustum point -> sorted position (its -1 entries were written by voxelise)
    int32_t dhw, hw;            // D*H*W and H*W: ranks_feat = (p / dhw) * hw + p % hw  (view_transformer.py:239-244)
};

// 19-comparator optimal sorting network for 8 keys (ascending)
__device__ __forceinline__ void sort8(int (&v)[8]) {
    cswap(v[0], v[1]); cswap(v[2], v[3]); cswap(v[4], v[5]); cswap(v[6], v[7]);
    cswap(v[0], v[2]); cswap(v[1], v[3]); cswap(v[4], v[6]); cswap(v[5], v[7]);
    cswap(v[1], v[2]); cswap(v[5], v[6]); cswap(v[0], v[4]); cswap(v[3], v[7]);
    cswap(v[1], v[5]); cswap(v[2], v[6]);
    cswap(v[1], v[4]); cswap(v[3], v[6]);
    cswap(v[2], v[4]); cswap(v[3], v[5]);
    cswap(v[3], v[4]);
}

constexpr int kLaneSortMax = 8;   // intervals up to this length are ordered by one lane in registers

template <bool kForward>
__global__ void __launch_bounds__(kSortThreads) order_segments_kernel(OrderArgs a) {
    __shared__ int s_sort[kSortThreads / 32][kSortSmem];
    const int n = *a.n_intervals;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = (blockIdx.x * kSortThreads + warp * 32); base < n; base += gridDim.x * kSortThreads) {
        const int k = base + lane;
        int s = 0, len = 0, bucket = 0;
        if (k < n) {
            s = a.iv_starts[k];
            len = a.iv_lengths[k];
            if (kForward) bucket = a.iv_bucket[k];
        }
        if (len > 0 && len <= kLaneSortMax) {
            int v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (j < len) ? a.sorted[s + j] : INT_MAX;
            if (len > 1) sort8(v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < len) {
                    if (len > 1) a.sorted[s + j] = v[j];
                    if (kForward) {
                        a.ranks_feat[s + j] = (v[j] / a.dhw) * a.hw + (v[j] % a.hw);
                        a.ranks_bev[s + j] = bucket;
                        a.pos2iv[s + j] = k;
                        a.pt2pos[v[j]] = s + j;
                    }
                }
            }
        }
        unsigned longmask = __ballot_sync(0xffffffffu, len > kLaneSortMax);
        while (longmask) {
            const int src = __ffs(longmask) - 1;
            longmask &= longmask - 1;
            const int ls = __shfl_sync(0xffffffffu, s, src);
            const int ll = __shfl_sync(0xffffffffu, len, src);
            const int lb = __shfl_sync(0xffffffffu, bucket, src);
            warp_sort_segment(a.sorted + ls, ll, s_sort[warp], lane);
            __syncwarp();
            if (kForward) {
                for (int j = lane; j < ll; j += 32) {
                    const int p = a.sorted[ls + j];
                    a.ranks_feat[ls + j] = (p / a.dhw) * a.hw + (p % a.hw);
                    a.ranks_bev[ls + j] = lb;
                    a.pos2iv[ls + j] = base + src;
                    a.pt2pos[p] = ls + j;
                }
            }
        }
    }
}

}  // namespace fo
