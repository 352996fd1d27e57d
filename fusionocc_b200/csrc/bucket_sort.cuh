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
//   scan     one-pass exclusive scan of cnt (tile aggregates come from the count pass, so CTAs are
//            independent), fused with run extraction: emits interval_starts / interval_lengths / bucket
//            ids for non-empty buckets, the totals, and (forward flavour) the per-tile first-interval
//            table of the forward plan
//   place    sorted[ cnt[key_i] + slot[i] ] = i
//   order    every bucket's segment is sorted ascending => the result is exactly the STABLE sort
//            (ties in ascending original index), independent of the atomics' arrival order
#pragma once

#include "common.cuh"

namespace fo {

#ifndef FO_SCAN_THREADS
#define FO_SCAN_THREADS 128
#endif
#ifndef FO_SCAN_ITEMS
#define FO_SCAN_ITEMS 16
#endif
constexpr int kScanThreads = FO_SCAN_THREADS;
constexpr int kScanItems   = FO_SCAN_ITEMS;               // buckets per thread (multiple of 4: int4 accesses)
constexpr int kScanTile    = kScanThreads * kScanItems;   // 2048 buckets per CTA

// Per-scan-tile aggregates (points << 32 | non-empty buckets) come from a separate reduce pass over the
// counter array (coalesced int4 reads, the array is L2-resident right after the count pass), so the scan
// kernel needs no inter-CTA communication at all — no decoupled look-back, no spinning: CTA t sums
// agg[0..t).  (Tried and rejected, profiles/r01: a decoupled look-back scan, 41 us, and aggregating with
// one atomic per element inside the count pass, 157 us of same-address contention.)
__global__ void __launch_bounds__(kScanThreads) tile_reduce_kernel(const int32_t *__restrict__ cnt, int64_t n_buckets,
                                                                   unsigned long long *__restrict__ agg,
                                                                   unsigned long long *__restrict__ agg_group) {
    __shared__ unsigned long long s_red[kScanThreads / 32];
    pdl_wait();
    pdl_launch();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    unsigned long long mine = 0;
#pragma unroll
    for (int j = 0; j < kScanItems / 4; ++j) {
        const int64_t i = base + ((int64_t)j * kScanThreads + tid) * 4;
        int4 v = make_int4(0, 0, 0, 0);
        if (i + 4 <= n_buckets) v = *reinterpret_cast<const int4 *>(cnt + i);
        else {
            if (i < n_buckets) v.x = cnt[i];
            if (i + 1 < n_buckets) v.y = cnt[i + 1];
            if (i + 2 < n_buckets) v.z = cnt[i + 2];
        }
        mine += ((unsigned long long)(unsigned)(v.x + v.y + v.z + v.w) << 32) |
                (unsigned long long)((v.x > 0) + (v.y > 0) + (v.z > 0) + (v.w > 0));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0) s_red[warp] = mine;
    __syncthreads();
    if (tid == 0) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; ++w) t += s_red[w];
        agg[blockIdx.x] = t;
        atomicAdd(agg_group + blockIdx.x / kScanGroup, t);   // ~kScanGroup adds per address: no contention
    }
}

struct ScanArgs {
    int32_t *cnt;            // in: per-bucket counts; out: exclusive point offsets (in place)
    int64_t  n_buckets;
    int32_t *iv_starts;      // compact outputs, one entry per non-empty bucket
    int32_t *iv_lengths;
    int32_t *iv_bucket;
    int32_t *bucket2iv;      // optional [n_buckets]: interval id of every bucket (forward flavour: the plan's vox2iv)
    int32_t *totals;         // totals[0] = number of points, totals[1] = number of non-empty buckets
    // forward flavour: first interval / first point of every 32-voxel sub-tile (nullptr to skip)
    int32_t *sub_iv;
    int32_t *sub_pt;
    uint32_t *sub_mask;      // optional: occupancy mask of every sub-tile (written by scan_buckets2_kernel only)
    int64_t  vox_per_sample;
    int32_t  subs_per_sample;
    int32_t  n_subs;
    FwdPlanHeader *fwd_hdr;  // may be nullptr
    BwdPlanHeader *bwd_hdr;  // may be nullptr
    const unsigned long long *agg;   // [ceil(n_buckets / kScanTile)] from the reduce pass
    const unsigned long long *agg_group;   // sums over groups of kScanGroup tiles
};

// One CTA = kScanTile consecutive buckets.
__global__ void __launch_bounds__(kScanThreads) scan_buckets_kernel(ScanArgs a) {
    __shared__ unsigned long long s_warp[kScanThreads / 32];
    __shared__ unsigned long long s_red[kScanThreads / 32];
    pdl_wait();
    pdl_launch();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int64_t base = (int64_t)tile * kScanTile + (int64_t)tid * kScanItems;

    // exclusive prefix of this tile = sum of the aggregates of all earlier tiles, two-level: whole groups of
    // kScanGroup tiles, then the earlier tiles of this tile's own group (a flat loop over all earlier tiles
    // was a 20-deep load chain and 25 MB of L2 reads at batch 8)
    unsigned long long pre = 0;
    const int g0 = tile / kScanGroup;
    for (int g = tid; g < g0; g += kScanThreads) pre += a.agg_group[g];
    if (tid < tile - g0 * kScanGroup) pre += a.agg[g0 * kScanGroup + tid];

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

    // block-wide exclusive scan of `mine`, block-wide sum of `pre`
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
    if (lane == 31) s_warp[warp] = incl;
    if (lane == 0) s_red[warp] = pre;
    __syncthreads();
    unsigned long long warp_off = 0, tile_prefix = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        if (w < warp) warp_off += s_warp[w];
        tile_prefix += s_red[w];
    }
    unsigned long long run = tile_prefix + warp_off + (incl - mine);
    unsigned pts = (unsigned)(run >> 32), ne = (unsigned)(run & 0xffffffffu);

    // position of this thread's first bucket inside its sample (for the tile table)
    int64_t vin = 0, sample = 0;
    const bool want_tiles = a.sub_iv != nullptr;
    if (want_tiles) { sample = base / a.vox_per_sample; vin = base - sample * a.vox_per_sample; }

    int o[kScanItems], iv[kScanItems];
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        const int64_t v = base + j;
        o[j] = (int)pts;
        iv[j] = (int)ne;
        if (v < a.n_buckets) {
            if (want_tiles) {
                if ((vin & (kSub - 1)) == 0) {
                    const int64_t u = sample * a.subs_per_sample + (vin >> kSubShift);
                    a.sub_iv[u] = (int)ne;
                    a.sub_pt[u] = (int)pts;
                }
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
    if (base + kScanItems <= a.n_buckets) {
        int4 *dst = reinterpret_cast<int4 *>(a.cnt + base);
#pragma unroll
        for (int j = 0; j < kScanItems / 4; ++j) dst[j] = make_int4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        if (a.bucket2iv) {
            int4 *d2 = reinterpret_cast<int4 *>(a.bucket2iv + base);
#pragma unroll
            for (int j = 0; j < kScanItems / 4; ++j)
                d2[j] = make_int4(iv[4 * j], iv[4 * j + 1], iv[4 * j + 2], iv[4 * j + 3]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < kScanItems; ++j)
            if (base + j < a.n_buckets) {
                a.cnt[base + j] = o[j];
                if (a.bucket2iv) a.bucket2iv[base + j] = iv[j];
            }
    }
    // the thread that owns the last bucket publishes the totals
    if (base <= a.n_buckets - 1 && a.n_buckets - 1 < base + kScanItems) {
        a.totals[0] = (int)pts;
        a.totals[1] = (int)ne;
        if (want_tiles) { a.sub_iv[a.n_subs] = (int)ne; a.sub_pt[a.n_subs] = (int)pts; }
        if (a.fwd_hdr) a.fwd_hdr->n_intervals = (int)ne;
        if (a.bwd_hdr) { a.bwd_hdr->n_bwd_intervals = (int)ne; a.bwd_hdr->n_points = (int)pts; }
    }
}

// ----------------------------------------------------------------------------------------------
// In-segment ordering.  One lane per interval for the short ones (len <= 8: sorting network in
// registers); longer intervals are handled one at a time by a whole warp in registers (bitonic networks over
// 1 / 2 / 4 registers per lane for len <= 32 / 64 / 128) or by a whole CTA in shared memory (len <= kSortSmemCta);
// beyond that (degenerate geometry) a single warp runs the network in global memory.
// ----------------------------------------------------------------------------------------------
constexpr int kSortThreads = 128;
constexpr int kSortSmemCta = 4096;   // ints of shared memory per CTA: the longest interval sorted on chip

__device__ __forceinline__ void cswap(int &a, int &b) {
    const int lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}

__device__ inline void warp_sort_segment(int32_t *seg, int len, int * /*unused*/, int lane) {
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
    if (len <= 64) {                                   // bitonic network over 2 registers per lane: 21 steps x 2
        int v[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) v[r] = (lane + 32 * r < len) ? seg[lane + 32 * r] : INT_MAX;
        bitonic_sort_regs<2>(v, lane);
#pragma unroll
        for (int r = 0; r < 2; ++r)
            if (lane + 32 * r < len) seg[lane + 32 * r] = v[r];
        return;
    }
    if (len <= 128) {                                  // ... over 4 registers per lane: 28 steps x 4 (rank-by-counting
        int v[4];                                      // needed 128 broadcasts x 4 compares)
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r] = (lane + 32 * r < len) ? seg[lane + 32 * r] : INT_MAX;
        bitonic_sort_regs<4>(v, lane);
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (lane + 32 * r < len) seg[lane + 32 * r] = v[r];
        return;
    }
    int n2 = 256;
    while (n2 < len) n2 <<= 1;
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
    FastDiv dhw, hw;            // D*H*W and H*W: ranks_feat = (p / dhw) * hw + p % hw  (view_transformer.py:239-244)
    // intervals longer than kLaneSortMax are queued here by the short pass and ordered, one warp each,
    // by the long pass — dense near-ego voxels are consecutive in voxel order, so without the queue a
    // few warps would inherit dozens of long intervals each (measured: 103 us -> tail-bound)
    // forward plans: the order pass also lists the dense sub-tiles (nullptr to skip)
    const int32_t *sub_pt;
    int32_t n_subs;
    int32_t *heavy_list;
    int32_t heavy_pts;          // dense sub-tile threshold (heavy_threshold(B, V))
    int32_t *heavy_n;           // zero-initialised
    int32_t *long_list;
    int32_t *long_count;        // zero-initialised; [0] = intervals of kLaneSortMax+1 .. kWarpSortMax points,
                                // [1] = longer ones, queued from the END of long_list (long_cap - 1 downwards)
    int32_t long_cap;
};

// feature row of frustum point p: (p / dhw) * hw + p % hw
__device__ __forceinline__ int feat_row_of(int p, const FastDiv &dhw, const FastDiv &hw) {
    const uint32_t n = (uint32_t)p;
    return (int)(fastdiv(n, dhw) * hw.d + (n - fastdiv(n, hw) * hw.d));
}

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

// Batcher's odd-even merge sort for 16 keys (63 comparators; verified exhaustively with the 0-1 principle)
__device__ __forceinline__ void sort16(int (&v)[16]) {
    cswap(v[0], v[1]); cswap(v[2], v[3]); cswap(v[0], v[2]); cswap(v[1], v[3]); cswap(v[1], v[2]); cswap(v[4], v[5]);
    cswap(v[6], v[7]); cswap(v[4], v[6]); cswap(v[5], v[7]); cswap(v[5], v[6]); cswap(v[0], v[4]); cswap(v[2], v[6]);
    cswap(v[2], v[4]); cswap(v[1], v[5]); cswap(v[3], v[7]); cswap(v[3], v[5]); cswap(v[1], v[2]); cswap(v[3], v[4]);
    cswap(v[5], v[6]); cswap(v[8], v[9]); cswap(v[10], v[11]); cswap(v[8], v[10]); cswap(v[9], v[11]); cswap(v[9], v[10]);
    cswap(v[12], v[13]); cswap(v[14], v[15]); cswap(v[12], v[14]); cswap(v[13], v[15]); cswap(v[13], v[14]); cswap(v[8], v[12]);
    cswap(v[10], v[14]); cswap(v[10], v[12]); cswap(v[9], v[13]); cswap(v[11], v[15]); cswap(v[11], v[13]); cswap(v[9], v[10]);
    cswap(v[11], v[12]); cswap(v[13], v[14]); cswap(v[0], v[8]); cswap(v[4], v[12]); cswap(v[4], v[8]); cswap(v[2], v[10]);
    cswap(v[6], v[14]); cswap(v[6], v[10]); cswap(v[2], v[4]); cswap(v[6], v[8]); cswap(v[10], v[12]); cswap(v[1], v[9]);
    cswap(v[5], v[13]); cswap(v[5], v[9]); cswap(v[3], v[11]); cswap(v[7], v[15]); cswap(v[7], v[11]); cswap(v[3], v[5]);
    cswap(v[7], v[9]); cswap(v[11], v[13]); cswap(v[1], v[2]); cswap(v[3], v[4]); cswap(v[5], v[6]); cswap(v[7], v[8]);
    cswap(v[9], v[10]); cswap(v[11], v[12]); cswap(v[13], v[14]);
}

constexpr int kLaneSortMax = 8;   // intervals up to this length are ordered by one lane in registers
constexpr int kWarpSortMax = 128; // ... up to this length by one warp in registers; longer ones by a whole CTA

template <bool kForward>
__global__ void __launch_bounds__(256) order_short_kernel(OrderArgs a) {
    pdl_wait();
    pdl_launch();
    const int n = *a.n_intervals;
    const int stride = gridDim.x * blockDim.x;
    if (kForward && a.heavy_list != nullptr)
        for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < a.n_subs; u += stride)
            if (a.sub_pt[u + 1] - a.sub_pt[u] > a.heavy_pts) a.heavy_list[atomicAdd(a.heavy_n, 1)] = u;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const int s = a.iv_starts[k], len = a.iv_lengths[k];
        if (len > kLaneSortMax) {
            if (len > kWarpSortMax) a.long_list[a.long_cap - 1 - atomicAdd(a.long_count + 1, 1)] = k;
            else a.long_list[atomicAdd(a.long_count, 1)] = k;
            continue;
        }
        if (!kForward && len == 1) continue;
        const int bucket = kForward ? a.iv_bucket[k] : 0;
        int v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (j < len) ? a.sorted[s + j] : INT_MAX;
        if (len > 1) sort8(v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < len) {
                if (len > 1) a.sorted[s + j] = v[j];
                if (kForward) {
                    if (a.ranks_feat) a.ranks_feat[s + j] = feat_row_of(v[j], a.dhw, a.hw);
                    a.ranks_bev[s + j] = bucket;
                }
            }
        }
    }
}

// CTA-wide bitonic sort of `len` distinct keys in shared memory (n2 = next power of two <= kSortSmemCta).
__device__ inline void cta_sort_segment(int32_t *seg, int len, int *smem) {
    int n2 = 256;
    while (n2 < len) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += kSortThreads) smem[i] = (i < len) ? seg[i] : INT_MAX;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n2 >> 1); t += kSortThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;      // the pair (i, i + j)
                const int x = smem[i], y = smem[l];
                const bool up = ((i & k) == 0);
                if ((x > y) == up) { smem[i] = y; smem[l] = x; }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < len; i += kSortThreads) seg[i] = smem[i];
    __syncthreads();
}

template <bool kForward>
__global__ void __launch_bounds__(kSortThreads) order_long_kernel(OrderArgs a) {
    __shared__ int s_sort[kSortSmemCta];
    pdl_wait();
    pdl_launch();
    const int n = a.long_count[0], n_big = a.long_count[1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = gridDim.x * (kSortThreads / 32);
    // (1) one warp per interval of up to kWarpSortMax points, all in registers
    for (int w = blockIdx.x * (kSortThreads / 32) + warp; w < n; w += nwarps) {
        const int k = a.long_list[w];
        const int ls = a.iv_starts[k], ll = a.iv_lengths[k];
        const int lb = kForward ? a.iv_bucket[k] : 0;
        warp_sort_segment(a.sorted + ls, ll, nullptr, lane);
        __syncwarp();
        if (kForward) {
            for (int j = lane; j < ll; j += 32) {
                const int p = a.sorted[ls + j];
                if (a.ranks_feat) a.ranks_feat[ls + j] = feat_row_of(p, a.dhw, a.hw);
                a.ranks_bev[ls + j] = lb;
            }
        }
        __syncwarp();
    }
    if (n_big == 0) return;
    // (2) one CTA per longer interval (a lone warp needs ~30 us for 256 keys: it was this kernel's tail)
    __syncthreads();
    for (int w = blockIdx.x; w < n_big; w += gridDim.x) {
        const int k = a.long_list[a.long_cap - 1 - w];
        const int ls = a.iv_starts[k], ll = a.iv_lengths[k];
        const int lb = kForward ? a.iv_bucket[k] : 0;
        if (ll <= kSortSmemCta) {
            cta_sort_segment(a.sorted + ls, ll, s_sort);
        } else {                                          // degenerate geometry: global-memory network, one warp
            if (warp == 0) warp_sort_segment(a.sorted + ls, ll, nullptr, lane);
            __threadfence_block();
            __syncthreads();
        }
        if (kForward) {
            for (int j = threadIdx.x; j < ll; j += kSortThreads) {
                const int p = a.sorted[ls + j];
                if (a.ranks_feat) a.ranks_feat[ls + j] = feat_row_of(p, a.dhw, a.hw);
                a.ranks_bev[ls + j] = lb;
            }
        }
        __syncthreads();
    }
}

}  // namespace fo
