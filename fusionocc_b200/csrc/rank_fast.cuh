// fusionocc_b200 — forward-flavour rank pipeline, second cut of the scan and order passes (sm_100a).
//
// Same one-digit bucket sort as bucket_sort.cuh (count -> scan -> place -> order within bucket) and the same
// bits out (profiles/r02_summary.md §7):
//
//   * the scan's per-interval outputs (interval_starts / interval_lengths / iv_vox) used to be written from the
//     sequential per-thread walk, one predicated 4-byte store per bucket and array; they are now staged in shared
//     memory in interval order and leave as full lines (33 -> 26 us at batch 8);
//   * the scan classifies the intervals while it has their lengths in registers and queues those of more than
//     kLaneSortMax points (9..128: one warp each; longer: one CTA each), so that ONE order launch does everything:
//     the queued intervals first, then one lane per short interval (the former order_short + order_long pair).
//
// Measured and NOT kept: a placement that writes all three rank arrays at the point's final position, with an
// order pass over only the >= 2-point intervals (82 % of the occupied voxels hold ONE point).  Every scattered
// 4-byte store per point costs a 32-byte sector write: place 16.5 -> 34.4 us, order 29.7 -> 19.5 us — a net loss
// (and 291 -> 352 us at 512x1408, where there are 2.5 points per voxel).  Also not kept: a slot-free count (reductions
// without a return value in the voxelise pass, no 12 MB slot array; the placement draws positions with a returning
// atomicAdd on the bucket offsets): rank precompute 95.7 -> 100.0 us at batch 8, 283.7 -> 291.9 us at 512x1408.  Nor the
// scan's tile aggregates summed inside the voxelise pass (match.any per warp, one 64-bit atomic per distinct tile, no
// tile_reduce launch): step 397.6 vs 397.7 us at batch 8, 939 vs 932 us at 512x1408.
#pragma once

#include "bucket_sort.cuh"

namespace fo {

constexpr int kScan2Threads = 256;
constexpr int kScan2Items   = kScanTile / kScan2Threads;      // 8 buckets per thread
static_assert(kScan2Items == 8 && kScanTile == 2048, "scan2 is written for 2048-bucket tiles, 256 threads x 8");

constexpr int kMidSortMax = 16;  // intervals of kLaneSortMax+1 .. kMidSortMax points: one lane each, 16-key network
struct ListArgs {
    int32_t *long_list;      // intervals of kMidSortMax+1 .. kWarpSortMax points from the front, longer ones from the END
    int32_t *mid_list;       // intervals of kLaneSortMax+1 .. kMidSortMax points
    int32_t *counts;         // zero-initialised: [0] mid, [1] long (front), [2] long (end)
    int32_t long_cap, mid_cap;
    int32_t mid_max;         // kMidSortMax, or kLaneSortMax to switch the mid queue off (FO_RANK_MID=0)
    int2 *ofiv;              // optional [n_buckets]: (point offset, interval id) of every bucket in ONE 8-byte entry — the
                             // placement then fetches both with a single gather (instead of cnt in place + bucket2iv)
};

__global__ void __launch_bounds__(kScan2Threads) scan_buckets2_kernel(ScanArgs a, ListArgs l) {
    __shared__ unsigned long long s_warp[kScan2Threads / 32];
    __shared__ unsigned long long s_red[kScan2Threads / 32];
    __shared__ int4 s_iv[kScanTile];         // (first point, points, voxel, -) of the tile's intervals, in interval order
    __shared__ int s_q[kScanTile / 8];       // tile-local queue of long intervals (> 16 points each, 2048 buckets)
    __shared__ int s_m[kScanTile / 8];       // ... of the 9..16-point intervals
    __shared__ int s_qn, s_qbase, s_mn, s_mbase;
    pdl_wait();
    pdl_launch();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int64_t base = (int64_t)tile * kScanTile + (int64_t)tid * kScan2Items;
    if (tid == 0) { s_qn = 0; s_mn = 0; }

    unsigned long long pre = 0;
    const int g0 = tile / kScanGroup;
    for (int g = tid; g < g0; g += kScan2Threads) pre += a.agg_group[g];
    if (tid < tile - g0 * kScanGroup) pre += a.agg[g0 * kScanGroup + tid];

    int c[kScan2Items];
    if (base + kScan2Items <= a.n_buckets) {
        const int4 *src = reinterpret_cast<const int4 *>(a.cnt + base);
        const int4 v0 = src[0], v1 = src[1];
        c[0] = v0.x; c[1] = v0.y; c[2] = v0.z; c[3] = v0.w; c[4] = v1.x; c[5] = v1.y; c[6] = v1.z; c[7] = v1.w;
    } else {
#pragma unroll
        for (int j = 0; j < kScan2Items; ++j) c[j] = (base + j < a.n_buckets) ? a.cnt[base + j] : 0;
    }
    unsigned long long mine = 0;                // (pts << 32) | ne
#pragma unroll
    for (int j = 0; j < kScan2Items; ++j)
        mine += ((unsigned long long)(unsigned)c[j] << 32) | (c[j] > 0 ? 1ull : 0ull);
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
    unsigned long long warp_off = 0, tile_prefix = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kScan2Threads / 32; ++w) {
        if (w < warp) warp_off += s_warp[w];
        tile_prefix += s_red[w];
        tile_total += s_warp[w];
    }
    const unsigned long long run = tile_prefix + warp_off + (incl - mine);
    unsigned pts = (unsigned)(run >> 32), ne = (unsigned)(run & 0xffffffffu);
    const int ne0 = (int)(tile_prefix & 0xffffffffu);          // first interval of this tile
    const int n_tile = (int)(tile_total & 0xffffffffu);        // intervals of this tile

    const bool want_tiles = a.sub_iv != nullptr;
    int o[kScan2Items], iv[kScan2Items];
    // fast path (every tile but a ragged last one, voxels per sample a multiple of 8): the thread's eight buckets lie in
    // one sample, a sub-tile can only start at its first bucket, all indices fit 32 bits
    const bool fast = (int64_t)(tile + 1) * kScanTile <= a.n_buckets && (a.vox_per_sample & 7) == 0 &&
                      a.n_buckets < (int64_t)INT_MAX;
    if (fast) {
        const int b32 = (int)base;
        // occupancy mask of the sub-tile: eight buckets per thread, four threads per sub-tile (aligned with the threads'
        // 8-bucket groups when voxels per sample is a multiple of 32; 0 = unknown otherwise)
        unsigned m8 = 0;
#pragma unroll
        for (int j = 0; j < kScan2Items; ++j) m8 |= (c[j] > 0 ? 1u : 0u) << j;
        unsigned m32 = m8;
        m32 |= __shfl_down_sync(0xffffffffu, m8, 1) << 8;
        m32 |= __shfl_down_sync(0xffffffffu, m8, 2) << 16;
        m32 |= __shfl_down_sync(0xffffffffu, m8, 3) << 24;
        if (want_tiles) {
            const int sample = b32 / (int)a.vox_per_sample, vin = b32 - sample * (int)a.vox_per_sample;
            if ((vin & (kSub - 1)) == 0) {
                const int u = sample * a.subs_per_sample + (vin >> kSubShift);
                a.sub_iv[u] = (int)ne;
                a.sub_pt[u] = (int)pts;
                if (a.sub_mask) a.sub_mask[u] = ((a.vox_per_sample & 31) == 0) ? m32 : 0u;
            }
        }
        // branch-free walk: one predicated 16-byte store per occupied bucket, the running sums advance unconditionally
        // (an empty bucket adds zero points) — the eight divergent branches per thread were half of this kernel's
        // instructions (ncu source view)
        int li = (int)ne - ne0;
#pragma unroll
        for (int j = 0; j < kScan2Items; ++j) {
            o[j] = (int)pts;
            iv[j] = ne0 + li;
            if (c[j] > 0) s_iv[li] = make_int4((int)pts, c[j], b32 + j, 0);
            pts += (unsigned)c[j];
            li += c[j] > 0 ? 1 : 0;
        }
        ne = (unsigned)(ne0 + li);
    } else {
        int64_t vin = 0, sample = 0;
        if (want_tiles) { sample = base / a.vox_per_sample; vin = base - sample * a.vox_per_sample; }
#pragma unroll
        for (int j = 0; j < kScan2Items; ++j) {
            const int64_t v = base + j;
            o[j] = (int)pts;
            iv[j] = (int)ne;
            if (v < a.n_buckets) {
                if (want_tiles) {
                    if ((vin & (kSub - 1)) == 0) {
                        const int64_t u = sample * a.subs_per_sample + (vin >> kSubShift);
                        a.sub_iv[u] = (int)ne;
                        a.sub_pt[u] = (int)pts;
                        if (a.sub_mask) a.sub_mask[u] = 0u;
                    }
                    if (++vin == a.vox_per_sample) { vin = 0; ++sample; }
                }
                if (c[j] > 0) {
                    const int li = (int)ne - ne0;
                    s_iv[li] = make_int4((int)pts, c[j], (int)v, 0);
                    pts += (unsigned)c[j];
                    ++ne;
                }
            }
        }
    }
    if (l.ofiv != nullptr) {
        if (base + kScan2Items <= a.n_buckets) {
            int4 *dst = reinterpret_cast<int4 *>(l.ofiv + base);
#pragma unroll
            for (int j = 0; j < kScan2Items / 2; ++j) dst[j] = make_int4(o[2 * j], iv[2 * j], o[2 * j + 1], iv[2 * j + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < kScan2Items; ++j)
                if (base + j < a.n_buckets) l.ofiv[base + j] = make_int2(o[j], iv[j]);
        }
    } else if (base + kScan2Items <= a.n_buckets) {
        int4 *dst = reinterpret_cast<int4 *>(a.cnt + base);
        dst[0] = make_int4(o[0], o[1], o[2], o[3]);
        dst[1] = make_int4(o[4], o[5], o[6], o[7]);
        if (a.bucket2iv) {
            int4 *d2 = reinterpret_cast<int4 *>(a.bucket2iv + base);
            d2[0] = make_int4(iv[0], iv[1], iv[2], iv[3]);
            d2[1] = make_int4(iv[4], iv[5], iv[6], iv[7]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < kScan2Items; ++j)
            if (base + j < a.n_buckets) {
                a.cnt[base + j] = o[j];
                if (a.bucket2iv) a.bucket2iv[base + j] = iv[j];
            }
    }
    if (base <= a.n_buckets - 1 && a.n_buckets - 1 < base + kScan2Items) {
        a.totals[0] = (int)pts;
        a.totals[1] = (int)ne;
        if (want_tiles) { a.sub_iv[a.n_subs] = (int)ne; a.sub_pt[a.n_subs] = (int)pts; }
        if (a.fwd_hdr) a.fwd_hdr->n_intervals = (int)ne;
        if (a.bwd_hdr) { a.bwd_hdr->n_bwd_intervals = (int)ne; a.bwd_hdr->n_points = (int)pts; }
    }
    __syncthreads();
    // compact per-interval outputs: full lines; classification of the intervals that need ordering
    for (int i = tid; i < n_tile; i += kScan2Threads) {
        const int4 e = s_iv[i];
        const int len = e.y;
        a.iv_starts[ne0 + i] = e.x;
        a.iv_lengths[ne0 + i] = len;
        if (a.iv_bucket) a.iv_bucket[ne0 + i] = e.z;
        if (len > kLaneSortMax) {
            if (len <= l.mid_max) {
                const int at = atomicAdd(&s_mn, 1);
                if (at < kScanTile / 8) s_m[at] = ne0 + i;
                else {                                   // a tile with more than 256 such intervals: straight to the global queue
                    const int g = atomicAdd(l.counts + 0, 1);
                    if (g < l.mid_cap) l.mid_list[g] = ne0 + i;
                }
            } else if (len <= kWarpSortMax) {
                const int at = atomicAdd(&s_qn, 1);
                if (at < kScanTile / 8) s_q[at] = ne0 + i;
                else {
                    const int g = atomicAdd(l.counts + 1, 1);
                    if (g < l.long_cap) l.long_list[g] = ne0 + i;
                }
            } else {
                const int at = atomicAdd(l.counts + 2, 1);
                if (at < l.long_cap) l.long_list[l.long_cap - 1 - at] = ne0 + i;
            }
        }
    }
    __syncthreads();
    const int nl = min(s_qn, kScanTile / 8), nm = min(s_mn, kScanTile / 8);
    if (tid == 0 && nl > 0) s_qbase = atomicAdd(l.counts + 1, nl);
    if (tid == 32 && nm > 0) s_mbase = atomicAdd(l.counts + 0, nm);
    __syncthreads();
    for (int i = tid; i < nl; i += kScan2Threads)
        if (s_qbase + i < l.long_cap) l.long_list[s_qbase + i] = s_q[i];
    for (int i = tid; i < nm; i += kScan2Threads)
        if (s_mbase + i < l.mid_cap) l.mid_list[s_mbase + i] = s_m[i];
}

// Placement with the interval id riding along: sorted[offset + slot] = p, and the point's voxel id in key[] is
// REPLACED by its interval id (the plan's pt2vox table becomes "pt2iv", FwdPlanHeader::structured = 2): the backward
// plan then needs no voxel -> interval gather (1.7 M scattered sector reads at batch 8) — interval ids order the
// points of a pixel exactly like voxel ids do.
__global__ void __launch_bounds__(256) place_iv_kernel(int32_t *__restrict__ key, const int32_t *__restrict__ slot,
                                                       const int2 *__restrict__ ofiv, int64_t n,
                                                       int32_t *__restrict__ sorted, FwdPlanHeader *hdr) {
    pdl_wait();
    pdl_launch();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid == 0) hdr->structured = 2;
    const int64_t nq = n >> 2;
    for (int64_t q = gtid; q < nq; q += stride) {
        const int4 k4 = *(reinterpret_cast<const int4 *>(key) + q);
        const int4 s4 = __ldcs(reinterpret_cast<const int4 *>(slot) + q);
        const int k[4] = {k4.x, k4.y, k4.z, k4.w}, s[4] = {s4.x, s4.y, s4.z, s4.w};
        int2 e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) e[j] = k[j] >= 0 ? __ldg(ofiv + k[j]) : make_int2(0, -1);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k[j] >= 0) sorted[e[j].x + s[j]] = (int)(q << 2) + j;
        *(reinterpret_cast<int4 *>(key) + q) = make_int4(e[0].y, e[1].y, e[2].y, e[3].y);
    }
    for (int64_t p = (nq << 2) + gtid; p < n; p += stride) {
        const int k = key[p];
        if (k >= 0) {
            const int2 e = ofiv[k];
            sorted[e.x + slot[p]] = (int)p;
            key[p] = e.y;
        }
    }
}

struct Order2Args {
    int32_t *sorted;            // ranks_depth (in/out)
    int32_t *ranks_feat, *ranks_bev;
    const int32_t *iv_starts, *iv_lengths, *iv_bucket;
    const int32_t *n_intervals; // device count
    ListArgs l;
    FastDiv dhw, hw;
    const int32_t *sub_pt;      // dense sub-tile list of the forward plan
    int32_t n_subs;
    int32_t *heavy_list, *heavy_n;
    int32_t heavy_pts;          // dense sub-tile threshold (heavy_threshold(B, V))
};

// One launch for the whole order pass: (1) queued long intervals, one warp each; (2) the forward plan's list of dense
// sub-tiles; (3) every interval of up to kLaneSortMax points, one lane each; (4) very long intervals, one CTA each.
template <bool MID>   // MID: the 9..16-point queue exists (dense frusta); without it the kernel needs 40 instead of 48 registers
__global__ void __launch_bounds__(kSortThreads) order2_kernel(Order2Args a) {
    __shared__ int s_sort[kSortSmemCta];
    pdl_wait();
    pdl_launch();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gtid = blockIdx.x * kSortThreads + threadIdx.x, stride = gridDim.x * kSortThreads;
    const int n_big = min(a.l.counts[2], a.l.long_cap);
    const int n_long = min(a.l.counts[1], a.l.long_cap - n_big);
    const int nwarps = gridDim.x * (kSortThreads / 32);
    for (int w = blockIdx.x * (kSortThreads / 32) + warp; w < n_long; w += nwarps) {
        const int k = a.l.long_list[w];
        const int ls = a.iv_starts[k], ll = a.iv_lengths[k], lb = a.iv_bucket[k];
        warp_sort_segment(a.sorted + ls, ll, nullptr, lane);
        __syncwarp();
        for (int j = lane; j < ll; j += 32) {
            a.ranks_feat[ls + j] = feat_row_of(a.sorted[ls + j], a.dhw, a.hw);
            a.ranks_bev[ls + j] = lb;
        }
        __syncwarp();
    }
    if (a.heavy_list != nullptr)
        for (int u = gtid; u < a.n_subs; u += stride)
            if (a.sub_pt[u + 1] - a.sub_pt[u] > a.heavy_pts) a.heavy_list[atomicAdd(a.heavy_n, 1)] = u;
    // 9..16-point intervals: one lane each, 16-key network in registers (a warp per such interval cost ~200 warp
    // instructions each: 83 k of them at 512x1408 batch 8)
    const int n_mid = MID ? min(a.l.counts[0], a.l.mid_cap) : 0;
    for (int i = gtid; MID && i < n_mid; i += stride) {
        const int k = a.l.mid_list[i];
        const int s = a.iv_starts[k], len = a.iv_lengths[k], bucket = a.iv_bucket[k];
        int v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (j < len) ? a.sorted[s + j] : INT_MAX;
        sort16(v);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (j < len) {
                a.sorted[s + j] = v[j];
                a.ranks_feat[s + j] = feat_row_of(v[j], a.dhw, a.hw);
                a.ranks_bev[s + j] = bucket;
            }
    }
    const int n = *a.n_intervals;
    // one lane per short interval; the next interval's (start, length, voxel) are fetched while this one is ordered
    // (two dependent round trips per iteration otherwise: interval record, then its points)
    int ns = 0, nlen = 0, nb = 0;
    if (gtid < n) { ns = a.iv_starts[gtid]; nlen = a.iv_lengths[gtid]; nb = a.iv_bucket[gtid]; }
    for (int k = gtid; k < n; k += stride) {
        const int s = ns, len = nlen, bucket = nb;
        if (k + stride < n) { ns = a.iv_starts[k + stride]; nlen = a.iv_lengths[k + stride]; nb = a.iv_bucket[k + stride]; }
        if (len > kLaneSortMax) continue;
        int v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (j < len) ? a.sorted[s + j] : INT_MAX;
        if (len > 1) sort8(v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < len) {
                if (len > 1) a.sorted[s + j] = v[j];
                a.ranks_feat[s + j] = feat_row_of(v[j], a.dhw, a.hw);
                a.ranks_bev[s + j] = bucket;
            }
    }
    if (n_big == 0) return;
    __syncthreads();
    for (int w = blockIdx.x; w < n_big; w += gridDim.x) {
        const int k = a.l.long_list[a.l.long_cap - 1 - w];
        const int ls = a.iv_starts[k], ll = a.iv_lengths[k], lb = a.iv_bucket[k];
        if (ll <= kSortSmemCta) {
            cta_sort_segment(a.sorted + ls, ll, s_sort);
        } else {
            if (warp == 0) warp_sort_segment(a.sorted + ls, ll, nullptr, lane);
            __threadfence_block();
            __syncthreads();
        }
        for (int j = threadIdx.x; j < ll; j += kSortThreads) {
            a.ranks_feat[ls + j] = feat_row_of(a.sorted[ls + j], a.dhw, a.hw);
            a.ranks_bev[ls + j] = lb;
        }
        __syncthreads();
    }
}

}  // namespace fo
