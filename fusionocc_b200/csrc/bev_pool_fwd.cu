// fusionocc_b200 — bev_pool_v2 forward (sm_100a).
//
// Reference: mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-48 (kernel) + bev_pool.py:27 (82 MB
// new_zeros) + bev_pool.py:91 (164 MB permute copy).  Here one kernel owns the DENSE OUTPUT: every warp
// takes a sub-tile of 32 consecutive voxels of one sample, reduces the points that land in it and writes
// the whole C x 32 block — zeros included — once, already in (B,C,Z,Y,X) order.  HBM-bound by design:
// the 81.92 MB/sample output is 94 % of the algorithmic bytes.
//
//   * the sorted rank arrays keep a sub-tile's points contiguous, so the plan's sub_pt table gives the
//     warp its point range and 32 points' (ranks_feat, ranks_depth, ranks_bev) arrive with three
//     coalesced loads; the depth values with one gather;
//   * lanes = channels: each point's 128-byte feature row is one coalesced load; the accumulator is the
//     reference's exact FFMA chain (psum = fmaf(feat, depth, psum) from +0.0f, in point order); when the
//     voxel id changes the finished accumulator is flushed to the warp's shared-memory stage;
//   * the stage is the output block's own layout, [C][32] with rows rotated by (c & 7) 16-byte chunks, so
//     the write-out is LDS.128 -> streaming STG.128, four full 128-byte lines per instruction;
//   * single-warp CTAs, no __syncthreads in the main kernel: up to 32 independent warps per SM walk the
//     3-round-trip dependency chain of their sub-tiles; the few very dense sub-tiles (listed by the plan) are
//     split over eight "front" CTAs each, which start first.
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"

namespace fo {

struct FwdArgs {
    const float *depth;
    const float *feat;
    const int32_t *rd, *rf, *rb, *starts, *lengths;
    int64_t n_points;
    int64_t n_intervals;            // capacity / host count
    const int32_t *n_intervals_dev; // optional live count
    int32_t C;
    int32_t B;
    int64_t V;                      // voxels per sample
    float *out;
    const FwdPlanHeader *hdr;
    const int32_t *sub_pt;
    int64_t out_bstride;            // (B,C,Z,Y,X) output: elements between samples (C_total * V; C * V when not a slice)
    int64_t out_rowstride;          // (B,Z,Y,X,C) output: elements between voxels   (C_total; C when not a slice)
    // dense sub-tiles (more than kHeavyPts points): listed by the plan, reduced by the "front" CTAs of the grid
    const int32_t *sub_iv;          // first interval of every sub-tile
    const int32_t *iv_vox;          // voxel id of every interval
    const int32_t *heavy_list;      // sub-tile ids
    const int32_t *n_heavy;         // device count
    int32_t front_y;                // grid rows (blockIdx.y) reserved for front CTAs; 0: everything inline
    int32_t sps;                    // sub-tiles per sample (host-known: saves a dependent load per CTA)
    int32_t check_flags;            // 0: the plan is trusted (FO_FWD_ASSUME_SORTED), skip the flag word
    int32_t store_evict_first;      // bulk-store write-out: L2 evict-first hint (sparse frusta; see forward_impl)
    int32_t heavy_pts;              // dense sub-tile threshold of this plan (heavy_threshold(B, V))
};

__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_zero4(unsigned addr) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "f"(0.f) : "memory");
}
__device__ __forceinline__ float4 lds_f4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// NACC = ceil(C / 32) accumulators per lane; EXACT: C == 32 * NACC (no channel predicates).
//
// Two things bound the first versions of this kernel (profiles/r01): instruction issue (~700 warp
// instructions per sub-tile) and the few dense near-ego sub-tiles (up to 844 points at the headline
// shape, 3 394 at 512x1408) whose strictly sequential FMA chain exposed one L2 round trip per four
// points.  Hence: point records are broadcast through a per-warp shared-memory array (LDS.128 = two
// records, no shuffles), shared-memory addresses are precomputed 32-bit values, feature rows are
// fetched in groups of U with the NEXT group already in flight while the current one is consumed
// (2*U rows in flight per lane), and the next 32 records are fetched while a batch is processed.
#ifndef FO_FWD_U
#define FO_FWD_U 8
#endif
#ifndef FO_FWD_MIN_CTAS
#define FO_FWD_MIN_CTAS 32          // <= 64 registers: 32 single-warp CTAs per SM (the hardware's CTA limit)
#endif
#ifndef FO_FRONT_GROUPS
#define FO_FRONT_GROUPS 8
#endif
#ifndef FO_FRONT_SLOTS
#define FO_FRONT_SLOTS 4096
#endif
constexpr int kFrontGroups = FO_FRONT_GROUPS;     // a dense sub-tile is split into this many front CTAs
constexpr int kFrontSlots  = FO_FRONT_SLOTS;  // front CTAs per launch (they loop over the list)

// The reduction of one warp: points [p_lo, p_hi) of the sorted rank arrays, all inside one sub-tile, are
// accumulated voxel by voxel — psum = fmaf(feat, depth, psum) from +0.0f in point order, the reference's
// FFMA chain (bev_pool_cuda.cu:39-43) — and every finished voxel is flushed into the shared-memory stage
// ([C][32] floats, rows rotated by (c & 7) 16-byte chunks).  lanes = channels.
//   * the three record arrays are fetched together; the first group of feature rows is requested as soon as
//     the row ids are there, before the (dependent) depth gather has returned;
//   * record words / depth values are broadcast through shared memory (LDS.128, no shuffles), feature rows
//     are fetched in groups of U with the NEXT group already in flight (2*U rows in flight per lane), and the
//     next 32 records are fetched while a batch is processed.
template <int NACC, bool EXACT, bool XSWZ = false, int U = (NACC <= 2 ? FO_FWD_U : 4)>   // U = feature rows per group
__device__ __forceinline__ void reduce_points(const FwdArgs &a, const int C, const int lane, const int p_lo,
                                              const int p_hi, const int bV, int *rx, float *rdv,
                                              const unsigned lane_row, const unsigned lane_rot) {
    int mx = 0, mr = -1;
    auto load_idx = [&](int i0) {
        mx = 0; mr = -1;
        if (i0 + lane < p_hi) {
            const int q = __ldg(a.rf + i0 + lane);
            const int v = __ldg(a.rb + i0 + lane) - bV;          // sub-tiles start at multiples of 32 in a sample
            mr = __ldg(a.rd + i0 + lane);
            mx = (q << kSubShift) | (v & (kSub - 1));
        }
    };
    load_idx(p_lo);
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
    int cur_v = -1;
    auto flush = [&]() {
        // rows rotated by (c & 7) 16-byte chunks, or (XSWZ) the TMA engine's 128-byte swizzle: chunk ^ (c & 7)
        const unsigned off = XSWZ ? ((((unsigned)cur_v << 2) ^ lane_rot) & 127u) : ((((unsigned)cur_v << 2) + lane_rot) & 127u);
#pragma unroll
        for (int k = 0; k < NACC; ++k)
            if (EXACT || lane + 32 * k < C) sts_f32(lane_row + off + 4096u * k, acc[k]);
    };
    auto load_group = [&](float (&f)[U][NACC], int (&r)[U], int j) {
#pragma unroll
        for (int t = 0; t < U; ++t) r[t] = rx[j + t];
#pragma unroll
        for (int t = 0; t < U; ++t) {
            const int row = (r[t] >> kSubShift) * C + lane;
#pragma unroll
            for (int k = 0; k < NACC; ++k) f[t][k] = (EXACT || lane + 32 * k < C) ? __ldg(a.feat + row + 32 * k) : 0.f;
        }
    };
    auto consume = [&](const float (&f)[U][NACC], const int (&r)[U], int j, int count) {
        float d[U];
#pragma unroll
        for (int t = 0; t < U; ++t) d[t] = rdv[j + t];
#pragma unroll
        for (int t = 0; t < U; ++t) {
            if (t < count) {                              // warp-uniform (compile-time true for full groups)
                const int v = r[t] & (kSub - 1);
                if (v != cur_v) {                         // warp-uniform: a new interval starts
                    if (cur_v >= 0) flush();
#pragma unroll
                    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
                    cur_v = v;
                }
#pragma unroll
                for (int k = 0; k < NACC; ++k) acc[k] = fmaf(f[t][k], d[t], acc[k]);
            }
        }
    };

    float md = (mr >= 0) ? __ldg(a.depth + mr) : 0.f;
    for (int i0 = p_lo; i0 < p_hi; i0 += 32) {
        const int n = min(32, p_hi - i0);
        __syncwarp();
        rx[lane] = mx;
        if (lane < U) { rx[32 + lane] = 0; rdv[32 + lane] = 0.f; }
        __syncwarp();
        const int nfull = n & ~(U - 1);
        float fa[U][NACC], fb[U][NACC];
        int ra[U], rb2[U];
        load_group(fa, ra, 0);                            // rows of padding records are row 0
        rdv[lane] = md;
        __syncwarp();
        const bool more = i0 + 32 < p_hi;
        if (more) load_idx(i0 + 32);                      // next batch's records fly during this batch
        if (nfull) {
            for (int j = 0; j < nfull; j += 2 * U) {
                const bool has_b = j + U < nfull;
                if (has_b) load_group(fb, rb2, j + U);
                consume(fa, ra, j, U);
                if (j == 0 && more) md = (mr >= 0) ? __ldg(a.depth + mr) : 0.f;
                if (has_b) {
                    if (j + 2 * U < nfull) load_group(fa, ra, j + 2 * U);
                    consume(fb, rb2, j + U, U);
                }
            }
            if (nfull < n) {                              // remainder group
                load_group(fa, ra, nfull);
                consume(fa, ra, nfull, n - nfull);
            }
        } else {
            consume(fa, ra, 0, n);
        }
    }
    if (cur_v >= 0) flush();
}

// One warp = one sub-tile = one CTA.  The dependent chain  {flag word, point range} -> rank records ->
// {feature rows, depth}  is three round trips: sub-tiles per sample is a kernel argument, so the plan's flag
// word and the point range are independent loads.
//
// Measured on a B200 (headline shape, batch 8; profiles/r01_summary.md "second session"): this kernel wants MANY
// SHORT-LIVED warps and a compact write window.  Runs of 2/4/8 sub-tiles per warp: 154/160/197 us; persistent
// warps with a 4-deep software pipeline across sub-tiles: 185 us; 2/4/8-warp CTAs with a cooperative
// >= 256-byte-per-plane write-out: 166/175/193 us; 21-25 instead of 32 resident CTAs per SM: 159-172 us;
// write-back / .cg / .wt instead of streaming stores: 169 us; this version: 149-152 us.
//
// Dense sub-tiles (more than kHeavyPts points; the plan lists them): their serial FMA chain (3 394 points at
// 512x1408, 844 at the headline shape) would be the critical path of a short launch and a tail of a long one.
// The first rows of the grid are FRONT CTAs: each takes one of kFrontGroups contiguous interval groups (about equal
// point counts) of a listed sub-tile, reduces it into its own stage and writes exactly the voxel columns from its
// first interval's voxel up to the next group's — so the groups tile the sub-tile and every element is still
// written once.  They start first and overlap with the bulk of the grid; regular CTAs skip listed sub-tiles.
// TMAST: the staged block (128-byte-swizzled rows) leaves through ONE cp.async.bulk.tensor store issued by lane 0
// instead of 8 LDS.128 + 8 STG.128 per lane (measured in profiles/r02_summary.md).
template <int NACC, bool EXACT, int LAYOUT, bool TMAST>
__global__ void __launch_bounds__(32, FO_FWD_MIN_CTAS) fwd_dense_kernel(FwdArgs a, const __grid_constant__ CUtensorMap tm) {
    extern __shared__ __align__(1024) float smem_raw[];  // stage [C][32]
    float *smem = TMAST ? reinterpret_cast<float *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023) : smem_raw;
    __shared__ __align__(16) int s_rx[32 + 8];           // (feature row << 5 | voxel slot) per point
    __shared__ __align__(16) float s_rd[32 + 8];         // depth value per point
    pdl_wait();
    pdl_launch();

    const int lane = threadIdx.x;
    const int C = EXACT ? 32 * NACC : a.C;
    // grid = (B, subs_per_sample): x-fastest block order interleaves the samples, so the dense near-ego
    // regions of all samples are reached at the same relative time
    const int sps = a.sps;
    const int64_t V = a.V;
    if ((int)blockIdx.y < a.front_y) {                    // ---- front CTA: one group of a dense sub-tile ----
        if (a.check_flags && (__ldg(&a.hdr->flags) & kFlagUnsorted)) return;
        const int n_tasks = __ldg(a.n_heavy) * kFrontGroups;
        const unsigned sb = (unsigned)__cvta_generic_to_shared(smem);
        for (int e = blockIdx.y * gridDim.x + blockIdx.x; e < n_tasks; e += a.front_y * gridDim.x) {
            const int u = __ldg(a.heavy_list + e / kFrontGroups), g = e % kFrontGroups;
            const int hb = u / sps, hsu = u - hb * sps;
            const int pa = __ldg(a.sub_pt + u), pb = __ldg(a.sub_pt + u + 1);
            const int ia = __ldg(a.sub_iv + u), ni = min(__ldg(a.sub_iv + u + 1) - ia, kSub);
            const int hv0 = hsu << kSubShift, hbV = hb * (int)V;
            const int nv = min(kSub, (int)V - hv0);
            int s_l = pb, v_l = nv;                       // first point / voxel slot of this lane's interval
            if (lane < ni) { s_l = __ldg(a.starts + ia + lane); v_l = __ldg(a.iv_vox + ia + lane) - hbV - hv0; }
            const int share = (int)(((long long)(s_l - pa) * kFrontGroups) / (pb - pa));
            // cut(g) = first interval whose share is >= g: group g owns intervals [cut(g), cut(g+1))
            const unsigned m0 = __ballot_sync(0xffffffffu, lane < ni && share >= g);
            const unsigned m1 = (g + 1 < kFrontGroups) ? __ballot_sync(0xffffffffu, lane < ni && share >= g + 1) : 0u;
            const int c0 = m0 ? __ffs(m0) - 1 : ni, c1 = m1 ? __ffs(m1) - 1 : ni;
            const int vs0 = __shfl_sync(0xffffffffu, v_l, min(c0, 31)), vs1 = __shfl_sync(0xffffffffu, v_l, min(c1, 31));
            const int ps0 = __shfl_sync(0xffffffffu, s_l, min(c0, 31)), ps1 = __shfl_sync(0xffffffffu, s_l, min(c1, 31));
            const int vstart = (g == 0) ? 0 : (c0 < ni ? vs0 : nv), vend = (c1 < ni) ? vs1 : nv;
            const int p_lo = (c0 < ni) ? ps0 : pb, p_hi = (c1 < ni) ? ps1 : pb;
            if (vstart >= vend) continue;                 // (warp-uniform) an empty group owns no voxels
            if (EXACT) {
#pragma unroll
                for (int i = 0; i < 8 * NACC; ++i) sts_zero4(sb + 16u * lane + 512u * i);
            } else {
                for (int i = lane; i < C * (kSub / 4); i += 32) sts_zero4(sb + 16u * i);
            }
            if (p_lo < p_hi)
                reduce_points<NACC, EXACT>(a, C, lane, p_lo, p_hi, hbV, s_rx, s_rd, sb + ((unsigned)lane << 7),
                                           ((unsigned)lane & 7u) << 4);
            __syncwarp();
            if (lane >= vstart && lane < vend) {          // one voxel column per lane, one channel plane per store
                float *dst = a.out + (int64_t)hb * a.out_bstride + hv0 + lane;
                for (int c = 0; c < C; ++c) __stcs(dst + (int64_t)c * V, smem[stage_index(c, lane)]);
            }
            __syncwarp();                                 // the stage is reused by the next task
        }
        return;
    }
    const int b = blockIdx.x;
    const int su = blockIdx.y - a.front_y;
    // one round trip: the plan's flag word and the point range are independent loads
    const int flags = a.check_flags ? __ldg(&a.hdr->flags) : 0;
    const int pa = __ldg(a.sub_pt + b * sps + su), pb = __ldg(a.sub_pt + b * sps + su + 1);
    if (flags & kFlagUnsorted) return;                   // the order-agnostic path runs instead
    const int bV = b * (int)V;                           // global voxel id of the sample's first voxel (< 2^31)
    const int v0 = su << kSubShift;
    const int nv = min(kSub, (int)V - v0);
    const bool vec_out = (LAYOUT == FO_LAYOUT_BCZYX) && ((V & 3) == 0);
    const int riq = lane >> 3, chunk = lane & 7;         // row within a quad of rows, 16-byte chunk
    float *stage = smem;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(stage);
    float *pl = a.out + (int64_t)b * a.out_bstride + (int64_t)riq * V + v0;      // row riq of this sub-tile's block

    if (pa >= pb && vec_out) {                            // empty sub-tile: stream zeros, no staging
        if (4 * chunk < nv) {
            float *dst = pl + 4 * chunk;
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (EXACT) {
#pragma unroll
                for (int i = 0; i < 8 * NACC; ++i) __stcs(reinterpret_cast<float4 *>(dst + (4 * i) * V), zero4);
            } else {
                for (int r = riq; r < C; r += 4, dst += 4 * V) __stcs(reinterpret_cast<float4 *>(dst), zero4);
            }
        }
        return;
    }
    if (a.front_y > 0 && pb - pa > a.heavy_pts) return;  // a dense sub-tile: the front CTAs own it
    if (EXACT) {
#pragma unroll
        for (int i = 0; i < 8 * NACC; ++i) sts_zero4(sbase + 16u * lane + 512u * i);
    } else {
        for (int e = lane; e < C * (kSub / 4); e += 32) sts_zero4(sbase + 16u * e);
    }
    reduce_points<NACC, EXACT, TMAST>(a, C, lane, pa, pb, bV, s_rx, s_rd, sbase + ((unsigned)lane << 7),
                                      ((unsigned)lane & 7u) << 4);
    if (TMAST) {
        fence_proxy_async_smem();                         // the flushes (generic proxy) -> visible to the bulk store
        __syncwarp();
        if (lane == 0) {
            // C rows x 128 bytes; voxels beyond V are not written
            if (a.store_evict_first) tma_store_3d_hint(&tm, v0, 0, b, sbase, l2_policy_evict_first());
            else tma_store_3d(&tm, v0, 0, b, sbase);
            bulk_commit();
            bulk_wait_read0();                            // the stage must outlive the store's reads
        }
        return;
    }
    __syncwarp();

    if (LAYOUT == FO_LAYOUT_BCZYX) {
        if (vec_out) {
            // lane -> (row r of a quad, 16-byte chunk): four full 128-byte lines per instruction.  The
            // logical chunk stored at smem position `chunk` of row r is (chunk - r) & 7: it alternates
            // between two values as r advances by 4.
            const int ck0 = (chunk - riq) & 7, ck1 = ck0 ^ 4;
            float *d0 = pl + 4 * ck0;
            float *d1 = pl + 4 * V + 4 * ck1;
            const unsigned sa = sbase + 16u * lane;       // row riq, position chunk
            const bool w0 = 4 * ck0 < nv, w1 = 4 * ck1 < nv;
            if (EXACT) {
#pragma unroll
                for (int i = 0; i < 4 * NACC; ++i) {
                    const float4 x0 = lds_f4(sa + 1024u * i);
                    const float4 x1 = lds_f4(sa + 1024u * i + 512u);
                    if (w0) __stcs(reinterpret_cast<float4 *>(d0 + (8 * i) * V), x0);
                    if (w1) __stcs(reinterpret_cast<float4 *>(d1 + (8 * i) * V), x1);
                }
            } else {
                for (int r = riq, i = 0; r < C; r += 8, ++i) {
                    const float4 x0 = lds_f4(sa + 1024u * i);
                    if (w0) __stcs(reinterpret_cast<float4 *>(d0 + (8 * i) * V), x0);
                    if (r + 4 < C) {
                        const float4 x1 = lds_f4(sa + 1024u * i + 512u);
                        if (w1) __stcs(reinterpret_cast<float4 *>(d1 + (8 * i) * V), x1);
                    }
                }
            }
        } else {
            float *plane0 = a.out + (int64_t)b * a.out_bstride + v0;
            for (int e = lane; e < C * kSub; e += 32) {
                const int c = e >> kSubShift, v = e & (kSub - 1);
                if (v < nv) __stcs(plane0 + (int64_t)c * V + v, stage[stage_index(c, v)]);
            }
        }
    } else {
        // (B,Z,Y,X,C): the sub-tile is nv rows of C floats
        float *dst = a.out + ((int64_t)bV + v0) * a.out_rowstride;
        for (int e = lane; e < nv * C; e += 32) {
            const int v = e / C, c = e - v * C;
            __stcs(dst + (int64_t)v * a.out_rowstride + c, stage[stage_index(c, v)]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Order-agnostic path: guarded zero fill + interval scatter.  Used (a) when the plan found the
// interval list not in canonical form (voxels not strictly increasing, intervals not back to back, or
// out of range), (b) by the source-compatible fo_compat_bev_pool_v2 (caller-zeroed (B,Z,Y,X,C) output,
// arbitrary interval order) and (c) for channel counts too large for the staged tile.  Invalid
// intervals are skipped instead of writing out of bounds.
// ------------------------------------------------------------------------------------------------
// zero-fills n_runs runs of run_len floats, run_stride floats apart (one run per sample of a (B,C,Z,Y,X) channel
// slice; one run per voxel of a (B,Z,Y,X,C) slice; a single run when the output is not a slice)
__global__ void __launch_bounds__(256) zero_if_flag_kernel(float *out, int64_t n_runs, int64_t run_len,
                                                           int64_t run_stride, int vec, const FwdPlanHeader *hdr,
                                                           int need_flag) {
    if (need_flag && !(hdr->flags & kFlagUnsorted)) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const int64_t per = run_len >> 2, total = n_runs * per;
        for (int64_t i = gtid; i < total; i += stride) {
            const int64_t r = i / per, e = i - r * per;
            __stcs(reinterpret_cast<float4 *>(out + r * run_stride) + e, make_float4(0.f, 0.f, 0.f, 0.f));
        }
    } else {
        const int64_t total = n_runs * run_len;
        for (int64_t i = gtid; i < total; i += stride) {
            const int64_t r = i / run_len, e = i - r * run_len;
            out[r * run_stride + e] = 0.f;
        }
    }
}

template <int LAYOUT>
__global__ void __launch_bounds__(256) fwd_scatter_kernel(FwdArgs a, int need_flag) {
    if (need_flag && !(a.hdr->flags & kFlagUnsorted)) return;
    const int64_t n = a.n_intervals_dev ? min((int64_t)max(*a.n_intervals_dev, 0), a.n_intervals) : a.n_intervals;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int C = a.C;
    const int64_t nvt = (int64_t)a.B * a.V;
    for (int64_t k = warp0; k < n; k += nwarps) {
        const int s = a.starts[k], len = a.lengths[k];
        if (s < 0 || len < 0 || s >= a.n_points || (int64_t)s + len > a.n_points) continue;
        const int v = a.rb[s];
        if (v < 0 || v >= nvt) continue;
        const int64_t b = v / a.V, vin = v - b * a.V;
        for (int c = lane; c < C; c += 32) {
            float psum = 0.f;
            for (int i = 0; i < len; ++i)
                psum = fmaf(a.feat[(int64_t)a.rf[s + i] * C + c], a.depth[a.rd[s + i]], psum);
            if (LAYOUT == FO_LAYOUT_BCZYX) a.out[b * a.out_bstride + c * a.V + vin] = psum;
            else a.out[(int64_t)v * a.out_rowstride + c] = psum;
        }
    }
}

}  // namespace fo

using namespace fo;

namespace {
template <int NACC, bool EXACT, int LAYOUT>
int launch_dense(const FwdArgs &a, int n_ctas, size_t smem, cudaStream_t stream, const CUtensorMap *tm) {
    if (tm != nullptr && LAYOUT == FO_LAYOUT_BCZYX) {
        auto kern = fwd_dense_kernel<NACC, EXACT, FO_LAYOUT_BCZYX, true>;
        const size_t sm2 = smem + 1024;
        if (sm2 > 48 * 1024) FO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        FO_CUDA(launch_pdl(kPdlFwd, kern, dim3(a.B, n_ctas + a.front_y), dim3(32), sm2, stream, a, *tm));
        return FO_OK;
    }
    auto kern = fwd_dense_kernel<NACC, EXACT, LAYOUT, false>;
    if (smem > 48 * 1024) FO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUtensorMap dummy{};
    FO_CUDA(launch_pdl(kPdlFwd, kern, dim3(a.B, n_ctas + a.front_y), dim3(32), smem, stream, a, dummy));
    return FO_OK;
}
template <int LAYOUT>
int launch_dense_any(const FwdArgs &a, int n_ctas, size_t smem, cudaStream_t stream, const CUtensorMap *tm) {
    const int nacc = (a.C + 31) / 32;
    if (a.C % 32 == 0) {
        switch (nacc) {
            case 1: return launch_dense<1, true, LAYOUT>(a, n_ctas, smem, stream, tm);
            case 2: return launch_dense<2, true, LAYOUT>(a, n_ctas, smem, stream, tm);
            case 3: return launch_dense<3, true, LAYOUT>(a, n_ctas, smem, stream, tm);
            case 4: return launch_dense<4, true, LAYOUT>(a, n_ctas, smem, stream, tm);
            default: break;
        }
    }
    switch (nacc) {
        case 1: return launch_dense<1, false, LAYOUT>(a, n_ctas, smem, stream, tm);
        case 2: return launch_dense<2, false, LAYOUT>(a, n_ctas, smem, stream, tm);
        case 3: return launch_dense<3, false, LAYOUT>(a, n_ctas, smem, stream, tm);
        case 4: return launch_dense<4, false, LAYOUT>(a, n_ctas, smem, stream, tm);
        case 5: return launch_dense<5, false, LAYOUT>(a, n_ctas, smem, stream, tm);
        case 6: return launch_dense<6, false, LAYOUT>(a, n_ctas, smem, stream, tm);
        case 7: return launch_dense<7, false, LAYOUT>(a, n_ctas, smem, stream, tm);
        default: return launch_dense<8, false, LAYOUT>(a, n_ctas, smem, stream, tm);
    }
}
}  // namespace

namespace {
int forward_impl(cudaStream_t stream, int32_t c, const float *depth, const float *feat, const int32_t *ranks_depth,
                 const int32_t *ranks_feat, const int32_t *ranks_bev, const int32_t *interval_starts,
                 const int32_t *interval_lengths, int64_t n_points, int64_t n_intervals,
                 const int32_t *n_intervals_dev, int32_t B, int64_t n_vox, float *out, int32_t out_layout,
                 int32_t c_total, int32_t c_offset, int32_t flags, const void *plan, size_t plan_bytes) {
    FO_CHECK_ARG(c >= 1, "channels must be positive (got %d)", c);
    FO_CHECK_ARG(c_offset >= 0 && c_total >= c && c_offset + c <= c_total,
                 "channel slice [%d, %d) does not fit %d channels", c_offset, c_offset + c, c_total);
    FO_CHECK_ARG(B >= 1 && n_vox >= 1, "B and voxels per sample must be positive");
    FO_CHECK_ARG(out != nullptr && ((uintptr_t)out & 15) == 0, "out must be non-NULL and 16-byte aligned");
    FO_CHECK_ARG(out_layout == FO_LAYOUT_BCZYX || out_layout == FO_LAYOUT_BZYXC, "unknown out_layout %d", out_layout);
    FO_CHECK_ARG(n_points >= 0 && n_intervals >= 0 && n_points < INT_MAX, "negative or oversized counts");
    FO_CHECK_ARG(n_intervals == 0 || (depth && feat && ranks_depth && ranks_feat && ranks_bev && interval_starts &&
                                      interval_lengths), "NULL input array");
    FwdPlanView pv; int64_t n_subs; int sps;
    if (int rc = open_fwd_plan_const(plan, plan_bytes, B, n_vox, n_points, &pv, &n_subs, &sps)) return rc;

    FwdArgs a;
    a.depth = depth; a.feat = feat; a.rd = ranks_depth; a.rf = ranks_feat; a.rb = ranks_bev;
    a.starts = interval_starts; a.lengths = interval_lengths;
    a.n_points = n_points; a.n_intervals = n_intervals; a.n_intervals_dev = n_intervals_dev;
    a.C = c; a.B = B; a.V = n_vox; a.hdr = pv.hdr; a.sub_pt = pv.sub_pt;
    // a channel slice [c_offset, c_offset + c) of a wider tensor: only the sample / voxel strides change
    a.out = out + (out_layout == FO_LAYOUT_BCZYX ? (int64_t)c_offset * n_vox : (int64_t)c_offset);
    a.out_bstride = (int64_t)c_total * n_vox;
    a.out_rowstride = c_total;
    a.sps = sps; a.check_flags = (flags & FO_FWD_ASSUME_SORTED) ? 0 : 1;
    a.heavy_pts = heavy_threshold(B, n_vox);

    // dense sub-tiles go to the front CTAs (contiguous (B,C,Z,Y,X) output only)
    const bool front_ok = out_layout == FO_LAYOUT_BCZYX && interval_starts != nullptr;
    a.sub_iv = pv.sub_iv; a.iv_vox = pv.iv_vox; a.heavy_list = pv.heavy_list; a.n_heavy = pv.hdr->fwd_heavy;
    a.front_y = front_ok ? (kFrontSlots + B - 1) / B : 0;
    const size_t smem = (size_t)kSub * c * sizeof(float);
    // 32-bit index arithmetic inside the kernel: feature rows * C and B*V must stay below 2^31 / 2^26
    const int n_ctas = sps;                                           // per sample (grid.y)
    const bool dense_ok = smem <= 200 * 1024 && c <= 256 && n_ctas + a.front_y <= 65535 && B <= 65535;
    if (dense_ok) {
        // Write-out through ONE bulk tensor store per sub-tile (FO_FWD_TMA=0 selects the LDS.128 + STG.128 path for
        // A/B runs).  Measured on a B200 (profiles/r02_summary.md): with the L2 evict-first hint 146.7 -> 141.8 us at
        // the headline shape batch 8 and 25.1 -> 21.0 us at batch 1, but 276 -> 282 us at 512x1408 (four times the
        // points per voxel: the kernel then lives on L2-resident index / feature reads); without the hint 272 -> 267 us
        // there and 146 -> 156 us at the headline shape.  Hence the hint only for sparse frusta.
        CUtensorMap tm;
        const CUtensorMap *tmp = nullptr;
        const char *te = getenv("FO_FWD_TMA");
        const int use_tma = (te && *te) ? atoi(te) : 1;
        a.store_evict_first = (n_points < (int64_t)B * n_vox) ? 1 : 0;   // fewer points than voxels (capacity or live count)
        if (use_tma >= 1 && out_layout == FO_LAYOUT_BCZYX && tmap_ok(a.out, n_vox, c, c_total)) {
            if (int rc2 = make_voxel_tmap(&tm, a.out, n_vox, c, c_total, B)) return rc2;
            tmp = &tm;
        }
        int rc = (out_layout == FO_LAYOUT_BCZYX) ? launch_dense_any<FO_LAYOUT_BCZYX>(a, n_ctas, smem, stream, tmp)
                                                 : launch_dense_any<FO_LAYOUT_BZYXC>(a, n_ctas, smem, stream, tmp);
        if (rc) return rc;
    }
    // order-agnostic path, guarded by the plan's flag on the device (no host sync); unconditional when
    // the staged tile does not fit shared memory.
    const int need_flag = dense_ok ? 1 : 0;
    if (dense_ok && (flags & FO_FWD_ASSUME_SORTED)) return FO_OK;
    int64_t n_runs, run_len, run_stride;
    if (c_total == c) { n_runs = 1; run_len = (int64_t)B * n_vox * c; run_stride = 0; }
    else if (out_layout == FO_LAYOUT_BCZYX) { n_runs = B; run_len = (int64_t)c * n_vox; run_stride = a.out_bstride; }
    else { n_runs = (int64_t)B * n_vox; run_len = c; run_stride = c_total; }
    const int vec = (run_len % 4 == 0) && (run_stride % 4 == 0) && (((uintptr_t)a.out & 15) == 0);
    zero_if_flag_kernel<<<sm_count() * 8, 256, 0, stream>>>(a.out, n_runs, run_len, run_stride, vec, pv.hdr, need_flag);
    FO_LAUNCH_CHECK("zero_if_flag_kernel");
    const int blocks = grid_for(n_intervals * 32, 256, 8);
    if (out_layout == FO_LAYOUT_BCZYX)
        fwd_scatter_kernel<FO_LAYOUT_BCZYX><<<blocks, 256, 0, stream>>>(a, need_flag);
    else
        fwd_scatter_kernel<FO_LAYOUT_BZYXC><<<blocks, 256, 0, stream>>>(a, need_flag);
    FO_LAUNCH_CHECK("fwd_scatter_kernel");
    return FO_OK;
}
}  // namespace

extern "C" int fo_bev_pool_v2_forward(fo_stream_t stream_, int32_t c, const float *depth, const float *feat,
                                      const int32_t *ranks_depth, const int32_t *ranks_feat,
                                      const int32_t *ranks_bev, const int32_t *interval_starts,
                                      const int32_t *interval_lengths, int64_t n_points, int64_t n_intervals,
                                      const int32_t *n_intervals_dev, int32_t B, int64_t n_vox, float *out,
                                      int32_t out_layout, int32_t flags, const void *plan, size_t plan_bytes) {
    return forward_impl((cudaStream_t)stream_, c, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts,
                        interval_lengths, n_points, n_intervals, n_intervals_dev, B, n_vox, out, out_layout, c, 0,
                        flags, plan, plan_bytes);
}

extern "C" int fo_bev_pool_v2_forward_slice(fo_stream_t stream_, int32_t c, const float *depth, const float *feat,
                                            const int32_t *ranks_depth, const int32_t *ranks_feat,
                                            const int32_t *ranks_bev, const int32_t *interval_starts,
                                            const int32_t *interval_lengths, int64_t n_points,
                                            int64_t n_intervals, const int32_t *n_intervals_dev, int32_t B,
                                            int64_t n_vox, float *out, int32_t out_layout, int32_t c_total,
                                            int32_t c_offset, int32_t flags, const void *plan, size_t plan_bytes) {
    return forward_impl((cudaStream_t)stream_, c, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts,
                        interval_lengths, n_points, n_intervals, n_intervals_dev, B, n_vox, out, out_layout, c_total,
                        c_offset, flags, plan, plan_bytes);
}

// Source-compatible launcher: semantics of bev_pool.cpp:7-9 / bev_pool_cuda.cu:125-131 — assign into a
// caller-zeroed (B,Z,Y,X,C) tensor, any interval order, legacy default stream, no status.
extern "C" void fo_compat_bev_pool_v2(int c, int n_intervals, const float *depth, const float *feat,
                                      const int *ranks_depth, const int *ranks_feat, const int *ranks_bev,
                                      const int *interval_starts, const int *interval_lengths, float *out) {
    if (n_intervals <= 0 || c <= 0) return;
    FwdArgs a;
    a.depth = depth; a.feat = feat; a.rd = ranks_depth; a.rf = ranks_feat; a.rb = ranks_bev;
    a.starts = interval_starts; a.lengths = interval_lengths;
    a.n_points = INT_MAX - 1; a.n_intervals = n_intervals; a.n_intervals_dev = nullptr;
    a.C = c; a.B = 1; a.V = INT_MAX - 1; a.out = out; a.hdr = nullptr; a.sub_pt = nullptr;
    a.sps = 0; a.check_flags = 0; a.out_bstride = 0; a.out_rowstride = c;
    a.sub_iv = nullptr; a.iv_vox = nullptr; a.heavy_list = nullptr; a.n_heavy = nullptr; a.front_y = 0;
    a.store_evict_first = 0; a.heavy_pts = kHeavyPts;
    fwd_scatter_kernel<FO_LAYOUT_BZYXC><<<grid_for((int64_t)n_intervals * 32, 256, 16), 256, 0, 0>>>(a, 0);
}
