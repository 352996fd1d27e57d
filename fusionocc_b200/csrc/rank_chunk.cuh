// fusionocc_b200 — rank precompute, round-2 pipeline: a two-level sort whose second level lives in shared memory.
//
// Replaces view_transformer.py:223-281 like the round-1 pipeline (bucket_sort.cuh), with the same outputs bit for
// bit.  Round 1 kept ONE counter per voxel in a 20 MB global array (batch 8): one global atomic per kept point, two
// full passes over 5.12 M counters for the scan, a scattered placement and a scattered order pass — every pass
// bound by 32-byte L2 sectors per 4-byte access (profiles/r01_summary.md: 112 us at 8 % of the HBM roofline).
// Here the voxel range is cut into CHUNKS of 1024 consecutive voxels of one sample (32 sub-tiles, 4.5 KB of counters):
//
//   A1  chunk_voxelize   voxel id of every frustum point (from coor, or from the calibration), written once
//                        (pt2vox, the plan keeps it); points per chunk counted in a per-CTA shared histogram and
//                        merged with one global atomic per (CTA, touched chunk); the last CTA to finish turns the
//                        chunk histogram into chunk offsets (no extra launch)
//   A2  chunk_scatter    multisplit: every kept point's (voxel id, point index) goes to its chunk's list — ranks
//                        inside the CTA by shared atomics, one global atomic per (CTA, touched chunk) reserves the
//                        CTA's run in the list
//   D1  chunk_distinct   non-empty voxels per chunk (a 1024-bit map per warp); the last CTA scans them: the global
//                        number of every chunk's first interval, so that the sort needs no inter-CTA communication
//   D2  chunk_sort<NW>   NW warps per chunk (1 for up to 512 points, 4 up to 4096, 16 above; three launches, every
//                        CTA checks its chunk's class): count in shared memory, scan by the 32 sub-tile owners
//                        (lane = sub-tile), placement into a shared stage, ascending order inside every voxel (= the
//                        stable sort; work lists keep all lanes busy), then COALESCED writes of ranks_depth /
//                        ranks_feat / ranks_bev, interval_starts / lengths and the plan tables (vox2iv, iv_vox,
//                        sub_iv, sub_pt, dense sub-tile list)
//
// Global atomics: ~4 * 10^5 on 40 000 addresses instead of 1.7 * 10^6 returning ones; no pass over the empty 78 % of
// the voxel grid touches DRAM.
//
// STATUS: bit-identical to the round-1 pipeline on every test, but NOT the default: measured on a B200 at the
// headline shape, batch 8 (profiles/r02_summary.md, profiles/r02_rank_chunk_ncu.txt) it takes 160 us against 108 us
// (785 against 296 us at 512x1408).  Every chunk costs ~2 000 warp instructions of dependent shared-memory work
// (zero, count, scan, place, emit 1024 voxel slots, order, write) at ~12 stall cycles each with 20 warps per SM, and
// the one-CTA prefix scans at the end of A1 / D1 are 15 us tails.  Selected with FO_RANK_IMPL=1.
#pragma once

#include "bucket_sort.cuh"

namespace fo {

constexpr int kChunkVox      = 1024;                         // voxels per chunk: 32 sub-tiles, one per lane of the owner warp
constexpr int kChunkCntWords = kChunkVox + (kChunkVox >> 5) * 4;   // counters padded by 4 words per 32 (bank-conflict-free rows)
constexpr int kChunkThreads  = 256;                          // A1 / A2
constexpr int kChunkPtsPerThread = 8;
constexpr int kChunkBlockPts = kChunkThreads * kChunkPtsPerThread;
constexpr int kChunkMaxPerSample = 8192;                     // per-CTA shared histogram bound (A1 / A2)
constexpr int kClassS = 512;                                 // points: one warp per chunk up to here
constexpr int kClassM = 4096;                                // four warps up to here, sixteen above
constexpr int kStageL = 24576;                               // staged points of the 16-warp class (denser chunks stage in global memory)
// per-chunk work lists of a STAGED chunk never overflow: at most STAGE / 2 voxels of >= 2 points, STAGE / 9 of >= 9
__host__ __device__ constexpr int short_cap(int stage) { return stage / 2; }
__host__ __device__ constexpr int long_cap(int stage) { return stage / 9 + 1; }

struct ChunkGeom {
    int32_t cps;             // chunks per sample
    int32_t n_chunks;        // B * cps
    int32_t sps_sub;         // 32-voxel sub-tiles per sample
    int64_t V;               // voxels per sample
};
__host__ inline ChunkGeom chunk_geom(int64_t V, int32_t B) {
    ChunkGeom g;
    g.V = V;
    g.cps = (int32_t)((V + kChunkVox - 1) / kChunkVox);
    g.n_chunks = g.cps * B;
    g.sps_sub = (int32_t)subs_per_sample(V);
    return g;
}

// Scratch:  [zeroed: ctrl(256) | hist | cursor]  chunk_off | iv_off | list
struct ChunkScratch {
    int32_t *ctrl;                   // [1] finished CTAs of A1, [2] finished CTAs of D1, [4] / [5] medium / dense chunks listed
    int32_t *hist;                   // [n_chunks] points per chunk
    int32_t *cursor;                 // [n_chunks] list positions handed out by A2
    int32_t *chunk_off;              // [n_chunks + 1] exclusive prefix of hist (= first list position = first sorted position)
    int32_t *iv_off;                 // [n_chunks + 1] D1: non-empty voxels per chunk, then their exclusive prefix
    int2 *list;                      // [P] (voxel id, point index), grouped by chunk
    int32_t *list_m, *list_l;        // [n_chunks] chunks of the medium / dense class (built by A1's last CTA)
    size_t zero_bytes, total_bytes;
};
__host__ inline int64_t chunk_bound(int64_t n_vox_total) { return n_vox_total / kChunkVox + 4097 + 1; }
__host__ inline ChunkScratch chunk_scratch_view(void *base, int64_t n_chunks_cap, int64_t P) {
    ChunkScratch s;
    char *p = (char *)base;
    s.ctrl = (int32_t *)p;                       p += 256;
    s.hist = (int32_t *)p;                       p += align_up(n_chunks_cap * 4, 256);
    s.cursor = (int32_t *)p;                     p += align_up(n_chunks_cap * 4, 256);
    s.zero_bytes = (size_t)(p - (char *)base);
    s.chunk_off = (int32_t *)p;                  p += align_up((n_chunks_cap + 1) * 4, 256);
    s.iv_off = (int32_t *)p;                     p += align_up((n_chunks_cap + 1) * 4, 256);
    s.list = (int2 *)p;                          p += align_up(P * 8, 256);
    s.list_m = (int32_t *)p;                     p += align_up(n_chunks_cap * 4, 256);
    s.list_l = (int32_t *)p;                     p += align_up(n_chunks_cap * 4, 256);
    s.total_bytes = (size_t)(p - (char *)base);
    return s;
}

struct ChunkArgs {
    VoxArgs vox;                     // coor / grid / pt2vox (key) / hdr ...; cnt and slot unused
    CalibArgs calib;
    ChunkGeom g;
    int32_t B, n_cams, dhw_pts;      // samples, cameras per sample, frustum points per camera (D*H*W)
    int64_t pps;                     // frustum points per sample
    int32_t *hist, *cursor, *chunk_off, *iv_off, *ctrl;
    int2 *list;
    int32_t *list_m, *list_l;
    int32_t *counts;                 // counts_dev: [0] kept points, [1] intervals
    // outputs of the sort kernels
    int32_t *rb, *rd, *rf, *iv_starts, *iv_lengths;
    int32_t *sub_iv, *sub_pt, *heavy_list, *vox2iv, *iv_vox;
    int32_t heavy_pts;           // dense sub-tile threshold (heavy_threshold(B, V))
    FwdPlanHeader *hdr;
    FastDiv dhw, hw;
    int32_t n_subs;
};

__device__ __forceinline__ int chunk_pad(int v) { return v + ((v >> 5) << 2); }

// exclusive scan of src[0..n) into dst[0..n], dst[n] = total, by ONE CTA of kChunkThreads threads; returns the total.
// Every thread owns a contiguous run of up to 16 elements, all loaded before anything is summed: one memory round
// trip per 4096 elements (a loop of 256-element rounds cost one round trip EACH, ~30 us for 5 000 chunks).
__device__ __forceinline__ int cta_exclusive_scan(const int32_t *src, int32_t *dst, int n) {
    __shared__ int s_warp[kChunkThreads / 32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += kChunkThreads * 16) {
        const int per = min(16, (n - base + kChunkThreads - 1) / kChunkThreads);
        const int i0 = base + tid * per;
        int v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = (k < per && i0 + k < n) ? __ldcg(src + i0 + k) : 0;
        int sum = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) sum += v[k];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int woff = 0;
#pragma unroll
        for (int w = 0; w < kChunkThreads / 32; ++w) woff += (w < warp) ? s_warp[w] : 0;
        const int carry = s_carry;
        int run = carry + woff + incl - sum;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (k < per && i0 + k < n) { dst[i0 + k] = run; run += v[k]; }
        __syncthreads();
        if (tid == kChunkThreads - 1) s_carry = carry + woff + incl;
        __syncthreads();
    }
    if (tid == 0) dst[n] = s_carry;
    return s_carry;
}

// Voxel index along one axis: trunc(RN(RN(c - lb) / itv)) (view_transformer.py:246-248; the cast truncates toward
// zero).  The IEEE division costs ~12 instructions; q' = t * RN(1 / itv) differs from RN(t / itv) by at most
// 3 * 2^-24 |q|, so whenever q' is farther than 2^-21 |q'| from the nearest integer both truncate to the same index
// and the division is skipped.  Non-finite values fail the comparison and take the exact path.
__device__ __forceinline__ long long axis_index(float c, float lb, float itv, float rcp) {
    const float t = __fsub_rn(c, lb);
    const float q = __fmul_rn(t, rcp);
    const float dist = fabsf(q - rintf(q));
    if (!(dist > fabsf(q) * 4.76837158e-7f)) return (long long)__fdiv_rn(t, itv);
    return (long long)q;
}

struct CamMats {
    float m[24];   // inv(post_rots) 9 | post_trans 3 | combine 9 | t_s2e 3
    float bd[12];  // bda 9 | translation 3
};
__device__ __forceinline__ void calib_point_cam(const CamMats &c, int mode, bool bda_has_t, float fx, float fy, float fz,
                                                float &ox, float &oy, float &oz) {
    const float *m = c.m, *bd = c.bd;
    const float x0 = __fsub_rn(fx, m[9]), y0 = __fsub_rn(fy, m[10]), z0 = __fsub_rn(fz, m[11]);
    const float x1 = dot3(mode, m[0], m[1], m[2], x0, y0, z0);
    const float y1 = dot3(mode, m[3], m[4], m[5], x0, y0, z0);
    const float z1 = dot3(mode, m[6], m[7], m[8], x0, y0, z0);
    const float x2 = __fmul_rn(x1, z1), y2 = __fmul_rn(y1, z1);
    const float x3 = __fadd_rn(dot3(mode, m[12], m[13], m[14], x2, y2, z1), m[21]);
    const float y3 = __fadd_rn(dot3(mode, m[15], m[16], m[17], x2, y2, z1), m[22]);
    const float z3 = __fadd_rn(dot3(mode, m[18], m[19], m[20], x2, y2, z1), m[23]);
    ox = dot3(mode, bd[0], bd[1], bd[2], x3, y3, z3);
    oy = dot3(mode, bd[3], bd[4], bd[5], x3, y3, z3);
    oz = dot3(mode, bd[6], bd[7], bd[8], x3, y3, z3);
    if (bda_has_t) { ox = __fadd_rn(ox, bd[9]); oy = __fadd_rn(oy, bd[10]); oz = __fadd_rn(oz, bd[11]); }
}

// ------------------------------------------------------------------------------------------------
// A1: voxel id of every point, chunk histogram, (last CTA) chunk offsets and class lists.
// grid = (blocks per camera, B * N): a CTA's points belong to ONE camera, so the calibration is read once per
// thread (36 uniform loads) instead of once per point, and the point -> (camera, frustum index) division disappears.
// ------------------------------------------------------------------------------------------------
template <bool CALIB>
__global__ void __launch_bounds__(kChunkThreads, 3) chunk_voxelize_kernel(ChunkArgs a) {
    extern __shared__ int s_hist[];                      // [cps]
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int bn = blockIdx.y;
    const int b = bn / a.n_cams;
    const VoxArgs &v = a.vox;
    if (blockIdx.x == 0 && bn == 0 && tid == 0) {
        v.hdr->flags = 0;
        v.hdr->n_subs = v.n_subs;
        v.hdr->subs_per_sample = v.subs_per_sample;
        v.hdr->structured = 1;
        v.hdr->fwd_heavy[0] = v.hdr->fwd_heavy[1] = v.hdr->fwd_heavy[2] = 0;
    }
    for (int i = tid; i < a.g.cps; i += kChunkThreads) s_hist[i] = 0;
    CamMats cm;
    if (CALIB) {
#pragma unroll
        for (int i = 0; i < 24; ++i) cm.m[i] = __ldg(a.calib.cam + 24 * bn + i);
#pragma unroll
        for (int i = 0; i < 12; ++i) cm.bd[i] = __ldg(a.calib.bda + 12 * b + i);
    }
    __syncthreads();
    const int dhw = a.dhw_pts;
    const int p_cam = bn * dhw;                          // first point of this camera (P < 2^31 is checked on the host)
    const int c0 = blockIdx.x * kChunkBlockPts;
    const int bV = (int)((int64_t)b * a.g.V);
    const float rx = __frcp_rn(v.ivx), ry = __frcp_rn(v.ivy), rz = __frcp_rn(v.ivz);
#pragma unroll
    for (int k = 0; k < kChunkPtsPerThread; ++k) {
        const int r = c0 + tid + k * kChunkThreads;
        if (r >= dhw) continue;
        const int p = p_cam + r;
        float x, y, z;
        if (CALIB) {
            const float *f = a.calib.frustum + 3 * r;
            calib_point_cam(cm, a.calib.mode, a.calib.bda_has_t != 0, __ldg(f), __ldg(f + 1), __ldg(f + 2), x, y, z);
            if (a.calib.coor_out) {
                float *o = a.calib.coor_out + 3 * (int64_t)p;
                o[0] = x; o[1] = y; o[2] = z;
            }
        } else {
            const float *cc = v.coor + 3 * (int64_t)p;
            x = __ldcs(cc); y = __ldcs(cc + 1); z = __ldcs(cc + 2);
        }
        const long long ix = axis_index(x, v.lbx, v.ivx, rx), iy = axis_index(y, v.lby, v.ivy, ry),
                        iz = axis_index(z, v.lbz, v.ivz, rz);
        int key = -1;
        if (ix >= 0 && ix < v.X && iy >= 0 && iy < v.Y && iz >= 0 && iz < v.Z) {
            const int vin = ((int)iz * v.Y + (int)iy) * v.X + (int)ix;
            key = bV + vin;
            atomicAdd(&s_hist[vin >> 10], 1);
        }
        v.key[p] = key;
    }
    __syncthreads();
    for (int i = tid; i < a.g.cps; i += kChunkThreads) {
        const int c = s_hist[i];
        if (c) atomicAdd(a.hist + b * a.g.cps + i, c);
    }
    // ---- the last CTA to get here scans the chunk histogram
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(a.ctrl + 1, 1) == (int)(gridDim.x * gridDim.y) - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int total = cta_exclusive_scan(a.hist, a.chunk_off, a.g.n_chunks);
    if (tid == 0) a.counts[0] = total;
    // the chunks that need more than one warp, by class (order within a list does not matter)
    for (int i0 = 0; i0 < a.g.n_chunks; i0 += kChunkThreads * 8) {
        int c[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = (i0 + tid + k * kChunkThreads < a.g.n_chunks) ? __ldcg(a.hist + i0 + tid + k * kChunkThreads) : 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (c[k] > kClassM) a.list_l[atomicAdd(a.ctrl + 5, 1)] = i0 + tid + k * kChunkThreads;
            else if (c[k] > kClassS) a.list_m[atomicAdd(a.ctrl + 4, 1)] = i0 + tid + k * kChunkThreads;
        }
    }
}
static_assert(kChunkVox == 1024, "the kernels shift by 10");

// ------------------------------------------------------------------------------------------------
// The same camera-major point loop feeding the global bucket sort of bucket_sort.cuh (its count pass): voxel id +
// arrival slot.  PPT points per thread, CTAs of 256 threads = 256 * PPT consecutive frustum points of ONE camera.
// ------------------------------------------------------------------------------------------------
#ifndef FO_VOXCAM_PPT
#define FO_VOXCAM_PPT 2
#endif
#ifndef FO_VOXCAM_MINB
#define FO_VOXCAM_MINB 6
#endif
// MODE: the matvec summation order (CalibArgs::mode) as a compile-time constant — with a run-time mode every one of
// the nine 3-term dot products carried its own four-way branch (62 BRA, ~250 instructions per point; the kernel is
// issue-bound).
template <bool CALIB, int MODE>
__global__ void __launch_bounds__(kChunkThreads, FO_VOXCAM_MINB) voxelize_count_cam_kernel(VoxArgs v, CalibArgs g, int n_cams,
                                                                                        int dhw) {
    constexpr int PPT = FO_VOXCAM_PPT;
    pdl_wait();
    pdl_launch();
    const int tid = threadIdx.x;
    const int bn = blockIdx.y;
    const int b = bn / n_cams;
    if (blockIdx.x == 0 && bn == 0 && tid == 0) {
        v.hdr->flags = 0;
        v.hdr->n_subs = v.n_subs;
        v.hdr->subs_per_sample = v.subs_per_sample;
        v.hdr->structured = 1;
        v.hdr->fwd_heavy[0] = v.hdr->fwd_heavy[1] = v.hdr->fwd_heavy[2] = 0;
    }
    CamMats cm;
    if (CALIB) {
#pragma unroll
        for (int i = 0; i < 24; ++i) cm.m[i] = __ldg(g.cam + 24 * bn + i);
#pragma unroll
        for (int i = 0; i < 12; ++i) cm.bd[i] = __ldg(g.bda + 12 * b + i);
    }
    const int p_cam = bn * dhw;
    const int c0 = blockIdx.x * (kChunkThreads * PPT);
    const int bV = b * (v.X * v.Y * v.Z);
    const float rx = __frcp_rn(v.ivx), ry = __frcp_rn(v.ivy), rz = __frcp_rn(v.ivz);
    int key[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int r = c0 + tid + k * kChunkThreads;
        key[k] = -2;
        if (r >= dhw) continue;
        const int p = p_cam + r;
        float x, y, z;
        if (CALIB) {
            const float *f = g.frustum + 3 * r;
            calib_point_cam(cm, MODE, g.bda_has_t != 0, __ldg(f), __ldg(f + 1), __ldg(f + 2), x, y, z);
            if (g.coor_out) {
                float *o = g.coor_out + 3 * (int64_t)p;
                o[0] = x; o[1] = y; o[2] = z;
            }
        } else {
            const float *cc = v.coor + 3 * (int64_t)p;
            x = __ldcs(cc); y = __ldcs(cc + 1); z = __ldcs(cc + 2);
        }
        const long long ix = axis_index(x, v.lbx, v.ivx, rx), iy = axis_index(y, v.lby, v.ivy, ry),
                        iz = axis_index(z, v.lbz, v.ivz, rz);
        key[k] = (ix >= 0 && ix < v.X && iy >= 0 && iy < v.Y && iz >= 0 && iz < v.Z)
                     ? bV + ((int)iz * v.Y + (int)iy) * v.X + (int)ix : -1;
    }
    int slot[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) slot[k] = key[k] >= 0 ? atomicAdd(v.cnt + key[k], 1) : 0;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        if (key[k] == -2) continue;
        const int p = p_cam + c0 + tid + k * kChunkThreads;
        v.key[p] = key[k];
        v.slot[p] = slot[k];
    }
}

// ------------------------------------------------------------------------------------------------
// A2: multisplit of the kept points into per-chunk lists.   grid = (blocks per camera, B * N)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kChunkThreads) chunk_scatter_kernel(ChunkArgs a) {
    extern __shared__ int s_bins[];                      // [cps] counts, then list base of this CTA's run per chunk
    const int tid = threadIdx.x;
    const int bn = blockIdx.y;
    const int b = bn / a.n_cams;
    for (int i = tid; i < a.g.cps; i += kChunkThreads) s_bins[i] = 0;
    __syncthreads();
    const int p_lo = bn * a.dhw_pts;
    const int c0 = blockIdx.x * kChunkBlockPts;
    const int bV = (int)((int64_t)b * a.g.V);
    int key[kChunkPtsPerThread], rank[kChunkPtsPerThread];
#pragma unroll
    for (int k = 0; k < kChunkPtsPerThread; ++k) {
        const int r = c0 + tid + k * kChunkThreads;
        key[k] = r < a.dhw_pts ? __ldg(a.vox.key + p_lo + r) : -1;
    }
#pragma unroll
    for (int k = 0; k < kChunkPtsPerThread; ++k) {
        rank[k] = 0;
        if (key[k] >= 0) rank[k] = atomicAdd(&s_bins[(key[k] - bV) >> 10], 1);
    }
    __syncthreads();
    for (int i = tid; i < a.g.cps; i += kChunkThreads) {
        const int c = s_bins[i];
        if (c) s_bins[i] = __ldg(a.chunk_off + b * a.g.cps + i) + atomicAdd(a.cursor + b * a.g.cps + i, c);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kChunkPtsPerThread; ++k)
        if (key[k] >= 0)
            a.list[s_bins[(key[k] - bV) >> 10] + rank[k]] = make_int2(key[k], p_lo + c0 + tid + k * kChunkThreads);
}

// ------------------------------------------------------------------------------------------------
// D1: non-empty voxels per chunk (one warp per chunk), (last CTA) their exclusive prefix = global interval numbers.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kChunkThreads) chunk_distinct_kernel(ChunkArgs a) {
    __shared__ unsigned s_map[kChunkThreads / 32][32];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ch = blockIdx.x * (kChunkThreads / 32) + warp;
    if (ch < a.g.n_chunks) {
        const int b = ch / a.g.cps, c = ch - b * a.g.cps;
        const int v_lo = (int)((int64_t)b * a.g.V) + c * kChunkVox;
        const int base = __ldg(a.chunk_off + ch), n = __ldg(a.chunk_off + ch + 1) - base;
        s_map[warp][lane] = 0u;
        __syncwarp();
        for (int i0 = 0; i0 < n; i0 += 512) {                    // sixteen loads in flight per lane
            int lv[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) lv[k] = (i0 + lane + 32 * k < n) ? __ldg(&a.list[base + i0 + lane + 32 * k].x) - v_lo : -1;
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (lv[k] >= 0) atomicOr(&s_map[warp][lv[k] >> 5], 1u << (lv[k] & 31));
        }
        __syncwarp();
        int ne = __popc(s_map[warp][lane]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ne += __shfl_xor_sync(0xffffffffu, ne, o);
        if (lane == 0) a.iv_off[ch] = ne;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(a.ctrl + 2, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int total = cta_exclusive_scan(a.iv_off, a.iv_off, a.g.n_chunks);     // in place: thread i reads its own element first
    if (tid == 0) {
        a.counts[1] = total;
        a.hdr->n_intervals = total;
        a.sub_iv[a.n_subs] = total;
        a.sub_pt[a.n_subs] = __ldcg(a.chunk_off + a.g.n_chunks);
    }
}

// ------------------------------------------------------------------------------------------------
// D2: NW warps sort one chunk.  NW == 1: four independent warps per CTA, warp-level synchronisation only.
// ------------------------------------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void chunk_sync() {
    if (NW == 1) __syncwarp();
    else __syncthreads();
}

// CACHE > 0: the class guarantees n <= CACHE * 32 * NW, every thread keeps its points in registers between the count
// and the placement pass (one list round trip per chunk instead of two)
template <int NW, int STAGE, int CACHE>
__device__ __forceinline__ void chunk_sort(const ChunkArgs &a, const int ch, int *smem, const int t /*thread within the group*/) {
    constexpr int T = 32 * NW;
    int *cnt = smem;                                             // [kChunkCntWords] counts -> cursors -> end offsets
    int *st_p = smem + kChunkCntWords;                           // [STAGE] point index at staged position
    unsigned short *st_v = reinterpret_cast<unsigned short *>(st_p + STAGE);      // [STAGE] voxel - v_lo
    int2 *wl_long = reinterpret_cast<int2 *>(st_v + STAGE);      // [long_cap] (start, len), len > 8
    int *wl_short = reinterpret_cast<int *>(wl_long + long_cap(STAGE));           // [short_cap] (start << 3 | len - 1), len 2..8
    int *wl_n = wl_short + short_cap(STAGE);                     // [2] list lengths
    const int lane = t & 31, wid = t >> 5;

    const int b = ch / a.g.cps, c = ch - b * a.g.cps;
    const int vin0 = c * kChunkVox;
    const int v_lo = (int)((int64_t)b * a.g.V) + vin0;
    const int nv = (int)min((int64_t)kChunkVox, a.g.V - vin0);
    const int pt_base = __ldg(a.chunk_off + ch);
    const int n = __ldg(a.chunk_off + ch + 1) - pt_base;
    const int iv_base = __ldg(a.iv_off + ch);
    const int2 *list = a.list + pt_base;
    const bool staged = n <= STAGE;

    // ---- zero the counters
    for (int i = t; i < kChunkCntWords / 4; i += T) reinterpret_cast<int4 *>(cnt)[i] = make_int4(0, 0, 0, 0);
    if (t == 0) { wl_n[0] = 0; wl_n[1] = 0; }
    chunk_sync<NW>();
    // ---- count
    int2 mine[CACHE > 0 ? CACHE : 1];
    if (CACHE > 0) {
#pragma unroll
        for (int k = 0; k < (CACHE > 0 ? CACHE : 1); ++k)
            mine[k] = (t + T * k < n) ? __ldg(&list[t + T * k]) : make_int2(-1, 0);
#pragma unroll
        for (int k = 0; k < (CACHE > 0 ? CACHE : 1); ++k)
            if (mine[k].x >= 0) atomicAdd(&cnt[chunk_pad(mine[k].x - v_lo)], 1);
    } else {
        for (int i0 = 0; i0 < n; i0 += 8 * T) {                  // eight loads in flight per thread
            int lv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) lv[k] = (i0 + t + T * k < n) ? __ldg(&list[i0 + t + T * k].x) - v_lo : -1;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (lv[k] >= 0) atomicAdd(&cnt[chunk_pad(lv[k])], 1);
        }
    }
    chunk_sync<NW>();
    // ---- scan by the owner warp: lane l owns sub-tile l = voxels [32 l, 32 l + 32); counters -> cursors
    int my_off = 0, my_ne = 0;
    if (wid == 0) {
        int4 x[8];
        const int4 *src = reinterpret_cast<const int4 *>(cnt + 36 * lane);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = src[i];
        int pts = 0, ne = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            pts += x[i].x + x[i].y + x[i].z + x[i].w;
            ne += (x[i].x > 0) + (x[i].y > 0) + (x[i].z > 0) + (x[i].w > 0);
        }
        int ip = pts, in = ne;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int xp = __shfl_up_sync(0xffffffffu, ip, o), xn = __shfl_up_sync(0xffffffffu, in, o);
            if (lane >= o) { ip += xp; in += xn; }
        }
        my_off = ip - pts; my_ne = in - ne;
        int run = my_off;
        int4 *dst = reinterpret_cast<int4 *>(cnt + 36 * lane);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int4 o;
            o.x = run; run += x[i].x;
            o.y = run; run += x[i].y;
            o.z = run; run += x[i].z;
            o.w = run; run += x[i].w;
            dst[i] = o;
        }
    }
    chunk_sync<NW>();
    // ---- placement (arrival order inside a voxel is arbitrary; ordered below)
    int *gp = a.rd + pt_base, *gv = a.rb + pt_base;              // global staging of a chunk that does not fit the stage
    if (CACHE > 0) {
#pragma unroll
        for (int k = 0; k < (CACHE > 0 ? CACHE : 1); ++k) {
            if (mine[k].x < 0) continue;
            const int pos = atomicAdd(&cnt[chunk_pad(mine[k].x - v_lo)], 1);
            st_p[pos] = mine[k].y; st_v[pos] = (unsigned short)(mine[k].x - v_lo);
        }
    } else {
        for (int i0 = 0; i0 < n; i0 += 8 * T) {
            int2 e[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) e[k] = (i0 + t + T * k < n) ? __ldg(&list[i0 + t + T * k]) : make_int2(-1, 0);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (e[k].x < 0) continue;
                const int pos = atomicAdd(&cnt[chunk_pad(e[k].x - v_lo)], 1);
                if (staged) { st_p[pos] = e[k].y; st_v[pos] = (unsigned short)(e[k].x - v_lo); }
                else { gp[pos] = e[k].y; gv[pos] = e[k].x; }
            }
        }
    }
    chunk_sync<NW>();
    // ---- per-voxel outputs by the owner warp: the cursors now are END offsets, so start(v) = end(v - 1)
    int *seg_base = staged ? st_p : gp;
    if (wid == 0 && 32 * lane < nv) {
        const int lv0 = 32 * lane;
        const int64_t u = (int64_t)b * a.g.sps_sub + ((vin0 + lv0) >> kSubShift);
        a.sub_iv[u] = iv_base + my_ne;
        a.sub_pt[u] = pt_base + my_off;
        int prev = my_off, ne = my_ne;
        const bool vec = lv0 + 32 <= nv && ((v_lo + lv0) & 3) == 0;
        const int4 *src = reinterpret_cast<const int4 *>(cnt + 36 * lane);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
            const int4 e4 = src[i4];
            const int ends[4] = {e4.x, e4.y, e4.z, e4.w};
            int ivs[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ci = ends[e] - prev;
                ivs[e] = iv_base + ne;
                if (ci > 0) {
                    const int k = iv_base + ne;
                    a.iv_starts[k] = pt_base + prev;
                    a.iv_lengths[k] = ci;
                    a.iv_vox[k] = v_lo + lv0 + 4 * i4 + e;
                    if (ci > 1 && staged) {                      // (a chunk too dense for the stage is ordered voxel by voxel below)
                        if (ci <= kLaneSortMax) wl_short[atomicAdd(&wl_n[0], 1)] = (prev << 3) | (ci - 1);
                        else wl_long[atomicAdd(&wl_n[1], 1)] = make_int2(prev, ci);
                    }
                    prev = ends[e]; ++ne;
                }
            }
            // interval id of every voxel (dense: 128 contiguous bytes per lane)
            if (vec) reinterpret_cast<int4 *>(a.vox2iv + v_lo + lv0)[i4] = make_int4(ivs[0], ivs[1], ivs[2], ivs[3]);
            else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (lv0 + 4 * i4 + e < nv) a.vox2iv[v_lo + lv0 + 4 * i4 + e] = ivs[e];
            }
        }
        if (prev - my_off > a.heavy_pts) a.heavy_list[atomicAdd(a.hdr->fwd_heavy, 1)] = (int)u;
    }
    chunk_sync<NW>();
    // ---- order inside the voxels: 2..8 points one lane each (sorting network in registers) ...
    if (staged) {
        const int ns = wl_n[0];
        for (int i = t; i < ns; i += T) {
            const int w = wl_short[i];
            const int st = w >> 3, len = (w & 7) + 1;
            int w8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) w8[k] = (k < len) ? st_p[st + k] : INT_MAX;
            sort8(w8);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < len) st_p[st + k] = w8[k];
        }
        // ... longer ones one warp each (register networks up to 128 points, an in-memory network above that)
        const int nl = wl_n[1];
        for (int i = wid; i < nl; i += NW) {
            const int2 w = wl_long[i];
            warp_sort_segment(st_p + w.x, w.y, nullptr, lane);
            __syncwarp();
        }
    } else {
        // degenerate density (more points in 1024 voxels than the stage holds): one warp per voxel, in global memory
        for (int v = wid; v < nv; v += NW) {
            const int end = cnt[chunk_pad(v)], start = (v == 0) ? 0 : cnt[chunk_pad(v - 1)];
            if (end - start > 1) warp_sort_segment(gp + start, end - start, nullptr, lane);
            __syncwarp();
        }
    }
    chunk_sync<NW>();
    // ---- the three rank arrays, coalesced
    if (staged) {
        for (int i = t; i < n; i += T) {
            const int p = st_p[i];
            a.rd[pt_base + i] = p;
            a.rf[pt_base + i] = feat_row_of(p, a.dhw, a.hw);
            a.rb[pt_base + i] = v_lo + (int)st_v[i];
        }
    } else {
        for (int i = t; i < n; i += T) a.rf[pt_base + i] = feat_row_of(gp[i], a.dhw, a.hw);
    }
}

__host__ __device__ constexpr size_t chunk_sort_smem(int stage) {
    return (size_t)kChunkCntWords * 4 + (size_t)stage * 6 + (size_t)long_cap(stage) * 8 + (size_t)short_cap(stage) * 4 + 16;
}

// one warp per chunk, chunks of up to kClassS points (the bulk); 4 chunks per CTA
__global__ void __launch_bounds__(128) chunk_sort_small_kernel(ChunkArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int warp = threadIdx.x >> 5;
    const int ch = blockIdx.x * 4 + warp;
    if (ch >= a.g.n_chunks) return;
    if (__ldg(a.hist + ch) > kClassS) return;
    constexpr size_t per = (chunk_sort_smem(kClassS) + 15) / 16 * 16;
    chunk_sort<1, kClassS, kClassS / 32>(a, ch, reinterpret_cast<int *>(s_raw + warp * per), threadIdx.x & 31);
}
// NW warps per chunk: the medium (LARGE = false) and the dense (LARGE = true) class; CTAs loop over the class list
template <int NW, int STAGE, bool LARGE>
__global__ void __launch_bounds__(32 * NW) chunk_sort_cta_kernel(ChunkArgs a) {
    constexpr int CACHE = LARGE ? 0 : STAGE / (32 * NW);
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int n_listed = __ldg(a.ctrl + (LARGE ? 5 : 4));
    const int32_t *lst = LARGE ? a.list_l : a.list_m;
    for (int i = blockIdx.x; i < n_listed; i += gridDim.x) {
        chunk_sort<NW, STAGE, CACHE>(a, __ldg(lst + i), reinterpret_cast<int *>(s_raw), threadIdx.x);
        __syncthreads();                                         // the shared arrays are reused by the next chunk
    }
}

}  // namespace fo
