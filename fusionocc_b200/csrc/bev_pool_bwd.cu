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
#include <stdlib.h>

#include "bwd_plan.cuh"
#include "common.cuh"
#include "tma.cuh"

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
    const uint32_t *sub_mask;       // occupancy mask per sub-tile (nullptr / 0: unknown — read iv_vox, fetch the whole sub-tile)
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


// =================================================================================================
// Round-2 backward: one launch, two roles, no plan-sized or G-sized round trip through DRAM.
//
// (1) PIXEL ROLE, several pixels per warp.  A gathered out_grad row is C floats; with 128-bit accesses it needs
//     only LPR = C/4 lanes (8 for C = 32), so one warp walks PPW = 32/LPR pixels in lock step and every row-loop
//     instruction moves PPW rows (4 x 128 bytes at C = 32) — the per-point instruction count drops ~3x against the
//     one-pixel-per-warp kernel above, which was latency/issue-bound (28 M warp instructions, 2.3 TB/s).
//     feat_grad[q][4s..4s+3] stays ONE sequential FMA chain over the pixel's points in plan order in one thread,
//     depth_grad[p] ONE sequential chain over c in one thread (bev_pool_cuda.cu:96-120): same bits as before.
// (2) GATHER ROLE with TMA.  The C x 32 block of out_grad of a sub-tile is a regular 2-D tile: one
//     cp.async.bulk.tensor (3-D map (V, C, B), box 32 x C x 1, 128-byte swizzle) per warp replaces 8 LDG.128 +
//     8 STS.128 per lane and their address arithmetic, and is in flight while the warp fetches its interval list.
// (3) NOT KEPT (second session, profiles/r02_summary.md §8): the rows through cp.async (LDGSTS) into a ring of three
//     tiles per warp, two tiles ahead of the consumer, no row registers: backward 192 -> 278 us with .cg copies
//     (every repeated row of a pixel — 1.5 points per voxel — goes back to L2), 211 us with .ca; ncu: 12 warps per SM,
//     73 % of the stall cycles on the copies' scoreboard, 1.3 TB/s.  And an upper bound for moving depth_grad out of
//     this kernel: with the tile stores and the depth-grad chains compiled OUT the kernel is no
//     faster (backward 197.7 vs 192.3 us, 72 registers, 4 or 6 CTAs per SM) — its time is the two dependent index
//     round trips and the scattered 128-byte row reads, not the chains.  Nor is it bytes in flight: a second register
//     set holding the rows of the tile after next (16 rows in flight per lane, 143 registers, 14 warps per SM) made it
//     SLOWER, 182.3 -> 189.9 us per backward (405 -> 446 us at 512x1408); ncu of the shipped kernel: L1/TEX throughput
//     72 %, DRAM 42 %, 14 warps per SM, first use of a tile's rows = 38 % of the stall samples.
// (4) NOT KEPT: one fused launch (gather units of sample b+1 interleaved with pixel units of sample b, device-side
//     completion counters, G consumed while L2-resident).  Measured on a B200 at the headline shape, batch 8
//     (profiles/r02_summary.md): two launches 163 us; fused and interleaved 225 us, with a two-stage box ring per
//     gather warp 195 us, same kernel with all gather units first 184 us; spin waits were negligible (max 42 polls).
//     The roles want different resources — the gather 6 CTAs x 8 warps at 24 registers to keep 6.5 TB/s of DRAM
//     reads in flight, the pixel role 96 registers — and one kernel can only have the worse of both.
// =================================================================================================
constexpr int kPix2Threads = 128;            // threads per CTA of the multi-pixel kernel (short CTAs: finer tail)
constexpr int kPix2Warps   = kPix2Threads / 32;
#ifndef FO_REC_CAP
#define FO_REC_CAP 64
#endif
constexpr int kRecCap      = FO_REC_CAP;     // records of one pixel staged per chunk (longer pixels take several chunks); multiple of 32

struct Pixel2Cfg {
    int32_t tile_stride;   // floats per row of the per-warp tile rows[32][tile_stride]: C + 4 or C + 8 (conflict-free LDS.128)
    int32_t warp_floats;   // floats of shared memory per warp
};
__host__ __device__ inline Pixel2Cfg pixel2_cfg(int C, int LPR) {
    Pixel2Cfg c;
    const int ppw = 32 / LPR;
    c.tile_stride = ((C >> 2) & 1) ? C + 8 : C + 4;
    // tile | feat rows of the PPW pixels | records (G row, depth bits) | depth indices
    c.warp_floats = 32 * c.tile_stride + ppw * C + 2 * ppw * kRecCap + ppw * kRecCap;
    return c;
}

// One warp, PPW = 32 / LPR backward intervals m0 .. m0 + PPW - 1 (lane group g = lane / LPR owns interval m0 + g).
//   phase 1  ALL records of the pixels (up to kRecCap per chunk) are fetched at once — entry, then depth value:
//            two dependent round trips per pixel instead of two per 8 points — and parked in shared memory;
//   phase 2  tiles of LPR points per pixel: the gathered rows go into feat_grad's FMA chain and into the tile;
//            the rows of the NEXT tile are requested before the depth-grad chains of the current one run.
// FULL: C == 4 * LPR (every lane carries four channels).
template <int LPR, bool FULL>
__device__ __forceinline__ void pixel_warp(const PixelArgs &a, float *wsm, const int m0, const int n, const int lane) {
    constexpr int PPW = 32 / LPR;
    constexpr int U = LPR < 8 ? LPR : 8;          // rows in flight per lane (128-bit each)
    const int g = lane / LPR, s = lane - g * LPR;
    const int C = FULL ? 4 * LPR : a.C, c4 = C >> 2;
    const Pixel2Cfg cfg = pixel2_cfg(C, LPR);
    const int S = cfg.tile_stride;
    float *rows = wsm;                                           // [32][S]
    float *fs = rows + 32 * S;                                   // [PPW][C]
    int2 *rec = reinterpret_cast<int2 *>(fs + PPW * C) + g * kRecCap;             // [PPW][kRecCap] (G row | -1, depth bits)
    int *recp = reinterpret_cast<int *>(fs + PPW * C + 2 * PPW * kRecCap) + g * kRecCap;   // [PPW][kRecCap] depth index
    const unsigned rows_s = (unsigned)__cvta_generic_to_shared(rows);
    const unsigned fs_s = (unsigned)__cvta_generic_to_shared(fs + g * C);
    const bool chan_ok = FULL || 4 * s < C;
    const int n_depth = (int)a.n_depth, n_rows_G = (int)a.n_rows_G, n_iv = (int)a.n_iv;

    const int m = m0 + g;
    int st = 0, len = 0, q = -1;
    if (m < n) {
        st = __ldg(a.bwd_starts + m); len = __ldg(a.bwd_lengths + m); q = __ldg(a.bwd_ids + m);
        if (len <= 0 || st < 0 || (int64_t)st + len > a.n_entries || q < 0 || q >= a.n_feat_rows) { len = 0; q = -1; }
    }
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    if (maxlen == 0) return;                                     // (feat_grad / depth_grad were zero-filled)
    const int32_t *ep = a.ent_p + st, *ei = a.ent_iv + st;
    // C == 32: every lane keeps its pixel's whole feature row in registers (the depth-grad chains then read only the
    // tile: the broadcast LDS.128 of the row were a third of the kernel's shared-memory wavefronts, and the L1 data
    // pipe is what bounds it — ncu, profiles/r02_summary.md); wider rows stay in shared memory
    constexpr bool FREG = FULL && LPR == 8;
    float4 fr[FREG ? LPR : 1];
    if (FREG) {
#pragma unroll
        for (int i = 0; i < (FREG ? LPR : 1); ++i)
            fr[i] = q >= 0 ? ldg4(a.feat + q * C + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (q >= 0 && chan_ok) {
        reinterpret_cast<float4 *>(fs + g * C)[s] = ldg4(a.feat + q * C + 4 * s);
    }
    float4 fg = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *Gs = a.G + 4 * s;
    const unsigned tile_col = rows_s + 4u * (unsigned)((g * LPR) * S + 4 * s);   // row g*LPR, my 4 channels
    auto load_row = [&](int row) -> float4 {
        const float *src = Gs + max(row, 0) * a.g_rowstride;     // invalid records read row 0 and are discarded
        if (!chan_ok) return make_float4(0.f, 0.f, 0.f, 0.f);
        return ldg4(src);
    };

    for (int base = 0; base < maxlen; base += kRecCap) {
        const int cnt = min(kRecCap, maxlen - base);             // warp-uniform
        __syncwarp();                                            // the previous chunk is done with rec / recp
        // ---- phase 1: records of this chunk.  RB batches of LPR entries per lane group = the whole chunk in ONE pass:
        //      all entry loads, then all depth loads (two dependent round trips per chunk; with RB = 4 at C = 32 a
        //      50-point pixel took two passes = four round trips)
#ifndef FO_PIX_RB
        constexpr int RB = kRecCap / LPR;
#else
        constexpr int RB = FO_PIX_RB;
#endif
        for (int i0 = 0; i0 < cnt; i0 += LPR * RB) {
            int p[RB], row[RB];
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                const int i = base + i0 + s + LPR * k;
                p[k] = -1; row[k] = -1;
                if (i < len) { p[k] = __ldg(ep + i); row[k] = __ldg(ei + i); }
            }
            if (a.row_map) {
#pragma unroll
                for (int k = 0; k < RB; ++k) row[k] = ((unsigned)row[k] < (unsigned)n_iv) ? __ldg(a.row_map + row[k]) : -1;
            }
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                const bool ok = (unsigned)p[k] < (unsigned)n_depth && (unsigned)row[k] < (unsigned)n_rows_G;
                const float d = ok ? __ldg(a.depth + p[k]) : 0.f;
                const int i = i0 + s + LPR * k;
                if (i < kRecCap) {
                    rec[i] = make_int2(ok ? row[k] : -1, __float_as_int(d));
                    recp[i] = ok ? p[k] : -1;
                }
            }
        }
        __syncwarp();
        // ---- phase 2: tiles of LPR points per pixel, rows of the next tile in flight during the chains
        float4 x[U];
        int2 r[U];
#pragma unroll
        for (int t = 0; t < U; ++t) { r[t] = rec[t]; x[t] = load_row(r[t].x); }
        for (int j0 = 0; j0 < cnt; j0 += LPR) {
#pragma unroll
            for (int t0 = 0; t0 < LPR; t0 += U) {
                if (t0 > 0) {                                    // LPR > U: later row groups of this tile
#pragma unroll
                    for (int t = 0; t < U; ++t) { r[t] = rec[j0 + t0 + t]; x[t] = load_row(r[t].x); }
                }
#pragma unroll
                for (int t = 0; t < U; ++t) {
                    if (chan_ok) sts_f4(tile_col + 4u * (unsigned)((t0 + t) * S), x[t]);
                    const float d = __int_as_float(r[t].y);
                    const float4 y = make_float4(fmaf(x[t].x, d, fg.x), fmaf(x[t].y, d, fg.y), fmaf(x[t].z, d, fg.z),
                                                 fmaf(x[t].w, d, fg.w));
                    if (r[t].x >= 0) fg = y;
                }
            }
            const int my_p = recp[j0 + s];
            if (j0 + LPR < cnt) {                                // first row group of the next tile
#pragma unroll
                for (int t = 0; t < U; ++t) { r[t] = rec[j0 + LPR + t]; x[t] = load_row(r[t].x); }
            }
            __syncwarp();
            // depth grad: lane owns tile row `lane` = its own record; one sequential chain over all C channels
            if (my_p >= 0) {
                const unsigned ra = rows_s + 4u * (unsigned)(lane * S);
                float sum = 0.f;
                if (FREG) {                                      // the pixel's feature row lives in registers
#pragma unroll
                    for (int i = 0; i < (FREG ? LPR : 1); ++i) {
                        const float4 gg = lds_f4(ra + 16u * i);
                        sum = fmaf(gg.x, fr[i].x, sum);
                        sum = fmaf(gg.y, fr[i].y, sum);
                        sum = fmaf(gg.z, fr[i].z, sum);
                        sum = fmaf(gg.w, fr[i].w, sum);
                    }
                } else {
#pragma unroll 4
                    for (int i = 0; i < c4; ++i) {
                        const float4 gg = lds_f4(ra + 16u * i), f = lds_f4(fs_s + 16u * i);
                        sum = fmaf(gg.x, f.x, sum);
                        sum = fmaf(gg.y, f.y, sum);
                        sum = fmaf(gg.z, f.z, sum);
                        sum = fmaf(gg.w, f.w, sum);
                    }
                }
                a.depth_grad[my_p] = sum;
            }
            __syncwarp();                                        // the tile is free for the next rows
        }
    }
    if (q >= 0 && chan_ok) *reinterpret_cast<float4 *>(a.feat_grad + q * C + 4 * s) = fg;
}

#ifndef FO_PIX2_MINB
#define FO_PIX2_MINB 4
#endif
template <int LPR, bool FULL>
__global__ void __launch_bounds__(kPix2Threads, FO_PIX2_MINB) bwd_pixel2_kernel(PixelArgs a) {
    extern __shared__ __align__(16) float psm[];
    pdl_wait();
    pdl_launch();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int PPW = 32 / LPR;
    const int n = a.n_bwd_dev ? min(max(*a.n_bwd_dev, 0), (int)a.n_bwd) : (int)a.n_bwd;
    const int m0 = (blockIdx.x * kPix2Warps + warp) * PPW;
    if (m0 >= n) return;
    pixel_warp<LPR, FULL>(a, psm + warp * pixel2_cfg(a.C, LPR).warp_floats, m0, n, lane);
}

// Gather role of one warp: sub-tile su of sample b.  `stage` is 1024-byte aligned, C rows of 128 bytes.
// Split in two so that a warp can have the boxes of several sub-tiles in flight before it consumes the first.
struct GatherTile {
    int ia, ni, my_v;
    int half;                       // 0: whole sub-tile staged (128-byte rows); 1 / 2: only its lower / upper 16 voxels (64-byte rows)
    unsigned mask;                  // occupancy mask of the sub-tile (0: unknown, my_v holds the voxels read from iv_vox)
};
// position of the n-th (0-based) set bit of m, 32 if there is none: binary search on prefix population counts
__device__ __forceinline__ int nth_set_bit(unsigned m, int n) {
    int pos = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1)
        if (__popc(m & ((1u << (pos + step)) - 1u)) <= n) pos += step;
    return ((m >> pos) & 1u) && __popc(m & ((1u << pos) - 1u)) == n ? pos : 32;
}

template <int NACC, bool EXACT>
__device__ __forceinline__ bool gather_issue(const GatherArgs &a, const CUtensorMap *tm, const CUtensorMap *tmh, float *stage,
                                             const unsigned bar, const int b, const int su, const int lane, GatherTile &t) {
    const int C = EXACT ? 32 * NACC : a.C;
    t.ni = 0;
    if (su >= a.sps) return false;
    const int u = b * a.sps + su;
    const int ia = __ldg(a.sub_iv + u), ib = __ldg(a.sub_iv + u + 1);
    const unsigned mask = a.sub_mask ? __ldg(a.sub_mask + u) : 0u;
    if (ib <= ia) return false;
    const int v0 = su << kSubShift;
    // a sub-tile with only one occupied 16-voxel half: fetch 64-byte rows (what the memory system fetches at least)
    t.half = mask == 0u ? 0 : ((mask >> 16) == 0u ? 1 : ((mask & 0xFFFFu) == 0u ? 2 : 0));
    if (lane == 0) {
        if (t.half) {
            mbar_expect_tx(bar, (unsigned)C * 64u);
            tma_load_3d((unsigned)__cvta_generic_to_shared(stage), tmh, v0 + (t.half == 2 ? kSub / 2 : 0), 0, b, bar);
        } else {
            mbar_expect_tx(bar, (unsigned)C * 128u);
            tma_load_3d((unsigned)__cvta_generic_to_shared(stage), tm, v0, 0, b, bar);
        }
    }
    const int nv = (int)min((int64_t)kSub, a.V - v0);
    t.ia = ia;
    t.ni = min(ib - ia, kSub);
    t.my_v = -1;
    t.mask = mask;
    if (mask != 0u) {
        // interval l of the sub-tile is its l-th occupied voxel: no second, dependent index load
        if (lane < t.ni) t.my_v = nth_set_bit(mask, lane);
        if ((unsigned)t.my_v >= (unsigned)nv) t.my_v = -1;
    } else if (lane < t.ni) {
        const int vbase = (int)((int64_t)b * a.V) + v0;
        t.my_v = __ldg(a.iv_vox + ia + lane) - vbase;
        if ((unsigned)t.my_v >= (unsigned)nv) t.my_v = -1;
    }
    return true;
}
template <int NACC, bool EXACT>
__device__ __forceinline__ void gather_emit(const GatherArgs &a, const float *stage, const unsigned bar,
                                            const unsigned parity, const int lane, const GatherTile &t) {
    const int C = EXACT ? 32 * NACC : a.C;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(stage);
    float *dst = a.G + (int64_t)t.ia * C + lane;
    mbar_wait(bar, parity);
    if (t.half) {
        const int vlo = t.half == 2 ? kSub / 2 : 0;
        for (int l = 0; l < t.ni; ++l, dst += C) {
            const int v = __shfl_sync(0xffffffffu, t.my_v, l) - vlo;
            if ((unsigned)v >= (unsigned)(kSub / 2)) continue;       // warp-uniform (never for a consistent plan)
#pragma unroll
            for (int k = 0; k < NACC; ++k)
                if (EXACT || lane + 32 * k < C) dst[32 * k] = lds_f32(sbase + swz64_off(lane + 32 * k, v));
        }
        return;
    }
    // four intervals per trip: the shuffles, the staged reads and the row stores of a trip are independent chains
    for (int l0 = 0; l0 < t.ni; l0 += 4, dst += 4 * C) {
        int v[4];
        float x[4][NACC];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __shfl_sync(0xffffffffu, t.my_v, (l0 + u) & 31);   // lanes >= ni hold -1
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < NACC; ++k)
                x[u][k] = (v[u] >= 0 && (EXACT || lane + 32 * k < C)) ? lds_f32(sbase + swz_off(lane + 32 * k, v[u])) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < NACC; ++k)
                if (v[u] >= 0 && l0 + u < t.ni && (EXACT || lane + 32 * k < C)) dst[u * C + 32 * k] = x[u][k];
    }
}

template <int NACC, bool EXACT>
__global__ void __launch_bounds__(256) bwd_gather_tma_kernel(GatherArgs a, const __grid_constant__ CUtensorMap tm,
                                                             const __grid_constant__ CUtensorMap tmh) {
    extern __shared__ __align__(1024) unsigned char gsm[];
    __shared__ __align__(8) unsigned long long s_bar[8];
    pdl_wait();
    pdl_launch();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (__ldg(&a.hdr->flags) & kFlagUnsorted) return;            // the per-interval gather runs instead
    const int C = EXACT ? 32 * NACC : a.C;
    const unsigned stage_bytes = ((unsigned)C * 128u + 1023u) & ~1023u;
    unsigned char *base = (unsigned char *)(((uintptr_t)gsm + 1023) & ~(uintptr_t)1023);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar[warp]);
    if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
    float *stage = reinterpret_cast<float *>(base + warp * stage_bytes);
    GatherTile t;
    if (gather_issue<NACC, EXACT>(a, &tm, &tmh, stage, bar, blockIdx.y, blockIdx.x * 8 + warp, lane, t))
        gather_emit<NACC, EXACT>(a, stage, bar, 0, lane, t);
}

// The gather with the BACKWARD PLAN riding along: one launch, two kinds of CTAs.  Every k-th CTA of the grid builds the
// inverse interval ordering of 16 image pixels (plan_pixel_bitonic, bwd_plan.cuh: comparison networks in registers,
// issue-bound); the others are gather CTAs (DRAM-bound, 32 % of the issue slots busy).  The plan depends only on the
// forward plan, so there is no ordering between the two kinds; interleaving them in block order puts about one plan
// CTA next to four gather CTAs on every SM for the whole launch.  (As separate launches on two streams the block
// scheduler ran them back to back; as a ninth warp inside every gather CTA the plan warp outlived its CTA's gather
// warps several times over and held their slots — profiles/r02_summary.md.)
struct PlanRideArgs {
    const FwdPlanHeader *fhdr;
    const int32_t *pt2vox, *vox2iv;
    int32_t D, HW, n_rows;
    int32_t n_plan_ctas, n_gather_ctas, gu;   // gu = gather CTAs per sample
    int32_t plan_first;                       // 1: the plan CTAs are the first blocks of the grid instead of interleaved
    int32_t small;                            // 1: (grid size) x (plan CTAs) < 2^31: the interleave map in 32-bit arithmetic
    FastDiv div_gu;                           // division by gu
    FastDiv div_T;                            // division by the grid size (the 64-bit divisions of the map were ~100
                                              // instructions for EVERY warp of the grid: 16 M of the kernel's 72 M)
    BwdPlanHeader *hdr;
    int32_t *ent_p, *ent_iv, *starts, *lengths, *ids;
    const int32_t *n_points_dev;
};
#ifndef FO_PLAN_PIX_PER_WARP
#define FO_PLAN_PIX_PER_WARP 2
#endif
constexpr int kPlanPixPerWarp = FO_PLAN_PIX_PER_WARP;
#ifndef FO_GRIDE_MINB
#define FO_GRIDE_MINB 6      // 40 registers; plan CTAs sort in the stage area, so six CTAs fit the SM's shared memory
#endif

template <int NACC, bool EXACT, int R>
__global__ void __launch_bounds__(256, FO_GRIDE_MINB) bwd_gather_plan_kernel(GatherArgs a, PlanRideArgs p,
                                                                 const __grid_constant__ CUtensorMap tm,
                                                                 const __grid_constant__ CUtensorMap tmh) {
    extern __shared__ __align__(1024) unsigned char gsm[];
    __shared__ __align__(8) unsigned long long s_bar[8];
    pdl_wait();
    pdl_launch();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // a plan CTA sorts in the (otherwise unused) stage area: one kind of shared memory per CTA, six CTAs per SM
    int *s_cmp = reinterpret_cast<int *>(gsm) + warp * (32 * R);
    const long long T = (long long)p.n_plan_ctas + p.n_gather_ctas;
    const int bid = blockIdx.x;
    int plans_before;
    bool is_plan;
    if (p.plan_first) {
        plans_before = min(bid, p.n_plan_ctas);
        is_plan = bid < p.n_plan_ctas;
    } else if (p.small) {
        plans_before = (int)fastdiv((uint32_t)bid * (uint32_t)p.n_plan_ctas, p.div_T);
        is_plan = (int)fastdiv((uint32_t)(bid + 1) * (uint32_t)p.n_plan_ctas, p.div_T) > plans_before;
    } else {
        plans_before = (int)(((long long)bid * p.n_plan_ctas) / T);
        is_plan = (int)(((long long)(bid + 1) * p.n_plan_ctas) / T) > plans_before;
    }
    if (is_plan) {
        if (plans_before == 0 && threadIdx.x == 0) {
            p.hdr->n_bwd_intervals = p.n_rows;
            p.hdr->n_points = p.n_points_dev ? *p.n_points_dev : 0;
            p.hdr->structured = 1;
        }
        const int q0 = (plans_before * 8 + warp) * kPlanPixPerWarp;
        const int32_t *v2i = plan_vox2iv(p.fhdr, p.vox2iv);
#pragma unroll 1
        for (int q = q0; q < min(p.n_rows, q0 + kPlanPixPerWarp); ++q)
            plan_pixel_bitonic<R>(p.pt2vox, v2i, p.D, p.HW, q, s_cmp, lane, p.ent_p, p.ent_iv, p.starts,
                                  p.lengths, p.ids);
        return;
    }
    const int gi = bid - plans_before;
    const int b = (int)fastdiv((uint32_t)gi, p.div_gu), blk = gi - b * p.gu;
    const int C = EXACT ? 32 * NACC : a.C;
    const unsigned stage_bytes = ((unsigned)C * 128u + 1023u) & ~1023u;
    unsigned char *base = (unsigned char *)(((uintptr_t)gsm + 1023) & ~(uintptr_t)1023);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar[warp]);
    if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
    float *stage = reinterpret_cast<float *>(base + warp * stage_bytes);
    GatherTile t;
    if (gather_issue<NACC, EXACT>(a, &tm, &tmh, stage, bar, b, blk * 8 + warp, lane, t))
        gather_emit<NACC, EXACT>(a, stage, bar, 0, lane, t);
}

// Order-agnostic gather for plans whose interval list is not canonical (kFlagUnsorted): one warp per interval,
// G[k, :] = out_grad[b, :, voxel(k)]; intervals without a valid voxel get a zero row (never read: their points
// carry row -1).
__global__ void __launch_bounds__(256) bwd_gather_flagged_kernel(GatherArgs a, int64_t n_intervals) {
    if (!(a.hdr->flags & kFlagUnsorted)) return;
    const int lane = threadIdx.x & 31;
    const int64_t n = min((int64_t)max(a.hdr->n_intervals, 0), n_intervals);
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp0; k < n; k += nwarps) {
        const int v = a.iv_vox[k];
        const int64_t b = v >= 0 ? v / a.V : 0, vin = v >= 0 ? v - b * a.V : 0;
        for (int c = lane; c < a.C; c += 32)
            a.G[k * a.C + c] = v >= 0 ? a.og[b * a.og_bstride + (int64_t)c * a.V + vin] : 0.f;
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
// FO_BWD_IMPL (debug / A-B measurements): 0 = round-1 kernels, 1 (default) = TMA gather + multi-pixel kernel
int bwd_impl_choice() {
    const char *e = getenv("FO_BWD_IMPL");
    return (e && *e) ? atoi(e) : 1;
}

// FO_BWD_RIDE (A/B): 0 = build a requested plan as its own launch, 1 = plan CTAs interleaved with the gather CTAs,
// 2 = plan CTAs first in the gather grid; default (-1): first for short gather grids (up to 8 waves of resident CTAs:
// 82.7 -> 76.5 us per step at batch 1, 128.5 -> 122.1 us at batch 2), interleaved for long ones (batch 8: 398.2 us
// interleaved, 401.9 first; 512x1408: 931.9 / 955.1)
int bwd_ride_choice() {
    const char *e = getenv("FO_BWD_RIDE");
    return (e && *e) ? atoi(e) : -1;
}

int launch_pixel2(const PixelArgs &pa, cudaStream_t stream) {
    const int64_t pixels = pa.n_bwd;
    if (pixels <= 0) return FO_OK;
    const int C = pa.C;
    const int lpr = C <= 32 ? 8 : (C <= 64 ? 16 : 32);
    const int ppw = 32 / lpr;
    const size_t smem = (size_t)kPix2Warps * pixel2_cfg(C, lpr).warp_floats * sizeof(float);
    const int64_t blocks = (pixels + kPix2Warps * ppw - 1) / (kPix2Warps * ppw);
    if (blocks >= INT_MAX) return set_error(FO_ERR_UNSUPPORTED, "too many backward intervals");
#define FO_PIX2(L, F)                                                                                             \
    do {                                                                                                          \
        if (smem > 48 * 1024)                                                                                     \
            FO_CUDA(cudaFuncSetAttribute(bwd_pixel2_kernel<L, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        FO_CUDA(launch_pdl(kPdlPixel, bwd_pixel2_kernel<L, F>, dim3((unsigned)blocks), dim3(kPix2Threads), smem, stream, pa)); \
    } while (0)
    const bool full = C == 4 * lpr;
    if (lpr == 8) { if (full) FO_PIX2(8, true); else FO_PIX2(8, false); }
    else if (lpr == 16) { if (full) FO_PIX2(16, true); else FO_PIX2(16, false); }
    else { if (full) FO_PIX2(32, true); else FO_PIX2(32, false); }
#undef FO_PIX2
    FO_LAUNCH_CHECK("bwd_pixel2_kernel");
    return FO_OK;
}
}  // namespace

namespace {
struct PlanRequest {          // build the (structured) backward plan inside this call
    const int32_t *n_points_dev;
    int32_t hw;
};
int backward_impl(cudaStream_t stream, int32_t c, const float *out_grad, int32_t og_layout, int32_t c_total,
                  int32_t c_offset, const float *depth, const float *feat, int64_t n_points, int64_t n_intervals,
                  int32_t B, int64_t n_vox, int64_t n_depth, int64_t n_feat_rows, float *depth_grad,
                  float *feat_grad, const void *fwd_plan, size_t fwd_plan_bytes, const void *bwd_plan,
                  size_t bwd_plan_bytes, void *scratch, size_t scratch_bytes, const PlanRequest *req = nullptr) {
    FO_CHECK_ARG(c >= 1 && c_offset >= 0 && c_total >= c && c_offset + c <= c_total,
                 "channel slice [%d, %d) does not fit %d channels", c_offset, c_offset + c, c_total);
    FO_CHECK_ARG(c >= 1 && B >= 1 && n_vox >= 1, "c, B and voxels per sample must be positive");
    FO_CHECK_ARG(og_layout == FO_LAYOUT_BCZYX || og_layout == FO_LAYOUT_BZYXC, "unknown og_layout %d", og_layout);
    FO_CHECK_ARG(depth_grad && feat_grad, "NULL gradient output");
    FO_CHECK_ARG(n_depth >= 0 && n_feat_rows >= 1 && n_points >= 0 && n_intervals >= 0, "negative size");
    FO_CHECK_ARG(n_points < INT_MAX && (int64_t)B * n_vox < INT_MAX && n_depth < INT_MAX, "sizes exceed int32 ranks");
    // zero-fill of both gradients (points outside the grid / pixels without points are never written below)
    if ((((uintptr_t)depth_grad | (uintptr_t)feat_grad) & 15) == 0 && n_depth % 4 == 0 && (n_feat_rows * c) % 4 == 0) {
        FO_CUDA(launch_pdl(kPdlZero, zero2_kernel, dim3(grid_for(n_depth / 4 + n_feat_rows * c / 4, 256, 8)), dim3(256), 0, stream,
                           (uint4 *)depth_grad, n_depth / 4, (uint4 *)feat_grad, n_feat_rows * c / 4));
    } else {
        FO_CUDA(cudaMemsetAsync(depth_grad, 0, (size_t)n_depth * 4, stream));
        FO_CUDA(cudaMemsetAsync(feat_grad, 0, (size_t)n_feat_rows * c * 4, stream));
    }
    // a requested plan rides along the gather when it can (below); otherwise it is built first, as a separate launch
    bool plan_pending = req != nullptr;
    auto build_plan_now = [&]() -> int {
        plan_pending = false;
        return fo_bwd_plan_build((fo_stream_t)stream, nullptr, nullptr, n_points, req->n_points_dev, n_depth, n_feat_rows,
                                 req->hw, FO_BWD_PLAN_STRUCTURED, fwd_plan, fwd_plan_bytes, B, n_vox,
                                 const_cast<void *>(bwd_plan), bwd_plan_bytes);
    };
    if (n_points == 0 || n_intervals == 0) return plan_pending ? build_plan_now() : FO_OK;
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

    const int impl = bwd_impl_choice();
    // the round-2 kernels: 128-bit rows (C % 4 == 0, C <= 128), 32-bit index arithmetic
    const bool v2_ok = impl >= 1 && vec && c <= 128 && n_feat_rows * c < INT_MAX && n_feat_rows < INT_MAX / 2 &&
                       bv.cap < INT_MAX;
    if (og_layout == FO_LAYOUT_BCZYX) {
        const size_t need = fo_bwd_scratch_bytes(n_intervals, c, og_layout);
        if (!scratch || scratch_bytes < need)
            return set_error(FO_ERR_SCRATCH, "backward scratch is %zu bytes, need %zu", scratch_bytes, need);
        FO_CHECK_ARG(((uintptr_t)scratch & 255) == 0, "scratch must be 256-byte aligned");
        float *G = (float *)scratch;
        GatherArgs ga;
        ga.og = out_grad + (int64_t)c_offset * n_vox; ga.og_bstride = (int64_t)c_total * n_vox; ga.C = c; ga.V = n_vox;
        ga.sps = sps;
        ga.hdr = pv.hdr; ga.sub_iv = pv.sub_iv; ga.iv_vox = pv.iv_vox; ga.sub_mask = pv.sub_mask; ga.G = G;
        pa.G = G; pa.row_map = nullptr; pa.n_rows_G = n_intervals; pa.g_rowstride = c;
        if (c > 256 || B > 65535) return set_error(FO_ERR_UNSUPPORTED, "channel or batch count too large for the gather kernel");
        // plans built from caller-supplied intervals may be flagged non-canonical on the device: the per-interval
        // gather fills G then (it exits at once otherwise), and the sub-tile gathers exit
        // (a plan request means the forward plan came from the rank precompute: always canonical, nothing to guard)
        if (req == nullptr) {
            bwd_gather_flagged_kernel<<<grid_for(n_intervals * 32, 256, 8), 256, 0, stream>>>(ga, n_intervals);
            FO_LAUNCH_CHECK("bwd_gather_flagged_kernel");
        }
        const bool tma = v2_ok && n_intervals * c < INT_MAX && tmap_ok(ga.og, n_vox, c, c_total);
        if (tma) {
            CUtensorMap tm, tmh;
            if (int rc = make_voxel_tmap(&tm, ga.og, n_vox, c, c_total, B)) return rc;
            if (int rc = make_voxel_tmap(&tmh, ga.og, n_vox, c, c_total, B, true)) return rc;
            // FO_BWD_HALF=0 (A/B): always fetch whole sub-tiles
            {
                const char *he = getenv("FO_BWD_HALF");
                if (he && *he && atoi(he) == 0) ga.sub_mask = nullptr;
            }
            const int nacc = (c + 31) / 32;
            const bool exact = c % 32 == 0;
            const size_t stage_bytes = ((size_t)c * 128 + 1023) & ~(size_t)1023;
            const size_t g_smem = 8 * stage_bytes + 1024;
            const int gu = (sps + 7) / 8;
            if (g_smem <= 200 * 1024 && gu <= 65535) {
                const int D = (req && n_feat_rows > 0) ? (int)(n_depth / n_feat_rows) : 0;
                const bool ride = plan_pending && D >= 1 && D <= 128 && (int64_t)B * n_vox < (1 << 24) && req->hw >= 1 &&
                                  n_feat_rows % req->hw == 0 && n_depth == n_feat_rows * D && n_depth <= pv.p_cap &&
                                  bv.cap >= n_depth && (int64_t)gu * B + n_feat_rows / 16 + 1 < INT_MAX &&
                                  g_smem <= 48 * 1024 &&      // wide channel counts leave two gather CTAs per SM: no room to share
                                  bwd_ride_choice() != 0;
                if (ride) {
                    PlanRideArgs pr;
                    pr.fhdr = pv.hdr; pr.pt2vox = pv.pt2vox; pr.vox2iv = pv.vox2iv; pr.D = D; pr.HW = req->hw; pr.n_rows = (int)n_feat_rows;
                    pr.gu = gu; pr.n_gather_ctas = gu * B;
                    pr.n_plan_ctas = (int)((n_feat_rows + 8 * kPlanPixPerWarp - 1) / (8 * kPlanPixPerWarp));
                    {
                        const int64_t T = (int64_t)pr.n_plan_ctas + pr.n_gather_ctas;
                        pr.small = (T + 1) * pr.n_plan_ctas < (1ll << 31) ? 1 : 0;
                        pr.div_gu = make_fastdiv((uint32_t)gu);
                        pr.div_T = make_fastdiv((uint32_t)(T < 1 ? 1 : (T > 0x7fffffff ? 0x7fffffff : T)));
                    }
                    const int ride = bwd_ride_choice();
                    pr.plan_first = ride == 2 || (ride < 0 && pr.n_gather_ctas <= 8 * FO_GRIDE_MINB * sm_count()) ? 1 : 0;
                    pr.hdr = bv.hdr; pr.ent_p = bv.ent_p; pr.ent_iv = bv.ent_iv; pr.starts = bv.starts;
                    pr.lengths = bv.lengths; pr.ids = bv.ids; pr.n_points_dev = req->n_points_dev;
                    const int R = D <= 32 ? 1 : (D <= 64 ? 2 : 4);
#define FO_GRIDE(NA, EX, RR)                                                                                          \
    do {                                                                                                              \
        if (g_smem > 48 * 1024)                                                                                       \
            FO_CUDA(cudaFuncSetAttribute(bwd_gather_plan_kernel<NA, EX, RR>,                                          \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem));                  \
        FO_CUDA(launch_pdl(kPdlGather, bwd_gather_plan_kernel<NA, EX, RR>, dim3((unsigned)(pr.n_gather_ctas + pr.n_plan_ctas)),   \
                           dim3(256), g_smem, stream, ga, pr, tm, tmh));                                                   \
    } while (0)
#define FO_GRIDE_R(NA, EX)                                                                                            \
    do {                                                                                                              \
        if (R == 1) FO_GRIDE(NA, EX, 1); else if (R == 2) FO_GRIDE(NA, EX, 2); else FO_GRIDE(NA, EX, 4);              \
    } while (0)
                    if (nacc == 1) { if (exact) FO_GRIDE_R(1, true); else FO_GRIDE_R(1, false); }
                    else if (nacc == 2) { if (exact) FO_GRIDE_R(2, true); else FO_GRIDE_R(2, false); }
                    else if (nacc == 3) { if (exact) FO_GRIDE_R(3, true); else FO_GRIDE_R(3, false); }
                    else { if (exact) FO_GRIDE_R(4, true); else FO_GRIDE_R(4, false); }
#undef FO_GRIDE_R
#undef FO_GRIDE
                    FO_LAUNCH_CHECK("bwd_gather_plan_kernel");
                    plan_pending = false;
                    return launch_pixel2(pa, stream);
                }
                if (plan_pending)
                    if (int rc = build_plan_now()) return rc;
#define FO_GTMA(NA, EX)                                                                                            \
    do {                                                                                                           \
        if (g_smem > 48 * 1024)                                                                                    \
            FO_CUDA(cudaFuncSetAttribute(bwd_gather_tma_kernel<NA, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)g_smem));                                                            \
        FO_CUDA(launch_pdl(kPdlGather, bwd_gather_tma_kernel<NA, EX>, dim3(gu, B), dim3(256), g_smem, stream, ga, tm, tmh));        \
    } while (0)
                if (nacc == 1) { if (exact) FO_GTMA(1, true); else FO_GTMA(1, false); }
                else if (nacc == 2) { if (exact) FO_GTMA(2, true); else FO_GTMA(2, false); }
                else if (nacc == 3) { if (exact) FO_GTMA(3, true); else FO_GTMA(3, false); }
                else { if (exact) FO_GTMA(4, true); else FO_GTMA(4, false); }
#undef FO_GTMA
                FO_LAUNCH_CHECK("bwd_gather_tma_kernel");
                return launch_pixel2(pa, stream);
            }
        }
        // round-1 sub-tile gather (LDG path): any V, any C <= 256
        if (plan_pending)
            if (int rc = build_plan_now()) return rc;
        const size_t smem = (size_t)kWarpsPerCta * kSub * c * sizeof(float);
        if (smem > 200 * 1024) return set_error(FO_ERR_UNSUPPORTED, "C=%d too large for the gather tile", c);
        const int n_ctas = (sps + kWarpsPerCta - 1) / kWarpsPerCta;
        if (n_ctas > 65535) return set_error(FO_ERR_UNSUPPORTED, "grid too large for the gather kernel");
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
    } else {
        if (plan_pending)
            if (int rc = build_plan_now()) return rc;
        pa.G = out_grad + c_offset; pa.row_map = pv.iv_vox; pa.n_rows_G = (int64_t)B * n_vox; pa.g_rowstride = c_total;
    }
    if (v2_ok && (((uintptr_t)pa.G & 15) == 0) && (pa.g_rowstride % 4 == 0) && pa.n_rows_G * pa.g_rowstride < INT_MAX)
        return launch_pixel2(pa, stream);
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

// Backward INCLUDING the construction of the (structured) backward plan: what fo_bwd_plan_build(FO_BWD_PLAN_STRUCTURED)
// followed by fo_bev_pool_v2_backward does, with the plan built by extra warps of the gather kernel whenever the shape
// allows it (contiguous (B,C,Z,Y,X) out_grad, D <= 128, B*Z*Y*X < 2^24), as a separate launch otherwise.
extern "C" int fo_bev_pool_v2_backward_with_plan(fo_stream_t stream_, int32_t c, const float *out_grad,
                                                 int32_t og_layout, const float *depth, const float *feat,
                                                 int64_t n_points, const int32_t *n_points_dev, int64_t n_intervals,
                                                 int32_t B, int64_t n_vox, int64_t n_depth, int64_t n_feat_rows,
                                                 int32_t hw_size, float *depth_grad, float *feat_grad,
                                                 const void *fwd_plan, size_t fwd_plan_bytes, void *bwd_plan,
                                                 size_t bwd_plan_bytes, void *scratch, size_t scratch_bytes) {
    FO_CHECK_ARG(bwd_plan != nullptr && ((uintptr_t)bwd_plan & 255) == 0, "backward plan must be non-NULL, 256-byte aligned");
    FO_CHECK_ARG(hw_size >= 1 && n_feat_rows >= 1 && n_feat_rows % hw_size == 0 && n_depth % n_feat_rows == 0,
                 "structured plan needs n_depth = n_feat_rows * D and n_feat_rows = B*N*hw_size");
    PlanRequest req{n_points_dev, hw_size};
    return backward_impl((cudaStream_t)stream_, c, out_grad, og_layout, c, 0, depth, feat, n_points, n_intervals, B,
                         n_vox, n_depth, n_feat_rows, depth_grad, feat_grad, fwd_plan, fwd_plan_bytes, bwd_plan,
                         bwd_plan_bytes, scratch, scratch_bytes, &req);
}
