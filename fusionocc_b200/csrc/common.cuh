// fusionocc_b200 — shared device/host helpers (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/fusionocc_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "fusionocc_b200 kernels are written for sm_100a (Blackwell B200) only"
#endif

namespace fo {

// ----------------------------------------------------------------------------------------------
// Tiling constants shared by the forward plan, the forward kernel and the backward gather.
// The unit of work is a SUB-TILE: kSub = 32 consecutive voxels of ONE sample in flattened (z,y,x)
// order, owned by one warp; in the (B,C,Z,Y,X) tensor that is C runs of 128 contiguous bytes (one full
// cache line each).  A CTA is ONE warp, with no block-level synchronisation at all.  Single-warp CTAs on
// purpose: sub-tile costs are very uneven (0 .. 844 points), and a CTA keeps its shared memory and
// registers until its slowest warp is done; with four warps per CTA a third of the resident warps were
// such "zombies" (achieved occupancy 40 % of 62 % theoretical).  Measured (profiles/r01, forward /
// backward at the headline shape, batch 8): 128 threads 185 / 245 us, 64 threads 173 / 243 us,
// 32 threads 152 / 208 us.
//
// Why (measured, profiles/r01): the first tile-per-CTA designs were issue-bound, ~2750 warp
// instructions per 128 voxels, because 8-lane groups each re-derived indices for one interval.  Points
// of a sub-tile are contiguous in the sorted rank arrays, so a warp loads 32 points' indices with one
// coalesced instruction each and then walks them with lanes = channels (one 128-byte feature row per
// point): ~230 instructions per sub-tile.
// ----------------------------------------------------------------------------------------------
#ifndef FO_HEAVY_PTS
#define FO_HEAVY_PTS 256
#endif
constexpr int kHeavyPts     = FO_HEAVY_PTS;   // sub-tiles with more points are "dense": listed in the plan, split by the forward
// The threshold is a function of the plan's identity (B, voxels per sample) only, so every entry point that lists or
// skips dense sub-tiles derives the same value.  Long grids hide the serial chain of a 256..512-point sub-tile behind
// the rest of the launch and prefer fewer front CTAs (measured, forward at batch 8: 512x1408 270.6 -> 260.8 us, stress
// 642.7 -> 617.9 us, headline shape unchanged); short grids need the split (stress batch 2: 160 us at 256, 171 us at
// 512; 512x1408 batch 1: 53 / 53 / 69 us at 256 / 512 / 1024).
__host__ __device__ inline int heavy_threshold(int B, int64_t n_vox) {
    return (int64_t)B * ((n_vox + 31) / 32) >= 131072 ? 2 * kHeavyPts : kHeavyPts;
}
constexpr int kSub          = 32;    // voxels per sub-tile (one warp)
constexpr int kSubShift     = 5;
#ifndef FO_TILE_THREADS
#define FO_TILE_THREADS 32
#endif
constexpr int kThreads      = FO_TILE_THREADS;    // threads per CTA of the tile kernels (see below)
constexpr int kWarpsPerCta  = kThreads / 32;

// Forward-plan flags (device side, FwdPlanHeader::flags)
constexpr int kFlagUnsorted   = 1;   // interval voxels not strictly increasing, or intervals not back to back in
                                     // the point arrays -> order-agnostic path
constexpr int kFlagOutOfRange = 2;   // some interval names a voxel / point range outside the tensors

struct __align__(16) FwdPlanHeader {
    int32_t flags;
    int32_t n_subs;          // number of sub-tiles = B * subs_per_sample
    int32_t subs_per_sample;
    int32_t n_intervals;     // live count (copied from n_intervals_dev or the host argument)
    int32_t structured;      // 1: pt2vox / vox2iv are valid (plan produced by fo_rank_prepare)
    int32_t fwd_heavy[3];    // [0] number of dense sub-tiles listed in heavy_list (written when the plan is built)
    int32_t reserved[8];
};
static_assert(sizeof(FwdPlanHeader) == 64, "header is one 64-byte block");

struct __align__(16) BwdPlanHeader {
    int32_t n_bwd_intervals; // live count of backward intervals (distinct ranks_feat values / pixels)
    int32_t n_points;        // live point count
    int32_t totals[2];       // scratch for the scan's totals
    int32_t structured;      // 1: interval m IS feature row m (fixed-stride rows of D entries, built per pixel)
    int32_t reserved[11];
};
static_assert(sizeof(BwdPlanHeader) == 64, "header is one 64-byte block");

__host__ __device__ inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline int64_t subs_per_sample(int64_t n_vox) { return (n_vox + kSub - 1) / kSub; }

// ----------------------------------------------------------------------------------------------
// Forward plan buffer:
//   [header(256) | sub_iv[bound+1] | sub_pt[bound+1] | heavy_list[bound+1] | sub_mask[bound+1] | vox2iv[NV] | pos2iv[P_cap] |
//    pt2vox[P_cap] | iv_vox[IV_cap]]
//   sub_iv[u]     first interval whose voxel lies in sub-tile u          (sub_iv[n_sub] = n_intervals)
//   sub_pt[u]     sorted position of that interval's first point         (sub_pt[n_sub] = end of points)
//   heavy_list[]  the dense sub-tiles (more than kHeavyPts points), in no particular order
//   sub_mask[u]   bit v set: voxel v of sub-tile u is occupied (0 = unknown: the reader falls back to iv_vox).  The
//                 backward's gather gets its interval -> voxel map from this word (interval l of the sub-tile is its
//                 l-th set bit) instead of a second, dependent load of iv_vox, and fetches only the occupied 64-byte
//                 halves of the out_grad rows — 64 bytes is what the memory system fetches at least
//                 (profiles/micro/sector_gran.cu).
//   vox2iv[v]     interval id of voxel v (rows of the backward's gathered out_grad); only when hdr.structured
//   pos2iv[i]     interval id of sorted position i; only for plans built from caller-supplied intervals
//   pt2vox[p]     voxel id of frustum point p, -1 if filtered; only when hdr.structured
//   iv_vox[k]     voxel id (= ranks_bev) of interval k
//   NV = B*Z*Y*X rounded up to 64, P_cap = point capacity rounded up to 64, IV_cap = min(P_cap, NV).
// The structured arrays replace a point -> position -> interval chain: the order pass of the rank pipeline
// then writes nothing but the three rank arrays (every scattered 4-byte access per point costs a 32-byte L2
// sector, ~10 us per access per point at batch 8 — profiles/r01_summary.md).
// ----------------------------------------------------------------------------------------------
struct FwdPlanView {
    FwdPlanHeader *hdr;
    int32_t *sub_iv;
    int32_t *sub_pt;
    int32_t *heavy_list;
    uint32_t *sub_mask;
    int32_t *vox2iv;
    int32_t *pos2iv;
    int32_t *iv_vox;
    int32_t *pt2vox;
    int64_t p_cap;           // point capacity of this buffer (multiple of 64)
    int64_t iv_cap;          // interval capacity
};
__host__ inline int64_t fwd_plan_subs_bound(int64_t n_vox_total) { return n_vox_total / kSub + 4096 + 1; }
__host__ inline size_t fwd_plan_sub_bytes(int64_t n_vox_total) {
    return (size_t)align_up((fwd_plan_subs_bound(n_vox_total) + 1) * 4, 256);
}
__host__ inline size_t fwd_plan_occ_bytes(int64_t n_vox_total) { return fwd_plan_sub_bytes(n_vox_total); }
__host__ inline size_t fwd_plan_bytes_for(int64_t n_vox_total, int64_t p_cap) {
    const int64_t pc = align_up(p_cap > 0 ? p_cap : 1, 64), nv = align_up(n_vox_total, 64);
    return 256 + 3 * fwd_plan_sub_bytes(n_vox_total) + fwd_plan_occ_bytes(n_vox_total) +
           (size_t)(4 * nv + 8 * pc + 4 * (pc < nv ? pc : nv));
}
// The layout is a pure function of (n_vox_total, plan_bytes): every entry point is handed the same
// plan_bytes the buffer was sized with and recovers the same pointers without reading the device.
__host__ inline bool fwd_plan_view(void *plan, int64_t n_vox_total, size_t plan_bytes, FwdPlanView *v) {
    const int64_t nv = align_up(n_vox_total, 64);
    const size_t fixed = 256 + 3 * fwd_plan_sub_bytes(n_vox_total) + fwd_plan_occ_bytes(n_vox_total) + (size_t)(4 * nv);
    if (plan_bytes < fixed + 64 * 12) return false;
    const int64_t rest = (int64_t)(plan_bytes - fixed);
    int64_t pc = rest / 12;
    if (pc > nv) pc = (rest - 4 * nv) / 8;
    pc = pc / 64 * 64;
    char *p = (char *)plan;
    v->hdr = (FwdPlanHeader *)p;             p += 256;
    v->sub_iv = (int32_t *)p;                p += fwd_plan_sub_bytes(n_vox_total);
    v->sub_pt = (int32_t *)p;                p += fwd_plan_sub_bytes(n_vox_total);
    v->heavy_list = (int32_t *)p;            p += fwd_plan_sub_bytes(n_vox_total);
    v->sub_mask = (uint32_t *)p;             p += fwd_plan_occ_bytes(n_vox_total);
    v->vox2iv = (int32_t *)p;                p += nv * 4;
    v->pos2iv = (int32_t *)p;                p += pc * 4;
    v->pt2vox = (int32_t *)p;                p += pc * 4;
    v->iv_vox = (int32_t *)p;
    v->p_cap = pc;
    v->iv_cap = pc < nv ? pc : nv;
    return true;
}

// ----------------------------------------------------------------------------------------------
// Backward plan buffer (the inverse interval ordering):
//   [header(256) | starts[rows] | lengths[rows] | ids[rows] | bucket counters (generic build) |
//    ent_p[cap] | ent_iv[cap] | pos[cap] | slot[cap]]   (bucket counters = cnt | tile aggregates | group aggregates | queue counters)
//   backward interval m (one image pixel) covers entries [starts[m], +lengths[m]); entry j names the
//   point's depth index ent_p[j] and its forward interval ent_iv[j]; ids[m] is the feature row.
//   Entries are in ascending forward position = the order bev_pool.py:47-49 produces.
//   pos / slot / counters are scratch of the generic (sort-based) build.
// ----------------------------------------------------------------------------------------------
struct BwdPlanView {
    BwdPlanHeader *hdr;
    int32_t *starts, *lengths, *ids;
    char *counters;          // bucket-sort zero region (cnt | scan state | tile counter)
    int32_t *ent_p, *ent_iv, *pos, *slot;
    int64_t cap;
};
constexpr int kScanGroup = 16;   // scan tiles per aggregate group (two-level tile prefix)
__host__ inline size_t bucket_zero_bytes(int64_t n_buckets) {
    const int64_t n_scan_tiles = (n_buckets + 2047) / 2048;
    const int64_t n_groups = n_scan_tiles / kScanGroup + 1;
    return (size_t)(align_up(n_buckets * 4, 256) + align_up(n_scan_tiles * 8, 256) + align_up(n_groups * 8, 256) + 256);
}
__host__ inline size_t bwd_plan_fixed_bytes(int64_t rows) {
    return 256 + 3 * (size_t)align_up(rows * 4, 256) + bucket_zero_bytes(rows);
}
__host__ inline size_t bwd_plan_bytes_for(int64_t cap, int64_t rows) {
    return bwd_plan_fixed_bytes(rows) + 16 * (size_t)align_up(cap > 0 ? cap : 1, 64);
}
__host__ inline bool bwd_plan_view(void *plan, int64_t rows, size_t plan_bytes, BwdPlanView *v) {
    const size_t fixed = bwd_plan_fixed_bytes(rows);
    if (plan_bytes < fixed + 16 * 64) return false;
    const int64_t cap = (int64_t)((plan_bytes - fixed) / 16) / 64 * 64;
    char *p = (char *)plan;
    v->hdr = (BwdPlanHeader *)p;  p += 256;
    v->starts = (int32_t *)p;     p += align_up(rows * 4, 256);
    v->lengths = (int32_t *)p;    p += align_up(rows * 4, 256);
    v->ids = (int32_t *)p;        p += align_up(rows * 4, 256);
    v->counters = p;              p += bucket_zero_bytes(rows);
    v->ent_p = (int32_t *)p;      p += cap * 4;
    v->ent_iv = (int32_t *)p;     p += cap * 4;
    v->pos = (int32_t *)p;        p += cap * 4;
    v->slot = (int32_t *)p;
    v->cap = cap;
    return true;
}

// ----------------------------------------------------------------------------------------------
// Error plumbing (thread-local message, integer status) — cabi.cu owns the storage.
// ----------------------------------------------------------------------------------------------
int set_error(int code, const char *fmt, ...);

#define FO_CHECK_ARG(cond, ...)                                             \
    do {                                                                    \
        if (!(cond)) return ::fo::set_error(FO_ERR_INVALID_ARG, __VA_ARGS__); \
    } while (0)

#define FO_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return ::fo::set_error(FO_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                \
                                   cudaGetErrorString(e__), __FILE__, __LINE__);               \
    } while (0)

#define FO_LAUNCH_CHECK(name)                                                                  \
    do {                                                                                       \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess)                                                                \
            return ::fo::set_error(FO_ERR_CUDA, "launch of %s failed: %s", name,               \
                                   cudaGetErrorString(e__));                                   \
    } while (0)

// opens + validates a forward plan buffer (defined in rank_prepare.cu)
int open_fwd_plan_const(const void *plan, size_t plan_bytes, int32_t B, int64_t n_vox, int64_t n_points,
                        FwdPlanView *pv, int64_t *n_subs, int *sps);

// number of SMs of the current device (148 on B200), queried once per device
inline int sm_count() {
    static int cached[64] = {0};
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
}

#ifndef FO_GRID_CTAS
#define FO_GRID_CTAS 16
#endif
inline int grid_for(int64_t work_items, int per_block, int ctas_per_sm = FO_GRID_CTAS) {
    int64_t b = (work_items + per_block - 1) / per_block;
    const int64_t cap = sm_count() * (int64_t)ctas_per_sm;   // a whole number of waves (148 SMs on a B200)
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}

// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch.  One step is ~14 short kernels on one stream; launched back to back each pays its
// launch latency and its fill / drain ramp (~3 us per boundary, a third of the rank precompute at batch 1).  Every
// hot-path kernel therefore starts with  pdl_wait(); pdl_launch();  and is launched with the programmatic-stream-
// serialization attribute: the NEXT kernel's CTAs become resident while this one drains and block in pdl_wait() until
// this grid has completed and its memory is visible — same ordering as a plain stream, without the bubble.
// FO_PDL=0 turns the attribute off (A/B runs); the device instructions are no-ops in a plain launch.
// ----------------------------------------------------------------------------------------------
// FO_PDL is a bit mask over kernel groups (A/B runs): 1 rank pipeline, 2 forward, 4 backward gather (+ plan),
// 8 backward pixel kernel, 16 zero-fill kernels; default: all.
enum { kPdlRank = 1, kPdlFwd = 2, kPdlGather = 4, kPdlPixel = 8, kPdlZero = 16 };
inline bool pdl_enabled(int group) {
    const char *e = getenv("FO_PDL");
    const int mask = (e && *e) ? atoi(e) : 31;
    return (mask & group) != 0;
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef FO_PDL_EARLY
#define FO_PDL_EARLY 0      // 1: every CTA releases the dependent grid at its start; 0: only at its exit (implicit)
#endif
__device__ __forceinline__ void pdl_launch() {
#if FO_PDL_EARLY
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int group, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled(group) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<Args &&>(args)...);
}

// zero-fills up to two 16-byte aligned regions (counters / gradient outputs) as a kernel of the same stream, so that
// it takes part in the programmatic launch chain (a memset node would serialise it)
static __global__ void __launch_bounds__(256) zero2_kernel(uint4 *a, int64_t na16, uint4 *b, int64_t nb16) {
    pdl_wait();
    pdl_launch();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t i = gtid; i < na16; i += stride) a[i] = z;
    for (int64_t i = gtid; i < nb16; i += stride) b[i] = z;
}
#endif

// ----------------------------------------------------------------------------------------------
// Division of a 32-bit unsigned by a run-time invariant divisor (Granlund & Montgomery, "Division by invariant
// integers using multiplication", fig. 4.1): q = (t + ((n - t) >> sh1)) >> sh2 with t = umulhi(m, n); exact for
// every n < 2^32 and 1 <= d < 2^31.  The order pass derives ranks_feat from the point index with two divisions per
// point; a hardware-emulated 32-bit division costs ~20 instructions.
// ----------------------------------------------------------------------------------------------
struct FastDiv {
    uint32_t m;
    int32_t sh1, sh2;
    uint32_t d;
};
__host__ inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    int l = 0;
    while ((1ull << l) < d) ++l;                      // l = ceil(log2 d)
    f.m = (uint32_t)((((1ull << l) - d) << 32) / d + 1);
    f.sh1 = l < 1 ? l : 1;
    f.sh2 = l > 1 ? l - 1 : 0;
    f.d = d;
    return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv &f) {
    const uint32_t t = __umulhi(f.m, n);
    return (t + ((n - t) >> f.sh1)) >> f.sh2;
}
#endif

// ----------------------------------------------------------------------------------------------
// Small device helpers
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__
template <int RS>   // ascending bitonic sort of 32*RS keys, element e = 32 r + lane
__device__ __forceinline__ void bitonic_sort_regs(int (&key)[RS], const int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * RS; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {                               // partner in another register of the same lane
#pragma unroll
                for (int r = 0; r < RS; ++r) {
                    const int pr = r ^ (j >> 5);
                    if (pr > r) {
                        const bool up = (((32 * r) & k) == 0);
                        const int lo = min(key[r], key[pr]), hi = max(key[r], key[pr]);
                        key[r] = up ? lo : hi;
                        key[pr] = up ? hi : lo;
                    }
                }
            } else {                                     // partner in lane ^ j, same register
#pragma unroll
                for (int r = 0; r < RS; ++r) {
                    const int other = __shfl_xor_sync(0xffffffffu, key[r], j);
                    const bool up = ((((32 * r) | lane) & k) == 0);
                    const bool lower = (lane & j) == 0;
                    key[r] = (lower == up) ? min(key[r], other) : max(key[r], other);
                }
            }
        }
    }
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// Per-warp shared-memory stage of one sub-tile: channel-major [C][kSub] floats = the layout of the
// (B,C,Z,Y,X) block itself, so it is drained / filled with LDS.128/STS.128 <-> 128-bit global
// accesses.  Row c is ROTATED by (c & 7) 16-byte chunks: a flush of one voxel's channels (lane =
// channel) then spreads over 8 bank groups instead of hammering one bank, the rotation costs three
// integer instructions, and 4-voxel chunks stay intact for the 128-bit accesses.
__device__ __forceinline__ int stage_index(int c, int v) {
    return (c << kSubShift) + ((v + ((c & 7) << 2)) & (kSub - 1));
}
#endif

}  // namespace fo
