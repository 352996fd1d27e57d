// fusionocc_b200 — shared device/host helpers (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fusionocc_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "fusionocc_b200 kernels are written for sm_100a (Blackwell B200) only"
#endif

namespace fo {

// ----------------------------------------------------------------------------------------------
// Tiling constants shared by the forward plan, the forward kernel and the backward gather.
// A tile is FO_TILE_VOXELS consecutive voxels of ONE sample in flattened (z,y,x) order; in the
// (B,C,Z,Y,X) output that is C contiguous runs of FO_TILE_VOXELS floats (512 B each).
// ----------------------------------------------------------------------------------------------
constexpr int kTile         = 128;   // voxels per tile
constexpr int kThreads      = 256;   // threads per CTA of the tile kernels
constexpr int kGroupLanes   = 8;     // lanes cooperating on one interval: 8 x float4 = 32 channels/pass
constexpr int kGroupsPerCta = kThreads / kGroupLanes;
constexpr int kMaxChunks    = 4;     // channels <= 8 lanes * 4 floats * 4 chunks = 128 on the vector path

// Forward-plan flags (device side, FwdPlanHeader::flags)
constexpr int kFlagUnsorted   = 1;   // interval voxels not strictly increasing -> order-agnostic path
constexpr int kFlagOutOfRange = 2;   // some interval names a voxel / point range outside the tensors

struct __align__(16) FwdPlanHeader {
    int32_t flags;
    int32_t n_tiles;
    int32_t tiles_per_sample;
    int32_t n_intervals;     // live count (copied from n_intervals_dev or the host argument)
    int32_t reserved[12];
};
static_assert(sizeof(FwdPlanHeader) == 64, "header is one 64-byte block");

struct __align__(16) BwdPlanHeader {
    int32_t n_bwd_intervals; // live count of distinct ranks_feat values (written by the scan)
    int32_t n_points;        // live point count
    int32_t reserved[14];
};
static_assert(sizeof(BwdPlanHeader) == 64, "header is one 64-byte block");

__host__ __device__ inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// Layout of the forward plan buffer:  [header | tile_off[n_tiles+1] | pos2iv[n_points_capacity]]
struct FwdPlanView {
    FwdPlanHeader *hdr;
    int32_t *tile_off;
    int32_t *pos2iv;
};
__host__ __device__ inline int64_t tiles_per_sample(int64_t n_vox) { return (n_vox + kTile - 1) / kTile; }
__host__ __device__ inline size_t fwd_plan_tile_bytes(int64_t n_tiles) {
    return (size_t)align_up((n_tiles + 1) * 4, 256);
}
inline FwdPlanView fwd_plan_view(void *plan, int64_t n_tiles) {
    char *p = (char *)plan;
    FwdPlanView v;
    v.hdr = (FwdPlanHeader *)p;
    v.tile_off = (int32_t *)(p + sizeof(FwdPlanHeader) + 192);          // 256-byte aligned
    v.pos2iv = (int32_t *)((char *)v.tile_off + fwd_plan_tile_bytes(n_tiles));
    return v;
}

// Layout of the backward plan buffer:
//   [header | bwd_pos[n_points_cap] | bwd_starts[n_rows] | bwd_lengths[n_rows] | bucket scratch...]
struct BwdPlanView {
    BwdPlanHeader *hdr;
    int32_t *bwd_pos;
    int32_t *bwd_starts;
    int32_t *bwd_lengths;
    int32_t *bucket_ids;    // ranks_feat value of each backward interval
    int32_t *cnt;           // [n_rows]   counters -> exclusive offsets
    int32_t *slot;          // [n_points] arrival slot of each position inside its bucket
    uint64_t *scan_state;   // decoupled look-back descriptors
    int32_t *scan_counter;
    size_t   zero_begin, zero_bytes;   // region that must be zeroed before a build
};

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

// ----------------------------------------------------------------------------------------------
// Small device helpers
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

__device__ __forceinline__ void fma4(float4 &acc, const float4 &a, float b) {
    acc.x = fmaf(a.x, b, acc.x);
    acc.y = fmaf(a.y, b, acc.y);
    acc.z = fmaf(a.z, b, acc.z);
    acc.w = fmaf(a.w, b, acc.w);
}

// streaming (evict-first) 32-bit store: the dense voxel tensor is written once and not re-read here
__device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream4(float *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }
#endif

}  // namespace fo
