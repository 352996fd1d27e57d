// fusionocc_b200 — Blackwell bulk-async (TMA) helpers: tensor-map construction on the host, mbarrier and
// cp.async.bulk(.tensor) wrappers on the device (sm_100a).
//
// The dense voxel tensor (forward output / backward out_grad) is (B,C,Z,Y,X): C planes of V = Z*Y*X floats per
// sample.  A 32-voxel sub-tile of it is C rows of 128 contiguous bytes, V*4 bytes apart — a regular 2-D tile, i.e.
// exactly what one cp.async.bulk.tensor instruction moves between global and shared memory.  The tensor map
// describes the tensor as (V, C, B) [fastest first]; a box is (32 voxels, C channels, 1 sample) and lands in shared
// memory as C rows of 128 bytes with the hardware's 128-byte swizzle (16-byte chunk index XOR (row & 7)), which is
// what makes the transposing accesses of the kernels (lane = channel, fixed voxel) conflict-light.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace fo {

// ----------------------------------------------------------------------------------------------
// Host: tensor map over a (B, C_total, V) fp32 tensor restricted to channels [c_offset, c_offset + C).
// cuTensorMapEncodeTiled is fetched through the runtime (no link-time dependency on libcuda).
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)p;
    return fn;
}

// true when the (B,C,Z,Y,X) tensor can be described by a tensor map (strides multiple of 16 bytes, box limits)
inline bool tmap_ok(const void *base, int64_t V, int32_t C, int32_t c_total) {
    return (V % 4 == 0) && (((uintptr_t)base & 15) == 0) && C >= 1 && C <= 256 && V < (1ll << 31) &&
           (int64_t)c_total * V * 4 < (1ll << 40);
}

// half = false: box of a whole sub-tile (32 voxels = 128-byte rows, 128-byte swizzle);
// half = true:  box of half a sub-tile (16 voxels = 64-byte rows, 64-byte swizzle) — the gather's sparse sub-tiles
inline int make_voxel_tmap(CUtensorMap *tm, const float *base, int64_t V, int32_t C, int32_t c_total, int32_t B,
                           bool half = false) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return set_error(FO_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[3] = {(cuuint64_t)V, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t gstr[2] = {(cuuint64_t)V * 4, (cuuint64_t)c_total * V * 4};
    const cuuint32_t box[3] = {(cuuint32_t)(half ? kSub / 2 : kSub), (cuuint32_t)C, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                           half ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(FO_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return FO_OK;
}

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------
// Device: the staged sub-tile.  Element (channel c, voxel v) of a box that landed at a 1024-byte aligned
// shared-memory address lives at byte offset  c*128 + (((v >> 2) ^ (c & 7)) << 4) + (v & 3)*4.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned swz_off(int c, int v) {
    // = c*128 + (((v >> 2) ^ (c & 7)) << 4) + (v & 3)*4: the XOR only touches the chunk bits of v, so it can be applied
    // to v itself with the row's constant (c & 7) << 2 — two instructions per access for a lane with a fixed row
    return ((unsigned)c << 7) + (((unsigned)v ^ (((unsigned)c & 7u) << 2)) << 2);
}

// Half box (16 voxels, 64-byte rows, 64-byte swizzle: address bits [5:4] ^= bits [8:7]): element (c, v), v < 16
__device__ __forceinline__ unsigned swz64_off(int c, int v) {
    return ((unsigned)c << 6) + (((unsigned)v ^ ((((unsigned)c >> 1) & 3u) << 2)) << 2);
}

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared box load, completion signalled on the mbarrier (bytes = box size, out-of-bounds part zero-filled)
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *tm, int x, int y, int z, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
// shared -> global box store (out-of-bounds part of the box is not written)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, int x, int y, int z, unsigned src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tm), "r"(x), "r"(y), "r"(z), "r"(src) : "memory");
}
// same with an L2 eviction-priority hint (the streaming stores of the voxel tensor want evict-first)
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap *tm, int x, int y, int z, unsigned src,
                                                  unsigned long long policy) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;"
                 ::"l"(tm), "r"(x), "r"(y), "r"(z), "r"(src), "l"(policy) : "memory");
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the bulk stores of this thread have finished READING shared memory (the stage may be reused / the CTA may exit)
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
// plain 1-D bulk copy global -> shared (16-byte aligned addresses, size multiple of 16)
__device__ __forceinline__ void bulk_load_1d(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
#endif

}  // namespace fo
