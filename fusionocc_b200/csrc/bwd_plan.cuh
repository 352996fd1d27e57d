// fusionocc_b200 — backward plan of ONE image pixel in registers (shared by the stand-alone plan kernel in
// rank_prepare.cu and by the plan warps riding along the backward's gather kernel in bev_pool_bwd.cu).
#pragma once

#include "common.cuh"

namespace fo {

// Plans produced by the rank precompute carry either pt2vox + vox2iv (structured == 1) or, when the placement already
// resolved the interval of every point, interval ids in the pt2vox table (structured == 2): no vox2iv gather then.
__device__ __forceinline__ const int32_t *plan_vox2iv(const FwdPlanHeader *fhdr, const int32_t *vox2iv) {
    return (__ldg(&fhdr->structured) == 2) ? nullptr : vox2iv;
}

// Bitonic network over packed 32-bit keys (voxel id << 7 | depth bin).  Needs B*Z*Y*X < 2^24 and D <= 128.
// Only ~57 % of a pixel's depth bins land inside the grid, so the valid keys are first COMPACTED (ballot + popc
// through a per-warp shared-memory row) and the network is sized to the live count: 32 / 64 / 128 elements =
// 15 / 21x2 / 28x4 compare-exchange steps (the kernel is issue-bound: 34 -> 2x us at the headline shape).
template <int RS>
__device__ __forceinline__ void plan_sort_emit(const int *cmp, const int n, const int lane, const int ebase,
                                               const int pbase, const int HW, const int32_t *__restrict__ vox2iv,
                                               int32_t *ent_p, int32_t *ent_iv) {
    int key[RS];
#pragma unroll
    for (int r = 0; r < RS; ++r) key[r] = (32 * r + lane < n) ? cmp[32 * r + lane] : INT_MAX;
    bitonic_sort_regs<RS>(key, lane);
#pragma unroll
    for (int r = 0; r < RS; ++r) {
        const int e = 32 * r + lane;
        if (e < n) {
            ent_p[ebase + e] = pbase + (key[r] & 127) * HW;
            ent_iv[ebase + e] = vox2iv ? __ldg(vox2iv + (key[r] >> 7)) : (key[r] >> 7);
        }
    }
}

// One warp, one pixel q: its <= D candidate points p = (bn*D + d)*HW + hw; entries in ascending (voxel id, d) =
// ascending forward position (the order of bev_pool.py:47-49).  `cmp` = 32*R ints of shared memory of this warp.
template <int R>   // R in {1, 2, 4}: 32 * R >= D
__device__ __forceinline__ void plan_pixel_bitonic(const int32_t *__restrict__ pt2vox, const int32_t *__restrict__ vox2iv,
                                                   const int D, const int HW, const int q, int *cmp, const int lane,
                                                   int32_t *ent_p, int32_t *ent_iv, int32_t *starts, int32_t *lengths,
                                                   int32_t *ids) {
    const unsigned lt = (1u << lane) - 1u;
    const int bn = q / HW, hw = q - bn * HW;
    const int pbase = bn * D * HW + hw;
    int key[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int d = lane + 32 * r;
        key[r] = -1;
        if (d < D) {
            const int v = __ldg(pt2vox + pbase + d * HW);
            if (v >= 0) key[r] = (v << 7) | d;
        }
    }
    int n = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const unsigned m = __ballot_sync(0xffffffffu, key[r] >= 0);
        if (key[r] >= 0) cmp[n + __popc(m & lt)] = key[r];
        n += __popc(m);
    }
    __syncwarp();
    if (n > 0) {
        if (n <= 32) plan_sort_emit<1>(cmp, n, lane, q * D, pbase, HW, vox2iv, ent_p, ent_iv);
        else if (R >= 2 && n <= 64) plan_sort_emit<(R >= 2 ? 2 : 1)>(cmp, n, lane, q * D, pbase, HW, vox2iv, ent_p, ent_iv);
        else plan_sort_emit<R>(cmp, n, lane, q * D, pbase, HW, vox2iv, ent_p, ent_iv);
    }
    if (lane == 0) { starts[q] = q * D; lengths[q] = n; ids[q] = q; }
    __syncwarp();                                        // the compaction row is reused by the next pixel
}

}  // namespace fo
