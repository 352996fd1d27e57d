// Microbenchmark: DRAM write efficiency of the forward's store pattern as a function of the contiguous
// run each CTA writes per channel plane, and of a delay that jumbles the temporal order of neighbouring
// lines (emulates occupied sub-tiles being written ~3 us after their empty neighbours).
//   out[b][c][v]: B=8, C=32, V=640000 floats (655 MB).  CTA (b, tile) writes C rows of L bytes.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_store(float *out, int V, int C, int Lf /* floats per row */, int warps, int delay_mod, int delay_cyc) {
    const int b = blockIdx.x, tile = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (delay_mod && (tile % delay_mod) != 0) { long long t0 = clock64(); while (clock64() - t0 < delay_cyc) {} }
    float *base = out + (long long)b * C * V + (long long)tile * Lf;
    const int chunks = Lf / 4;                       // float4 per row
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    // warp w takes rows w, w+warps, ...; lanes stride the row's 16-byte chunks
    for (int c = warp; c < C; c += warps)
        for (int i = lane; i < chunks; i += 32) __stcs(reinterpret_cast<float4 *>(base + (long long)c * V) + i, z);
}
// same bytes, but one warp writes 4 rows x 128 B per instruction (the shipped kernel's pattern), Lf = 32
__global__ void k_store_quad(float *out, int V, int C, int delay_mod, int delay_cyc) {
    const int b = blockIdx.x, tile = blockIdx.y, lane = threadIdx.x;
    if (delay_mod && (tile % delay_mod) != 0) { long long t0 = clock64(); while (clock64() - t0 < delay_cyc) {} }
    float *base = out + (long long)b * C * V + (long long)tile * 32 + (long long)(lane >> 3) * V + 4 * (lane & 7);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) __stcs(reinterpret_cast<float4 *>(base + (long long)(4 * i) * V), z);
}
int main() {
    const int B = 8, C = 32, V = 640000;
    float *out; size_t bytes = (size_t)B * C * V * 4;
    cudaMalloc(&out, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](auto f, const char *name) {
        for (int i = 0; i < 3; ++i) f();
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) f();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-48s %7.1f us  %6.0f GB/s  (%s)\n", name, ms * 100, bytes / (ms / 10 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    timeit([&] { cudaMemsetAsync(out, 0, bytes); }, "cudaMemset");
    for (int delay = 0; delay <= 6000; delay += 6000) {
        char nm[128];
        snprintf(nm, sizeof nm, "quad pattern 1 warp, L=128B, delay=%d", delay);
        timeit([&] { k_store_quad<<<dim3(B, V / 32), 32>>>(out, V, C, delay ? 4 : 0, delay); }, nm);
        for (int L = 128; L <= 8192; L *= 2) {
            const int Lf = L / 4;
            for (int warps : {1, 4, 16}) {
                if (warps > 1 && L < 512) continue;
                snprintf(nm, sizeof nm, "rows L=%dB warps=%d delay=%d", L, warps, delay);
                timeit([&] { k_store<<<dim3(B, V / Lf), 32 * warps>>>(out, V, C, Lf, warps, delay ? 4 : 0, delay); }, nm);
            }
        }
    }
    return 0;
}
