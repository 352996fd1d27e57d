// Microbenchmark: DRAM throughput of reading 128-byte rows at pseudo-random positions of a buffer that does not fit L2
// (the backward's pixel kernel reads the gathered out_grad rows G[N_i][32] in the order of the image pixels' rays).
//   buffer: n_rows x 128 B (1.13 GB, so that nothing survives in the 126 MB L2 between launches); every group of 8 lanes reads
//   `run` consecutive rows starting at a hashed position; a warp has U x 4 rows in flight.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int U>
__global__ void __launch_bounds__(128) k_rows(const float4 *G, unsigned n_rows, int run, int iters, float *sink) {
    const int lane = threadIdx.x & 31, g = lane >> 3, s = lane & 7;
    const unsigned wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
        float4 x[U];
#pragma unroll
        for (int t = 0; t < U; ++t) {
            const unsigned j = (unsigned)(it * U + t);
            // `run` consecutive rows per hashed start; the four lane groups of a warp read neighbouring runs
            const unsigned start = hash32(wid * 7919u + (j / run) * 104729u) % (n_rows - 4u * run - 8u);
            const unsigned row = start + g * run + (j % run);
            x[t] = __ldg(G + (size_t)row * 8 + s);
        }
#pragma unroll
        for (int t = 0; t < U; ++t) { acc.x += x[t].x; acc.y += x[t].y; acc.z += x[t].z; acc.w += x[t].w; }
    }
    if (acc.x == 123.f) *sink = acc.y + acc.z + acc.w;
}
int main() {
    const unsigned n_rows = 8u * 1106717u;       // 1.13 GB: nothing survives in L2 between launches
    float4 *G; float *sink;
    cudaMalloc(&G, (size_t)n_rows * 128); cudaMalloc(&sink, 4);
    cudaMemset(G, 0, (size_t)n_rows * 128);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 8;                                  // rows per lane group = iters * U
    for (int run : {1, 2, 4, 8, 16}) {
        for (int ctas_per_sm : {4, 8, 16}) {
            const int blocks = 148 * ctas_per_sm * 4;      // 4 waves
            auto launch = [&] { k_rows<8><<<blocks, 128>>>(G, n_rows, run, iters, sink); };
            launch(); launch();
            cudaEventRecord(e0);
            for (int r = 0; r < 5; ++r) launch();
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)blocks * 4 /*warps*/ * 4 /*groups*/ * iters * 8 * 128.0;
            printf("run=%2d rows (%5d B contiguous per group)  %2d CTAs/SM resident target: %7.1f us  %6.0f GB/s  (%s)\n", run, run * 128,
                   ctas_per_sm, ms * 200, bytes / (ms / 5 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
