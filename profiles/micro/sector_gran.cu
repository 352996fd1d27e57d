// Microbenchmark: what does the memory system fetch from DRAM when a kernel touches only SOME 32-byte sectors of a
// 128-byte line?  (The backward's gather reads the (B,C,Z,Y,X) out_grad of occupied voxels only: at the headline
// shape 73 % of the 128-byte lines hold an occupied voxel, 60 % of the 64-byte halves, 50 % of the 32-byte sectors.)
//   buffer: 1 GiB of floats; every thread reads ONE 16-byte chunk of "its" line; the pattern selects which lines /
//   sectors are touched.  Run under  ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum  to get the bytes.
#include <cstdio>
#include <cuda_runtime.h>
template <int HINT>
__device__ __forceinline__ float4 ld(const float4 *p) {
    float4 v;
    if (HINT == 0) asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (HINT == 1) asm volatile("ld.global.nc.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (HINT == 2) asm volatile("ld.global.nc.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (HINT == 3) asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (HINT == 4) asm volatile("ld.global.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// sectors_per_line: how many of the four 32-byte sectors of every line are read (1, 2 or 4); one 16-byte chunk each
template <int HINT>
__global__ void k_read(const float4 *buf, long long n_lines, int sectors_per_line, float *sink) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long line = t / sectors_per_line;
    const int s = (int)(t - line * sectors_per_line);
    if (line >= n_lines) return;
    const float4 v = ld<HINT>(buf + line * 8 + s * 2);
    if (v.x == 123.456f) *sink = v.y;
}
int main() {
    const size_t bytes = 1ull << 30;
    float4 *buf; float *sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, bytes);
    const long long n_lines = bytes / 128;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](auto kern, int spl, const char *name) {
        const long long threads = n_lines * spl;
        const int blocks = (int)((threads + 255) / 256);
        kern<<<blocks, 256>>>(buf, n_lines, spl, sink);
        cudaEventRecord(e0);
        kern<<<blocks, 256>>>(buf, n_lines, spl, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s sectors/line=%d  %8.1f us  useful(32B sectors) %6.0f GB/s  (%s)\n", name, spl, ms * 1e3,
               n_lines * spl * 32.0 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    for (int spl : {1, 2, 4}) {
        run(k_read<0>, spl, "ld.global.nc");
        run(k_read<1>, spl, "ld.global.nc.L2::64B");
        run(k_read<2>, spl, "ld.global.nc.L2::128B");
        run(k_read<3>, spl, "ld.global.cs");
        run(k_read<4>, spl, "ld.L1::no_allocate.L2::64B");
    }
    return 0;
}
