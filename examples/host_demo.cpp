// A plain C++ host of the C ABI (include/fusionocc_b200.h): no Python, no torch.  It drives the whole step through
// fo_view_transform_host with pinned host buffers — rank precompute, forward, backward — and checks the results
// against a straightforward CPU evaluation of the same definition (view_transformer.py:246-265 for the voxel index,
// bev_pool_cuda.cu:39-47 / :96-120 for the sums; double accumulation, so the comparison is a tolerance, not bits).
//
//   nvcc -std=c++17 -I include examples/host_demo.cpp -L fusionocc_b200/lib -lfusionocc_b200 \
//        -Xlinker -rpath -Xlinker $PWD/fusionocc_b200/lib -o examples/host_demo && examples/host_demo
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "fusionocc_b200.h"

#define CK(x)                                                                        \
    do {                                                                             \
        cudaError_t e_ = (x);                                                        \
        if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } \
    } while (0)

int main() {
    const int B = 2, N = 3, D = 24, H = 8, W = 11, C = 32, X = 40, Y = 36, Z = 4;
    const float lb[3] = {-8.f, -7.2f, -1.f}, itv[3] = {0.4f, 0.4f, 0.8f};
    const long long P = 1LL * B * N * D * H * W, rows = 1LL * B * N * H * W, V = 1LL * X * Y * Z;
    std::printf("%s (abi %d)\n", fo_build_info(), fo_abi_version());

    float *coor, *depth, *feat, *og, *out, *dg, *fg;
    int32_t *counts;
    CK(cudaMallocHost(&coor, P * 3 * sizeof(float)));
    CK(cudaMallocHost(&depth, P * sizeof(float)));
    CK(cudaMallocHost(&feat, rows * C * sizeof(float)));
    CK(cudaMallocHost(&og, B * C * V * sizeof(float)));
    CK(cudaMallocHost(&out, B * C * V * sizeof(float)));
    CK(cudaMallocHost(&dg, P * sizeof(float)));
    CK(cudaMallocHost(&fg, rows * C * sizeof(float)));
    CK(cudaMallocHost(&counts, 4 * sizeof(int32_t)));
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> ux(-9.f, 9.f), uz(-1.5f, 2.6f), u01(0.f, 1.f);
    std::normal_distribution<float> nrm(0.f, 1.f);
    for (long long p = 0; p < P; ++p) { coor[3 * p] = ux(rng); coor[3 * p + 1] = ux(rng); coor[3 * p + 2] = uz(rng); depth[p] = u01(rng); }
    for (long long i = 0; i < rows * C; ++i) feat[i] = nrm(rng);
    for (long long i = 0; i < B * C * V; ++i) og[i] = nrm(rng);

    const size_t ws_bytes = fo_view_transform_host_workspace_bytes(B, N, D, H, W, C, X, Y, Z, 1);
    void *ws;
    CK(cudaMalloc(&ws, ws_bytes));
    cudaStream_t s, up;
    CK(cudaStreamCreate(&s));
    CK(cudaStreamCreate(&up));
    const int rc = fo_view_transform_host(s, coor, depth, feat, og, B, N, D, H, W, C, lb, itv, X, Y, Z, out, dg, fg, counts,
                                          ws, ws_bytes, up);
    if (rc != FO_OK) { std::fprintf(stderr, "fo_view_transform_host: %s\n", fo_last_error()); return 3; }
    CK(cudaStreamSynchronize(s));

    // CPU evaluation of the definition
    std::vector<double> r_out(B * C * V, 0.0), r_dg(P, 0.0), r_fg(rows * C, 0.0);
    long long kept = 0;
    for (long long p = 0; p < P; ++p) {
        const long long ix = (long long)((coor[3 * p] - lb[0]) / itv[0]), iy = (long long)((coor[3 * p + 1] - lb[1]) / itv[1]),
                        iz = (long long)((coor[3 * p + 2] - lb[2]) / itv[2]);
        if (ix < 0 || ix >= X || iy < 0 || iy >= Y || iz < 0 || iz >= Z) continue;
        ++kept;
        const long long b = p / (1LL * N * D * H * W), q = (p / (1LL * D * H * W)) * (H * W) + p % (H * W);
        const long long v = (iz * Y + iy) * X + ix;
        double dsum = 0.0;
        for (int c = 0; c < C; ++c) {
            const long long o = (b * C + c) * V + v;
            r_out[o] += (double)feat[q * C + c] * depth[p];
            dsum += (double)og[o] * feat[q * C + c];
            r_fg[q * C + c] += (double)og[o] * depth[p];
        }
        r_dg[p] = dsum;
    }
    auto max_err = [](const float *a, const std::vector<double> &b) {
        double m = 0.0;
        for (size_t i = 0; i < b.size(); ++i) m = std::fmax(m, std::fabs((double)a[i] - b[i]) / (1.0 + std::fabs(b[i])));
        return m;
    };
    const double e_out = max_err(out, r_out), e_dg = max_err(dg, r_dg), e_fg = max_err(fg, r_fg);
    std::printf("kept %lld of %lld points (library: %d), %d intervals; max rel err: out %.2e depth_grad %.2e feat_grad %.2e\n",
                kept, P, counts[0], counts[1], e_out, e_dg, e_fg);
    const bool ok = kept == counts[0] && e_out < 1e-5 && e_dg < 1e-5 && e_fg < 1e-5;
    std::printf(ok ? "host_demo: OK\n" : "host_demo: MISMATCH\n");
    return ok ? 0 : 1;
}
