// Micro-benchmark: (1) dependent fp64 latency, (2) throughput of the radix-16 pass
// (butterfly + twiddle application) of mlv_fft.cuh at 8 and 16 warps per SM.
#include <cstdio>
#include "../../melvin.py_b200/csrc/mlv_fft.cuh"
using namespace mlv;

__global__ void k_lat(double* out, int iters) {
    double a = threadIdx.x, m = 1.0000001, c = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = fma(a, m, c);
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / (16.0 * iters);
}
__global__ void k_lat_add(double* out, int iters) {
    double a = threadIdx.x, c = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = a + c;
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / (16.0 * iters);
}

template <int NT>
__global__ void __launch_bounds__(NT, 1) k_pass(double* out, int iters) {
    cplx v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = mk(threadIdx.x + j, 0.5 * j);
    cplx wb[4] = {mk(0.999, 0.01), mk(0.998, 0.02), mk(0.997, 0.04), mk(0.99, 0.08)};
    for (int i = 0; i < iters; ++i) {
        bfly16<false>(v);
        tw_apply<16, false>(v, 0, wb);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += v[j].x + v[j].y;
    out[blockIdx.x * NT + threadIdx.x] = s;
}

template <int NT>
float run_pass(double* out, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_pass<NT><<<148, NT>>>(out, iters);
    cudaEventRecord(a);
    for (int i = 0; i < 3; ++i) k_pass<NT><<<148, NT>>>(out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 3;
}

int main() {
    double* out; cudaMalloc(&out, 148 * 1024 * sizeof(double));
    double h;
    k_lat<<<1, 32>>>(out, 1000); cudaMemcpy(&h, out + 64, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.2f cycles\n", h);
    k_lat_add<<<1, 32>>>(out, 1000); cudaMemcpy(&h, out + 64, 8, cudaMemcpyDeviceToHost);
    printf("dependent DADD latency: %.2f cycles\n", h);
    const int iters = 2000;
    float t8 = run_pass<256>(out, iters), t16 = run_pass<512>(out, iters), t4 = run_pass<128>(out, iters);
    // cycles per (bfly16 + twiddle) per warp-pass, and SM-level rate
    printf(" 4 warps/SM: %.3f ms -> %.0f cycles per pass per scheduler-warp\n", t4, t4 * 1e-3 * 1.965e9 / iters);
    printf(" 8 warps/SM: %.3f ms -> %.0f cycles per pass (2 warps per scheduler)\n", t8, t8 * 1e-3 * 1.965e9 / iters);
    printf("16 warps/SM: %.3f ms -> %.0f cycles per pass (4 warps per scheduler)\n", t16, t16 * 1e-3 * 1.965e9 / iters);
    printf("error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
