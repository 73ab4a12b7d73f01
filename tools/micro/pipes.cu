// Micro-benchmark: do fp64 arithmetic and shared-memory traffic overlap on sm_100a?
// Three kernels, 16 warps per CTA, 1 CTA per SM on all SMs:
//   A: every warp runs a DFMA stream        B: every warp runs an LDS/STS stream
//   C: even warps DFMA, odd warps LDS/STS (each with the SAME per-warp work as in A/B)
// If the pipes are independent, time(C) ~= max(A, B)/1 per warp-half ... see printout.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dp_stream(double& a0, double& a1, double& a2, double& a3,
                                          double& a4, double& a5, double& a6, double& a7, int iters) {
    const double m = 1.0000001, c = 0.5;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
}

__device__ __forceinline__ double2 lds_stream(double2* sm, int lane, int iters) {
    double2 acc = make_double2(0.0, 0.0);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            double2 v = sm[(lane + 32 * u + i) & 511];      // LDS.128, conflict free
            acc.x += v.x; acc.y += v.y;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) sm[(lane + 32 * u + i) & 511] = acc;   // STS.128
    }
    return acc;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, int it_dp, int it_ls) {
    extern __shared__ double2 sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* mine = sm + warp * 512;
    for (int i = lane; i < 512; i += 32) mine[i] = make_double2(i, warp);
    __syncthreads();
    double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    double2 r = make_double2(0, 0);
    const bool do_dp = MODE == 0 || (MODE == 2 && (warp & 1) == 0);
    const bool do_ls = MODE == 1 || (MODE == 2 && (warp & 1) == 1);
    if (do_dp) dp_stream(a0, a1, a2, a3, a4, a5, a6, a7, it_dp);
    if (do_ls) r = lds_stream(mine, lane, it_ls);
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + r.x + r.y;
}

template <int MODE>
float run(double* out, int it_dp, int it_ls) {
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 16);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148, 512, 16 * 512 * 16>>>(out, it_dp, it_ls);
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) k<MODE><<<148, 512, 16 * 512 * 16>>>(out, it_dp, it_ls);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    double* out;
    cudaMalloc(&out, 148 * 512 * sizeof(double));
    const int it_dp = 4000, it_ls = 2000;
    float A = run<0>(out, it_dp, it_ls), B = run<1>(out, it_dp, it_ls), C = run<2>(out, it_dp, it_ls);
    // per SM: A issues 16 warps * it_dp*64 DFMA; at 1 warp-DFMA / 2 clk / SMSP (4 SMSP)
    double dp_instr = 16.0 * it_dp * 64, ls_wf = 16.0 * it_ls * 32 * 4;
    printf("A (all DFMA)      %.3f ms  -> %.2f warp-DFMA/clk/SM (peak 2.0 if 64 lanes/SM)\n", A, dp_instr / (A * 1e-3 * 1.965e9));
    printf("B (all LDS/STS)   %.3f ms  -> %.2f wavefronts/clk/SM (peak 1.0)\n", B, ls_wf / (B * 1e-3 * 1.965e9));
    printf("C (half/half)     %.3f ms  ; A/2 = %.3f, B/2 = %.3f, A/2+B/2 = %.3f, max = %.3f\n", C, A / 2, B / 2, (A + B) / 2, (A > B ? A : B) / 2);
    printf("error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
