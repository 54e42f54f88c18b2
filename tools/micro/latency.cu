// Dependent-issue latencies on one warp (sm_100a): DFMA, DADD, 64-bit SHFL, LDS, and a mock of the
// chained Gauss-Seidel step.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, long long *cyc, int n, double a, double b) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = fma(x, b, a);
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
    out[threadIdx.x] = x;
}
__global__ void k_dadd(double *out, long long *cyc, int n, double a) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = x + a;
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
    out[threadIdx.x] = x;
}
__global__ void k_ffma(float *out, long long *cyc, int n, float a, float b) {
    float x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = fmaf(x, b, a);
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
    out[threadIdx.x] = x;
}
__global__ void k_shfl(double *out, long long *cyc, int n, double a) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
    out[threadIdx.x] = x;
}
__global__ void k_lds(double *out, long long *cyc, int n) {
    __shared__ int s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 7 + 1) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) p = s[p];
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
    out[threadIdx.x] = p;
}
// independent DFMAs (throughput per warp): 8 accumulators
__global__ void k_dfma_tp(double *out, long long *cyc, int n, double a, double b) {
    double x[8];
    for (int k = 0; k < 8; ++k) x[k] = a + k;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = fma(x[k], b, a);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
    double s = 0;
    for (int k = 0; k < 8; ++k) s += x[k];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
}
// mock chain step, B = 4: 8 shuffles (64-bit) + 8 DFMA in 4 accumulators + 3 DADD, matrices in registers
__global__ void k_step4(double *out, long long *cyc, int n, double a) {
    const int lane = threadIdx.x & 31, gb = lane & ~3;
    double ml[4], mu[4];
    for (int c = 0; c < 4; ++c) { ml[c] = 1e-3 * (c + 1 + lane); mu[c] = 2e-3 * (c + 1); }
    double xprev = a;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        double upv = __shfl_up_sync(0xffffffffu, xprev, 4);
        double a0 = a, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double xl = __shfl_sync(0xffffffffu, xprev, gb + c);
            const double xu = __shfl_sync(0xffffffffu, upv, gb + c);
            if (c & 1) { a1 = fma(ml[c], xl, a1); a3 = fma(mu[c], xu, a3); }
            else { a0 = fma(ml[c], xl, a0); a2 = fma(mu[c], xu, a2); }
        }
        xprev = (a0 + a1) + (a2 + a3);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
    out[threadIdx.x] = xprev;
}

int main() {
    double *out; long long *cyc; float *outf;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&outf, 1 << 16); cudaMalloc(&cyc, 64);
    long long h;
    const int n = 4096;
    auto rep = [&](const char *nm, int ops) {
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-28s %8.1f cycles/op\n", nm, (double)h / ops);
    };
    for (int pass = 0; pass < 2; ++pass) {
        k_dfma<<<1, 32>>>(out, cyc, n, 1.0, 0.999); if (pass) rep("DFMA dependent (1 warp)", n);
        k_dadd<<<1, 32>>>(out, cyc, n, 1.0); if (pass) rep("DADD dependent", n);
        k_ffma<<<1, 32>>>(outf, cyc, n, 1.0f, 0.999f); if (pass) rep("FFMA dependent", n);
        k_shfl<<<1, 32>>>(out, cyc, n, 1.0); if (pass) rep("SHFL.64 dependent", n);
        k_lds<<<1, 32>>>(out, cyc, n); if (pass) rep("LDS dependent", n);
        k_dfma_tp<<<1, 32>>>(out, cyc, n, 1.0, 0.999); if (pass) rep("DFMA indep x8 (1 warp)", n * 8);
        k_dfma_tp<<<1, 128>>>(out, cyc, n, 1.0, 0.999); if (pass) rep("DFMA indep x8 (4 warps)", n * 8);
        k_dfma_tp<<<1, 512>>>(out, cyc, n, 1.0, 0.999); if (pass) rep("DFMA indep x8 (16 warps)", n * 8);
        k_dfma_tp<<<1, 1024>>>(out, cyc, n, 1.0, 0.999); if (pass) rep("DFMA indep x8 (32 warps)", n * 8);
        k_step4<<<1, 32>>>(out, cyc, n, 1.0); if (pass) rep("mock chain step b=4", n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
