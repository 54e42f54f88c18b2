// DMMA vs DFMA for the element-block contraction of the p=5 assembly (north_star: "tensor cores (FP64 DMMA) only
// where ncu shows they beat FP64 CUDA cores"):  C[b x b] = X^T[b x q] Y[q x b],  b = 36, q = 64  (the stiffness
// volume integral of one element, dgfem/element.py:181-199), X and Y resident in shared memory.
//   v0  one thread per entry of C, two shared-memory loads per DFMA (the round-1 assembly kernel's inner loop)
//   v1  one thread per 4 x 4 tile of C, 16-byte loads: 0.25 LDS.128 per DFMA
//   v2  DMMA m8n8k4 (mma.sync ... f64): one warp per row of five 8 x 8 tiles (b padded to 40)
// usage: contract [reps]    prints ns per contraction and GFLOP/s per variant; run under ncu for the pipe metrics
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o contract contract.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int B = 36, Q = 64, BP = 40;      // BP: b padded to a multiple of 8

__device__ void fill(double *X, double *Y, int stride) {
    for (int t = threadIdx.x; t < Q * stride; t += blockDim.x) {
        const int q = t / stride, k = t % stride;
        X[t] = k < B ? 1e-2 * ((q * 7 + k * 3) % 11 - 5) : 0.0;
        Y[t] = k < B ? 1e-2 * ((q * 5 + k * 2) % 13 - 6) : 0.0;
    }
}

__global__ void __launch_bounds__(256) k_v0(double *out, int reps) {
    __shared__ double X[Q * B], Y[Q * B];
    fill(X, Y, B);
    __syncthreads();
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int r = 0; r < reps; ++r) {
        int s = 0;
        for (int kl = threadIdx.x; kl < B * B; kl += 256, ++s) {
            const int k = kl / B, l = kl - k * B;
            double a = 0.0;
#pragma unroll 8
            for (int q = 0; q < Q; ++q) a = fma(X[q * B + k], Y[q * B + l], a);
            acc[s] += a;
        }
    }
    int s = 0;
    for (int kl = threadIdx.x; kl < B * B; kl += 256, ++s) out[(size_t)blockIdx.x * B * B + kl] = acc[s];
}

__global__ void __launch_bounds__(96) k_v1(double *out, int reps) {
    __shared__ __align__(16) double X[Q * B], Y[Q * B];
    fill(X, Y, B);
    __syncthreads();
    const int t = threadIdx.x;
    double c[4][4] = {};
    if (t < 81) {
        const int k0 = (t / 9) * 4, l0 = (t % 9) * 4;
        for (int r = 0; r < reps; ++r) {
#pragma unroll 4
            for (int q = 0; q < Q; ++q) {
                const double2 x0 = *(const double2 *)&X[q * B + k0], x1 = *(const double2 *)&X[q * B + k0 + 2];
                const double2 y0 = *(const double2 *)&Y[q * B + l0], y1 = *(const double2 *)&Y[q * B + l0 + 2];
                const double xv[4] = {x0.x, x0.y, x1.x, x1.y}, yv[4] = {y0.x, y0.y, y1.x, y1.y};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) c[i][j] = fma(xv[i], yv[j], c[i][j]);
            }
        }
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) out[(size_t)blockIdx.x * B * B + (k0 + i) * B + l0 + j] = c[i][j];
    }
}

__global__ void __launch_bounds__(160) k_v2(double *out, int reps) {
    __shared__ double X[Q * BP], Y[Q * BP];
    fill(X, Y, BP);
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;       // warp w: rows 8w .. 8w+7 of C
    const int m = lane >> 2, kk = lane & 3;
    double c[5][2] = {};
    for (int r = 0; r < reps; ++r) {
#pragma unroll 4
        for (int q0 = 0; q0 < Q; q0 += 4) {
            const double a = X[(q0 + kk) * BP + 8 * w + m];           // A[m][kk] = X[q0+kk][k0+m]
#pragma unroll
            for (int n = 0; n < 5; ++n) {
                const double b = Y[(q0 + kk) * BP + 8 * n + m];       // B[kk][n'] with n' = lane / 4
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(c[n][0]), "+d"(c[n][1]) : "d"(a), "d"(b));
            }
        }
    }
    for (int n = 0; n < 5; ++n)
        for (int j = 0; j < 2; ++j) {
            const int row = 8 * w + m, col = 8 * n + 2 * kk + j;
            if (row < B && col < B) out[(size_t)blockIdx.x * B * B + row * B + col] = c[n][j];
        }
}

int main(int argc, char **argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 200;
    const int grid = 148 * 8;
    double *out;
    cudaMalloc(&out, sizeof(double) * (size_t)grid * B * B);
    double *h0 = (double *)malloc(sizeof(double) * B * B), *h1 = (double *)malloc(sizeof(double) * B * B);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double flop = 2.0 * B * B * Q * (double)reps * grid;
    for (int v = 0; v < 3; ++v) {
        for (int pass = 0; pass < 2; ++pass) {
            cudaEventRecord(e0);
            if (v == 0) k_v0<<<grid, 256>>>(out, reps);
            if (v == 1) k_v1<<<grid, 96>>>(out, reps);
            if (v == 2) k_v2<<<grid, 160>>>(out, reps);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(v == 0 ? h0 : h1, out, sizeof(double) * B * B, cudaMemcpyDeviceToHost);
        double err = 0;
        if (v > 0) for (int i = 0; i < B * B; ++i) { double d = h1[i] - h0[i]; if (d < 0) d = -d; if (d > err) err = d; }
        printf("v%d %-44s %8.3f ms  %8.1f GFLOP/s  %7.1f ns per contraction per SM-slot  max|diff vs v0| %.2e  (%s)\n", v,
               v == 0 ? "thread per entry, 2 LDS per DFMA" : v == 1 ? "4x4 register tiles, 16-byte LDS" : "DMMA m8n8k4",
               ms, flop / ms / 1e6, ms * 1e6 / ((double)reps * grid / 148.0), err, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
