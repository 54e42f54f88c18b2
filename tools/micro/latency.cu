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

__device__ __forceinline__ double2 ldsv2(const void *p) {
    double2 v;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y)
                 : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
// STS -> LDS round trip through shared memory inside one warp (same address, __syncwarp between)
__global__ void k_stslds(double *out, long long *cyc, int n) {
    __shared__ double s[64];
    s[threadIdx.x] = 1.0;
    __syncwarp();
    double v = 0.0;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < n; ++i) {
        v = *(volatile double *)&s[(threadIdx.x + 1) & 31];     // another lane's last store
        __syncwarp();
        *(volatile double *)&s[threadIdx.x] = v + 1.0;
        __syncwarp();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
    out[threadIdx.x] = v;
}
// issue cost of independent LDS.128 for one warp: NL loads, then one use of all of them
template <int NL>
__global__ void k_lds_issue(double *out, long long *cyc, int n) {
    __shared__ double2 s[32 * 16];
    for (int i = threadIdx.x; i < 32 * 16; i += blockDim.x) s[i] = make_double2(i, 1.0);
    __syncthreads();
    double acc = 0.0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        double2 v[NL];
#pragma unroll
        for (int k = 0; k < NL; ++k) v[k] = ldsv2(&s[((threadIdx.x + i) & 31) + 32 * (k & 15)]);
#pragma unroll
        for (int k = 0; k < NL; ++k) acc += v[k].x;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
    out[threadIdx.x] = acc;
}
// mock chain step through shared memory, b = 4, 8 rows per warp: lane (g, r); p = own row's previous value, u = the
// previous value of the row above; REC: 0 matrices in registers, 1 matrices re-read from shared memory every step
template <int REC>
__global__ void k_step4_smem(double *out, long long *cyc, int n, double a) {
    __shared__ __align__(16) double ring[2][8][4];
    __shared__ __align__(16) double rec[8][4][8 + 2];
    const int lane = threadIdx.x & 31, g = lane >> 2, r = lane & 3;
    for (int c = 0; c < 8; ++c) rec[g][r][c] = 1e-3 * (c + 1 + lane);
    rec[g][r][8] = a; rec[g][r][9] = 2 * a;
    ring[0][g][r] = a; ring[1][g][r] = a;
    __syncwarp();
    double ml[4], mu[4], c0 = a;
    for (int c = 0; c < 4; ++c) { ml[c] = rec[g][r][c]; mu[c] = rec[g][r][4 + c]; }
    const int gu = g > 0 ? g - 1 : 0;
    double x = a;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
        const int sl = i & 1;
        const double2 p0 = ldsv2(&ring[sl][g][0]), p1 = ldsv2(&ring[sl][g][2]);
        const double2 u0 = ldsv2(&ring[sl][gu][0]), u1 = ldsv2(&ring[sl][gu][2]);
        if (REC) {
            const double2 m0 = ldsv2(&rec[g][r][0]), m1 = ldsv2(&rec[g][r][2]);
            const double2 m2 = ldsv2(&rec[g][r][4]), m3 = ldsv2(&rec[g][r][6]);
            const double2 cd = ldsv2(&rec[g][r][8]);
            ml[0] = m0.x; ml[1] = m0.y; ml[2] = m1.x; ml[3] = m1.y;
            mu[0] = m2.x; mu[1] = m2.y; mu[2] = m3.x; mu[3] = m3.y; c0 = cd.x;
        }
        double a0 = fma(ml[0], p0.x, c0), a1 = ml[1] * p0.y, a2 = mu[0] * u0.x, a3 = mu[1] * u0.y;
        a0 = fma(ml[2], p1.x, a0); a1 = fma(ml[3], p1.y, a1); a2 = fma(mu[2], u1.x, a2); a3 = fma(mu[3], u1.y, a3);
        x = (a0 + a1) + (a2 + a3);
        *(volatile double *)&ring[sl ^ 1][g][r] = x;
        __syncwarp();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
    out[threadIdx.x] = x;
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
        k_stslds<<<1, 32>>>(out, cyc, n); if (pass) rep("LDS->DADD->STS->LDS loop", n);
        k_lds_issue<4><<<1, 32>>>(out, cyc, n); if (pass) rep("4 indep LDS.128 + use", n);
        k_lds_issue<8><<<1, 32>>>(out, cyc, n); if (pass) rep("8 indep LDS.128 + use", n);
        k_lds_issue<12><<<1, 32>>>(out, cyc, n); if (pass) rep("12 indep LDS.128 + use", n);
        k_step4_smem<0><<<1, 32>>>(out, cyc, n, 1e-3); if (pass) rep("smem chain step b=4 (regs)", n);
        k_step4_smem<1><<<1, 32>>>(out, cyc, n, 1e-3); if (pass) rep("smem chain step b=4 (rec LDS)", n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
