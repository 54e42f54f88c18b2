// dgb_dense.cu -- coarse-grid direct solve: dense inverse of a (small) BSR operator + dense mat-vec.
//
// Reference: Solver.solve_directly = scipy.sparse.linalg.spsolve(grid.BSR.tocsr(), RHS)
// (dgfem/solver.py:56-59), used on the coarsest level when `coarse grid solver: direct`
// (dgfem/solver.py:199-200).  The coarsest level of a DG hierarchy has a few hundred unknowns: the operator is
// expanded to a dense n x n matrix and inverted once per hierarchy (Gauss-Jordan, partial pivoting, one CTA);
// every V-cycle then needs one dense mat-vec, u = A^-1 rhs (one warp per row, fixed summation order).
#include "dgb_common.cuh"

namespace dgb {

constexpr int kDenseMaxN = 4096;

__global__ void __launch_bounds__(256)
k_dense_expand(const double *__restrict__ data, const int32_t *__restrict__ indices,
               const int32_t *__restrict__ indptr, int n_brow, int b, double *dense) {
    const long long n = (long long)n_brow * b;
    const long long total = (long long)indptr[n_brow] * b * b;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
        const int jj = (int)(t / (b * b));
        const int rc = (int)(t - (long long)jj * b * b);
        const int r = rc / b, c = rc - r * b;
        // block row of stored block jj: binary search in indptr
        int lo = 0, hi = n_brow;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (indptr[mid] <= jj) lo = mid; else hi = mid;
        }
        // duplicate blocks (tiny periodic grids) are summed, as scipy's tocsr() does
        atomicAdd(&dense[((long long)lo * b + r) * n + (long long)indices[jj] * b + c], data[t]);
    }
}

// in-place Gauss-Jordan inverse with partial pivoting (first maximum wins => deterministic), one CTA
__global__ void __launch_bounds__(1024)
k_dense_inverse(double *a, int n, int32_t *piv, int32_t *info) {
    extern __shared__ double s_buf[];       // [n] pivot row, [n] pivot column
    double *s_row = s_buf, *s_col = s_buf + n;
    __shared__ double s_best[32];
    __shared__ int s_bi[32];
    __shared__ int s_p;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, w = tid >> 5;
    for (int k = 0; k < n; ++k) {
        double best = -1.0;
        int bi = k;
        for (int i = k + tid; i < n; i += nt) {
            const double v = fabs(a[(size_t)i * n + k]);
            if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { s_best[w] = best; s_bi[w] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int q = 1; q < (nt + 31) / 32; ++q)
                if (s_best[q] > best || (s_best[q] == best && s_bi[q] < bi)) { best = s_best[q]; bi = s_bi[q]; }
            s_p = bi;
            piv[k] = bi;
            if (best == 0.0 || !(best == best)) atomicCAS(info, 0, k + 1);
        }
        __syncthreads();
        const int p = s_p;
        // swap rows k and p while staging the (new) pivot row; stage the pivot column
        for (int c = tid; c < n; c += nt) {
            const double vk = a[(size_t)k * n + c], vp = a[(size_t)p * n + c];
            s_row[c] = vp;
            if (p != k) a[(size_t)p * n + c] = vk;
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt) s_col[i] = (i == k) ? s_row[k] : a[(size_t)i * n + k];
        __syncthreads();
        const double pinv = 1.0 / s_row[k];
        for (int c = tid; c < n; c += nt) {
            const double v = (c == k) ? pinv : s_row[c] * pinv;
            s_row[c] = v;
            a[(size_t)k * n + c] = v;
        }
        __syncthreads();
        for (size_t t = tid; t < (size_t)n * n; t += nt) {
            const int i = (int)(t / n), c = (int)(t - (size_t)i * n);
            if (i == k) continue;
            const double f = s_col[i];
            a[t] = (c == k) ? -f * pinv : fma(-f, s_row[c], a[t]);
        }
        __syncthreads();
    }
    // undo the row interchanges as column interchanges, in reverse order
    for (int k = n - 1; k >= 0; --k) {
        const int p = piv[k];
        if (p != k) {
            for (int i = tid; i < n; i += nt) {
                const double t = a[(size_t)i * n + k];
                a[(size_t)i * n + k] = a[(size_t)i * n + p];
                a[(size_t)i * n + p] = t;
            }
        }
        __syncthreads();
    }
}

// u = M rhs: one warp per row, lanes stride the columns, fixed-shape shuffle reduction
__global__ void __launch_bounds__(256)
k_dense_matvec(const double *__restrict__ M, int n, const double *__restrict__ rhs, double *__restrict__ u) {
    const int row = (int)(((size_t)blockIdx.x * 256 + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const double *m = M + (size_t)row * n;
    double acc = 0.0;
    for (int c = lane; c < n; c += 32) acc = fma(m[c], rhs[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) u[row] = acc;
}

// C = A B on C's given block structure: one thread per scalar entry of C; blocks of A's row in stored order
__global__ void __launch_bounds__(256)
k_bsr_spgemm(int b, int n_brow, const int32_t *__restrict__ a_indptr, const int32_t *__restrict__ a_indices,
             const double *__restrict__ a_data, const int32_t *__restrict__ b_indptr,
             const int32_t *__restrict__ b_indices, const double *__restrict__ b_data,
             const int32_t *__restrict__ c_indptr, const int32_t *__restrict__ c_indices, double *__restrict__ c_data) {
    const int bb = b * b;
    const long long total = (long long)c_indptr[n_brow] * bb;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
        const int jj = (int)(t / bb);
        const int rc = (int)(t - (long long)jj * bb);
        const int r = rc / b, c = rc - r * b;
        int lo = 0, hi = n_brow;                       // block row of stored block jj of C
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (c_indptr[mid] <= jj) lo = mid; else hi = mid;
        }
        const int I = lo, J = c_indices[jj];
        double acc = 0.0;
        for (int ka = a_indptr[I]; ka < a_indptr[I + 1]; ++ka) {
            const int K = a_indices[ka];
            for (int kb = b_indptr[K]; kb < b_indptr[K + 1]; ++kb) {
                if (b_indices[kb] != J) continue;
                const double *ar = a_data + (size_t)ka * bb + (size_t)r * b;
                const double *bc = b_data + (size_t)kb * bb + c;
                for (int m = 0; m < b; ++m) acc = fma(ar[m], bc[(size_t)m * b], acc);
            }
        }
        c_data[t] = acc;
    }
}

__global__ void __launch_bounds__(256) k_fill_sentinel(double *v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        v[i] = __longlong_as_double(-1LL);
}

}  // namespace dgb

using namespace dgb;

extern "C" {

int dgb_dense_inverse(const double *data, const int32_t *indices, const int32_t *indptr, int32_t n_brow,
                      int32_t b, double *inverse, int32_t *info, void *stream) {
    DGB_ARG(data && indices && indptr && inverse && info && n_brow > 0 && b > 0);
    const long long n = (long long)n_brow * b;
    if (n > kDenseMaxN) {
        set_error("dgb_dense_inverse: %lld unknowns on the coarsest level (limit %d): coarsen further or use "
                  "`coarse grid solver: smoother`", n, kDenseMaxN);
        return 2;
    }
    cudaStream_t st = (cudaStream_t)stream;
    DGB_CUDA_OK(cudaMemsetAsync(inverse, 0, sizeof(double) * n * n, st));
    DGB_CUDA_OK(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    k_dense_expand<<<sm_count() * 4, 256, 0, st>>>(data, indices, indptr, n_brow, b, inverse);
    DGB_LAUNCH_OK();
    int32_t *piv = nullptr;
    DGB_CUDA_OK(cudaMallocAsync(&piv, sizeof(int32_t) * n, st));
    const size_t smem = sizeof(double) * 2 * n;
    DGB_CUDA_OK(cudaFuncSetAttribute(k_dense_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_dense_inverse<<<1, 1024, smem, st>>>(inverse, (int)n, piv, info);
    DGB_LAUNCH_OK();
    DGB_CUDA_OK(cudaFreeAsync(piv, st));
    return 0;
}

int dgb_dense_solve(const double *inverse, int32_t n, const double *rhs, double *u, void *stream) {
    DGB_ARG(inverse && rhs && u && n > 0 && rhs != u);
    k_dense_matvec<<<(n * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(inverse, n, rhs, u);
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_bsr_spgemm(int32_t b, int32_t n_brow, const int32_t *a_indptr, const int32_t *a_indices, const double *a_data,
                   const int32_t *b_indptr, const int32_t *b_indices, const double *b_data, const int32_t *c_indptr,
                   const int32_t *c_indices, double *c_data, void *stream) {
    DGB_ARG(b > 0 && n_brow > 0 && a_indptr && a_indices && a_data && b_indptr && b_indices && b_data && c_indptr &&
            c_indices && c_data);
    k_bsr_spgemm<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(b, n_brow, a_indptr, a_indices, a_data, b_indptr,
                                                                    b_indices, b_data, c_indptr, c_indices, c_data);
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_fill_sentinel(double *v, int64_t n, void *stream) {
    DGB_ARG(v && n > 0);
    int64_t g = (n + 255) / 256;
    if (g > sm_count() * 8) g = sm_count() * 8;
    k_fill_sentinel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(v, n);
    DGB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
