// dgb_common.cuh -- shared helpers for libdgb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dgb200.h"

namespace dgb {

// Grid cap for kernels that emit one partial sum per CTA (fixed-order second stage).
constexpr int kMaxPartials = 4096;

void set_error(const char *fmt, ...);
int sm_count();
extern long long g_launches;   // kernels launched by this library (dgb_launch_count)
int gs_pyamg(const dgb_operator *op, const double *rhs, double *u, int32_t direction, int32_t max_iterations,
             int32_t mode, int32_t check_residual, dgb_smoother_ctl *ctl, double *partials, double *sumsq,
             double *r_keep, void *stream, void *event_after_last_pass = nullptr, bool u_is_zero = false,
             bool entry_primed = false);
// the smoother call on this operator opens with the fused entry residual (dgb_block_gs_entry_residual)
bool gs_entry_fused(const dgb_operator *op);

#define DGB_CUDA_OK(expr)                                                               \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            dgb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,                 \
                           cudaGetErrorString(_e));                                     \
            return -1;                                                                  \
        }                                                                               \
    } while (0)

#define DGB_LAUNCH_OK()                                                                 \
    do {                                                                                \
        ++dgb::g_launches;                                                              \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess) {                                                        \
            dgb::set_error("%s:%d launch -> %s", __FILE__, __LINE__,                    \
                           cudaGetErrorString(_e));                                     \
            return -1;                                                                  \
        }                                                                               \
    } while (0)

#define DGB_ARG(cond)                                                                   \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            dgb::set_error("%s:%d bad argument: %s", __FILE__, __LINE__, #cond);        \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

// Block sizes with a compiled specialisation: (p+1)^2 for p = 0..5, plus the Stokes
// local-order block 2*9+4 = 22 (p_u=2, p_p=1).
#define DGB_DISPATCH_B(b, ...)                                                          \
    switch (b) {                                                                        \
    case 1: { constexpr int B = 1; __VA_ARGS__; } break;                                \
    case 4: { constexpr int B = 4; __VA_ARGS__; } break;                                \
    case 9: { constexpr int B = 9; __VA_ARGS__; } break;                                \
    case 16: { constexpr int B = 16; __VA_ARGS__; } break;                              \
    case 22: { constexpr int B = 22; __VA_ARGS__; } break;                              \
    case 25: { constexpr int B = 25; __VA_ARGS__; } break;                              \
    case 36: { constexpr int B = 36; __VA_ARGS__; } break;                              \
    default:                                                                            \
        dgb::set_error("unsupported block size b=%d", (int)(b));                        \
        return 2;                                                                       \
    }

// The row-per-thread kernels (apply / residual / relaxation / block inverse) also exist for the block sizes SciPy's
// blocksize heuristic picks for the global-order Stokes blocks (2, 3, 6).
#define DGB_DISPATCH_B_ANY(b, ...)                                                      \
    switch (b) {                                                                        \
    case 1: { constexpr int B = 1; __VA_ARGS__; } break;                                \
    case 2: { constexpr int B = 2; __VA_ARGS__; } break;                                \
    case 3: { constexpr int B = 3; __VA_ARGS__; } break;                                \
    case 4: { constexpr int B = 4; __VA_ARGS__; } break;                                \
    case 6: { constexpr int B = 6; __VA_ARGS__; } break;                                \
    case 9: { constexpr int B = 9; __VA_ARGS__; } break;                                \
    case 16: { constexpr int B = 16; __VA_ARGS__; } break;                              \
    case 22: { constexpr int B = 22; __VA_ARGS__; } break;                              \
    case 25: { constexpr int B = 25; __VA_ARGS__; } break;                              \
    case 36: { constexpr int B = 36; __VA_ARGS__; } break;                              \
    default:                                                                            \
        dgb::set_error("unsupported block size b=%d", (int)(b));                        \
        return 2;                                                                       \
    }

// sum_c a[c] * x[c], c = 0..B-1 in order (scipy's bsr_matvec accumulation order), with the matrix row read in
// aligned 16-byte pieces.  Rows of an odd-b block start at 8 (mod 16) every other row: the aligned window then
// begins one double early / ends one double late, and the stray entry meets a zero factor.  (A warp of these row
// reads touches every 128-byte line once per 16-byte piece instead of once per double -- the L1 tag stage, not
// HBM, bounded the 8-byte version.)  The stray entry is always inside the array or its 16-byte slack.
template <int B>
__device__ __forceinline__ double row_dot(const double *__restrict__ a, const double *__restrict__ x) {
    double t = 0.0;
    if (B == 1) {
        t = a[0] * x[0];
    } else if (B % 2 == 0) {
        const double2 *a2 = reinterpret_cast<const double2 *>(a);
#pragma unroll
        for (int k = 0; k < B / 2; ++k) {
            const double2 v = a2[k];
            t = fma(v.x, x[2 * k], t);
            t = fma(v.y, x[2 * k + 1], t);
        }
    } else {
        const int mis = (int)((reinterpret_cast<uintptr_t>(a) >> 3) & 1);
        const double2 *a2 = reinterpret_cast<const double2 *>(a - mis);
#pragma unroll
        for (int k = 0; k < (B + 1) / 2; ++k) {
            const double2 v = a2[k];
            const int c0 = 2 * k - mis, c1 = c0 + 1;
            const double x0 = (c0 >= 0 && c0 < B) ? x[c0] : 0.0;
            const double x1 = (c1 < B) ? x[c1] : 0.0;
            t = fma(v.x, x0, t);
            t = fma(v.y, x1, t);
        }
    }
    return t;
}

// a row of B doubles into registers with aligned 16-byte loads (same window trick as row_dot: an odd-b row that starts
// at 8 mod 16 is read from one double earlier, the stray entries are dropped)
template <int B>
__device__ __forceinline__ void load_row(const double *__restrict__ a, double (&v)[B]) {
    if (B == 1) {
        v[0] = a[0];
    } else if (B % 2 == 0) {
        const double2 *a2 = reinterpret_cast<const double2 *>(a);
#pragma unroll
        for (int k = 0; k < B / 2; ++k) {
            const double2 t = a2[k];
            v[2 * k] = t.x;
            v[2 * k + 1] = t.y;
        }
    } else {
        const int mis = (int)((reinterpret_cast<uintptr_t>(a) >> 3) & 1);
        const double2 *a2 = reinterpret_cast<const double2 *>(a - mis);
        double w[B + 1];
#pragma unroll
        for (int k = 0; k < (B + 1) / 2; ++k) {
            const double2 t = a2[k];
            w[2 * k] = t.x;
            w[2 * k + 1] = t.y;
        }
#pragma unroll
        for (int c = 0; c < B; ++c) v[c] = mis ? w[c + 1] : w[c];
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// CTA-wide sum (fixed shape => deterministic); result valid in thread 0.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *s_red /* [32] */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_red[w] = v;
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
        t = (lane < (NT + 31) / 32) ? s_red[lane] : 0.0;
        t = warp_sum(t);
    }
    __syncthreads();
    return t;
}

}  // namespace dgb
