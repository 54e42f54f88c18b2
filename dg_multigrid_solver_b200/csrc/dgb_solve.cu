// dgb_solve.cu -- solve-phase kernels (apply / residual / smoothers / transfers): row-per-thread
// streaming kernels for any BSR structure; the lexicographic Gauss-Seidel order runs in the single-launch
// kernels of dgb_chain.cu / dgb_stream.cu where the operator has the DG 5-point stencil.
//
// Reference semantics restated here:
//   y = A x            scipy bsr_matvec               <- dgfem/solver.py:117,119,150
//   block GS pass      pyamg amg_core.block_gauss_seidel <- dgfem/pyamg_relaxation.py:252-255
//   block Jacobi / GS  dgfem/relaxation.py:123-195
//   transfers          dgfem/solver.py:152-193
#include <stdarg.h>

#include "dgb_common.cuh"

namespace dgb {

static thread_local char g_err[512] = "";
long long g_launches = 0;
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    }
    return n;
}

// ---------------------------------------------------------------------------------------
// element selection for the relaxation kernels
struct Sel {
    int mode;   // 0: all elements; 1: colour class (i+j)&1 == sel; 2: anti-diagonal i+j == sel;
                // 3: colour class, compact: one item per pair (2m, 2m+1) of a row, every item has work
    int sel;
    int Ni, Nj;
    int i_lo;   // mode 2: first i on the diagonal
    int count;  // number of candidate items (mode 0/1: Ni*Nj, mode 2: elements on the diagonal)
    const int32_t *list;   // mode 4: explicit block-row list (one dependency level of a generic Gauss-Seidel pass)
    int keep_partner;      // mode 3 residual: leave the residual rows of the other colour alone (default: written as zero)
};

__device__ __forceinline__ int sel_element(const Sel &s, int idx) {
    if (s.mode == 0) return idx;
    if (s.mode == 1) {     // colour class; i_lo doubles as a parity shift (global row offset of a slab)
        const int i = idx % s.Ni, j = idx / s.Ni;
        return (((i + j + s.i_lo) & 1) == s.sel) ? idx : -1;
    }
    if (s.mode == 4) return s.list[idx];
    if (s.mode == 3) {     // item = (row j, pair m): the element of the pair with (i + j + shift) & 1 == sel
        const int np = (s.Ni + 1) >> 1;
        const int j = idx / np, m = idx - j * np;
        const int i = 2 * m + ((s.sel + j + s.i_lo) & 1);
        return i < s.Ni ? j * s.Ni + i : -1;
    }
    const int i = s.i_lo + idx;
    return (s.sel - i) * s.Ni + i;
}
// mode 3: the other element of the item's pair (the colour that is not selected), -1 if outside the row
__device__ __forceinline__ int sel_partner(const Sel &s, int idx) {
    const int np = (s.Ni + 1) >> 1;
    const int j = idx / np, m = idx - j * np;
    const int i = 2 * m + 1 - ((s.sel + j + s.i_lo) & 1);
    return i < s.Ni ? j * s.Ni + i : -1;
}

// MODE_RESIDUAL_RELAX: the entry residual of a 2-colour smoother call fused with the relaxation of its first colour --
// a row of that colour reads only its own x and the other colour's, so the in-place update races with nothing and the
// four neighbour blocks are read once for both (r -> r_out, x relaxed in place unless *skip)
enum { MODE_APPLY = 0, MODE_RESIDUAL = 1, MODE_RELAX = 2, MODE_RESIDUAL_RELAX = 3 };

template <int B>
struct RowCfg {
    // lanes per scalar row: rows of 128 bytes and more are shared by two adjacent lanes that take alternate
    // 16-byte pieces, so one warp-wide load touches half as many cache lines (the L1 tag stage bounds these)
    static constexpr int LP = (B >= 16 && B % 2 == 0) ? 2 : 1;
    static constexpr int EPB = (256 / (B * LP)) > 0 ? (256 / (B * LP)) : 1;  // elements per CTA
    static constexpr int NT = ((EPB * B * LP + 31) / 32) * 32;               // threads per CTA
};

// my share of sum_c a[c] x[c]: all of it (LP == 1, scipy's order), or the 16-byte pieces h, h+2, ... (LP == 2)
template <int B, int LP>
__device__ __forceinline__ double row_dot_part(const double *__restrict__ a, const double *__restrict__ x, int h) {
    if (LP == 1) return row_dot<B>(a, x);
    const double2 *a2 = reinterpret_cast<const double2 *>(a);
    const double2 *x2 = reinterpret_cast<const double2 *>(x);
    double t = 0.0;
#pragma unroll
    for (int kk = 0; kk < (B / 2 + 1) / 2; ++kk) {
        const int k = 2 * kk + h;
        if (k < B / 2) {
            const double2 v = a2[k], w = x2[k];
            t = fma(v.x, w.x, t);
            t = fma(v.y, w.y, t);
        }
    }
    return t;
}

// One thread (LP == 1) or two adjacent lanes (LP == 2) per scalar row (element e, row r).  Blocks are read
// straight from global memory in aligned 16-byte pieces (the L1 serves the neighbouring columns).
template <int B, int MODE>
__global__ void __launch_bounds__(RowCfg<B>::NT)
k_rows(const double *__restrict__ data, const int32_t *__restrict__ indices,
       const int32_t *__restrict__ indptr, const double *__restrict__ dinv,
       const double *__restrict__ rhs, const double *x_in, double *x_out, double *partials,
       double omega, Sel sel, const int32_t *__restrict__ skip, double *r_out = nullptr) {
    constexpr int EPB = RowCfg<B>::EPB;
    constexpr int NT = RowCfg<B>::NT;
    constexpr int LP = RowCfg<B>::LP;
    constexpr bool RELAXES = MODE == MODE_RELAX || MODE == MODE_RESIDUAL_RELAX;
    const bool frozen = skip != nullptr && *skip != 0;
    if (frozen && MODE != MODE_RESIDUAL_RELAX) return;
    __shared__ __align__(16) double s_rsum[RELAXES ? EPB * B : 2];
    __shared__ double s_red[32];
    const int rowid = threadIdx.x / LP, h = threadIdx.x - rowid * LP;
    const int el = rowid / B;
    const int r = rowid - el * B;
    const bool lane_ok = el < EPB;
    const bool writer = h == 0;               // the lane that owns the row's result
    const int ntiles = (sel.count + EPB - 1) / EPB;
    double sumsq = 0.0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int idx = tile * EPB + el;
        int e = -1;
        if (lane_ok && idx < sel.count) e = sel_element(sel, idx);
        if (MODE == MODE_RESIDUAL && sel.mode == 3 && !sel.keep_partner && lane_ok && idx < sel.count && writer &&
            x_out != nullptr) {
            const int ep = sel_partner(sel, idx);  // a row of the colour relaxed last: its residual is zero (to rounding)
            if (ep >= 0 && indptr[ep] != indptr[ep + 1]) x_out[(size_t)ep * B + r] = 0.0;
        }
        double acc = 0.0, acc_diag = 0.0;
        if (e >= 0) {
            const int j0 = indptr[e], j1 = indptr[e + 1];
            if (j0 == j1) e = -1;        // empty block row (ghost row of a slab): not part of this rank's operator
            for (int jj = j0; jj < j1; ++jj) {
                const int col = indices[jj];
                if (MODE == MODE_RELAX && col == e) continue;
                const double t = row_dot_part<B, LP>(data + ((size_t)jj * B + r) * B, x_in + (size_t)col * B, h);
                if (MODE == MODE_RESIDUAL_RELAX && col == e) acc_diag = t;
                else acc += t;
            }
        }
        if (LP == 2) acc += __shfl_xor_sync(0xffffffffu, acc, 1);      // the two halves of the row
        if (LP == 2 && MODE == MODE_RESIDUAL_RELAX) acc_diag += __shfl_xor_sync(0xffffffffu, acc_diag, 1);
        if (MODE == MODE_APPLY) {
            if (e >= 0 && writer) x_out[(size_t)e * B + r] = acc;
        } else if (MODE == MODE_RESIDUAL) {
            if (e >= 0 && writer) {
                const double res = rhs[(size_t)e * B + r] - acc;
                if (x_out != nullptr) x_out[(size_t)e * B + r] = res;
                sumsq = fma(res, res, sumsq);
            }
        } else {
            if (e >= 0 && writer) {
                const double rest = rhs[(size_t)e * B + r] - acc;
                s_rsum[el * B + r] = rest;
                if (MODE == MODE_RESIDUAL_RELAX) {
                    const double res = rest - acc_diag;
                    if (r_out != nullptr) r_out[(size_t)e * B + r] = res;
                    sumsq = fma(res, res, sumsq);
                }
            }
            __syncthreads();
            double t = 0.0;
            if (e >= 0) t = row_dot_part<B, LP>(dinv + ((size_t)e * B + r) * B, s_rsum + el * B, h);
            if (LP == 2) t += __shfl_xor_sync(0xffffffffu, t, 1);
            if (e >= 0 && writer && !frozen) {
                const double xo = x_in[(size_t)e * B + r];
                x_out[(size_t)e * B + r] = (omega == 1.0) ? t : omega * t + (1.0 - omega) * xo;
            }
            __syncthreads();
        }
    }
    if (MODE == MODE_RESIDUAL || MODE == MODE_RESIDUAL_RELAX) {
        const double t = block_sum<NT>(sumsq, s_red);
        if (threadIdx.x == 0) partials[blockIdx.x] = t;
    }
}

// fixed-order sum of the per-CTA partials -> *out
__global__ void __launch_bounds__(1024)
k_sum_partials(const double *__restrict__ partials, int n, double *out,
               const int32_t *__restrict__ skip) {
    if (skip != nullptr && *skip != 0) return;
    __shared__ double s_red[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) v += partials[i];
    const double t = block_sum<1024>(v, s_red);
    if (threadIdx.x == 0) *out = t;
}

__global__ void __launch_bounds__(256)
k_sumsq(const double *__restrict__ v, int64_t n, double *partials) {
    __shared__ double s_red[32];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        s = fma(v[i], v[i], s);
    const double t = block_sum<256>(s, s_red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// CTAs per SM the kernel can actually keep resident (registers): the grid is ONE wave of persistent CTAs
template <int B, int MODE>
static int rows_occupancy() {
    static int occ = 0;
    if (occ == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rows<B, MODE>, RowCfg<B>::NT, 0) != cudaSuccess || occ < 1)
            occ = 1;
        if (occ > 8) occ = 8;
    }
    return occ;
}
static int rows_grid(int count, int epb, int occ = 8) {
    const int ntiles = (count + epb - 1) / epb;
    int g = sm_count() * occ;
    if (g > kMaxPartials) g = kMaxPartials;
    if (g > ntiles) g = ntiles;
    return g < 1 ? 1 : g;
}

// ---------------------------------------------------------------------------------------
// smoother control (device-side early exit; relaxation.py:202-216)
__device__ __forceinline__ void smoother_begin(dgb_smoother_ctl *ctl, double sumsq, double n) {
    ctl->res0 = sqrt(sumsq / n);
    ctl->ratio = 1.0;
    ctl->skip = ctl->diverged;      // `diverged` is sticky (only the host clears it): nothing runs after a divergence
    ctl->iters = 0;
    ctl->calls += 1;
}
__device__ __forceinline__ void smoother_check(dgb_smoother_ctl *ctl, double sumsq, double n) {
    if (ctl->skip) return;
    const double ratio = sqrt(sumsq / n) / ctl->res0;
    ctl->ratio = ratio;
    ctl->iters += 1;
    if (ratio < 1e-6) {
        ctl->skip = 1;
    } else if (ratio > 1e10) {
        ctl->diverged = 1;
        ctl->skip = 1;
    }
}
__global__ void k_smoother_begin(dgb_smoother_ctl *ctl, const double *sumsq, double n) { smoother_begin(ctl, *sumsq, n); }
__global__ void k_smoother_check(dgb_smoother_ctl *ctl, const double *sumsq, double n) { smoother_check(ctl, *sumsq, n); }

// k_sum_partials followed by the smoother's test on the sum (MODE 1: dgb_smoother_begin, 2: dgb_smoother_check) in one
// launch: the chained smoother runs this pair after its entry residual and after every iteration, 61 times per V-cycle
// of C3, and a kernel of a few microseconds costs its launch gap
template <int MODE>
__global__ void __launch_bounds__(1024)
k_sum_partials_ctl(const double *__restrict__ partials, int n, double *out, const int32_t *__restrict__ skip,
                   dgb_smoother_ctl *ctl, double ndof) {
    if (skip != nullptr && *skip != 0) return;      // after an early exit both halves are no-ops
    __shared__ double s_red[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) v += partials[i];
    const double t = block_sum<1024>(v, s_red);
    if (threadIdx.x == 0) {
        *out = t;
        if (MODE == 1) smoother_begin(ctl, t, ndof);
        else smoother_check(ctl, t, ndof);
    }
}

// ---------------------------------------------------------------------------------------
// K6: in-place Gauss-Jordan inverse with partial pivoting, one warp per block, matrix in smem
template <int B>
__global__ void __launch_bounds__(128)
k_block_diag_inverse(const double *__restrict__ data, const int32_t *__restrict__ indices,
                     const int32_t *__restrict__ indptr, int n_brow, double *dinv, int32_t *info) {
    constexpr int WPB = 4;
    __shared__ double s_a[WPB][B * B];
    __shared__ int s_piv[WPB][B];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * WPB + w;
    if (e >= n_brow) return;
    double *a = s_a[w];
    if (indptr[e] == indptr[e + 1]) {       // empty (ghost) row
        for (int t = lane; t < B * B; t += 32) dinv[(size_t)e * B * B + t] = 0.0;
        return;
    }
    // gather (sum of) the diagonal block(s) of row e
    for (int t = lane; t < B * B; t += 32) a[t] = 0.0;
    __syncwarp();
    for (int jj = indptr[e]; jj < indptr[e + 1]; ++jj) {
        if (indices[jj] != e) continue;
        for (int t = lane; t < B * B; t += 32) a[t] += data[(size_t)jj * B * B + t];
        __syncwarp();
    }
    for (int k = 0; k < B; ++k) {
        // pivot search over rows k..B-1 (first maximum wins => deterministic)
        double best = -1.0;
        int bi = k;
        for (int i = k + lane; i < B; i += 32) {
            const double v = fabs(a[i * B + k]);
            if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) s_piv[w][k] = bi;
        if (best == 0.0 || !(best == best)) {
            if (lane == 0) atomicCAS(info, 0, e + 1);
            // leave a zero block behind (the reference's pinv would return the pseudo-inverse)
            for (int t = lane; t < B * B; t += 32) dinv[(size_t)e * B * B + t] = 0.0;
            return;
        }
        if (bi != k) {
            for (int c = lane; c < B; c += 32) {
                const double t = a[k * B + c];
                a[k * B + c] = a[bi * B + c];
                a[bi * B + c] = t;
            }
        }
        __syncwarp();
        const double pinv = 1.0 / a[k * B + k];
        __syncwarp();
        for (int c = lane; c < B; c += 32) a[k * B + c] = (c == k) ? pinv : a[k * B + c] * pinv;
        __syncwarp();
        // eliminate column k from all other rows: items (i, c)
        for (int t = lane; t < B * B; t += 32) {
            const int i = t / B, c = t - i * B;
            if (i == k) continue;
            const double f = a[i * B + k];
            // column k itself must be handled last for each row; use the saved factor
            if (c != k) a[t] = fma(-f, a[k * B + c], a[t]);
        }
        __syncwarp();
        for (int i = lane; i < B; i += 32)
            if (i != k) a[i * B + k] = -a[i * B + k] * pinv;
        __syncwarp();
    }
    // undo the row interchanges as column interchanges, in reverse order
    for (int k = B - 1; k >= 0; --k) {
        const int p = s_piv[w][k];
        if (p != k) {
            for (int i = lane; i < B; i += 32) {
                const double t = a[i * B + k];
                a[i * B + k] = a[i * B + p];
                a[i * B + p] = t;
            }
        }
        __syncwarp();
    }
    for (int t = lane; t < B * B; t += 32) dinv[(size_t)e * B * B + t] = a[t];
}

template <int B>
__global__ void __launch_bounds__(256)
k_build_gs_stream(const double *__restrict__ data, const int32_t *__restrict__ indices,
                  const int32_t *__restrict__ indptr, const double *__restrict__ dinv, int n_brow,
                  double *gs) {
    // one warp per block row
    const int e = (int)(((size_t)blockIdx.x * 256 + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (e >= n_brow) return;
    for (int jj = indptr[e]; jj < indptr[e + 1]; ++jj) {
        const double *src = (indices[jj] == e) ? dinv + (size_t)e * B * B : data + (size_t)jj * B * B;
        for (int t = lane; t < B * B; t += 32) gs[(size_t)jj * B * B + t] = src[t];
    }
}

// ---------------------------------------------------------------------------------------
// K9: transfers.  One thread per output scalar; R/P (<= 16x36 doubles) staged in smem.
// H gather (solver.py:164): flat fine element = A0*(4*Nj_c) + a1*(2*Nj_c) + 2*A2 + a3 with
// coarse row K = A0*Nj_c + A2 and child slot a1*2 + a3 -- the reference's literal
// reshape((Ni_c,2,Nj_c,2,b)).transpose(0,2,1,3,4), valid as geometry for square grids only.
__device__ __forceinline__ size_t h_fine_elem(int K, int child, int Nj_c) {
    const int A0 = K / Nj_c, A2 = K - A0 * Nj_c;
    const int a1 = child >> 1, a3 = child & 1;
    return (size_t)A0 * (4 * (size_t)Nj_c) + (size_t)a1 * (2 * (size_t)Nj_c) + 2 * (size_t)A2 + a3;
}

// Slab variant (kind DGB_TRANSFER_H_SLAB): children of coarse element (I, J) are the fine elements
// (2I+a_i, 2J+a_j), child slot a_j*2+a_i -- what the reference's gather means on the square global grid --
// with ghost-row offsets: coarse rows [gc, Nj_c - gc_hi) are active, fine row = gf + 2*(J - gc) + a_j.
struct SlabH {
    int Ni_c, gc, gf;     // coarse elements per row, ghost rows below on the coarse / fine level
};
__device__ __forceinline__ size_t slab_fine_elem(int K, int child, SlabH h) {
    const int J = K / h.Ni_c, I = K - J * h.Ni_c;
    const int aj = child >> 1, ai = child & 1;
    return (size_t)(h.gf + 2 * (J - h.gc) + aj) * (2 * (size_t)h.Ni_c) + 2 * (size_t)I + ai;
}

__global__ void __launch_bounds__(256)
k_restrict(int kind, const double *__restrict__ R, int nc, int nf, int Nj_c, int64_t n_coarse_el,
           const double *__restrict__ fine, double *__restrict__ coarse, SlabH slab, int64_t first_el) {
    extern __shared__ double s_R[];
    for (int t = threadIdx.x; t < nc * nf; t += blockDim.x) s_R[t] = R[t];
    __syncthreads();
    const int64_t total = n_coarse_el * nc;
    for (int64_t o0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o0 < total;
         o0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t K = o0 / nc + first_el;           // first_el: first active coarse element of a slab
        const int a = (int)(o0 % nc);
        const int64_t o = K * nc + a;
        double acc = 0.0;
        if (kind == DGB_TRANSFER_P) {
            const double *f = fine + K * nf;
            for (int c = 0; c < nf; ++c) acc = fma(s_R[a * nf + c], f[c], acc);
        } else {
            const int bf = nf / 4;
            for (int child = 0; child < 4; ++child) {
                const size_t fe = kind == DGB_TRANSFER_H ? h_fine_elem((int)K, child, Nj_c) : slab_fine_elem((int)K, child, slab);
                const double *f = fine + fe * bf;
                for (int d = 0; d < bf; ++d) acc = fma(s_R[a * nf + child * bf + d], f[d], acc);
            }
        }
        coarse[o] = acc;
    }
}

__global__ void __launch_bounds__(256)
k_prolong_add(int kind, const double *__restrict__ P, int nc, int nf, int Nj_c, int64_t n_coarse_el,
              const double *__restrict__ coarse, double *__restrict__ fine, SlabH slab, int64_t first_el) {
    extern __shared__ double s_P[];
    for (int t = threadIdx.x; t < nc * nf; t += blockDim.x) s_P[t] = P[t];
    __syncthreads();
    const int64_t total = n_coarse_el * nf;
    for (int64_t o0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o0 < total;
         o0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t K = o0 / nf + first_el;
        const int f = (int)(o0 % nf);
        const double *uc = coarse + K * nc;
        double acc = 0.0;
        for (int a = 0; a < nc; ++a) acc = fma(s_P[f * nc + a], uc[a], acc);
        if (kind == DGB_TRANSFER_P) {
            fine[K * nf + f] += acc;
        } else {
            const int bf = nf / 4;
            const int child = f / bf, d = f - child * bf;
            const size_t fe = kind == DGB_TRANSFER_H ? h_fine_elem((int)K, child, Nj_c) : slab_fine_elem((int)K, child, slab);
            fine[fe * bf + d] += acc;
        }
    }
}

}  // namespace dgb

namespace dgb {
// single-launch lexicographic smoother kernels (dgb_stream.cu, dgb_chain.cu)
extern int g_kernel_path;
bool stream_supported(int b);
int gs_rows_launch(int b, const double *gs, const double *rhs, double *x, double *mbox, int Ni, int Nj, int flags,
                   int dir, double omega, const int32_t *skip, cudaStream_t st);
// chained lexicographic GS (dgb_chain.cu)
bool chain_supported(int b, int flags);
int gs_chain_pass(const dgb_operator *op, const double *rhs, double *x, int dir, bool have_c, const int32_t *skip,
                  cudaStream_t st);
bool chain_c_recurrence(int flags);
int gs_chain_helper_residual(const dgb_operator *op, const double *rhs, const double *x, int dir, double *r,
                             double *partials, int *grid_out, cudaStream_t st, bool x_zero = false);
bool chain_residual_supported(int b);
int gs_chain_residual(const dgb_operator *op, const double *x, int last_dir, double *r, double *partials,
                      int *grid_out, const int32_t *skip, cudaStream_t st);
extern int g_gs_variant;

// kernels that need the closed-form DG stencil (k_gs_rows)
static bool use_stream(const dgb_operator *op) {
    return g_kernel_path == 0 && op->stencil >= 0 && stream_supported(op->b);
}
static int check_op(const dgb_operator *op) {
    DGB_ARG(op != nullptr);
    DGB_ARG(op->data && op->indices && op->indptr && op->Ni > 0 && op->Nj > 0 && op->b > 0);
    return 0;
}
}  // namespace dgb

extern "C" {
int dgb_bsr_residual_colour(const dgb_operator *op, const double *rhs, const double *x, double *r, int32_t relaxed,
                            int32_t shift, double *partials, double *sumsq, const int32_t *skip, void *stream);
int dgb_block_gs_colour_entry(const dgb_operator *op, const double *rhs, double *x, double *r, int32_t first,
                              int32_t shift, double *partials, double *sumsq, const int32_t *frozen, void *stream);
static int lexicographic_pass(const dgb_operator *op, const double *rhs, double *x, double omega, int direction,
                              const int32_t *skip, cudaStream_t st, bool have_c = false);
}

namespace dgb {
bool gs_entry_fused(const dgb_operator *op) {
    return use_stream(op) && op->gs_chain != nullptr && op->gs_mailbox != nullptr && chain_supported(op->b, op->stencil) &&
           chain_c_recurrence(op->stencil);
}
// Relaxation.block_gauss_seidel_pyamg on the device (dgfem/relaxation.py:198-218).  r_keep (optional): every
// residual test also stores the residual vector, so that after the call r_keep == rhs - A u for the u the
// smoother returns (the tests after an early exit are no-ops and u no longer changes) -- the V-cycle reuses it
// for the restriction instead of evaluating the same residual again (dgfem/solver.py:150).
int gs_pyamg(const dgb_operator *op, const double *rhs, double *u, int32_t direction, int32_t max_iterations,
             int32_t mode, int32_t check_residual, dgb_smoother_ctl *ctl, double *partials, double *sumsq,
             double *r_keep, void *stream, void *event_after_last_pass, bool u_is_zero, bool entry_primed) {
    int rc = check_op(op);
    if (rc) return rc;
    DGB_ARG(ctl && partials && sumsq);
    DGB_ARG(direction == 0 || direction == 1 || direction == -1);
    const int64_t n = (int64_t)op->Ni * op->Nj * op->b;
    DGB_ARG(op->dinv && rhs && u);
    int last_dir = 0;      // direction of the previous lexicographic pass of this call (u untouched since)
    int entry_colour = -1; // 2-colour mode: the colour the entry residual kernel already relaxed
    // every pass of this call runs in the chained kernel with the c-recurrence (no ghost rows): the records of the
    // opposite direction are complete after each pass
    const bool chained_loop = mode == DGB_GS_LEXICOGRAPHIC && max_iterations > 0 && use_stream(op) &&
                              op->gs_chain != nullptr && op->gs_mailbox != nullptr &&
                              chain_supported(op->b, op->stencil) && chain_c_recurrence(op->stencil);
    bool begun = false;    // the entry test already ran inside the reduction kernel
    if (check_residual) {
        const int first_dir = direction >= 0 ? +1 : -1;
        const bool chained = mode == DGB_GS_LEXICOGRAPHIC && max_iterations > 0 && use_stream(op) &&
                             op->gs_chain != nullptr && op->gs_mailbox != nullptr &&
                             chain_supported(op->b, op->stencil) && chain_c_recurrence(op->stencil);
        if (chained && entry_primed) {
            // dgb_vcycle_ex: the caller ran dgb_block_gs_entry_residual on (rhs, u) itself -- *sumsq, r_keep and the
            // first pass's right-hand sides are in place
            last_dir = -first_dir;
        } else if (chained) {
            // the entry residual shares its block reads with the dependency-free part of the first pass
            int grid = 1;
            rc = gs_chain_helper_residual(op, rhs, u, first_dir, r_keep, partials, &grid, (cudaStream_t)stream,
                                          u_is_zero && g_gs_variant != 42);
            if (rc) return rc;
            k_sum_partials_ctl<1><<<1, 1024, 0, (cudaStream_t)stream>>>(partials, grid, sumsq, nullptr, ctl, (double)n);
            DGB_LAUNCH_OK();
            last_dir = -first_dir;          // the first pass finds its c in place
            begun = true;
        } else if (mode == DGB_GS_REDBLACK && max_iterations > 0 && op->stencil >= 0 && g_gs_variant != 43) {
            // the first colour of the first pass is relaxed by the kernel that evaluates its entry residual
            entry_colour = first_dir > 0 ? 0 : 1;
            rc = dgb_block_gs_colour_entry(op, rhs, u, r_keep, entry_colour, 0, partials, sumsq, &ctl->diverged, stream);
            if (rc) return rc;
        } else {
            rc = dgb_bsr_residual(op, rhs, u, r_keep, partials, sumsq, nullptr, stream);
            if (rc) return rc;
        }
        if (!begun && (rc = dgb_smoother_begin(ctl, sumsq, n, stream))) return rc;
    }
    const int32_t *skip = check_residual ? &ctl->skip : nullptr;
    // 2-colour mode: relaxing a colour twice in a row with nothing in between recomputes the same values bit for bit
    // (x_e = Dinv_e (rhs_e - sum A x_other colour)), so the second of two adjacent passes over one colour is dropped:
    // a symmetric iteration 0,1 | 1,0 runs 0,1,0 and the next one starts at 1
    int last_colour = entry_colour;
    for (int it = 0; it < max_iterations; ++it) {
        for (int dir = +1; dir >= -1; dir -= 2) {
            if ((dir > 0 && direction < 0) || (dir < 0 && direction > 0)) continue;
            if (mode == DGB_GS_LEXICOGRAPHIC) {
                rc = ::lexicographic_pass(op, rhs, u, 1.0, dir, skip, (cudaStream_t)stream, last_dir == -dir);
                last_dir = dir;
            } else {
                for (int k = 0; k < 2 && rc == 0; ++k) {
                    const int colour = dir > 0 ? k : 1 - k;
                    if (colour == last_colour) continue;
                    rc = dgb_block_gs_colour(op, rhs, u, colour, 0, skip, stream);
                    last_colour = colour;
                }
            }
            if (rc) return rc;
        }
        if (it == max_iterations - 1 && event_after_last_pass != nullptr)       // u is final from here on
            DGB_CUDA_OK(cudaEventRecord((cudaEvent_t)event_after_last_pass, (cudaStream_t)stream));
        if (check_residual) {
            // after a chained pass the opposite direction's records hold everything the residual needs but the
            // diagonal block: 2 b^2 + 2 b doubles per element instead of 5 b^2 (k_residual_rec, dgb_chain.cu)
            if (chained_loop && last_dir != 0 && chain_residual_supported(op->b) && g_gs_variant != 41) {
                int grid = 1;
                rc = gs_chain_residual(op, u, last_dir, r_keep, partials, &grid, skip, (cudaStream_t)stream);
                if (rc) return rc;
                k_sum_partials_ctl<2><<<1, 1024, 0, (cudaStream_t)stream>>>(partials, grid, sumsq, skip, ctl, (double)n);
                DGB_LAUNCH_OK();
                continue;                   // the test ran inside the reduction kernel
            } else if (mode == DGB_GS_REDBLACK && last_colour >= 0 && op->stencil >= 0) {
                rc = dgb_bsr_residual_colour(op, rhs, u, r_keep, last_colour, 0, partials, sumsq, skip, stream);
                if (rc) return rc;
            } else {
                rc = dgb_bsr_residual(op, rhs, u, r_keep, partials, sumsq, skip, stream);
                if (rc) return rc;
            }
            rc = dgb_smoother_check(ctl, sumsq, n, stream);
            if (rc) return rc;
        }
    }
    return 0;
}
}  // namespace dgb

using namespace dgb;

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int dgb_abi_version(void) { return DGB_ABI_VERSION; }
const char *dgb_last_error(void) { return g_err; }
int dgb_sm_count(void) { return sm_count(); }
int dgb_partials_len(void) { return kMaxPartials; }
long long dgb_launch_count(int32_t reset) {
    const long long n = g_launches;
    if (reset) g_launches = 0;
    return n;
}

int dgb_bsr_apply(const dgb_operator *op, const double *x, double *y, void *stream) {
    int rc = check_op(op);
    if (rc) return rc;
    DGB_ARG(x && y);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = op->Ni * op->Nj;
    Sel sel{0, 0, N, 1, 0, N, nullptr};
    DGB_DISPATCH_B_ANY(op->b, k_rows<B, MODE_APPLY><<<rows_grid(N, RowCfg<B>::EPB, rows_occupancy<B, MODE_APPLY>()), RowCfg<B>::NT, 0, st>>>(
                              op->data, op->indices, op->indptr, nullptr, nullptr, x, y, nullptr, 1.0, sel, nullptr));
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_bsr_residual(const dgb_operator *op, const double *rhs, const double *x, double *r,
                     double *partials, double *sumsq, const int32_t *skip, void *stream) {
    int rc = check_op(op);
    if (rc) return rc;
    DGB_ARG(x && rhs && partials && sumsq);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = op->Ni * op->Nj;
    int grid = 1;
    Sel sel{0, 0, N, 1, 0, N, nullptr};
    DGB_DISPATCH_B_ANY(op->b, grid = rows_grid(N, RowCfg<B>::EPB, rows_occupancy<B, MODE_RESIDUAL>());
                   k_rows<B, MODE_RESIDUAL><<<grid, RowCfg<B>::NT, 0, st>>>(
                       op->data, op->indices, op->indptr, nullptr, rhs, x, r, partials, 1.0, sel, skip));
    DGB_LAUNCH_OK();
    k_sum_partials<<<1, 1024, 0, st>>>(partials, grid, sumsq, skip);
    DGB_LAUNCH_OK();
    return 0;
}

// Residual after a 2-colour pass that relaxed colour `relaxed` last: those rows satisfy their equations (their
// residual is rounding noise, written as zero), only the rows of the other colour are evaluated -- half the traffic.
int dgb_bsr_residual_colour(const dgb_operator *op, const double *rhs, const double *x, double *r, int32_t relaxed,
                            int32_t shift, double *partials, double *sumsq, const int32_t *skip, void *stream) {
    int rc = check_op(op);
    if (rc) return rc;
    DGB_ARG(x && rhs && partials && sumsq && (relaxed == 0 || relaxed == 1));
    cudaStream_t st = (cudaStream_t)stream;
    int grid = 1;
    Sel sel{3, 1 - relaxed, op->Ni, op->Nj, shift & 1, ((op->Ni + 1) / 2) * op->Nj, nullptr};
    DGB_DISPATCH_B_ANY(op->b, grid = rows_grid(sel.count, RowCfg<B>::EPB, rows_occupancy<B, MODE_RESIDUAL>());
                   k_rows<B, MODE_RESIDUAL><<<grid, RowCfg<B>::NT, 0, st>>>(
                       op->data, op->indices, op->indptr, nullptr, rhs, x, r, partials, 1.0, sel, skip));
    DGB_LAUNCH_OK();
    k_sum_partials<<<1, 1024, 0, st>>>(partials, grid, sumsq, skip);
    DGB_LAUNCH_OK();
    return 0;
}

// Entry of a 2-colour smoother call: r = rhs - A x and its sum of squares over all rows, and colour `first` relaxed in
// place.  Two launches: the rows of the other colour (residual only; they read the first colour's x, so they go
// first), then the rows of `first` (residual + relaxation from one read of their blocks).
int dgb_block_gs_colour_entry(const dgb_operator *op, const double *rhs, double *x, double *r, int32_t first,
                              int32_t shift, double *partials, double *sumsq, const int32_t *frozen, void *stream) {
    int rc = check_op(op);
    if (rc) return rc;
    DGB_ARG(op->dinv && x && rhs && partials && sumsq && (first == 0 || first == 1));
    if (op->stencil < 0) return DGB_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    int g0 = 1, g1 = 1;
    Sel other{3, 1 - first, op->Ni, op->Nj, shift & 1, ((op->Ni + 1) / 2) * op->Nj, nullptr, 1};
    Sel mine{3, first, op->Ni, op->Nj, shift & 1, ((op->Ni + 1) / 2) * op->Nj, nullptr, 1};
    DGB_DISPATCH_B_ANY(op->b, g0 = rows_grid(other.count, RowCfg<B>::EPB, rows_occupancy<B, MODE_RESIDUAL>());
                   k_rows<B, MODE_RESIDUAL><<<g0, RowCfg<B>::NT, 0, st>>>(
                       op->data, op->indices, op->indptr, nullptr, rhs, x, r, partials, 1.0, other, nullptr));
    DGB_LAUNCH_OK();
    DGB_DISPATCH_B_ANY(op->b, g1 = rows_grid(mine.count, RowCfg<B>::EPB, rows_occupancy<B, MODE_RESIDUAL_RELAX>());
                   k_rows<B, MODE_RESIDUAL_RELAX><<<g1, RowCfg<B>::NT, 0, st>>>(
                       op->data, op->indices, op->indptr, op->dinv, rhs, x, x, partials + g0, 1.0, mine, frozen, r));
    DGB_LAUNCH_OK();
    k_sum_partials<<<1, 1024, 0, st>>>(partials, g0 + g1, sumsq, nullptr);
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_sumsq(const double *v, int64_t n, double *partials, double *sumsq, void *stream) {
    DGB_ARG(v && partials && sumsq && n > 0);
    cudaStream_t st = (cudaStream_t)stream;
    int grid = (int)((n + 255) / 256);
    if (grid > 1024) grid = 1024;
    k_sumsq<<<grid, 256, 0, st>>>(v, n, partials);
    DGB_LAUNCH_OK();
    k_sum_partials<<<1, 1024, 0, st>>>(partials, grid, sumsq, nullptr);
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_block_diag_inverse(const double *data, const int32_t *indices, const int32_t *indptr,
                           int32_t n_brow, int32_t b, double *dinv, int32_t *info, void *stream) {
    DGB_ARG(data && indices && indptr && dinv && info && n_brow > 0);
    cudaStream_t st = (cudaStream_t)stream;
    DGB_CUDA_OK(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    DGB_DISPATCH_B_ANY(b, k_block_diag_inverse<B><<<(n_brow + 3) / 4, 128, 0, st>>>(data, indices, indptr,
                                                                              n_brow, dinv, info));
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_build_gs_stream(const double *data, const int32_t *indices, const int32_t *indptr,
                        const double *dinv, int32_t n_brow, int32_t b, double *gs_data,
                        void *stream) {
    DGB_ARG(data && indices && indptr && dinv && gs_data && n_brow > 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)(((size_t)n_brow * 32 + 255) / 256);
    DGB_DISPATCH_B_ANY(b, k_build_gs_stream<B><<<grid, 256, 0, st>>>(data, indices, indptr, dinv, n_brow,
                                                                gs_data));
    DGB_LAUNCH_OK();
    return 0;
}

// one generic relaxation launch over a selection
static int relax_launch(const dgb_operator *op, const double *rhs, const double *x_in, double *x_out,
                        double omega, Sel sel, const int32_t *skip, cudaStream_t st) {
    if (sel.count <= 0) return 0;
    DGB_DISPATCH_B_ANY(op->b, k_rows<B, MODE_RELAX><<<rows_grid(sel.count, RowCfg<B>::EPB, rows_occupancy<B, MODE_RELAX>()), RowCfg<B>::NT, 0, st>>>(
                              op->data, op->indices, op->indptr, op->dinv, rhs, x_in, x_out, nullptr, omega, sel, skip));
    DGB_LAUNCH_OK();
    return 0;
}

static int wavefront_pass(const dgb_operator *op, const double *rhs, double *x, double omega, int direction,
                          const int32_t *skip, cudaStream_t st) {
    const int Ni = op->Ni, Nj = op->Nj;
    const int ndiag = Ni + Nj - 1;
    for (int k = 0; k < ndiag; ++k) {
        const int c = direction > 0 ? k : ndiag - 1 - k;
        const int i_lo = c - (Nj - 1) > 0 ? c - (Nj - 1) : 0;
        const int i_hi = c < Ni - 1 ? c : Ni - 1;
        Sel sel{2, c, Ni, Nj, i_lo, i_hi - i_lo + 1, nullptr};
        int rc = relax_launch(op, rhs, x, x, omega, sel, skip, st);
        if (rc) return rc;
    }
    return 0;
}

// Arbitrary (structurally symmetric) BSR: the rows of one dependency level of the lexicographic sweep have no
// coupling among themselves, so a level is one relaxation launch over its row list (schedule built by the caller,
// dgb_operator.gs_rows / h_gs_offsets); the levels run in order.  Exactly pyamg's sequential sweep.
static int level_pass(const dgb_operator *op, const double *rhs, double *x, double omega, int direction,
                      const int32_t *skip, cudaStream_t st) {
    const int N = op->Ni * op->Nj;
    const int nl = direction > 0 ? op->gs_nlevels_fwd : op->gs_nlevels_bwd;
    const int32_t *off = op->h_gs_offsets + (direction > 0 ? 0 : op->gs_nlevels_fwd + 1);
    const int32_t *rows = op->gs_rows + (direction > 0 ? 0 : N);
    for (int l = 0; l < nl; ++l) {
        Sel sel{4, 0, N, 1, 0, off[l + 1] - off[l], rows + off[l]};
        int rc = relax_launch(op, rhs, x, x, omega, sel, skip, st);
        if (rc) return rc;
    }
    return 0;
}

// exact lexicographic order, either kernel family
// have_c: see gs_chain_pass (only the smoother loop below, which knows the pass sequence, sets it)
static int lexicographic_pass(const dgb_operator *op, const double *rhs, double *x, double omega, int direction,
                              const int32_t *skip, cudaStream_t st, bool have_c) {
    if (use_stream(op) && omega == 1.0 && op->gs_chain != nullptr && op->gs_mailbox != nullptr &&
        chain_supported(op->b, op->stencil))
        return gs_chain_pass(op, rhs, x, direction, have_c, skip, st);
    if (use_stream(op) && op->gs_data != nullptr && op->gs_mailbox != nullptr)
        return gs_rows_launch(op->b, op->gs_data, rhs, x, op->gs_mailbox, op->Ni, op->Nj, op->stencil, direction,
                              omega, skip, st);
    if (op->stencil < 0) {
        if (op->gs_rows == nullptr || op->h_gs_offsets == nullptr) {
            set_error("lexicographic Gauss-Seidel on an arbitrary BSR matrix needs a level schedule (dgb_operator.gs_rows)");
            return 3;
        }
        return level_pass(op, rhs, x, omega, direction, skip, st);
    }
    return wavefront_pass(op, rhs, x, omega, direction, skip, st);
}

int dgb_block_gs_pass(const dgb_operator *op, const double *rhs, double *x, int32_t direction,
                      int32_t mode, const int32_t *skip, void *stream) {
    int rc = check_op(op);
    if (rc) return rc;
    DGB_ARG(op->dinv && rhs && x);
    DGB_ARG(direction == 1 || direction == -1);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == DGB_GS_REDBLACK) {
        for (int k = 0; k < 2; ++k) {
            const int colour = direction > 0 ? k : 1 - k;
            Sel sel{3, colour, op->Ni, op->Nj, 0, ((op->Ni + 1) / 2) * op->Nj, nullptr};
            rc = relax_launch(op, rhs, x, x, 1.0, sel, skip, st);
            if (rc) return rc;
        }
        return 0;
    }
    return lexicographic_pass(op, rhs, x, 1.0, direction, skip, st);
}

int dgb_block_gs_pass_seq(const dgb_operator *op, const double *rhs, double *x, int32_t direction,
                          int32_t prev_direction, const int32_t *skip, void *stream) {
    int rc = check_op(op);
    if (rc) return rc;
    DGB_ARG(op->dinv && rhs && x);
    DGB_ARG(direction == 1 || direction == -1);
    DGB_ARG(prev_direction == 0 || prev_direction == 1 || prev_direction == -1);
    return lexicographic_pass(op, rhs, x, 1.0, direction, skip, (cudaStream_t)stream, prev_direction == -direction);
}

int dgb_block_gs_entry_residual(const dgb_operator *op, const double *rhs, const double *x, int32_t first_direction,
                                double *r, double *partials, double *sumsq, void *stream) {
    int rc = check_op(op);
    if (rc) return rc;
    DGB_ARG(op->dinv && rhs && x && partials && sumsq);
    DGB_ARG(first_direction == 1 || first_direction == -1);
    cudaStream_t st = (cudaStream_t)stream;
    if (!(use_stream(op) && op->gs_chain != nullptr && op->gs_mailbox != nullptr && chain_supported(op->b, op->stencil)))
        return DGB_UNSUPPORTED;
    int grid = 1;
    rc = gs_chain_helper_residual(op, rhs, x, first_direction, r, partials, &grid, st);
    if (rc) return rc;
    k_sum_partials<<<1, 1024, 0, st>>>(partials, grid, sumsq, nullptr);
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_block_gs_residual_after_pass(const dgb_operator *op, const double *x, int32_t last_direction, double *r,
                                     double *partials, double *sumsq, const int32_t *skip, void *stream) {
    int rc = check_op(op);
    if (rc) return rc;
    DGB_ARG(x && partials && sumsq && (last_direction == 1 || last_direction == -1));
    cudaStream_t st = (cudaStream_t)stream;
    if (!(use_stream(op) && op->gs_chain != nullptr && op->gs_mailbox != nullptr && chain_supported(op->b, op->stencil) &&
          chain_c_recurrence(op->stencil) && chain_residual_supported(op->b)))
        return DGB_UNSUPPORTED;
    int grid = 1;
    rc = gs_chain_residual(op, x, last_direction, r, partials, &grid, skip, st);
    if (rc) return rc;
    k_sum_partials<<<1, 1024, 0, st>>>(partials, grid, sumsq, skip);
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_block_gs_colour(const dgb_operator *op, const double *rhs, double *x, int32_t colour, int32_t shift,
                        const int32_t *skip, void *stream) {
    int rc = check_op(op);
    if (rc) return rc;
    DGB_ARG(op->dinv && rhs && x && (colour == 0 || colour == 1));
    Sel sel{3, colour, op->Ni, op->Nj, shift & 1, ((op->Ni + 1) / 2) * op->Nj, nullptr};
    return relax_launch(op, rhs, x, x, 1.0, sel, skip, (cudaStream_t)stream);
}

int dgb_block_relax_sweep(const dgb_operator *op, const double *rhs, const double *x_in,
                          double *x_out, double omega, void *stream) {
    int rc = check_op(op);
    if (rc) return rc;
    DGB_ARG(op->dinv && rhs && x_in && x_out);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = op->Ni * op->Nj;
    if (x_in == x_out) return lexicographic_pass(op, rhs, x_out, omega, +1, nullptr, st);
    Sel sel{0, 0, op->Ni, op->Nj, 0, N, nullptr};
    return relax_launch(op, rhs, x_in, x_out, omega, sel, nullptr, st);
}

int dgb_smoother_begin(dgb_smoother_ctl *ctl, const double *sumsq, int64_t n, void *stream) {
    DGB_ARG(ctl && sumsq && n > 0);
    k_smoother_begin<<<1, 1, 0, (cudaStream_t)stream>>>(ctl, sumsq, (double)n);
    DGB_LAUNCH_OK();
    return 0;
}
int dgb_smoother_check(dgb_smoother_ctl *ctl, const double *sumsq, int64_t n, void *stream) {
    DGB_ARG(ctl && sumsq && n > 0);
    k_smoother_check<<<1, 1, 0, (cudaStream_t)stream>>>(ctl, sumsq, (double)n);
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_block_gauss_seidel_pyamg(const dgb_operator *op, const double *rhs, double *u,
                                 int32_t direction, int32_t max_iterations, int32_t mode,
                                 int32_t check_residual, dgb_smoother_ctl *ctl, double *partials,
                                 double *sumsq, void *stream) {
    return dgb::gs_pyamg(op, rhs, u, direction, max_iterations, mode, check_residual, ctl, partials, sumsq, nullptr,
                         stream);
}

static int transfer_launch(bool restrict_dir, int32_t kind, const double *M, int32_t nc, int32_t nf, int32_t Ni_c,
                           int32_t Nj_c, int32_t ghost_c_lo, int32_t ghost_c_hi, int32_t ghost_f_lo,
                           const double *src, double *dst, void *stream) {
    DGB_ARG(M && src && dst && nc > 0 && nf > 0 && Ni_c > 0 && Nj_c > 0);
    DGB_ARG(kind == DGB_TRANSFER_P || ((kind == DGB_TRANSFER_H || kind == DGB_TRANSFER_H_SLAB) && nf % 4 == 0));
    DGB_ARG(ghost_c_lo >= 0 && ghost_c_hi >= 0 && Nj_c - ghost_c_lo - ghost_c_hi > 0);
    // only the active coarse rows are written / read (ghost rows of a slab belong to the neighbour rank)
    const int64_t nel = (int64_t)Ni_c * (Nj_c - ghost_c_lo - ghost_c_hi);
    const int64_t first = (int64_t)Ni_c * ghost_c_lo;
    SlabH slab{Ni_c, ghost_c_lo, ghost_f_lo};
    int64_t g = (nel * (restrict_dir ? nc : nf) + 255) / 256;
    if (g > sm_count() * 16) g = sm_count() * 16;
    if (restrict_dir)
        k_restrict<<<(int)g, 256, sizeof(double) * nc * nf, (cudaStream_t)stream>>>(kind, M, nc, nf, Nj_c, nel, src, dst,
                                                                                  slab, first);
    else
        k_prolong_add<<<(int)g, 256, sizeof(double) * nc * nf, (cudaStream_t)stream>>>(kind, M, nc, nf, Nj_c, nel, src,
                                                                                     dst, slab, first);
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_restrict(int32_t kind, const double *R, int32_t nc, int32_t nf, int32_t Ni_c,
                 int32_t Nj_c, const double *fine, double *coarse, void *stream) {
    DGB_ARG(kind == DGB_TRANSFER_P || kind == DGB_TRANSFER_H);
    return transfer_launch(true, kind, R, nc, nf, Ni_c, Nj_c, 0, 0, 0, fine, coarse, stream);
}

int dgb_prolong_add(int32_t kind, const double *P, int32_t nc, int32_t nf, int32_t Ni_c,
                    int32_t Nj_c, const double *coarse, double *fine, void *stream) {
    DGB_ARG(kind == DGB_TRANSFER_P || kind == DGB_TRANSFER_H);
    return transfer_launch(false, kind, P, nc, nf, Ni_c, Nj_c, 0, 0, 0, coarse, fine, stream);
}

int dgb_restrict_slab(int32_t kind, const double *R, int32_t nc, int32_t nf, int32_t Ni_c, int32_t Nj_c,
                      int32_t ghost_c_lo, int32_t ghost_c_hi, int32_t ghost_f_lo, const double *fine,
                      double *coarse, void *stream) {
    return transfer_launch(true, kind == DGB_TRANSFER_H ? DGB_TRANSFER_H_SLAB : kind, R, nc, nf, Ni_c, Nj_c,
                           ghost_c_lo, ghost_c_hi, ghost_f_lo, fine, coarse, stream);
}

int dgb_prolong_add_slab(int32_t kind, const double *P, int32_t nc, int32_t nf, int32_t Ni_c, int32_t Nj_c,
                         int32_t ghost_c_lo, int32_t ghost_c_hi, int32_t ghost_f_lo, const double *coarse,
                         double *fine, void *stream) {
    return transfer_launch(false, kind == DGB_TRANSFER_H ? DGB_TRANSFER_H_SLAB : kind, P, nc, nf, Ni_c, Nj_c,
                           ghost_c_lo, ghost_c_hi, ghost_f_lo, coarse, fine, stream);
}

}  // extern "C"
