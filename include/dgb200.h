/*
 * dgb200.h -- C ABI of libdgb200.so: the B200-native (sm_100a) implementation of the
 * dgfem hot path (DG Poisson/Stokes assembly + multigrid V-cycle).
 *
 * The reference (thmsdelange/dg-multigrid-solver, pure Python) has no FFI; its hot path
 * sits behind Python-level seams.  Each entry point below names the
 * reference interface it replaces (file:line relative to the reference tree) -- the
 * ctypes stub a maintainer would add on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name starts with `h_`;
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns
 *     without synchronising, unless documented otherwise;
 *   - return value: 0 = ok, <0 = CUDA/launch error (dgb_last_error() has the text),
 *     >0 = argument error;
 *   - one host thread drives a GPU at a time (the reference is single-threaded): the smoother kernels
 *     share one device-side ticket counter and error flag PER DEVICE (allocated on first use on the
 *     current device), so two smoother calls must not run concurrently on different streams of the
 *     same device; one process may drive several devices (cudaSetDevice before each call);
 *   - all reals are IEEE fp64, all indices int32 (scipy's default index type, which the
 *     reference's sp.bsr_array uses);
 *   - BSR layout is scipy's: data[nnzb][b][b] row-major blocks, indices[nnzb] ascending
 *     within a row, indptr[n_brow+1]   (dgfem/discrete_system.py:145);
 *   - element numbering m = j*Ni + i (utils/helpers.py:3-14), mode numbering
 *     n = j_s*(p+1)+i_r, point numbering i_r + N_int*i_s (dgfem/interpolation.py:133-140).
 */
#ifndef DGB200_H
#define DGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGB_ABI_VERSION 6

/* ---- status ------------------------------------------------------------------------- */
int dgb_abi_version(void);
const char *dgb_last_error(void);
/* number of SMs of the current device (used by callers to size workspaces) */
int dgb_sm_count(void);

/* Number of doubles a `partials` workspace must hold (per-CTA partial sums of the fused
 * norm reductions; reduced in a fixed order => bitwise reproducible norms). */
int dgb_partials_len(void);

/* Kernels launched by this library since the last reset (bench.py's gpu_launches). */
long long dgb_launch_count(int32_t reset);

/* ---- device-side smoother control block ------------------------------------------------
 * Carries the early-exit / divergence state of Relaxation.block_gauss_seidel_pyamg
 * (dgfem/relaxation.py:202-216) on the device, so a smoother call needs no host sync:
 * once `skip` is set the remaining sweeps of that call are no-ops. */
typedef struct dgb_smoother_ctl {
    double res0;      /* RMS residual at smoother entry            (relaxation.py:202) */
    double ratio;     /* last normalised residual                  (relaxation.py:208) */
    int32_t skip;     /* ratio < 1e-6 seen -> stop sweeping        (relaxation.py:211-213) */
    int32_t diverged; /* ratio > 1e10 seen -> caller raises/exits  (relaxation.py:214-216); STICKY: only the
                         host clears it, and while it is set every later smoother call on this block is
                         skipped (the reference has left the process at this point) */
    int32_t iters;    /* completed iterations of this call */
    int32_t calls;    /* smoother calls since the block was zeroed */
} dgb_smoother_ctl;

/* ---- the operator of one level ---------------------------------------------------------
 * HOST struct of DEVICE pointers, passed by pointer to the solve-phase entry points.
 * `stencil` selects the kernel family:
 *   -1  arbitrary BSR matrix (Ni*Nj block rows): generic row-per-thread kernels;
 *   >=0 the DG 5-point block stencil on the Ni x Nj element grid, value = DGB_FLAG_PERIODIC_I|J
 *       (| DGB_FLAG_GHOST_LO|HI) bits, structure verified once with dgb_check_stencil: the
 *       single-launch lexicographic smoother kernels are used (gs_chain or gs_data, plus
 *       gs_mailbox; else the generic kernels run).
 * `data`, `dinv` (and gs_data) need 16 bytes of readable slack behind the last block: all kernels
 * read matrix rows in aligned 16-byte pieces. */
typedef struct dgb_operator {
    int32_t Ni, Nj;          /* element grid; N = Ni*Nj block rows                        */
    int32_t b, nnzb;         /* block size, number of stored blocks                        */
    int32_t stencil;         /* see above                                                  */
    int32_t reserved;
    const double *data;      /* [nnzb][b][b]                  (dgfem/discrete_system.py:145) */
    const int32_t *indices;  /* [nnzb]                                                     */
    const int32_t *indptr;   /* [N+1]                                                      */
    const double *dinv;      /* [N][b][b] inverse diagonal blocks (NULL for apply/residual) */
    const double *gs_data;   /* [nnzb][b][b] smoother stream, or NULL                      */
    double *gs_mailbox;      /* [N*b] row hand-over scratch of the lexicographic GS kernel:
                                every 8 bytes 0xFF (all-ones NaN) outside a pass, or NULL   */
    double *gs_chain;        /* [dgb_gs_chain_len()] record stream of the chained lexicographic GS
                                kernel (dgb_build_gs_chain), or NULL                        */
    /* stencil == -1 only: level schedule of the lexicographic sweep on an arbitrary, structurally symmetric BSR
     * matrix (the global-order Stokes blocks).  gs_rows (device): the N block rows grouped by forward dependency
     * level, then the N rows grouped by backward level; h_gs_offsets (HOST): fwd offsets [nlevels_fwd + 1], then
     * bwd offsets [nlevels_bwd + 1].  level(k) = 1 + max level of the rows j < k (fwd; j > k bwd) stored in row k. */
    const int32_t *gs_rows;
    const int32_t *h_gs_offsets;
    int32_t gs_nlevels_fwd, gs_nlevels_bwd;
} dgb_operator;

#define DGB_FLAG_PERIODIC_I 1
#define DGB_FLAG_PERIODIC_J 2
#define DGB_FLAG_MINV 4
/* element-slab partitioning (one slab of whole j-rows per GPU): the local Ni x Nj grid carries one
 * ghost element row below / above (the neighbour slab's edge row).  Ghost rows have vector entries
 * (the halo) but no matrix rows; every kernel leaves them untouched. */
#define DGB_FLAG_GHOST_LO 8
#define DGB_FLAG_GHOST_HI 16

/* 0 = auto (single-launch smoother kernels where the operator allows), 1 = generic kernels only.
 * Tuning values (tests and probes): 100+v experiment switch of the smoother kernels, 300+mask block sizes the
 * chained Gauss-Seidel kernel is used for (bit 0..4 = b 4, 9, 16, 25, 36; default all), 400+n CTAs per
 * thread-block cluster of that kernel.
 * Returns the previous setting of the 0/1 switch. */
int dgb_set_kernel_path(int32_t path);
/* Error flag of the asynchronous kernels of the current device (0 ok, 1 = TMA/mbarrier wait timed out,
 * 2 = row dependency wait timed out).  Synchronises the device.  A non-zero flag means the pass that raised it
 * (and every later one until the flag is reset) returned early: u is partly updated and the level's
 * gs_mailbox may hold stale values -- reset the flag, refill every gs_mailbox with dgb_fill_sentinel and treat
 * the solve as failed. */
int dgb_device_error(int32_t reset);
/* v[0..n) = all-ones NaN (the "not delivered" mark of gs_mailbox). */
int dgb_fill_sentinel(double *v, int64_t n, void *stream);

/* ---- K5: block-sparse operator apply / residual ---------------------------------------
 * replaces  grid.BSR @ u  (scipy bsr_matvec) -- dgfem/solver.py:117,119,150;
 *           dgfem/relaxation.py:202,208; utils/helpers.py:39 */
int dgb_bsr_apply(const dgb_operator *h_op, const double *x, double *y, void *stream);

/* r = rhs - A x (r may be NULL: norm only) and sum(r^2) into *sumsq (device scalar),
 * replaces  RHS - grid.BSR @ u  + compute_Lp_norm(.,2)  (utils/helpers.py:16-39).
 * `partials` is a workspace of dgb_partials_len() doubles.  skip: optional device flag
 * (dgb_smoother_ctl.skip); the call is a no-op when *skip != 0. */
int dgb_bsr_residual(const dgb_operator *h_op, const double *rhs, const double *x, double *r,
                     double *partials, double *sumsq, const int32_t *skip, void *stream);

/* The same right after a 2-colour pass that relaxed colour `relaxed` ((i + j + shift) & 1) last: those rows satisfy
 * their equations -- their residual (rounding noise) is written as zero -- and only the rows of the other colour
 * are evaluated: half the traffic of dgb_bsr_residual.  DG 5-point stencil only. */
int dgb_bsr_residual_colour(const dgb_operator *h_op, const double *rhs, const double *x, double *r, int32_t relaxed,
                            int32_t shift, double *partials, double *sumsq, const int32_t *skip, void *stream);

/* C = A B for BSR matrices with b x b blocks on a given structure of C (c_indptr / c_indices, any order inside a
 * row): every stored block of C is the sum over k of A_ik B_kj; replaces  grid.BSR_block_D @ grid.BSR_block_G
 * (dgfem/relaxation.py:240). */
int dgb_bsr_spgemm(int32_t b, int32_t n_brow, const int32_t *a_indptr, const int32_t *a_indices, const double *a_data,
                   const int32_t *b_indptr, const int32_t *b_indices, const double *b_data, const int32_t *c_indptr,
                   const int32_t *c_indices, double *c_data, void *stream);

/* sum(v^2) of a plain vector into *sumsq (device scalar). */
int dgb_sumsq(const double *v, int64_t n, double *partials, double *sumsq, void *stream);

/* ---- K6: batched inverse of the diagonal blocks ---------------------------------------
 * replaces  pyamg.util.utils.get_block_diag(A, blocksize, inv_flag=True)
 *           (dgfem/pyamg_relaxation.py:230-231; recomputed per call there, once here)
 * and the per-row np.linalg.solve(D_i, .) of dgfem/relaxation.py:148,194.
 * dinv[n_brow][b][b]; *info (device int) is set to 1+row of the first singular block. */
int dgb_block_diag_inverse(const double *data, const int32_t *indices, const int32_t *indptr,
                           int32_t n_brow, int32_t b, double *dinv, int32_t *info, void *stream);

/* Build the smoother stream: a copy of `data` with every diagonal block replaced by its
 * inverse (so one directional pass streams exactly nnzb blocks). */
int dgb_build_gs_stream(const double *data, const int32_t *indices, const int32_t *indptr,
                        const double *dinv, int32_t n_brow, int32_t b, double *gs_data,
                        void *stream);

/* Chained lexicographic Gauss-Seidel (non-periodic DG stencil, b in {4, 9, 16, 25}): per sweep direction
 * one record {-Dinv_e A_e,row-predecessor, -Dinv_e A_e,previous-row, c_e} per element, so that the
 * dependency chain of pyamg's block_gauss_seidel (dgfem/pyamg_relaxation.py:252-255) only needs two
 * b x b products per element; everything else (c_e) is computed by a dependency-free kernel first.
 * dgb_gs_chain_len: doubles to allocate for h_op->gs_chain (0 = not supported for this operator, the
 * row-pipelined kernel on gs_data is used); dgb_build_gs_chain fills it from data / dinv. */
int64_t dgb_gs_chain_len(int32_t b, int32_t Ni, int32_t Nj, int32_t stencil);
int dgb_build_gs_chain(const dgb_operator *h_op, void *stream);

/* *mismatch (device int) = number of block rows whose (indptr, indices) differ from the
 * closed-form 5-point stencil of an Ni x Nj DG grid with the given periodicity flags. */
int dgb_check_stencil(const int32_t *indices, const int32_t *indptr, int32_t Ni, int32_t Nj,
                      int32_t flags, int32_t *mismatch, void *stream);

/* ---- K7: block Gauss-Seidel, one directional pass ---------------------------------------
 * replaces  pyamg.amg_core.block_gauss_seidel(Ap,Aj,Ax,x,b,Dinv,row_start,row_stop,row_step,bs)
 *           (dgfem/pyamg_relaxation.py:252-255).
 * direction: +1 forward (rows 0..N-1), -1 backward (rows N-1..0).
 * mode: DGB_GS_LEXICOGRAPHIC reproduces the lexicographic order exactly (row-pipelined kernel,
 *       or the anti-diagonal wavefront c=i+j with the generic kernels; both need the 5-point
 *       block stencil the DG operator has); DGB_GS_REDBLACK is the 2-colour multicolour
 *       variant (forward = colour 0 then 1, backward = 1 then 0).
 * skip: optional device flag (ctl->skip); when non-zero the pass is a no-op. */
#define DGB_GS_LEXICOGRAPHIC 0
#define DGB_GS_REDBLACK 1
int dgb_block_gs_pass(const dgb_operator *h_op, const double *rhs, double *x, int32_t direction,
                      int32_t mode, const int32_t *skip, void *stream);

/* The same lexicographic pass as part of a sequence: prev_direction is the direction of the immediately
 * preceding lexicographic pass on the SAME rhs, with x untouched since except for the ghost rows of a slab
 * (0 = none / unknown).  When it is the opposite of `direction`, the chained kernel finds its right-hand
 * sides where the previous pass left them and skips the dependency-free launch (slabs: only the ghost-row
 * terms are refreshed). */
int dgb_block_gs_pass_seq(const dgb_operator *h_op, const double *rhs, double *x, int32_t direction,
                          int32_t prev_direction, const int32_t *skip, void *stream);

/* Entry residual of a smoother call fused with the dependency-free part of its first lexicographic pass
 * (direction first_direction): r = rhs - A x (r may be NULL), *sumsq = sum r^2, and the chained kernel's
 * right-hand sides for that pass -- follow with dgb_block_gs_pass_seq(..., direction = first_direction,
 * prev_direction = -first_direction).  Returns DGB_UNSUPPORTED (nothing launched) when the operator has no
 * chained kernel; the caller then uses dgb_bsr_residual. */
#define DGB_UNSUPPORTED 100
int dgb_block_gs_entry_residual(const dgb_operator *h_op, const double *rhs, const double *x,
                                int32_t first_direction, double *r, double *partials, double *sumsq,
                                void *stream);

/* Residual right after a chained lexicographic pass (direction last_direction) on the same rhs, x untouched since:
 * r = rhs - A x (r may be NULL) and *sumsq = sum r^2, evaluated from the record stream the pass left behind plus the
 * diagonal blocks (2 b^2 + 2 b + b^2 doubles per element instead of the 5 blocks of dgb_bsr_residual).  Replaces
 * the residual test of every smoother iteration (dgfem/relaxation.py:208).  Returns DGB_UNSUPPORTED (nothing
 * launched) when the operator has no chained kernel, lives on a slab with ghost rows, or b is not 4, 9 or 16. */
int dgb_block_gs_residual_after_pass(const dgb_operator *h_op, const double *x, int32_t last_direction, double *r,
                                     double *partials, double *sumsq, const int32_t *skip, void *stream);

/* One colour class of the 2-colour sweep: rows with ((i + j + shift) & 1) == colour are relaxed in place.
 * `shift` carries the global row parity of a slab so that all ranks colour the global grid alike; the
 * caller exchanges halos between the two colours. */
int dgb_block_gs_colour(const dgb_operator *h_op, const double *rhs, double *x, int32_t colour, int32_t shift,
                        const int32_t *skip, void *stream);

/* Entry of a 2-colour smoother call (the residual before the first iteration, dgfem/relaxation.py:202-204, together
 * with the first colour of the first pass): r = rhs - A x over all rows (r may be NULL), *sumsq = sum r^2, and the
 * rows of colour `first` relaxed in place unless *frozen (device flag, may be NULL) is set.  The rows of `first`
 * read their four neighbour blocks once for both results (5.5 b^2 doubles per element instead of 7.5).  Needs the DG
 * stencil layout (stencil >= 0), otherwise DGB_UNSUPPORTED and nothing is launched. */
int dgb_block_gs_colour_entry(const dgb_operator *h_op, const double *rhs, double *x, double *r, int32_t first,
                              int32_t shift, double *partials, double *sumsq, const int32_t *frozen, void *stream);

/* ---- K8: one block-row relaxation sweep, x_out_i = omega*Dinv_i(rhs_i - sum_{j!=i} A_ij x_in_j)
 *                                                   + (1-omega) x_in_i
 * x_out != x_in : block-Jacobi            (first iteration of dgfem/relaxation.py:123-150)
 * x_out == x_in : forward block-GS, lexicographic order (dgfem/relaxation.py:170-195 and
 *                 iterations >= 2 of block_jacobi, whose `u = u_new` aliases the buffers). */
int dgb_block_relax_sweep(const dgb_operator *h_op, const double *rhs, const double *x_in,
                          double *x_out, double omega, void *stream);

/* ---- smoother control ------------------------------------------------------------------ */
/* ctl->res0 = sqrt(*sumsq / n); skip = diverged (sticky); iters = 0; calls += 1 */
int dgb_smoother_begin(dgb_smoother_ctl *ctl, const double *sumsq, int64_t n, void *stream);
/* ratio = sqrt(*sumsq / n) / res0; apply the 1e-6 / 1e10 tests; iters += 1 (unless skipped) */
int dgb_smoother_check(dgb_smoother_ctl *ctl, const double *sumsq, int64_t n, void *stream);

/* Whole smoother call on the device:
 * replaces  Relaxation.block_gauss_seidel_pyamg(grid, RHS, u, direction, omega, max_iterations)
 *           (dgfem/relaxation.py:198-218); u is updated IN PLACE (the host wrapper copies).
 * direction: 0 symmetric, +1 forward, -1 backward.  `check_residual` = 0 drops the
 * per-iteration residual norms (and with them the early exit) -- NOT reference semantics. */
int dgb_block_gauss_seidel_pyamg(const dgb_operator *h_op, const double *rhs, double *u,
                                 int32_t direction, int32_t max_iterations, int32_t mode,
                                 int32_t check_residual, dgb_smoother_ctl *ctl, double *partials,
                                 double *sumsq, void *stream);

/* ---- K9: level transfer ----------------------------------------------------------------
 * replaces the einsum('ij,kj->ki', R|P, .) transfers of Solver.multigrid_V_cycle
 * (dgfem/solver.py:152-193) with the dense operators built in
 * DGFEM.assemble_multigrid_operators (dgfem/dgfem.py:303-372).
 * kind: DGB_TRANSFER_P  per-element  coarse[e] = R[nc x nf] fine[e]
 *       DGB_TRANSFER_H  2x2 children gathered with the reference's reshape/transpose
 *                       (solver.py:164): R is [b x 4b], Ni_c x Nj_c coarse elements.
 * restrict: coarse = R * gather(fine);  prolong_add: fine += scatter(P * coarse). */
#define DGB_TRANSFER_P 1
#define DGB_TRANSFER_H 2
int dgb_restrict(int32_t kind, const double *R, int32_t nc, int32_t nf, int32_t Ni_c,
                 int32_t Nj_c, const double *fine, double *coarse, void *stream);
int dgb_prolong_add(int32_t kind, const double *P, int32_t nc, int32_t nf, int32_t Ni_c,
                    int32_t Nj_c, const double *coarse, double *fine, void *stream);

/* Slab variants (multi-GPU): Nj_c counts the coarse level's local rows INCLUDING its ghost rows
 * (ghost_c_lo / ghost_c_hi in {0,1}); only active coarse rows are written (restrict) / read (prolong).
 * For DGB_TRANSFER_H the children of coarse element (I, J) are the fine elements (2I+a_i, 2J+a_j) with
 * child slot a_j*2+a_i -- what the reference's gather (dgfem/solver.py:164) means on the square global
 * grid -- and fine local row = ghost_f_lo + 2*(J - ghost_c_lo) + a_j. */
#define DGB_TRANSFER_H_SLAB 3
int dgb_restrict_slab(int32_t kind, const double *R, int32_t nc, int32_t nf, int32_t Ni_c, int32_t Nj_c,
                      int32_t ghost_c_lo, int32_t ghost_c_hi, int32_t ghost_f_lo, const double *fine,
                      double *coarse, void *stream);
int dgb_prolong_add_slab(int32_t kind, const double *P, int32_t nc, int32_t nf, int32_t Ni_c, int32_t Nj_c,
                         int32_t ghost_c_lo, int32_t ghost_c_hi, int32_t ghost_f_lo, const double *coarse,
                         double *fine, void *stream);

/* ---- V-cycle driver ---------------------------------------------------------------------
 * replaces  Solver.multigrid_V_cycle(k, RHS, u)  (dgfem/solver.py:141-207).
 * levels[0] is the coarsest grid, levels[nlevels-1] the finest (the order of
 * Solver.grids); levels[k].R/P map between level k (coarse) and k+1 (fine). */
typedef struct dgb_level {
    dgb_operator op;
    double *rhs;             /* [Ni*Nj*b] work: right-hand side of this level  */
    double *u;               /* [Ni*Nj*b] work: iterate of this level          */
    double *r;               /* [Ni*Nj*b] work: residual                       */
    /* transfer between this level (coarse side) and the next finer one */
    int32_t transfer_kind;   /* 0 on the finest level                          */
    int32_t nc, nf;          /* R is [nc x nf], P is [nf x nc]                 */
    int32_t pad0;
    const double *R;
    const double *P;
    /* smoother settings of the coarsening that owns this level (paramfile.yml:20-65): the pre-smoother
     * (also the coarsest level's 'smoother' solver, dgfem/solver.py:202-204) ... */
    int32_t smoother;        /* DGB_SMOOTHER_*                                 */
    int32_t direction;       /* 0 symmetric, +1 forward, -1 backward           */
    int32_t pre_iterations, post_iterations;
    double omega;
    /* ... and the post-smoother, resolved independently (dgfem/solver.py:144,196) */
    int32_t post_smoother;   /* DGB_SMOOTHER_*                                 */
    int32_t post_direction;
    double post_omega;
} dgb_level;

#define DGB_SMOOTHER_BLOCK_GS_PYAMG 0
#define DGB_SMOOTHER_BLOCK_JACOBI 1
#define DGB_SMOOTHER_BLOCK_GS 2

typedef struct dgb_vcycle_opts {
    int32_t gs_mode;            /* DGB_GS_LEXICOGRAPHIC | DGB_GS_REDBLACK       */
    int32_t check_residual;     /* 1 = reference semantics (early exit active)  */
    int32_t coarse_iterations;  /* 10 in the reference (solver.py:204)          */
    int32_t coarse_solver;      /* DGB_COARSE_SMOOTHER | DGB_COARSE_DIRECT      */
    void *u_final_event;        /* optional cudaEvent_t, recorded on `stream` right after the last kernel
                                   that writes the finest level's u (the post-smoother's closing residual
                                   test only reads it): a caller may copy u out on another stream from
                                   there on.  NULL = not used.                  */
    const double *coarse_inverse; /* DGB_COARSE_DIRECT: dense inverse [n x n] of levels[0]'s operator,
                                   n = Ni*Nj*b (dgb_dense_inverse)              */
} dgb_vcycle_opts;

/* coarse grid solver (paramfile.yml:23, dgfem/solver.py:199-204) */
#define DGB_COARSE_SMOOTHER 0
#define DGB_COARSE_DIRECT 1

/* Coarse-grid direct solve, replaces  splin.spsolve(grid.BSR.tocsr(), RHS)  (dgfem/solver.py:56-59,199-200):
 * dgb_dense_inverse expands the BSR operator of a (small) level into a dense n x n matrix, n = n_brow*b, and
 * inverts it in place by Gauss-Jordan elimination with partial pivoting (once per hierarchy); *info (device
 * int) = 1 + column of the first zero pivot.  dgb_dense_solve then is u = inverse * rhs. */
int dgb_dense_inverse(const double *data, const int32_t *indices, const int32_t *indptr, int32_t n_brow,
                      int32_t b, double *inverse, int32_t *info, void *stream);
int dgb_dense_solve(const double *inverse, int32_t n, const double *rhs, double *u, void *stream);

/* One V-cycle on the finest level: levels[n-1].u is updated in place from levels[n-1].rhs.
 * ctl: array of nlevels control blocks (device); partials/sumsq: workspaces. */
int dgb_vcycle(const dgb_level *h_levels, int32_t nlevels, const dgb_vcycle_opts *h_opts,
               dgb_smoother_ctl *ctl, double *partials, double *sumsq, void *stream);

/* The same cycle for a caller that has just evaluated the finest level's residual itself -- Solver.solve_multigrid
 * does, after every cycle (dgfem/solver.py:117-119), and that vector is exactly the entry residual the next cycle's
 * pre-smoother opens with (dgfem/relaxation.py:202).  flags = DGB_VCYCLE_ENTRY_PRIMED: the caller ran
 * dgb_block_gs_entry_residual(&levels[n-1].op, levels[n-1].rhs, levels[n-1].u, first direction of the pre-smoother,
 * levels[n-1].r, partials, sumsq, stream) on the current u; the cycle then starts with the first pass (one residual
 * evaluation of the finest level saved per cycle of a solve).  Returns DGB_UNSUPPORTED (nothing launched) when the
 * finest level's pre-smoother does not open that way (other smoother / 2-colour mode / no residual tests / no
 * chained kernel); flags = 0 is dgb_vcycle. */
#define DGB_VCYCLE_ENTRY_PRIMED 1
int dgb_vcycle_ex(const dgb_level *h_levels, int32_t nlevels, const dgb_vcycle_opts *h_opts,
                  dgb_smoother_ctl *ctl, double *partials, double *sumsq, void *stream, int32_t flags);

/* ---- K0-K3: DG assembly (Poisson) ------------------------------------------------------- */
/* Basis / quadrature / geometry-operator tables of one level, uploaded once
 * (replaces Grid.initialize_interpolation, dgfem/grid.py:178-213).  All arrays are HOST
 * pointers, row-major; copied into a device-resident table block owned by the returned
 * handle.  Layout documented in DESIGN.md section "Tables". */
typedef struct dgb_tables dgb_tables;
typedef struct dgb_tables_desc {
    int32_t Pg;      /* geometry degree, ng = (Pg+1)^2 nodes per element          */
    int32_t p;       /* solution degree, b = (p+1)^2                              */
    int32_t nq1;     /* N_int (points per direction)                              */
    int32_t cf;      /* coarsening factor of this level (1 = fine)                */
    const double *h_V, *h_Vr, *h_Vs;            /* [nq1^2][b] volume tables       */
    const double *h_w2;                         /* [nq1^2] 2-D weights (r fastest)*/
    const double *h_w1;                         /* [nq1]                          */
    const double *h_Vf, *h_Vrf, *h_Vsf;         /* [4][nq1][b] traces: iL,iR,jL,jR*/
    /* geometry operators: element nodes (F-order) -> x, x_r, x_s at points.
     * volume: [nq1^2][ng]; faces [4][nq1][ng] in the order imin,imax,jmin,jmax.
     * For cf > 1 every point also carries the fine sub-element offset it samples
     * (dgfem/element.py:273-310): sub_vol[nq1^2][2], sub_face[4][nq1][2]. */
    const double *h_GX, *h_GR, *h_GS;
    const double *h_FX, *h_FR, *h_FS;
    const int32_t *h_sub_vol, *h_sub_face;
} dgb_tables_desc;
int dgb_tables_create(const dgb_tables_desc *h_desc, dgb_tables **out);
void dgb_tables_destroy(dgb_tables *t);

/* K0: geometric terms of all Ni x Nj elements of a level
 * replaces Element.compute_geometric_terms / metric_xy_rs (dgfem/element.py:52-130) and
 * CoarseElement._init_coarse_element (dgfem/element.py:242-356).
 * xn, yn: node coordinates in Plot3D file order [jl][il] (i fastest; dgfem/grid.py:49-51),
 * il = Ni_fine*Pg+1.  Outputs (element-major, m = j*Ni+i):
 *   vol [N][7][nq]   : J, rx, sx, ry, sy, x, y at the volume points
 *   face[N][4][8][nq1]: per face (imin,imax,jmin,jmax): J_f, alpha, beta, x, y, nx, ny, 0
 *                       with alpha = rx*nx+ry*ny, beta = sx*nx+sy*ny (so d_n phi = Vr*alpha+Vs*beta)
 *   area[N]          : A = sum J w                                   (dgfem/element.py:30) */
int dgb_metrics(const dgb_tables *t, const double *xn, const double *yn, int32_t il,
                int32_t Ni, int32_t Nj, double *vol, double *face, double *area, void *stream);

/* K1+K2: assemble the Poisson operator of a level straight into BSR
 * replaces Poisson.assemble_BSR_Poisson (dgfem/discrete_system.py:54-145) with
 * Element.compute_momentum_laplace_volume_integral / compute_mass_matrix
 * (dgfem/element.py:181-199,132-133) and Face.compute_momentum_laplace_SIP_terms
 * (dgfem/face.py:115-280).  flags: bit0 periodic in i (O-grid), bit1 periodic in j,
 * bit2 multiply by the inverse mass matrix.
 * Outputs: indptr[N+1], indices[nnzb], data[nnzb][b][b], minv[N][b][b]. */
int64_t dgb_poisson_nnzb(int32_t Ni, int32_t Nj, int32_t flags);
int dgb_assemble_poisson(const dgb_tables *t, const double *vol, const double *face,
                         const double *area, int32_t Ni, int32_t Nj, double nu, double sigma,
                         int32_t flags, int32_t *indptr, int32_t *indices, double *data,
                         double *minv, void *stream);

/* K3: right-hand side
 * replaces Poisson.assemble_RHS_Poisson (dgfem/discrete_system.py:355-403).
 * f_vol[N][nq]: source at the volume points; g_face[N][4][nq1]: Dirichlet data at the
 * face points (only domain-boundary faces are read). */
int dgb_assemble_rhs(const dgb_tables *t, const double *vol, const double *face,
                     const double *area, const double *minv, const double *f_vol,
                     const double *g_face, int32_t Ni, int32_t Nj, double nu, double sigma,
                     int32_t flags, double *rhs, void *stream);

/* ---- K4: Stokes, local ordering ---------------------------------------------------------
 * replaces Stokes.assemble_BSR_Stokes_local_order (dgfem/discrete_system.py:812-965) and
 * Stokes.assemble_RHS_Stokes (dgfem/discrete_system.py:967-1028): one (2 b_u + b_p)^2 block per element
 * pair, rows (x-momentum, y-momentum, continuity), columns (u, v, p), no inverse-mass scaling.
 * Four table sets: velocity basis @ velocity points (t_uu), pressure basis @ velocity points (t_pu),
 * velocity basis @ pressure points (t_up), pressure basis @ pressure points (t_pp); dgb_metrics is run
 * once per point set (vol_u/face_u with t_uu, vol_p/face_p with t_up); area comes from the velocity points
 * (dgfem/element.py:30).  pin_pressure != 0 reproduces the direct solver's pressure pin
 * (discrete_system.py:946).  indptr[N+1], indices[nnzb], data[nnzb][bt][bt] with nnzb =
 * dgb_poisson_nnzb(Ni, Nj, flags) (same block structure as the Poisson operator). */
int dgb_assemble_stokes(const dgb_tables *t_uu, const dgb_tables *t_pu, const dgb_tables *t_up,
                        const dgb_tables *t_pp, const double *vol_u, const double *face_u,
                        const double *vol_p, const double *face_p, const double *area, int32_t Ni,
                        int32_t Nj, double nu, double sigma, double gamma, int32_t flags,
                        int32_t pin_pressure, int32_t *indptr, int32_t *indices, double *data,
                        void *stream);
/* f_mom[N][2][nq_u]: momentum source (x, y) at the velocity points; f_cont[N][nq_p]: continuity source
 * at the pressure points; g_u[N][4][2][nq1_u], g_p[N][4][2][nq1_p]: exact velocity (u, v) at the face
 * points of both point sets (only domain-boundary faces are read). */
int dgb_assemble_rhs_stokes(const dgb_tables *t_uu, const dgb_tables *t_pu, const dgb_tables *t_up,
                            const dgb_tables *t_pp, const double *vol_u, const double *face_u,
                            const double *vol_p, const double *face_p, const double *area,
                            const double *f_mom, const double *f_cont, const double *g_u,
                            const double *g_p, int32_t Ni, int32_t Nj, double nu, double sigma,
                            double gamma, int32_t flags, double *rhs, void *stream);

/* ---- multi-GPU: element slabs over peer memory (NVLink / NVSwitch) -------------------------------------
 * The reference is a single process; these entry points carry its V-cycle across slabs of
 * whole element rows, one process per GPU.  Every rank creates a communicator with an arena of the SAME size, the
 * 64-byte export handles are exchanged out of band (any bootstrap: torch.distributed, MPI, a file) and every rank
 * connects.  Allocation inside the arena is the caller's, with one rule: identical offsets on every rank.
 * All collectives are stream-ordered kernel launches (no host synchronisation); the waits inside them are bounded
 * (20 s) and report through dgb_comm_error (3 = a peer did not arrive).  Ranks must issue the same sequence of
 * collectives.  One rank per device: kernels of different ranks wait for each other. */
typedef struct dgb_comm dgb_comm;
int dgb_comm_create(int32_t rank, int32_t world, int64_t arena_bytes, dgb_comm **out);
int dgb_comm_handle_bytes(void);                          /* 64 */
int dgb_comm_export(dgb_comm *c, void *h_handle);         /* h_handle: dgb_comm_handle_bytes() bytes          */
int dgb_comm_connect(dgb_comm *c, const void *h_handles); /* world handles, rank order (own slot ignored)      */
void *dgb_comm_arena(dgb_comm *c, int64_t *bytes);        /* device pointer of the arena's allocatable part    */
int dgb_comm_error(dgb_comm *c, int32_t reset);           /* synchronises                                      */
void dgb_comm_destroy(dgb_comm *c);

/* Halo exchange of one vector.  `block` (in the arena, same offset on every rank) holds rows + 2 element rows of
 * row_doubles doubles: [ghost row below | rows owned rows | ghost row above]; after the call the ghost rows hold the
 * neighbour slabs' edge rows (rank - 1's last, rank + 1's first owned row).  Replaces nothing in the reference: it is
 * what makes  grid.BSR @ u  (dgfem/solver.py:150, relaxation.py:202,208) and the smoother passes see the neighbour
 * slab. */
int dgb_halo_exchange(dgb_comm *c, double *block, int64_t row_doubles, int32_t rows, void *stream);

/* *value (device scalar) <- sum over ranks, added in rank order on every rank (identical bits everywhere).
 * mode 0: only that; 1: then dgb_smoother_begin(ctl, value, n_global); 2: then dgb_smoother_check(...). */
int dgb_allreduce_sum(dgb_comm *c, double *value, int32_t mode, dgb_smoother_ctl *ctl, int64_t n_global, void *stream);

/* dst_block (in the arena, same offset everywhere, world*chunk doubles) <- [src of rank 0 | src of rank 1 | ...]
 * on every rank ("coarsest level gathered", north_star; here every rank holds the gathered level). */
int dgb_allgather(dgb_comm *c, const double *src, double *dst_block, int64_t chunk, void *stream);

/* One level of a slab hierarchy: `lev` as in dgb_vcycle (op carries the DGB_FLAG_GHOST_* bits, Nj counts the ghost
 * rows; lev.R / lev.P / transfer_kind link this level, as the coarse side, to the next finer slab level);
 * lev.u must be u_block + (1 - ghost_lo) * Ni * b. */
typedef struct dgb_slab_level {
    dgb_level lev;
    double *u_block;           /* [rows + 2][Ni * b] in the arena                                   */
    int32_t ghost_lo, ghost_hi;
    int32_t colour_shift;      /* (global row of local row 0) & 1: all ranks colour the global grid   */
    int32_t pad;
    int64_t n_global;          /* scalar unknowns of the level over all ranks (residual norms)        */
} dgb_slab_level;

#define DGB_GS_SLAB_LEXICOGRAPHIC 2   /* lexicographic inside a slab, neighbour rows as before the pass */

typedef struct dgb_slab_opts {
    int32_t gs_mode;           /* DGB_GS_REDBLACK | DGB_GS_SLAB_LEXICOGRAPHIC                          */
    int32_t check_residual;
    /* link between the coarsest slab level (fine side) and the finest replicated level (coarse side) */
    int32_t link_kind, link_nc, link_nf;
    int32_t link_Ni_c, link_rows_c;   /* this rank's un-ghosted chunk of that coarse level             */
    int32_t n_coarse;
    const double *link_R, *link_P;
    double *link_rhs_local;    /* [link_rows_c * link_Ni_c * link_nc] scratch                          */
    /* the replicated hierarchy below (every rank runs it on the gathered level: identical bits);
     * coarse_levels[n_coarse - 1].rhs is the all-gather destination and must lie in the arena */
    const dgb_level *coarse_levels;
    dgb_smoother_ctl *coarse_ctl;
    dgb_vcycle_opts coarse_opts;
} dgb_slab_opts;

/* One V-cycle on the finest slab level (levels[nlevels - 1]): Solver.multigrid_V_cycle (dgfem/solver.py:141-207)
 * with the smoother of dgfem/relaxation.py:198-218 across slabs.  ctl: nlevels control blocks. */
int dgb_vcycle_slab(dgb_comm *comm, const dgb_slab_level *h_levels, int32_t nlevels, const dgb_slab_opts *h_opts,
                    dgb_smoother_ctl *ctl, double *partials, double *sumsq, void *stream);

/* ---- post-processing --------------------------------------------------------------------------
 * replaces the per-element loop of DGFEM.solve (dgfem/dgfem.py:188-232): u_nodal[e] = V_DOF_grid @ u_e at the
 * element's (Pg+1)^2 geometry nodes (node a_i + (Pg+1)*a_j, i fastest), minus the exact solution there, and
 *   sums[0] = sum |u_nodal - u_exact|,  sums[1] = sum (u_nodal - u_exact)^2   over all N*ng element nodes
 * (L1 = sums[0]/(N ng), L2 = sqrt(sums[1]/(N ng)), dgfem.py:220-221).
 * V_grid[ng][b] (Grid.initialize_interpolation's V_DOF_grid); exact_nodes[jl][il]: the exact solution at the grid
 * nodes in Plot3D file order; u_nodal[N][ng] optional (NULL = sums only); partials: dgb_partials_len() doubles. */
int dgb_nodal_error(const double *V_grid, int32_t ng, int32_t b, int32_t Pg, int32_t Ni, int32_t Nj, int32_t il,
                    const double *u, const double *exact_nodes, double *u_nodal, double *partials, double *sums,
                    void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DGB200_H */
