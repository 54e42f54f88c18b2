// dgb_vcycle.cu -- host-side V-cycle scheduler: enqueues the kernels of one multigrid V-cycle
// on a stream with no host synchronisation (the smoother's early-exit test lives on the
// device, see dgb_smoother_ctl).
//
// Restates Solver.multigrid_V_cycle (dgfem/solver.py:141-207):
//   k>1 : pre-smooth, r = rhs - A u, restrict, recurse from u_c = 0, prolong + correct, post-smooth
//   k==1: coarse grid solver 'smoother' = the pre-smoother with max_iterations = 10
#include "dgb_common.cuh"

namespace dgb {

// *have_residual (optional, out): L.r holds rhs - A u of the returned u (the smoother's own last residual test)
// (smoother, direction, omega): the pre- or the post-smoother's settings, resolved independently
// (dgfem/solver.py:143-147,196)
static int smooth(const dgb_level &L, int smoother, int direction, double omega, const dgb_vcycle_opts &o,
                  int iterations, dgb_smoother_ctl *ctl, double *partials, double *sumsq, void *stream,
                  bool *have_residual = nullptr, void *u_final_event = nullptr, bool u_is_zero = false,
                  bool entry_primed = false) {
    if (have_residual) *have_residual = false;
    if (iterations <= 0 || smoother != DGB_SMOOTHER_BLOCK_GS_PYAMG) {
        // no in-smoother hook: the event is recorded by the caller after the smoother returns
        if (u_final_event != nullptr && iterations <= 0)
            DGB_CUDA_OK(cudaEventRecord((cudaEvent_t)u_final_event, (cudaStream_t)stream));
    }
    if (iterations <= 0) return 0;
    const size_t nbytes = sizeof(double) * (size_t)L.op.Ni * L.op.Nj * L.op.b;
    switch (smoother) {
    case DGB_SMOOTHER_BLOCK_GS_PYAMG:
        if (have_residual && o.check_residual) {
            *have_residual = true;
            return gs_pyamg(&L.op, L.rhs, L.u, direction, iterations, o.gs_mode, 1, ctl, partials, sumsq, L.r, stream,
                            nullptr, u_is_zero, entry_primed);
        }
        return gs_pyamg(&L.op, L.rhs, L.u, direction, iterations, o.gs_mode, o.check_residual, ctl, partials, sumsq,
                        nullptr, stream, u_final_event, u_is_zero);
    case DGB_SMOOTHER_BLOCK_JACOBI: {
        // relaxation.py:123-150: iteration 1 is Jacobi into a fresh buffer, then `u = u_new`
        // aliases the two, so the remaining iterations are in-place forward sweeps.
        int rc = dgb_block_relax_sweep(&L.op, L.rhs, L.u, L.r, omega, stream);
        for (int it = 1; it < iterations && rc == 0; ++it)
            rc = dgb_block_relax_sweep(&L.op, L.rhs, L.r, L.r, omega, stream);
        if (rc) return rc;
        DGB_CUDA_OK(cudaMemcpyAsync(L.u, L.r, nbytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return 0;
    }
    case DGB_SMOOTHER_BLOCK_GS: {
        int rc = 0;
        for (int it = 0; it < iterations && rc == 0; ++it)   // relaxation.py:170-195 (forward only)
            rc = dgb_block_relax_sweep(&L.op, L.rhs, L.u, L.u, omega, stream);
        return rc;
    }
    default:
        set_error("unknown smoother id %d", smoother);
        return 3;
    }
}

// u_zero: this level's iterate was just zeroed (solver.py:171) -- its first smoother call need not read the operator
// for its entry residual (r = rhs) and first right-hand sides (c = Dinv rhs): same bits, a sixth of the bytes
// entry_primed (finest level only): the caller has evaluated the pre-smoother's entry residual itself
// (dgb_block_gs_entry_residual on this level's rhs, u, r and `sumsq`) -- dgb_vcycle_ex
static int vcycle(const dgb_level *lv, int k, const dgb_vcycle_opts &o, dgb_smoother_ctl *ctl,
                  double *partials, double *sumsq, void *stream, bool top = false, bool u_zero = false,
                  bool entry_primed = false) {
    const dgb_level &L = lv[k];
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    void *ev = top ? o.u_final_event : nullptr;
    if (k == 0 && o.coarse_solver == DGB_COARSE_DIRECT) {   // solver.py:199-200
        rc = dgb_dense_solve(o.coarse_inverse, L.op.Ni * L.op.Nj * L.op.b, L.rhs, L.u, stream);
        if (rc == 0 && ev != nullptr) DGB_CUDA_OK(cudaEventRecord((cudaEvent_t)ev, st));
        return rc;
    }
    if (k == 0) {   // solver.py:201-204
        rc = smooth(L, L.smoother, L.direction, L.omega, o, o.coarse_iterations, ctl + k, partials, sumsq, stream, nullptr, ev,
                    u_zero);
        if (rc == 0 && ev != nullptr && L.smoother != DGB_SMOOTHER_BLOCK_GS_PYAMG)
            DGB_CUDA_OK(cudaEventRecord((cudaEvent_t)ev, st));
        return rc;
    }
    const dgb_level &C = lv[k - 1];
    bool have_r = false;
    if ((rc = smooth(L, L.smoother, L.direction, L.omega, o, L.pre_iterations, ctl + k, partials, sumsq, stream, &have_r,
                     nullptr, u_zero, entry_primed)))
        return rc;
    // residual = RHS - BSR @ u (solver.py:150); the pre-smoother's last residual test already evaluated exactly
    // this vector (same kernel, same inputs), so it is not computed twice
    if (!have_r && (rc = dgb_bsr_residual(&L.op, L.rhs, L.u, L.r, partials, sumsq, nullptr, stream))) return rc;
    if ((rc = dgb_restrict(C.transfer_kind, C.R, C.nc, C.nf, C.op.Ni, C.op.Nj, L.r, C.rhs, stream))) return rc;
    DGB_CUDA_OK(cudaMemsetAsync(C.u, 0, sizeof(double) * (size_t)C.op.Ni * C.op.Nj * C.op.b, st));   // solver.py:171
    if ((rc = vcycle(lv, k - 1, o, ctl, partials, sumsq, stream, false, true))) return rc;
    if ((rc = dgb_prolong_add(C.transfer_kind, C.P, C.nc, C.nf, C.op.Ni, C.op.Nj, C.u, L.u, stream))) return rc;
    rc = smooth(L, L.post_smoother, L.post_direction, L.post_omega, o, L.post_iterations, ctl + k, partials, sumsq,
                stream, nullptr, ev);
    if (rc == 0 && ev != nullptr && L.post_smoother != DGB_SMOOTHER_BLOCK_GS_PYAMG && L.post_iterations > 0)
        DGB_CUDA_OK(cudaEventRecord((cudaEvent_t)ev, st));
    return rc;
}

// the same cycle for the slab driver (dgb_comm.cu): the replicated coarse hierarchy below the distributed levels
int vcycle_entry(const dgb_level *lv, int k, const dgb_vcycle_opts &o, dgb_smoother_ctl *ctl, double *partials,
                 double *sumsq, void *stream, bool u_zero) {
    return vcycle(lv, k, o, ctl, partials, sumsq, stream, false, u_zero);
}

static int vcycle_args_ok(const dgb_level *h_levels, int32_t nlevels, const dgb_vcycle_opts *h_opts,
                          dgb_smoother_ctl *ctl, double *partials, double *sumsq) {
    DGB_ARG(h_levels && h_opts && ctl && partials && sumsq && nlevels >= 1);
    DGB_ARG(h_opts->coarse_solver == DGB_COARSE_SMOOTHER ||
            (h_opts->coarse_solver == DGB_COARSE_DIRECT && h_opts->coarse_inverse != nullptr));
    for (int k = 0; k < nlevels; ++k) {
        const dgb_level &L = h_levels[k];
        DGB_ARG(L.op.data && L.op.indices && L.op.indptr && L.op.dinv && L.rhs && L.u && L.r);
        if (k < nlevels - 1) {
            DGB_ARG(L.R && L.P && (L.transfer_kind == DGB_TRANSFER_P || L.transfer_kind == DGB_TRANSFER_H));
            const dgb_level &F = h_levels[k + 1];
            if (L.transfer_kind == DGB_TRANSFER_P) {
                DGB_ARG(L.nc == L.op.b && L.nf == F.op.b && L.op.Ni == F.op.Ni && L.op.Nj == F.op.Nj);
            } else {
                DGB_ARG(L.nc == L.op.b && L.nf == 4 * F.op.b && 2 * L.op.Ni == F.op.Ni && 2 * L.op.Nj == F.op.Nj);
            }
        }
    }
    return 0;
}

}  // namespace dgb

extern "C" int dgb_vcycle(const dgb_level *h_levels, int32_t nlevels, const dgb_vcycle_opts *h_opts,
                          dgb_smoother_ctl *ctl, double *partials, double *sumsq, void *stream) {
    int rc = dgb::vcycle_args_ok(h_levels, nlevels, h_opts, ctl, partials, sumsq);
    if (rc) return rc;
    return dgb::vcycle(h_levels, nlevels - 1, *h_opts, ctl, partials, sumsq, stream, true);
}

extern "C" int dgb_vcycle_ex(const dgb_level *h_levels, int32_t nlevels, const dgb_vcycle_opts *h_opts,
                             dgb_smoother_ctl *ctl, double *partials, double *sumsq, void *stream, int32_t flags) {
    int rc = dgb::vcycle_args_ok(h_levels, nlevels, h_opts, ctl, partials, sumsq);
    if (rc) return rc;
    DGB_ARG((flags & ~DGB_VCYCLE_ENTRY_PRIMED) == 0);
    const bool primed = (flags & DGB_VCYCLE_ENTRY_PRIMED) != 0;
    if (primed) {
        // only where the pre-smoother would itself open with dgb_block_gs_entry_residual
        const dgb_level &L = h_levels[nlevels - 1];
        if (!(nlevels >= 2 && L.smoother == DGB_SMOOTHER_BLOCK_GS_PYAMG && L.pre_iterations > 0 && h_opts->check_residual &&
              h_opts->gs_mode == DGB_GS_LEXICOGRAPHIC && dgb::gs_entry_fused(&L.op)))
            return DGB_UNSUPPORTED;
    }
    return dgb::vcycle(h_levels, nlevels - 1, *h_opts, ctl, partials, sumsq, stream, true, false, primed);
}
