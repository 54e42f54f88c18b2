"""Distributive Gauss-Seidel for the Stokes system, `lsq` splitting (TEST INFRASTRUCTURE).

Restates  Relaxation.distributive_gauss_seidel(..., splitting='lsq')   dgfem/relaxation.py:221-283
on the global-order blocks of  Stokes.assemble_BSR_Stokes_global_order  dgfem/discrete_system.py:416-745:
    [A G; D 0] [u; p] = [f_mom; f_cont],   DG = D @ G  (relaxation.py:240)
per outer iteration (relaxation.py:241-268)
    RHS_mom  = f_mom - A u - G p
    du*      = one symmetric block-GS iteration on A  from 0          (block size = A.blocksize[0], App. B.6)
    RHS_cont = f_cont - D (u + du*)
    dp*      = one symmetric block-GS iteration on DG from 0
    du       = du* + G dp*
    dp       = one symmetric block-GS iteration on DG from 0 with right-hand side -D A G dp*
    u += du ; p += dp ; stop when ||RHS - K [u; p]|| / ||RHS - K [u0; p0]|| < 1e-6
The matrices are inputs (scipy BSR arrays as the reference built them); the inner smoother is the oracle's
restatement of Relaxation.block_gauss_seidel_pyamg, including its own residual tests.
"""
import numpy as np

from . import relax


def _bsr(M):
    """scipy BSR (square blocks) -> the oracle's container, stored order of the blocks kept."""
    assert M.blocksize[0] == M.blocksize[1]
    return relax.BSR(np.asarray(M.data), np.asarray(M.indices), np.asarray(M.indptr))


def distributive_gauss_seidel_lsq(A, D, G, RHS, u=None, max_iterations=1000000, DG=None, sweeps=1):
    """A, D, G (and optionally DG): scipy sparse BSR arrays.  Returns (u, normalised residual history)."""
    n_u = A.shape[0]
    u = np.zeros_like(RHS) if u is None else u.copy()
    DG = (D @ G) if DG is None else DG
    A_o, DG_o = _bsr(A), _bsr(DG)

    def full_residual(v):
        return np.concatenate([RHS[:n_u] - A @ v[:n_u] - G @ v[n_u:], RHS[n_u:] - D @ v[:n_u]])
    residual_0 = relax.lp_norm(full_residual(u), 2)
    history = []
    n = 0
    while n < max_iterations:
        u_k, p_k = u[:n_u], u[n_u:]
        f_mom, f_cont = RHS[:n_u], RHS[n_u:]
        RHS_mom = f_mom - A @ u_k - G @ p_k
        du_star = relax.block_gauss_seidel_pyamg(A_o, RHS_mom, np.zeros_like(u_k), "symmetric", 1, sweeps)
        RHS_cont = f_cont - D @ (u_k + du_star)
        dp_star = relax.block_gauss_seidel_pyamg(DG_o, RHS_cont, np.zeros_like(p_k), "symmetric", 1, sweeps)
        du = du_star + G @ dp_star
        RHS_DG = -(D @ (A @ (G @ dp_star)))        # (-D @ A @ G) @ dp*: sparse products first in the reference
        dp = relax.block_gauss_seidel_pyamg(DG_o, RHS_DG, np.zeros_like(p_k), "symmetric", 1, sweeps)
        u[:n_u] += du
        u[n_u:] += dp
        residual = relax.lp_norm(full_residual(u), 2) / residual_0
        history.append(residual)
        if residual < 1e-6:
            break
        if residual > 1e10:
            raise relax.SmootherDiverged(f"diverging, residual={residual:.6e}")
        n += 1
    return u, np.array(history)
