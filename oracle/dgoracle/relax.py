"""Smoothers (TEST INFRASTRUCTURE).

Restates
  Relaxation.block_gauss_seidel_pyamg   dgfem/relaxation.py:198-218
  pyamg_relaxation.block_gauss_seidel   dgfem/pyamg_relaxation.py:175-255  (wrapper, vendored in the reference)
  pyamg.util.utils.get_block_diag       (un-vendored, pyamg 5.0.1; pseudo-inverse of the diagonal blocks)
  pyamg.amg_core.block_gauss_seidel     (un-vendored C++; oracle/csrc/dgoracle.c)
  Relaxation.block_jacobi               dgfem/relaxation.py:123-150   (incl. the u = u_new aliasing, App. B.1)
  Relaxation.block_gauss_seidel         dgfem/relaxation.py:170-195
  utils.helpers.compute_Lp_norm         utils/helpers.py:16-26
"""
import numpy as np

from . import native


class SmootherDiverged(SystemExit):
    """relaxation.py:214-216 calls exit() -> SystemExit."""


def lp_norm(delta, p=2):
    return (np.sum(abs(delta) ** p) / delta.size) ** (1 / p)


class BSR:
    """Minimal BSR container (data[nnzb,b,b], indices, indptr) with scipy's matvec order."""

    def __init__(self, data, indices, indptr):
        self.data = np.ascontiguousarray(data, dtype=np.float64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        self.b = self.data.shape[1]
        self.N = len(self.indptr) - 1
        self._dinv = None

    def matvec(self, x):
        return native.bsr_matvec(self.indptr, self.indices, self.data, x, self.b)

    def __matmul__(self, x):
        return self.matvec(x)

    def block_diag(self):
        rows = np.repeat(np.arange(self.N), np.diff(self.indptr))
        sel = np.nonzero(self.indices == rows)[0]
        D = np.zeros((self.N, self.b, self.b))
        np.add.at(D, rows[sel], self.data[sel])
        return D

    def dinv(self):
        """get_block_diag(A, blocksize, inv_flag=True); the reference recomputes it on every
        smoother call (pyamg_relaxation.py:230-231) -- same values every time, cached here."""
        if self._dinv is None:
            self._dinv = np.ascontiguousarray(np.linalg.pinv(self.block_diag()))
        return self._dinv

    def to_scipy(self):
        import scipy.sparse as sp
        n = self.N * self.b
        return sp.bsr_array((self.data, self.indices, self.indptr), shape=(n, n))


def gs_pass(A, x, b, direction):
    """One directional pass of pyamg's block_gauss_seidel (pyamg_relaxation.py:240-255)."""
    Dinv = A.dinv()
    if direction == "forward":
        native.block_gauss_seidel(A.indptr, A.indices, A.data, x, b, Dinv, 0, A.N, 1, A.b)
    elif direction == "backward":
        native.block_gauss_seidel(A.indptr, A.indices, A.data, x, b, Dinv, A.N - 1, -1, -1, A.b)
    elif direction == "symmetric":
        gs_pass(A, x, b, "forward")
        gs_pass(A, x, b, "backward")
    else:
        raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')


def slab_gs_pass(A, x, b, direction, slabs):
    """One directional pass of the product's `slab_lexicographic` multi-GPU mode (NOT the reference's
    iteration): the block rows are split into `slabs` contiguous, equally sized slabs (one per GPU); every
    slab runs pyamg's lexicographic pass over its own rows, reading the rows of the other slabs as they were
    BEFORE the pass (the halo exchanged ahead of the pass) -- block-Jacobi coupling between slabs."""
    if direction == "symmetric":
        slab_gs_pass(A, x, b, "forward", slabs)
        slab_gs_pass(A, x, b, "backward", slabs)
        return
    if direction not in ("forward", "backward"):
        raise ValueError('valid sweep directions: "forward", "backward", and "symmetric"')
    assert A.N % slabs == 0
    rows = A.N // slabs
    Dinv = A.dinv()
    frozen = x.copy()
    for s in range(slabs):
        xs = frozen.copy()
        r0, r1 = s * rows, (s + 1) * rows
        if direction == "forward":
            native.block_gauss_seidel(A.indptr, A.indices, A.data, xs, b, Dinv, r0, r1, 1, A.b)
        else:
            native.block_gauss_seidel(A.indptr, A.indices, A.data, xs, b, Dinv, r1 - 1, r0 - 1, -1, A.b)
        x[r0 * A.b:r1 * A.b] = xs[r0 * A.b:r1 * A.b]


def block_gauss_seidel_pyamg(A, RHS, u=None, direction="symmetric", omega=1, max_iterations=1000,
                             info=None, slabs=1):
    """relaxation.py:198-218 (omega is accepted and ignored, as in the reference).  slabs > 1: the
    product's slab_lexicographic variant (slab_gs_pass), same residual tests."""
    u = np.zeros_like(RHS) if not isinstance(u, np.ndarray) else u.copy()
    residual_0 = lp_norm(RHS - A @ u, 2)
    n = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        while n < max_iterations:
            if slabs > 1:
                slab_gs_pass(A, u, RHS, direction, slabs)
            else:
                gs_pass(A, u, RHS, direction)
            residual = lp_norm(RHS - A @ u, 2) / residual_0
            if residual < 1e-6:
                if info is not None:
                    info["early_exit_after"] = n + 1
                break
            elif residual > 1e10:
                raise SmootherDiverged(f"diverging, residual={residual:.6e}")
            n += 1
    return u


def red_black_gauss_seidel(A, RHS, Ni, Nj, u=None, direction="symmetric", max_iterations=1, info=None):
    """Two-colour variant of block_gauss_seidel_pyamg (NOT in the reference; checks the product's
    multicolour mode).  forward = colour 0 then 1; backward = colour 1 then 0."""
    u = np.zeros_like(RHS) if not isinstance(u, np.ndarray) else u.copy()
    Dinv = A.dinv()

    def one(colour):
        native.block_gauss_seidel_colour(A.indptr, A.indices, A.data, u, RHS, Dinv, Ni, Nj, 2, colour, A.b)
    residual_0 = lp_norm(RHS - A @ u, 2)
    n = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        while n < max_iterations:
            if direction in ("forward", "symmetric"):
                one(0); one(1)
            if direction in ("backward", "symmetric"):
                one(1); one(0)
            residual = lp_norm(RHS - A @ u, 2) / residual_0
            if residual < 1e-6:
                if info is not None:
                    info["early_exit_after"] = n + 1
                break
            elif residual > 1e10:
                raise SmootherDiverged(f"diverging, residual={residual:.6e}")
            n += 1
    return u


def _row_solve_sweep(A, RHS, u_read, u_write, omega):
    """One pass over block rows of relaxation.py:131-148 / 177-194:
    u_write_i = omega * solve(D_i, E_i u_read + F_i u_read + RHS_i) + (1-omega) u_read_i
    with E, F the NEGATED strict lower/upper blocks (split_block_EDF, relaxation.py:444-492)."""
    b = A.b
    D = A.block_diag()
    for i in range(A.N):
        acc = np.zeros(b)
        lo, hi = A.indptr[i], A.indptr[i + 1]
        cols = A.indices[lo:hi]
        # data_E @ u[j_E] then data_F @ u[j_F]: lower columns first, then upper (relaxation.py:148)
        accE = np.zeros(b)
        accF = np.zeros(b)
        selE = [k for k in range(lo, hi) if A.indices[k] < i]
        selF = [k for k in range(lo, hi) if A.indices[k] > i]
        if selE:
            dE = -A.data[selE].transpose(1, 0, 2).reshape(b, -1)
            accE = dE @ np.concatenate([u_read[c * b:(c + 1) * b] for c in A.indices[selE]])
        if selF:
            dF = -A.data[selF].transpose(1, 0, 2).reshape(b, -1)
            accF = dF @ np.concatenate([u_read[c * b:(c + 1) * b] for c in A.indices[selF]])
        rhs = accE + accF + RHS[i * b:(i + 1) * b]
        u_write[i * b:(i + 1) * b] = omega * np.linalg.solve(D[i], rhs) + (1 - omega) * u_read[i * b:(i + 1) * b]
        del acc, cols


def block_jacobi(A, RHS, u=None, direction=None, omega=1, max_iterations=1000):
    """relaxation.py:123-150.  `u = u_new` aliases the buffers, so iteration 1 is block-Jacobi and
    iterations >= 2 are in-place forward block-GS.  direction is ignored."""
    u = np.zeros_like(RHS) if not isinstance(u, np.ndarray) else u.copy()
    u_new = np.zeros_like(u)
    for _ in range(int(max_iterations)):
        _row_solve_sweep(A, RHS, u, u_new, omega)
        u = u_new
    return u


def block_gauss_seidel(A, RHS, u=None, direction="forward", omega=1, max_iterations=1000):
    """relaxation.py:170-195 (forward only; direction ignored)."""
    u = np.zeros_like(RHS) if not isinstance(u, np.ndarray) else u.copy()
    for _ in range(int(max_iterations)):
        _row_solve_sweep(A, RHS, u, u, omega)
    return u
