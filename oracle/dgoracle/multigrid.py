"""Multigrid hierarchy, transfer operators and V-cycle (TEST INFRASTRUCTURE).

Restates
  DGFEM.assemble_multigrid_operators   dgfem/dgfem.py:269-376   (p- and h-levels, R/P operators)
  DGFEM.initialize (assembly per level) dgfem/dgfem.py:103-151
  Solver.solve_multigrid               dgfem/solver.py:114-139
  Solver.multigrid_V_cycle             dgfem/solver.py:141-207
"""
import numpy as np

from . import assemble, geometry, relax
from .mms import PoissonMMS
from .tables import LevelTables


def p_restriction(p_coarse, p_fine):
    """dgfem.py:306-317: zero-padded identity selecting modes (i<=p_c, j<=p_c); P = R^T."""
    N_fine, N_coarse = (p_fine + 1) ** 2, (p_coarse + 1) ** 2
    R = np.eye(N_coarse)
    for i in range(p_coarse):
        R = np.insert(R, (i + 1) * (p_coarse + 1) + i * (p_fine - p_coarse),
                      np.zeros((p_fine - p_coarse, N_coarse)), axis=1)
    R = np.append(R, np.zeros((N_coarse, N_fine - N_coarse - (p_fine - p_coarse) * p_coarse)), axis=1)
    return R


def h_restriction():
    """dgfem.py:362-367: L2 projection of the 2x2 children onto the parent, p=1; P = 4 R^T."""
    s3 = np.sqrt(3)
    R = np.array([
        np.array([1., 0., 0., 0., 1., 0., 0., 0., 1., 0., 0., 0., 1., 0., 0., 0.]) / 4.,
        np.array([-s3, 1., 0., 0., s3, 1., 0., 0., -s3, 1., 0., 0., s3, 1., 0., 0.]) / 8.,
        np.array([-s3, 0., 1., 0., -s3, 0., 1., 0., s3, 0., 1., 0., s3, 0., 1., 0.]) / 8.,
        np.array([3., -s3, -s3, 1., -3., -s3, s3, 1., -3., s3, -s3, 1., 3., s3, s3, 1.]) / 16.])
    return R, R.T * 4.


class Level:
    def __init__(self, Ni, Nj, p, sigma, T, G, A, rhs, cf=None):
        self.Ni, self.Nj, self.p, self.sigma, self.T, self.G, self.A, self.RHS, self.cf = \
            Ni, Nj, p, sigma, T, G, A, rhs, cf
        self.b = (p + 1) ** 2


class Hierarchy:
    """grids = [h-coarse (cf descending) ..., p_min, ..., p_max]   (dgfem.py:269-376)."""

    def __init__(self, x, y, Pg, p_levels, h_factors, sigma_mult=1.0, nu=1.0, O_grid=False,
                 fully_periodic=False, exact_u="-2*sin(pi*x)**2*sin(pi*y)*cos(pi*y)", factor=3,
                 multiply_inverse_mass=True, fast=True, rhs_all_levels=True):
        self.x, self.y, self.Pg = x, y, Pg
        Ni, Nj = (x.shape[0] - 1) // Pg, (x.shape[1] - 1) // Pg
        self.mms = PoissonMMS(exact_u, nu)
        self.levels, self.R, self.P, self.types = [], [], [], []
        p_levels = sorted(p_levels)
        h_factors = sorted(h_factors, reverse=True)
        asm = assemble.assemble_bsr_fast if fast else assemble.assemble_bsr

        def build(G, T, sigma, want_rhs):
            data, indices, indptr, Minv = asm(G, T, nu, sigma, O_grid, fully_periodic, multiply_inverse_mass)
            rhs = None
            if want_rhs:
                f = self.mms.source(G.vol["x"], G.vol["y"])
                gf = {k: self.mms.solution(G.face[k]["x"], G.face[k]["y"]) for k in geometry.FACES}
                Minv_ij = Minv if Minv.ndim == 4 else Minv.reshape(G.Nj, G.Ni, T.b, T.b).transpose(1, 0, 2, 3)
                rhs = assemble.assemble_rhs(G, T, nu, sigma, f, gf, Minv_ij, O_grid, fully_periodic,
                                            multiply_inverse_mass)
            return relax.BSR(data, indices, indptr), rhs

        T_min = None
        for k, p in enumerate(p_levels):
            T = LevelTables(Pg, p, factor)
            sigma = (p + 1) ** 2 * sigma_mult                      # dgfem.py:298
            G = geometry.fine_geometry(x, y, Ni, Nj, T)
            A, rhs = build(G, T, sigma, rhs_all_levels or k == len(p_levels) - 1)
            self.levels.append(Level(Ni, Nj, p, sigma, T, G, A, rhs))
            if k == 0:
                T_min, sigma_min = T, sigma
        for k in range(len(p_levels) - 1):
            self.R.append(p_restriction(p_levels[k], p_levels[k + 1]))
            self.P.append(self.R[-1].T)
            self.types.append("polynomial")
        if h_factors:
            if p_levels[0] != 1:
                raise ValueError("h-coarsening requires p=1 at the bottom p-level (SURVEY App. A.7)")
            hl = []
            for cf in h_factors:
                G = geometry.coarse_geometry(x, y, Ni, Nj, T_min, cf)
                A, rhs = build(G, T_min, sigma_min, rhs_all_levels)
                hl.append(Level(G.Ni, G.Nj, p_levels[0], sigma_min, T_min, G, A, rhs, cf=cf))
            self.levels[0:0] = hl
            Rh, Ph = h_restriction()
            self.R[0:0] = [Rh for _ in h_factors]
            self.P[0:0] = [Ph for _ in h_factors]
            self.types[0:0] = ["geometric" for _ in h_factors]


class Schedule:
    """The smoother settings of paramfile.yml:20-65 (defaults = shipped values).  The post smoother is
    resolved independently of the pre smoother (dgfem/solver.py:143-147,196): post_smoother / post_direction /
    post_omega default to the pre smoother's values.
    gs_mode: 'lexicographic' (the reference), 'redblack', or 'slab_lexicographic' with `world` slabs -- the
    product's multi-GPU variants (dg_multigrid_solver_b200/parallel.py); levels with fewer than `min_rows`
    element rows per slab are gathered to one GPU and use the lexicographic order."""

    def __init__(self, smoother="block_gauss_seidel_pyamg", direction="symmetric", pre=2, post=1,
                 coarse_iterations=10, omega=1.0, coarse_solver="smoother", gs_mode="lexicographic",
                 post_smoother=None, post_direction=None, post_omega=None, world=1, min_rows=8):
        self.smoother, self.direction, self.pre, self.post = smoother, direction, pre, post
        self.coarse_iterations, self.omega, self.coarse_solver = coarse_iterations, omega, coarse_solver
        self.gs_mode = gs_mode
        self.post_smoother = smoother if post_smoother is None else post_smoother
        self.post_direction = direction if post_direction is None else post_direction
        self.post_omega = omega if post_omega is None else post_omega
        self.world, self.min_rows = world, min_rows


def slabs_of_level(H, k, sched):
    """Number of slabs level index k (0 = coarsest) is smoothed in under `sched` (1 = gathered / one GPU):
    restates parallel.distributed_levels -- p-levels are always distributed; an h-level (coarsening factor
    cf) is distributed iff Nj % (world*cf) == 0, it keeps >= min_rows rows per rank, and every finer h-level
    is distributed too."""
    if sched.gs_mode != "slab_lexicographic" or sched.world <= 1:
        return 1
    Nj = H.levels[-1].Nj
    lev = H.levels[k]
    if lev.cf is None:
        return sched.world
    for cf in sorted(L.cf for L in H.levels if L.cf is not None):
        ok = Nj % (sched.world * cf) == 0 and Nj // sched.world // cf >= sched.min_rows
        if not ok:
            return 1
        if cf == lev.cf:
            return sched.world
    return 1


def _smooth(level, sched, RHS, u, iterations, post=False, slabs=1):
    smoother = sched.post_smoother if post else sched.smoother
    direction = sched.post_direction if post else sched.direction
    omega = sched.post_omega if post else sched.omega
    if smoother == "block_gauss_seidel_pyamg":
        if sched.gs_mode == "redblack":
            return relax.red_black_gauss_seidel(level.A, RHS, level.Ni, level.Nj, u, direction, iterations)
        return relax.block_gauss_seidel_pyamg(level.A, RHS, u, direction, omega, iterations, slabs=slabs)
    if smoother == "block_jacobi":
        return relax.block_jacobi(level.A, RHS, u, direction, omega, iterations)
    if smoother == "block_gauss_seidel":
        return relax.block_gauss_seidel(level.A, RHS, u, direction, omega, iterations)
    raise AttributeError(smoother)


def restrict(H, k, residual):
    """solver.py:152-168 for operator index k-2 (fine level index k-1, coarse k-2)."""
    R = H.R[k - 2]
    if H.types[k - 2] == "geometric":
        c = H.levels[k - 2]
        residual = residual.reshape((c.Ni, 2, c.Nj, 2, c.b)).transpose((0, 2, 1, 3, 4))
    residual = residual.reshape((-1, R.shape[1]))
    return np.ravel(np.einsum("ij,kj->ki", R, residual))


def prolong(H, k, u_coarse):
    """solver.py:174-190."""
    P = H.P[k - 2]
    v = np.einsum("ij,kj->ki", P, u_coarse.reshape((-1, P.shape[1])))
    if H.types[k - 2] == "geometric":
        c = H.levels[k - 2]
        v = v.reshape((c.Ni, c.Nj, 2, 2, c.b)).transpose((0, 2, 1, 3, 4))
    return np.ravel(v)


def v_cycle(H, sched, k, RHS, u):
    """solver.py:141-207."""
    lev = H.levels[k - 1]
    slabs = slabs_of_level(H, k - 1, sched)
    if k > 1:
        u = _smooth(lev, sched, RHS, u, sched.pre, slabs=slabs)
        residual = RHS - lev.A @ u
        RHS_c = restrict(H, k, residual)
        u_c = v_cycle(H, sched, k - 1, RHS_c, np.zeros_like(RHS_c))
        u = u + prolong(H, k, u_c)                 # the reference does u += ... on the smoother's fresh copy
        u = _smooth(lev, sched, RHS, u, sched.post, post=True, slabs=slabs)
    else:
        if sched.coarse_solver == "direct":
            import scipy.sparse.linalg as splin
            u = splin.spsolve(lev.A.to_scipy().tocsr(), RHS)
        else:
            u = _smooth(lev, sched, RHS, u, sched.coarse_iterations, slabs=slabs)
    return u


def solve_multigrid(H, sched, tol=1e-6, max_cycles=1000, u=None, on_cycle=None):
    """solver.py:114-126.  Returns (u, residual history)."""
    fine = H.levels[-1]
    RHS = fine.RHS
    u = np.zeros_like(RHS) if u is None else u
    history = []
    n = 0
    residual_0 = relax.lp_norm(RHS - fine.A @ u, 2)
    while n < max_cycles:
        residual = relax.lp_norm(RHS - fine.A @ u, 2) / residual_0
        history.append(residual)
        if residual < tol or np.isnan(residual) or np.isinf(residual):
            break
        u = v_cycle(H, sched, len(H.levels), RHS, u)
        n += 1
        if on_cycle is not None:
            on_cycle(n, u)
    return u, np.array(history)
