"""Level-wide basis / quadrature tables (TEST INFRASTRUCTURE, see package docstring).

Restates dgfem/interpolation.py and Grid.initialize_interpolation (dgfem/grid.py:178-213).
Conventions: mode n = j_s*(p+1) + i_r, point index = i_r + len(r)*i_s
(r fastest), orthonormal Legendre basis.
"""
from math import factorial

import numpy as np
from scipy.special import eval_jacobi, gamma, roots_jacobi


def jacobi_orthonormal(x, alpha, beta, P):
    """interpolation.py:29-44 (Jacobi.evaluate_polynomial)."""
    x = np.asarray(x, dtype=np.float64)
    norm = 2 ** (alpha + beta + 1) * gamma(P + alpha + 1) * gamma(P + beta + 1) / (
        (2 * P + alpha + beta + 1) * gamma(P + alpha + beta + 1) * factorial(P))
    return eval_jacobi(P, alpha, beta, x) / np.sqrt(norm)


def legendre(x, P):
    """interpolation.py:46-50."""
    return jacobi_orthonormal(x, 0, 0, P)


def grad_legendre(x, P):
    """interpolation.py:52-59 (alpha=beta=0)."""
    x = np.asarray(x, dtype=np.float64)
    if P == 0:
        return np.zeros_like(x)
    return np.sqrt(P * (P + 1)) * jacobi_orthonormal(x, 1, 1, P - 1)


def gauss_legendre(N):
    """interpolation.py:75-86."""
    return roots_jacobi(N, 0, 0)


def lgl(N):
    """interpolation.py:88-110 (Legendre-Gauss-Lobatto nodes, N = P+1 points)."""
    P = N - 1
    if P < 1:
        raise ValueError("The polynomial order P must be a positive integer")
    xi = np.zeros(P + 1)
    xi[0], xi[-1] = -1, 1
    if P > 1:
        xi[1:-1], _ = roots_jacobi(P - 1, 1, 1)
    return xi


def vandermonde2D(N, r, s):
    """interpolation.py:118-142; returns [len(r)*len(s), N*N]."""
    r = np.asarray(r, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    V = np.zeros((len(r) * len(s), N * N))
    n = 0
    for j in range(N):
        for i in range(N):
            V[:, n] = np.ravel(np.outer(legendre(r, i), legendre(s, j)), order="F")
            n += 1
    return V


def grad_vandermonde2D(N, r, s):
    """interpolation.py:150-170."""
    r = np.asarray(r, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    Vr = np.zeros((len(r) * len(s), N * N))
    Vs = np.zeros((len(r) * len(s), N * N))
    n = 0
    for j in range(N):
        for i in range(N):
            Vr[:, n] = np.ravel(np.outer(grad_legendre(r, i), legendre(s, j)), order="F")
            Vs[:, n] = np.ravel(np.outer(legendre(r, i), grad_legendre(s, j)), order="F")
            n += 1
    return Vr, Vs


def n_int(p, factor=3):
    """grid.py:107."""
    return factor * p // 2 + 1


class LevelTables:
    """grid.py:178-213 for one variable ('u'); Pg = geometry degree, p = solution degree."""

    def __init__(self, Pg, p, factor=3, N_int=None):
        self.Pg, self.p = Pg, p
        self.N_grid = Pg + 1
        self.N_sol = p + 1
        self.b = self.N_sol ** 2
        self.N_int = n_int(p, factor) if N_int is None else N_int
        self.r_grid = lgl(self.N_grid)
        self.r_int, self.w_int = gauss_legendre(self.N_int)
        self.w_int_2D = np.outer(self.w_int, self.w_int)
        Ng, Ns, ri = self.N_grid, self.N_sol, self.r_int
        self.V_grid_grid = vandermonde2D(Ng, self.r_grid, self.r_grid)
        self.V_grid_int = vandermonde2D(Ng, ri, ri)
        self.Vr_grid_int, self.Vs_grid_int = grad_vandermonde2D(Ng, ri, ri)
        self.Vr_grid_face, self.Vs_grid_face = {}, {}
        self.Vr_grid_face["imin"], self.Vs_grid_face["imin"] = grad_vandermonde2D(Ng, [-1], ri)
        self.Vr_grid_face["imax"], self.Vs_grid_face["imax"] = grad_vandermonde2D(Ng, [1], ri)
        self.Vr_grid_face["jmin"], self.Vs_grid_face["jmin"] = grad_vandermonde2D(Ng, ri, [-1])
        self.Vr_grid_face["jmax"], self.Vs_grid_face["jmax"] = grad_vandermonde2D(Ng, ri, [1])
        self.V_DOF_int = vandermonde2D(Ns, ri, ri)
        self.Vr_DOF_int, self.Vs_DOF_int = grad_vandermonde2D(Ns, ri, ri)
        # traces: L side evaluated at +1, R side at -1 (grid.py:203-210)
        self.V_face, self.Vr_face, self.Vs_face = {}, {}, {}
        for name, (rr, ss) in {"iL": ([1], ri), "iR": ([-1], ri), "jL": (ri, [1]), "jR": (ri, [-1])}.items():
            self.V_face[name] = vandermonde2D(Ns, rr, ss)
            self.Vr_face[name], self.Vs_face[name] = grad_vandermonde2D(Ns, rr, ss)
        self.V_DOF_grid = vandermonde2D(Ns, self.r_grid, self.r_grid)
        # geometry operators (element.py:76-77,122): nodes (F-order) -> values at points
        Vgg = self.V_grid_grid
        self.L_gg = (np.linalg.inv(Vgg.T) @ Vgg.T).T            # metric_xy_rs at the grid nodes themselves
        self.L_int = (np.linalg.inv(Vgg.T) @ self.V_grid_int.T).T
        self.Dr_int = (np.linalg.inv(Vgg).T @ self.Vr_grid_int.T).T
        self.Ds_int = (np.linalg.inv(Vgg).T @ self.Vs_grid_int.T).T
        self.Dr_face = {f: (np.linalg.inv(Vgg).T @ self.Vr_grid_face[f].T).T for f in self.Vr_grid_face}
        self.Ds_face = {f: (np.linalg.inv(Vgg).T @ self.Vs_grid_face[f].T).T for f in self.Vs_grid_face}
        self.L_face = {}
        for f, (rr, ss) in {"imin": ([-1], ri), "imax": ([1], ri), "jmin": (ri, [-1]), "jmax": (ri, [1])}.items():
            self.L_face[f] = (np.linalg.inv(Vgg.T) @ vandermonde2D(Ng, rr, ss).T).T

    def point_ops(self, r, s):
        """Geometry operators at one arbitrary reference point (used by the coarse-element
        sampling, element.py:292-293): returns (L, Dr, Ds) rows of shape [1, N_grid^2]."""
        Vgg = self.V_grid_grid
        L = (np.linalg.inv(Vgg.T) @ vandermonde2D(self.N_grid, [r], [s]).T).T
        Vr, Vs = grad_vandermonde2D(self.N_grid, [r], [s])
        Dr = (np.linalg.inv(Vgg).T @ Vr.T).T
        Ds = (np.linalg.inv(Vgg).T @ Vs.T).T
        return L, Dr, Ds
