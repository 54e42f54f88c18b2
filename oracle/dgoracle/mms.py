"""Manufactured-solution source and Dirichlet data (TEST INFRASTRUCTURE).

Restates DGFEM.compute_exact_solution (dgfem/dgfem.py:410-483) for the Poisson problem,
evaluated once over all points instead of once per element."""
import numpy as np
import sympy as sym


class PoissonMMS:
    def __init__(self, expr_u, nu=1.0):
        x, y = sym.symbols("x y")
        self.u = sym.sympify(expr_u)
        grad_u = [nu * sym.diff(self.u, x), nu * sym.diff(self.u, y)]          # dgfem.py:461
        self.f = -(sym.diff(grad_u[0], x) + sym.diff(grad_u[1], y))             # dgfem.py:462
        self._u = self._lam(self.u, x, y)
        self._f = self._lam(self.f, x, y)

    @staticmethod
    def _lam(expr, x, y):
        if isinstance(expr, sym.Number):                                        # dgfem.py:438-440,477-479
            val = float(expr)
            return lambda X, Y: np.full_like(np.asarray(X, dtype=np.float64), val)
        return sym.lambdify((x, y), expr)

    def solution(self, X, Y):
        return np.asarray(self._u(X, Y), dtype=np.float64) + 0.0 * X

    def source(self, X, Y):
        return np.asarray(self._f(X, Y), dtype=np.float64) + 0.0 * X
