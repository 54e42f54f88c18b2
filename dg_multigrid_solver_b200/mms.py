"""Manufactured solution: exact fields and source terms from the paramfile's expression strings.

Mirrors DGFEM.compute_exact_solution (dgfem/dgfem.py:410-483): the exact solution is a string
(paramfile.yml:79-81), differentiated symbolically; f = -div(nu grad u).  The reference calls
sympy.lambdify once per element and per boundary face; here each expression is compiled once
and evaluated over ALL points of a level in one call -- on the device (torch elementwise ops as
the lambdify namespace) for CUDA tensors, with NumPy for host arrays.
"""
import math

import numpy as np
import sympy as sym


def _torch_namespace():
    import torch
    ns = {k: getattr(torch, k) for k in ("sin", "cos", "tan", "exp", "log", "sqrt", "sinh", "cosh", "tanh",
                                          "asin", "acos", "atan", "atan2", "abs")}
    ns.update({"Abs": torch.abs, "pi": math.pi, "E": math.e, "arcsin": torch.asin, "arccos": torch.acos,
               "arctan": torch.atan, "arctan2": torch.atan2})
    return ns


class Field:
    """One scalar expression in (x, y), callable on numpy arrays or torch tensors."""
    CHUNK = 1 << 22           # entries per slice of a device evaluation

    def __init__(self, expr):
        self.expr = sym.sympify(expr)
        x, y = sym.symbols("x y")
        self._const = float(self.expr) if isinstance(self.expr, sym.Number) or not self.expr.free_symbols else None
        if self._const is None:
            self._np = sym.lambdify((x, y), self.expr, "numpy")
            self._th = None
            self._xy = (x, y)

    def __call__(self, X, Y):
        if isinstance(X, np.ndarray):
            if self._const is not None:
                return np.full_like(X, self._const, dtype=np.float64)
            return np.asarray(self._np(X, Y), dtype=np.float64)
        import torch
        if self._const is not None:
            return torch.full_like(X, self._const)
        if self._th is None:
            self._th = sym.lambdify(self._xy, self.expr, modules=[_torch_namespace()])
        # the generated expression makes a fresh temporary per operation: on a whole level (2048^2 elements x 16
        # points = 537 MB each) every one of them is a cudaMalloc.  Slices along the element axis keep the temporaries
        # in blocks the caching allocator hands back at once; same values, entry by entry
        per = max(1, X[0].numel()) if X.dim() > 0 else 1
        if X.dim() == 0 or X.numel() <= self.CHUNK:
            return self._th(X, Y)
        out = torch.empty(X.shape, dtype=torch.float64, device=X.device)
        step = max(1, self.CHUNK // per)
        for i0 in range(0, X.shape[0], step):
            out[i0:i0 + step] = self._th(X[i0:i0 + step], Y[i0:i0 + step])
        return out


class PoissonMMS:
    def __init__(self, settings):
        x, y = sym.symbols("x y")
        nu = settings.problem.kinematic_viscosity
        u = sym.sympify(settings.problem.exact_solution.u)
        self.u = Field(u)
        gx, gy = nu * sym.diff(u, x), nu * sym.diff(u, y)          # dgfem.py:461
        self.f = Field(-(sym.diff(gx, x) + sym.diff(gy, y)))      # dgfem.py:462

    def solution(self, X, Y):
        return self.u(X, Y)

    def source(self, X, Y):
        return self.f(X, Y)
