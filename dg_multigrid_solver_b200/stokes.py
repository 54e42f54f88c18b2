"""Stokes (pressure-robust DG, local DOF ordering): assembly of the operator and right-hand side on the
device.  Same surface as the reference's `Stokes` problem class (dgfem/discrete_system.py:405-414,
812-1028): `DiscreteSystem(settings).problem.assemble(grid)` fills grid.BSR / grid.RHS.

The reference has no Stokes multigrid (README "future work"; settings.py:33-36), so this path ends at the
assembled system, `grid.BSR @ u` and the single-level block smoothers (b = 2 b_u + b_p blocks)."""
import numpy as np
import sympy as sym

from . import _lib
from .discrete_system import prepare_smoother_data, stencil_flags
from .grid import padded_blocks, upload_tables
from .mms import Field
from .tables import Tables


class StokesMMS:
    """dgfem/dgfem.py:410-483 for problem == 'Stokes' (momentum source = -div(nu grad u) + grad p)."""

    def __init__(self, settings):
        x, y = sym.symbols("x y")
        nu = settings.problem.kinematic_viscosity
        ex = settings.problem.exact_solution
        u, v, p = sym.sympify(ex.u), sym.sympify(ex.v), sym.sympify(ex.p)
        f_cont = sym.diff(u, x) + sym.diff(v, y)
        if settings.solution.manufactured_solution and not sym.simplify(f_cont).is_zero:
            raise ValueError("Manufactured solution is not divergence-free")          # dgfem.py:427-429
        lap = lambda w: -(sym.diff(nu * sym.diff(w, x), x) + sym.diff(nu * sym.diff(w, y), y))   # noqa: E731
        self.u, self.v = Field(u), Field(v)
        self.fx, self.fy = Field(lap(u) + sym.diff(p, x)), Field(lap(v) + sym.diff(p, y))   # dgfem.py:466-469
        self.fc = Field(f_cont)


class Stokes:
    def __init__(self, settings):
        self.settings = settings

    def assemble(self, grid):
        order = self.settings.solution.ordering.lower()
        if order != "local":
            raise NotImplementedError("only the local DOF ordering is accelerated (SURVEY.md section 8 a17); the "
                                      "global ordering feeds distributive_gauss_seidel, which is 'next' (8f-2)")
        self.assemble_BSR_Stokes_local_order(grid)
        self.assemble_RHS_Stokes(grid)

    def _setup(self, grid):
        """Tables and metrics at the velocity and at the pressure quadrature points."""
        torch = _lib.require_cuda()
        if getattr(grid, "_stokes", None) is not None:
            return grid._stokes
        s = self.settings
        pu, pp = grid.P_sol["u"], grid.P_sol["p"]
        fu = s.solution.u.integration_polynomial_degree_factor
        fp = s.solution.p.integration_polynomial_degree_factor
        n1u, n1p = fu * pu // 2 + 1, fp * pp // 2 + 1                                  # grid.py:107
        T = dict(uu=Tables(grid.P_grid, pu, nq1=n1u), pu=Tables(grid.P_grid, pp, nq1=n1u),
                 up=Tables(grid.P_grid, pu, nq1=n1p), pp=Tables(grid.P_grid, pp, nq1=n1p))
        H = {k: upload_tables(v) for k, v in T.items()}
        xn, yn = grid.geometry.device_nodes()
        N = grid.Ni * grid.Nj
        st = _lib.stream_ptr()
        geo = {}
        for key, tab in (("u", "uu"), ("p", "up")):
            nq1 = T[tab].nq1
            vol = torch.empty((N, 7, nq1 * nq1), dtype=torch.float64, device="cuda")
            face = torch.empty((N, 4, 8, nq1), dtype=torch.float64, device="cuda")
            area = torch.empty((N,), dtype=torch.float64, device="cuda")
            _lib.call("dgb_metrics", H[tab], xn, yn, grid.il, grid.Ni, grid.Nj, vol, face, area, st)
            geo[key] = (vol, face, area)
        grid._stokes = dict(T=T, H=H, geo=geo, bu=T["uu"].b, bp=T["pp"].b)
        return grid._stokes

    def assemble_BSR_Stokes_local_order(self, grid):
        torch = _lib.require_cuda()
        L = _lib.load()
        S = self._setup(grid)
        s = self.settings
        flags = stencil_flags(grid, s) & ~_lib.FLAG_MINV        # no inverse-mass scaling (discrete_system.py:941)
        if grid.fully_periodic_boundaries:
            raise NotImplementedError("the Stokes assembly of the reference has no fully periodic branch")
        bt = 2 * S["bu"] + S["bp"]
        N = grid.Ni * grid.Nj
        nnzb = int(L.dgb_poisson_nnzb(grid.Ni, grid.Nj, flags))
        grid.d_data = padded_blocks(nnzb, bt)
        grid.d_indices = torch.empty(nnzb, dtype=torch.int32, device="cuda")
        grid.d_indptr = torch.empty(N + 1, dtype=torch.int32, device="cuda")
        (vu, fu, area), (vp, fp, _) = S["geo"]["u"], S["geo"]["p"]
        H = S["H"]
        pin = 1 if s.get("solver.method") == "direct" else 0                     # discrete_system.py:946
        _lib.call("dgb_assemble_stokes", H["uu"], H["pu"], H["up"], H["pp"], vu, fu, vp, fp, area, grid.Ni, grid.Nj,
                  float(s.problem.kinematic_viscosity), float(grid.sigma), float(grid.gamma), flags, pin,
                  grid.d_indptr, grid.d_indices, grid.d_data, _lib.stream_ptr())
        grid.d_area = area
        grid.flags, grid.nnzb, grid._BSR = flags, nnzb, None
        grid.stencil = flags
        if not pin:
            # block smoothers need the inverse diagonal blocks; the pinned direct-solve matrix does not
            prepare_smoother_data(grid)

    def assemble_RHS_Stokes(self, grid):
        torch = _lib.require_cuda()
        S = self._setup(grid)
        s = self.settings
        if s.problem.include_pressure_BC:
            raise NotImplementedError("`include pressure BC: True` is not accelerated")
        mms = StokesMMS(s)
        (vu, fu, area), (vp, fp, _) = S["geo"]["u"], S["geo"]["p"]
        xu, yu = vu[:, 5, :], vu[:, 6, :]
        f_mom = torch.stack([mms.fx(xu, yu), mms.fy(xu, yu)], dim=1).contiguous()                 # [N,2,nqu]
        f_cont = mms.fc(vp[:, 5, :], vp[:, 6, :]).contiguous()
        g_u = torch.stack([mms.u(fu[:, :, 3, :], fu[:, :, 4, :]), mms.v(fu[:, :, 3, :], fu[:, :, 4, :])], dim=2).contiguous()
        g_p = torch.stack([mms.u(fp[:, :, 3, :], fp[:, :, 4, :]), mms.v(fp[:, :, 3, :], fp[:, :, 4, :])], dim=2).contiguous()
        bt = 2 * S["bu"] + S["bp"]
        grid.d_rhs = torch.empty(grid.Ni * grid.Nj * bt, dtype=torch.float64, device="cuda")
        H = S["H"]
        _lib.call("dgb_assemble_rhs_stokes", H["uu"], H["pu"], H["up"], H["pp"], vu, fu, vp, fp, area, f_mom, f_cont,
                  g_u, g_p, grid.Ni, grid.Nj, float(s.problem.kinematic_viscosity), float(grid.sigma),
                  float(grid.gamma), grid.flags, grid.d_rhs, _lib.stream_ptr())
        grid._RHS = None
