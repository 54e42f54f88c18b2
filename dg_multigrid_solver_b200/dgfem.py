"""DGFEM: the paramfile-driven driver (settings -> geometry -> levels -> assembly -> solve ->
error norms), same flow and observable outputs as the reference's dgfem/dgfem.py:19-266, with
the level hierarchy of DGFEM.assemble_multigrid_operators (dgfem/dgfem.py:269-376):

    grids = [h-coarse (coarsening factor descending) ..., p_min, ..., p_max]

Every level lives on the device (grid.py); assembly, smoothing and the V-cycle run in
libdgb200.so.  Out of scope here: sigma ("penalty parameter")
coarsening, FVM coarse levels, Gram-Schmidt bases, VTK/plots.
"""
import os

import numpy as np

from . import _lib
from .discrete_system import DiscreteSystem
from .grid import CoarseGrid, Geometry, Grid
from .mms import PoissonMMS
from .settings import Settings, load_params
from .solver import Solver, compute_Lp_norm
from .tables import h_restriction, p_restriction
from .timer import Timer


def _int_list(v):
    return [int(v)] if isinstance(v, int) else [int(t) for t in str(v).split(",")]


class DGFEM:
    def __init__(self, **kwargs):
        self.settings = kwargs.get("settings") or Settings(load_params())
        self.settings.update_settings(kwargs)
        for key, arg in kwargs.items():
            if key.startswith("solve_") and arg:
                self.settings.update_setting("solver.method", key[len("solve_"):])     # dgfem.py:34-36
        self.solver = Solver(self.settings.solver.method, self.settings)
        self.timings = {}

        geometry = kwargs.get("geometry")
        if geometry is None:
            path = os.path.join(os.getcwd(), self.settings.grid.folder, self.settings.grid.filename)
            geometry = Geometry(path, self.settings)
        self.geometry = geometry

        if self.settings.problem.type == "Poisson":
            self.vars = ["u"]
            self.P_sol = {"u": self.settings.solution.u.polynomial_degree}
            self.exact_sol = {"u": self.settings.problem.exact_solution.u}
        elif self.settings.problem.type == "Stokes":
            self.vars = ["u", "p"]
            self.P_sol = {v: getattr(getattr(self.settings.solution, v), "polynomial_degree") for v in self.vars}
            self.exact_sol = {v: getattr(self.settings.problem.exact_solution, v) for v in ("u", "v", "p")}
        else:
            raise NotImplementedError(f"There exists no implementation for the {self.settings.problem.type} "
                                      "equation(s), possible equation(s) are: Poisson|Stokes")
        self.settings._validate_settings(self.settings)

        name = os.path.basename(str(geometry.filepath or "synthetic.xyz"))
        self.grid_filename = name[:name.rfind(".xyz")] if ".xyz" in name else name
        self.write_results = kwargs.get("write_results", True)
        if self.write_results:
            folder = f"exact_sol_{self.settings.problem.exact_solution.tag}" + \
                f"_sigmamul{self.settings.problem.SIP_penalty_parameter_multiplier}".replace(".", "_")
            self.results_dir = os.path.join("results", self.settings.problem.type, f"grid_{self.grid_filename}", folder)
            os.makedirs(self.results_dir, exist_ok=True)
            self.solution_summary_filepath = os.path.join(self.results_dir, "summary.txt")
        with Timer() as t:
            self.initialize()
        self.timings["initialize"] = t.elapsed()
        if self.write_results:
            with open(self.solution_summary_filepath, "w") as f:          # dgfem.py:85-101
                f.write("############################################\n###          SIMULATION SUMMARY          ###\n")
                f.write("############################################\n\n")
                f.write(f"### grid={self.grid_filename}\n### exact solution={self.exact_sol}\n")
                f.write(f"### Ni={self.geometry.Ni}, Nj={self.geometry.Nj}\n")
                f.write(f"### P grid={self.settings.grid.polynomial_degree}\n### P sol={self.P_sol}\n")
                f.write(f"### epsilon multiplier={self.settings.problem.SIP_penalty_parameter_multiplier}\n###\n")
                f.write(f"### solver={'multigrid' if self.settings.solver.method == 'multigrid' else 'direct'}\n\n")
                f.write("############################################\n\n")

    # ------------------------------------------------------------------------------------
    def initialize(self):
        """dgfem.py:103-151."""
        s = self.settings
        self.sigma = s.problem.SIP_penalty_parameter if s.problem.SIP_penalty_parameter else \
            (self.P_sol["u"] + 1) ** 2 * s.problem.SIP_penalty_parameter_multiplier
        self.grids = []
        if s.solver.method == "multigrid":
            self.assemble_multigrid_operators()
        else:
            self.grids.append(Grid(self.geometry, self.vars, s.solver.discretization).initialize(self.P_sol, self.sigma))
        discrete_system = DiscreteSystem(s)
        from .discrete_system import SETUP_TIMINGS
        SETUP_TIMINGS["smoother_setup"] = 0.0
        with Timer() as t:
            if s.solver.method == "multigrid":
                for grid in self.grids:                      # dgfem.py:121-123 (every level, RHS included)
                    discrete_system.problem.assemble(grid)
            else:
                discrete_system.problem.assemble(self.grids[-1])
            _lib.require_cuda().cuda.synchronize()
        # `assemble` = what the reference's assemble() does (metrics, operator, right-hand side); the inverse
        # diagonal blocks and smoother records are this implementation's own setup phase
        self.timings["smoother_setup"] = SETUP_TIMINGS["smoother_setup"]
        self.timings["assemble"] = t.elapsed() - self.timings["smoother_setup"]
        self.solver.grids = self.grids

    def assemble_multigrid_operators(self):
        """dgfem.py:269-376 for polynomial + geometric coarsening."""
        s = self.settings
        mg = s.solver.multigrid
        R_ops, P_ops, types = [], [], []
        if mg.penalty_parameter_coarsening.enabled:
            raise NotImplementedError("penalty-parameter coarsening is out of scope (SURVEY.md section 2.1 row 20)")
        if mg.polynomial_coarsening.enabled:
            p_levels = sorted(_int_list(mg.polynomial_coarsening.levels.u))          # dgfem.py:291
            mult = s.problem.SIP_penalty_parameter_multiplier
            for p in p_levels:                                                       # dgfem.py:298-301
                self.grids.append(Grid(self.geometry, self.vars).initialize({"u": p}, (p + 1) ** 2 * mult))
            for k in range(len(p_levels) - 1):                                       # dgfem.py:306-317
                R = p_restriction(p_levels[k], p_levels[k + 1])
                R_ops.append(R)
                P_ops.append(R.T.copy())
                types.append("polynomial")
        if mg.geometric_coarsening.enabled:
            if not self.grids:
                self.grids.append(Grid(self.geometry, self.vars).initialize(self.P_sol, self.sigma))
            if mg.geometric_coarsening.use_FVM:
                raise NotImplementedError("FVM coarse levels are out of scope (SURVEY.md section 2.1 row 19)")
            factors = sorted(_int_list(mg.geometric_coarsening.coarsening_factors), reverse=True)   # dgfem.py:337
            base = self.grids[0]
            if base.P_sol["u"] != 1:
                raise ValueError("geometric coarsening needs p=1 on the lowest polynomial level (4x16 transfer, dgfem.py:362)")
            coarse = [CoarseGrid(self.geometry, base, self.vars).initialize(coarsening_factor=cf) for cf in factors]
            self.grids[0:0] = coarse                                                 # dgfem.py:361
            Rh, Ph = h_restriction()
            R_ops[0:0] = [Rh for _ in factors]
            P_ops[0:0] = [Ph for _ in factors]
            types[0:0] = ["geometric" for _ in factors]
        self.solver.restriction_operators = R_ops
        self.solver.prolongation_operators = P_ops
        self.solver.multigrid_type = types

    # ------------------------------------------------------------------------------------
    def solve(self):
        """dgfem.py:153-266 (residual report, modal -> nodal, L1/L2 error norms, summary.txt)."""
        torch = _lib.require_cuda()
        g = self.grids[-1]
        u_modal = self.solver.solve()
        self.u_modal = u_modal
        with Timer() as t:
            d_u = torch.from_numpy(np.ascontiguousarray(u_modal)).cuda()
            from .relaxation import residual_norm
            n = g.d_rhs.numel()
            if getattr(g, "ordering", "local") == "global":       # Stokes, global ordering: the regrouped matrix
                r = g.d_rhs - g.BSR_global.apply(d_u)
                self.residual = float(torch.sqrt((r * r).sum() / n).item())
            else:
                sumsq, _ = residual_norm(g, g.d_rhs, d_u)
                self.residual = float(np.sqrt(sumsq.item() / n))
            residual_0 = compute_Lp_norm(g.RHS, 2)
            self.residual_normalized = self.residual / residual_0
            if self.settings.problem.type == "Poisson":
                self.u_nodal, self.L1_error_u, self.L2_error_u = self._nodal_error(g, d_u)
        self.timings["postprocess"] = t.elapsed()
        if self.write_results and self.settings.problem.type == "Poisson":
            if self.settings.get("visualization.export", False) and self.u_nodal is not None:   # dgfem.py:236-247
                from .visualization import modal_to_vtk
                self.solution_visualization_filepath = os.path.join(self.results_dir, "solution")
                modal_to_vtk(self.solution_visualization_filepath, self, self.u_nodal, self.u_exact_nodes)
            with open(self.solution_summary_filepath, "a") as f:
                f.write(f"Residual={self.residual}\nL1 error={self.L1_error_u}\nL2 error={self.L2_error_u}\n")
        return u_modal

    def _nodal_error(self, g, d_u, want_nodal=True):
        """dgfem.py:188-232 in one kernel (dgb_nodal_error): u at the geometry nodes of every element
        (V_DOF_grid @ u_e, [N, (Pg+1)^2], node index a + N1*c), minus the exact solution there (evaluated once per
        grid node by the MMS expression), and the L1 / L2 error norms.  Returns (u_nodal or None, L1, L2)."""
        torch = _lib.require_cuda()
        T = g.tables
        Vg = torch.from_numpy(np.ascontiguousarray(T.V_DOF_grid, dtype=np.float64)).cuda()      # [ng, b]
        ng = int(Vg.shape[0])
        xn, yn = self.geometry.device_nodes()                                                    # [jl, il]
        exact = PoissonMMS(self.settings).solution(xn, yn).contiguous()                         # dgfem.py:114
        N = g.Ni * g.Nj
        u_nodal = torch.empty((N, ng), dtype=torch.float64, device="cuda") if want_nodal else None
        partials = torch.zeros(_lib.load().dgb_partials_len(), dtype=torch.float64, device="cuda")
        sums = torch.zeros(2, dtype=torch.float64, device="cuda")
        _lib.call("dgb_nodal_error", Vg, ng, int(T.b), int(g.P_grid), int(g.Ni), int(g.Nj), int(g.il),
                  d_u.contiguous(), exact, u_nodal, partials, sums, _lib.stream_ptr())
        s1, s2 = (float(v) for v in sums.cpu())
        self.u_exact_nodes = exact
        return u_nodal, s1 / (N * ng), float(np.sqrt(s2 / (N * ng)))                             # dgfem.py:220-221
