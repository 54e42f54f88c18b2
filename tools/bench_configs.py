#!/usr/bin/env python3
"""Full-size runs of the BASELINE.json configurations that are not the bench.py workload:
  C4  CircleInCircle 1024x1024 nPoly5 (O-grid, sigma-mult 2), p=5 single level: assembly, operator apply,
      block-Jacobi sweep, one symmetric lexicographic block-GS iteration (36x36 blocks, 37.7 M DOFs)
  C5  Rectangle 1024x1024 nPoly2, Stokes local ordering (p_u=2, p_p=1): assembly + operator apply (22x22 blocks)
usage: bench_configs.py {c4|c5} [N]      -> one JSON line (device-timed with CUDA events)"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from dg_multigrid_solver_b200 import _lib  # noqa: E402
from dg_multigrid_solver_b200.dgfem import DGFEM  # noqa: E402
from dg_multigrid_solver_b200.grid import Geometry  # noqa: E402
from dg_multigrid_solver_b200.relaxation import Relaxation, bsr_apply  # noqa: E402
from dg_multigrid_solver_b200.settings import Settings  # noqa: E402
from dg_multigrid_solver_b200.tables import gauss_lobatto_nodes  # noqa: E402


def lgl_line(edges, P):
    xi = gauss_lobatto_nodes(P + 1)
    out = np.empty((len(edges) - 1) * P + 1)
    for e in range(len(edges) - 1):
        out[e * P:(e + 1) * P + 1] = edges[e] + (edges[e + 1] - edges[e]) * (xi + 1.0) / 2.0
    return out


def circle_nodes_file_order(n, P, r_in=0.1, r_out=1.0):
    """CircleInCircle_{n}X{n}_nPoly{P}: i = angle (clockwise), j = radius with element widths
    in geometric progression of ratio 10^(1/(n-1)); Plot3D file order [jl][il]."""
    q = 10.0 ** (1.0 / (n - 1))
    widths = (r_out - r_in) * (q - 1.0) / (q ** n - 1.0) * q ** np.arange(n)
    redges = r_in + np.concatenate([[0.0], np.cumsum(widths)])
    redges[-1] = r_out
    th = lgl_line(-2.0 * np.pi * np.arange(n + 1) / n, P)
    rr = lgl_line(redges, P)
    x = np.cos(th)[None, :] * rr[:, None]
    y = np.sin(th)[None, :] * rr[:, None]
    x[:, -1], y[:, -1] = x[:, 0], y[:, 0]          # close the O-grid exactly (grid.py:56-57)
    return np.ascontiguousarray(x), np.ascontiguousarray(y)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c) / reps


def main():
    which = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    peak, _ = bench.measured_peak()
    st = _lib.stream_ptr()
    if which == "c4":
        p = 5
        prm = bench.make_params(n, p, "lexicographic", True)
        prm["grid"].update({"O grid": True, "circular": True, "filename": f"synthetic_CircleInCircle_{n}X{n}_nPoly5.xyz"})
        prm["problem"]["SIP penalty parameter multiplier"] = 2.0
        s = Settings(prm)
        xn, yn = circle_nodes_file_order(n, p)
        t0 = time.perf_counter()
        d = DGFEM(settings=s, geometry=Geometry(None, s, nodes=(xn, yn)), solve_smoother=True,
                  smoother="block_gauss_seidel_pyamg", write_results=False)
        torch.cuda.synchronize()
        setup = time.perf_counter() - t0
        g = d.grids[-1]
        N, b, nnzb = g.Ni * g.Nj, g.b, int(g.d_indices.numel())
        ab = bench.algorithmic_bytes(nnzb, N, b)
        x = torch.randn(N * b, dtype=torch.float64, device="cuda")
        y = torch.empty_like(x)
        op = g.operator()
        t_apply = timed(lambda: _lib.call("dgb_bsr_apply", op, x, y, st))
        t_jac = timed(lambda: _lib.call("dgb_block_relax_sweep", op, g.d_rhs, x, y, 1.0, st))
        xg = torch.zeros_like(x)
        L = _lib.load()
        ctl = torch.zeros(32, dtype=torch.uint8, device="cuda")
        part = torch.zeros(L.dgb_partials_len(), dtype=torch.float64, device="cuda")
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")

        def smoother(iters):      # the smoother as Solver.solve_smoother calls it, without its residual tests
            _lib.call("dgb_block_gauss_seidel_pyamg", op, g.d_rhs, xg, 0, iters, 0, 0, ctl, part, ss, st)
        t1 = timed(lambda: smoother(1), reps=2)
        t3 = timed(lambda: smoother(3), reps=2)
        t_gs = (t3 - t1) / 2.0                       # one symmetric iteration inside a longer call
        t_gs_first = t1
        # the reference's `-s --smoother block_gauss_seidel_pyamg` run: 100 symmetric iterations with residual tests
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d.solver.solve()
        torch.cuda.synchronize()
        t_s100 = time.perf_counter() - t0
        out = {"config": f"C4 CircleInCircle {n}x{n} p=5 O-grid", "elements": N, "dofs": N * b, "b": b, "nnzb": nnzb,
               "smoother_run_100_iterations_s": t_s100,
               "operator_GB": nnzb * b * b * 8 / 1e9, "setup_s": setup, "assemble_s": d.timings.get("assemble"),
               "assembly_elements_per_s": N / d.timings["assemble"],
               "apply_ms": t_apply, "apply_GBs": ab["apply"] / t_apply / 1e6, "apply_frac": ab["apply"] / t_apply / 1e6 / peak,
               "apply_dof_per_s": N * b / (t_apply * 1e-3),
               "block_jacobi_sweep_ms": t_jac, "block_jacobi_GBs": ab["gs_pass"] / t_jac / 1e6,
               "block_jacobi_dof_per_s": N * b / (t_jac * 1e-3),
               "gs_first_symmetric_iteration_ms": t_gs_first,
               "gs_symmetric_iteration_ms": t_gs, "gs_pass_GBs": 2 * ab["gs_pass"] / t_gs / 1e6,
               "gs_sweep_dof_per_s": 2 * N * b / (t_gs * 1e-3), "device_error": _lib.load().dgb_device_error(1)}
    else:
        prm = bench.make_params(n, 2, "lexicographic", True)
        prm["problem"]["type"] = "Stokes"
        prm["problem"]["include pressure BC"] = False
        prm["solution"]["p"]["polynomial degree"] = 1
        prm["solution"]["ordering"] = "local"
        s = Settings(prm)
        xn, yn = bench.rectangle_nodes_file_order(n, 2)
        t0 = time.perf_counter()
        d = DGFEM(settings=s, geometry=Geometry(None, s, nodes=(xn, yn)), solve_direct=True, write_results=False)
        torch.cuda.synchronize()
        setup = time.perf_counter() - t0
        g = d.grids[-1]
        N, b, nnzb = g.Ni * g.Nj, int(g.d_data.shape[1]), int(g.d_indices.numel())
        ab = bench.algorithmic_bytes(nnzb, N, b)
        x = torch.randn(N * b, dtype=torch.float64, device="cuda")
        y = torch.empty_like(x)
        op = g.operator()
        t_apply = timed(lambda: _lib.call("dgb_bsr_apply", op, x, y, st))
        out = {"config": f"C5 Rectangle {n}x{n} Stokes p_u=2 p_p=1 local order", "elements": N, "dofs": N * b, "b": b,
               "nnzb": nnzb, "operator_GB": nnzb * b * b * 8 / 1e9, "setup_s": setup,
               "assemble_s": d.timings.get("assemble"), "assembly_elements_per_s": N / d.timings["assemble"],
               "apply_ms": t_apply, "apply_GBs": ab["apply"] / t_apply / 1e6, "apply_frac": ab["apply"] / t_apply / 1e6 / peak,
               "apply_dof_per_s": N * b / (t_apply * 1e-3)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
