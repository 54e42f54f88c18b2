#!/usr/bin/env python3
"""Wall-clock timings of the UNMODIFIED reference (/root/reference) on this container's CPU: SURVEY.md section 8(d),
"CPU baseline, timed beside it", item (1) -- `python -m dgfem -m` end to end (DGFEM(...) = grids + assembly of every
level, then Solver.solve) on C1, C2 and synthetic Rectangle 8^2 .. 32^2 p=2 grids.

TEST / MEASUREMENT INFRASTRUCTURE: runs only where /root/reference exists (not on the GPU box); the product never
imports it.  Same recipe as oracle/gen_golden.py (import shims; `pyamg` restated in C, oracle/csrc/dgoracle.c).
The reference is single-threaded (SciPy bsr_matvec, pyamg's sweep, per-element Python loops).

Usage:  python oracle/time_reference.py > profiles/rNN_reference_cpu_timings.jsonl
        python oracle/time_reference.py --case rect32      (worker mode)
"""
import argparse
import copy
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"

# name -> (grid file or synthetic size, P_grid = p_u, O-grid, sigma multiplier, p levels, h factors)
CASES = {
    "c1": dict(grid="Rectangle_4X4_nPoly2.xyz", p=2, ogrid=False, sigmul=1.0, levels="2,1", factors="2"),
    "c2": dict(grid="CircleInCircle_8X8_nPoly5.xyz", p=5, ogrid=True, sigmul=2.0, levels="5,3,1", factors="2"),
    "rect8": dict(grid="Rectangle_8X8_nPoly2.xyz", p=2, ogrid=False, sigmul=1.0, levels="2,1", factors="2,4"),
    "rect16": dict(n=16, p=2, ogrid=False, sigmul=1.0, levels="2,1", factors="2,4"),
    "rect32": dict(n=32, p=2, ogrid=False, sigmul=1.0, levels="2,1", factors="2,4,8"),
}


def worker(name):
    import numpy as np
    case = CASES[name]
    w = tempfile.mkdtemp(prefix="dgref_time_")
    for d in ("input", "logs", "results", "cache/grid", "cache/discrete_system", "postprocessing/pickles/relaxation"):
        os.makedirs(os.path.join(w, d), exist_ok=True)
    shutil.copy(os.path.join(REF, "input", "paramfile.yml"), os.path.join(w, "input", "paramfile.yml"))
    if "grid" in case:
        fname = case["grid"]
        shutil.copy(os.path.join(REF, "input", fname), os.path.join(w, "input", fname))
    else:                                   # synthetic rectangle by the rule of the shipped grids
        sys.path.insert(0, REPO)
        import bench
        from dg_multigrid_solver_b200.visualization import write_plot3d
        fname = f"Rectangle_{case['n']}X{case['n']}_nPoly{case['p']}.xyz"
        xn, yn = bench.rectangle_nodes_file_order(case["n"], case["p"])
        write_plot3d(os.path.join(w, "input", fname), xn, yn)
    os.chdir(w)
    sys.path[:0] = [os.path.join(HERE, "shims"), REF]
    from input import params
    params = copy.deepcopy(params)
    params["grid"]["filename"] = fname
    params["grid"]["polynomial degree"] = case["p"]
    params["grid"]["O grid"] = case["ogrid"]
    params["grid"]["circular"] = case["ogrid"]
    params["solution"]["u"]["polynomial degree"] = case["p"]
    params["problem"]["SIP penalty parameter multiplier"] = case["sigmul"]
    params["visualization"]["automatically open paraview"] = False
    params["visualization"]["export"] = False
    params["logging"]["loglevel"] = "ERROR"
    params["solver"]["multigrid"]["polynomial coarsening"]["levels"]["u"] = case["levels"]
    params["solver"]["multigrid"]["geometric coarsening"]["coarsening factors"] = case["factors"]
    from dgfem.settings import Settings
    from dgfem.dgfem import DGFEM
    t0 = time.perf_counter()
    d = DGFEM(settings=Settings(params), solve_multigrid=True)      # grids, elements, faces, assembly of every level
    t_init = time.perf_counter() - t0
    fine = d.grids[-1]
    u = np.sin(0.37 * np.arange(fine.RHS.size))
    t0 = time.perf_counter()
    for _ in range(20):
        fine.BSR @ u
    t_apply = (time.perf_counter() - t0) / 20
    d.solver.residuals = []
    t0 = time.perf_counter()
    d.solve()
    t_solve = time.perf_counter() - t0
    cycles = len(d.solver.residuals) - 1
    nel = sum(g.Ni * g.Nj for g in d.grids)
    print(json.dumps({
        "case": name, "grid": fname, "p": case["p"], "levels": [(int(g.Ni), int(g.Nj), int(g.P_sol["u"])) for g in d.grids],
        "fine_elements": int(fine.Ni * fine.Nj), "fine_dofs": int(fine.RHS.size), "elements_all_levels": int(nel),
        "init_s (grids + assembly of all levels)": t_init, "assembly_elements_per_s": float(nel / t_init),
        "solve_s (residual tests + V-cycles + post-processing)": t_solve, "cycles": cycles,
        "s_per_cycle": t_solve / max(cycles, 1), "vcycles_per_s": max(cycles, 1) / t_solve,
        "apply_s": t_apply, "apply_dof_per_s": fine.RHS.size / t_apply,
        "final_normalised_residual": float(d.solver.residuals[-1]), "L2_error": float(d.L2_error_u)}), flush=True)
    shutil.rmtree(w, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    args = ap.parse_args()
    if args.case:
        return worker(args.case)
    import platform
    cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")]
    print(json.dumps({"host": "build container (no GPU)", "cpu": cpu[0] if cpu else platform.processor(),
                      "cores": os.cpu_count(), "threads_used": 1, "python": platform.python_version(),
                      "note": "unmodified reference from /root/reference under oracle/shims (pyamg's sweep restated in C)"}),
          flush=True)
    for name in CASES:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name], capture_output=True, text=True)
        if r.returncode != 0:
            print(json.dumps({"case": name, "error": r.stderr[-400:]}), flush=True)
        else:
            print(r.stdout.strip().splitlines()[-1], flush=True)


if __name__ == "__main__":
    main()
