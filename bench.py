#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout.

Workload (config.workload): BASELINE.json configs[2] -- synthetic Rectangle 2048x2048 Plot3D grid
(byte-identical generator rule to the shipped Rectangle_*_nPoly2 files), p=2 Poisson MMS,
multigrid levels p = 2,1 + geometric coarsening factors 2..512 (coarsest 4x4), smoother schedule of
the shipped paramfile (symmetric block-GS, 2 pre / 1 post, 10 sweeps on the coarsest level).
A "step" is one multigrid V-cycle on the finest level (Solver.multigrid_V_cycle).

  value      V-cycles/s, device-timed (CUDA events), operator + vectors resident in HBM
  e2e        the same cycle through Solver.multigrid_V_cycle with HOST (pinned) RHS/u in and u out:
             host->device and device->host copies inside the timed region
  roofline   the dominant kernel family (one directional block-GS pass over the fine level), with
             algorithmic bytes from SURVEY.md section 8d
  cpu_baseline  the oracle (CPU restatement of the reference: scipy-order BSR matvec + restated
             pyamg block-GS in C, single thread like the reference) on a bounded sample

--impl reference times that CPU restatement alone (the reference is pure Python + pyamg/scipy
native code; it cannot run the 2048^2 case).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

MMS_U = "-2*sin(pi*x)**2*sin(pi*y)*cos(pi*y)"


def h_factors(n, coarsest=4):
    f, c = [], 2
    while n // c >= coarsest and n % c == 0:
        f.append(c)
        c *= 2
    return f


def rectangle_nodes_file_order(n, P):
    """Uniform elements on [-1,1]^2, LGL interior nodes (the rule the shipped Rectangle_* grids follow,
    SURVEY.md App. A.9).  Returns xn, yn in Plot3D file order [jl][il]."""
    from dg_multigrid_solver_b200.tables import gauss_lobatto_nodes
    xi = gauss_lobatto_nodes(P + 1)
    edges = np.linspace(-1.0, 1.0, n + 1)
    line = np.empty(n * P + 1)
    for e in range(n):                      # same arithmetic as the generator the survey validated
        a, b = edges[e], edges[e + 1]
        line[e * P:(e + 1) * P + 1] = a + (b - a) * (xi + 1.0) / 2.0
    xn = np.repeat(line[None, :], line.size, axis=0)
    yn = np.repeat(line[:, None], line.size, axis=1)
    return xn, yn


def make_params(n, p, gs_mode, check_residual):
    from dg_multigrid_solver_b200.settings import load_params
    prm = load_params(os.path.join(REPO, "input", "paramfile.yml"))
    prm["grid"].update({"filename": f"synthetic_Rectangle_{n}X{n}_nPoly{p}.xyz", "polynomial degree": p,
                        "O grid": False, "circular": False})
    prm["solution"]["u"]["polynomial degree"] = p
    prm["problem"]["SIP penalty parameter multiplier"] = 1.0
    prm["problem"]["exact solution"]["u"] = MMS_U
    mg = prm["solver"]["multigrid"]
    mg["polynomial coarsening"]["levels"]["u"] = ",".join(str(q) for q in ([p, 1] if p > 1 else [1]))
    mg["geometric coarsening"]["coarsening factors"] = ",".join(str(c) for c in h_factors(n))
    prm["solver"]["b200"] = {"gs mode": gs_mode, "check residual": check_residual}
    return prm


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(nnzb, N, b):
    """SURVEY.md section 8d."""
    apply_b = nnzb * (8 * b * b + 4) + 4 * (N + 1) + 16 * b * N
    return {"apply": apply_b, "residual": apply_b + 8 * b * N, "gs_pass": apply_b + 8 * b * N}


def measured_peak():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference_vcycle(n_sample, p, steps, warmup):
    """Oracle V-cycle on an n_sample^2 grid with the same level structure; returns (seconds per
    V-cycle, sample DOFs, setup seconds)."""
    from dgoracle import multigrid, plot3d
    t0 = time.perf_counter()
    x, y = plot3d.rectangle_nodes(n_sample, n_sample, p)
    H = multigrid.Hierarchy(x, y, p, [1, p] if p > 1 else [1], h_factors(n_sample), exact_u=MMS_U,
                            rhs_all_levels=False)
    setup = time.perf_counter() - t0
    fine = H.levels[-1]
    sched = multigrid.Schedule()
    u = np.zeros_like(fine.RHS)
    for _ in range(warmup):
        u = multigrid.v_cycle(H, sched, len(H.levels), fine.RHS, u)
    t0 = time.perf_counter()
    for _ in range(steps):
        u = multigrid.v_cycle(H, sched, len(H.levels), fine.RHS, u)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, fine.RHS.size, setup


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, p = args.size, args.p
    full_dofs = n * n * (p + 1) ** 2
    dt, sample_dofs, setup = cpu_reference_vcycle(args.cpu_sample, p, args.steps, min(args.warmup, 1))
    value = (1.0 / dt) * (sample_dofs / full_dofs)
    sample = (f"oracle V-cycle on Rectangle {args.cpu_sample}x{args.cpu_sample} p={p} (same level structure), "
              f"{sample_dofs} DOFs, scaled by DOFs to {n}x{n}; setup {setup:.1f}s not timed")
    line = {"impl": "reference", "metric": "multigrid_vcycles_per_s", "value": value, "unit": "V-cycles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": "V-cycles/s", "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "V-cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    n, p = args.size, args.p
    return {"workload": f"synthetic Rectangle {n}x{n} Plot3D grid, p={p} Poisson MMS multigrid V-cycle "
                        f"(BASELINE.json configs[2])",
            "elements": n * n, "fine_dofs": n * n * (p + 1) ** 2, "p_levels": [p, 1] if p > 1 else [1],
            "h_factors": h_factors(n), "smoother": "block_gauss_seidel_pyamg symmetric 2 pre / 1 post, 10 coarse",
            "gs_mode": args.gs_mode, "check_residual": bool(args.check_residual),
            "l2_policy": "inputs larger than L2 (fine operator 13.6 GB >> 126 MB); no flush needed"}


def run_b200(args):
    import torch
    from dg_multigrid_solver_b200 import _lib
    from dg_multigrid_solver_b200.dgfem import DGFEM
    from dg_multigrid_solver_b200.grid import Geometry
    from dg_multigrid_solver_b200.settings import Settings
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from dg_multigrid_solver_b200.parallel import run_bench_multi_gpu
        return run_bench_multi_gpu(args, sys.modules[__name__])
    torch.cuda.set_device(local_rank)
    L = _lib.load()
    n, p = args.size, args.p
    settings = Settings(make_params(n, p, args.gs_mode, bool(args.check_residual)))
    xn, yn = rectangle_nodes_file_order(n, p)
    geo = Geometry(None, settings, nodes=(xn, yn))
    del xn, yn
    t0 = time.perf_counter()
    d = DGFEM(settings=settings, geometry=geo, solve_multigrid=True, write_results=False)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    for g in d.grids:
        g.release_geometry()
    torch.cuda.empty_cache()
    solver, fine = d.solver, d.grids[-1]
    nlev = len(d.grids)
    b = fine.d_data.shape[1]
    N = fine.Ni * fine.Nj
    n_dof = N * b
    H = solver.hierarchy()
    rhs_k, u_k, r_k = H["vecs"][nlev - 1]
    rhs_k.copy_(fine.d_rhs)
    u_k.zero_()
    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731

    # ---- device-resident V-cycles ---------------------------------------------------------
    for _ in range(args.warmup):
        solver._vcycle_device(nlev)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    L.dgb_launch_count(1)
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        solver._vcycle_device(nlev)
    e1.record()
    torch.cuda.synchronize()
    launches = int(L.dgb_launch_count(0))
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step
    solver._check_divergence()

    # convergence sanity of what was timed: the normalised residual after warmup+steps cycles
    from dg_multigrid_solver_b200.relaxation import residual_norm
    ss, _ = residual_norm(fine, fine.d_rhs, u_k)
    res_after = float(np.sqrt(ss.item() / n_dof))
    ss0, _ = residual_norm(fine, fine.d_rhs, torch.zeros_like(u_k))
    res0 = float(np.sqrt(ss0.item() / n_dof))

    # ---- end to end through the reference-facing call, host buffers --------------------------
    h_rhs = torch.empty(n_dof, dtype=torch.float64, pin_memory=True)
    h_u = torch.zeros(n_dof, dtype=torch.float64, pin_memory=True)
    h_out = torch.empty(n_dof, dtype=torch.float64, pin_memory=True)
    h_rhs.copy_(fine.d_rhs)
    torch.cuda.synchronize()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(1):
        solver.multigrid_V_cycle(nlev, h_rhs, h_u, out=h_out)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(e2e_steps):
        solver.multigrid_V_cycle(nlev, h_rhs, h_u, out=h_out)     # H2D rhs,u ; V-cycle ; D2H u
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    e2e = {"value": 1e3 / e2e_ms, "unit": "V-cycles/s", "h2d_bytes_per_step": 2 * n_dof * 8,
           "d2h_bytes_per_step": n_dof * 8, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "api": "Solver.multigrid_V_cycle(k, RHS_host, u_host) -> u_host (pinned buffers)"}

    # ---- per-kernel timings on the fine level (roofline) -----------------------------------
    nnzb = int(fine.d_indices.numel())
    ab = algorithmic_bytes(nnzb, N, b)
    st = _lib.stream_ptr()
    ws_part, ws_sum = H["partials"], H["sumsq"]
    y = torch.empty_like(u_k)

    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        L.dgb_launch_count(1)
        a, c = ev(), ev()
        a.record()
        for _ in range(reps):
            fn()
        c.record(); torch.cuda.synchronize()
        return a.elapsed_time(c) / reps, int(L.dgb_launch_count(0)) // reps
    op = fine.operator()
    k_apply = timed(lambda: _lib.call("dgb_bsr_apply", op, u_k, y, st))
    k_resid = timed(lambda: _lib.call("dgb_bsr_residual", op, fine.d_rhs, u_k, None, ws_part, ws_sum, None, st))
    mode = _lib.GS_REDBLACK if args.gs_mode == "redblack" else _lib.GS_LEXICOGRAPHIC
    xg = u_k.clone()
    k_gs = timed(lambda: _lib.call("dgb_block_gs_pass", op, fine.d_rhs, xg, 1, mode, None, st), reps=3)
    # the smoother as the V-cycle calls it (symmetric sweeps, no residual test here): with the chained kernel only
    # the first pass of a call launches the dependency-free helper, every later pass is the chain kernel alone
    ctl0 = H["ctl"]
    sm1 = timed(lambda: _lib.call("dgb_block_gauss_seidel_pyamg", op, fine.d_rhs, xg, 0, 1, mode, 0, ctl0, ws_part,
                                  ws_sum, st), reps=3)
    sm3 = timed(lambda: _lib.call("dgb_block_gauss_seidel_pyamg", op, fine.d_rhs, xg, 0, 3, mode, 0, ctl0, ws_part,
                                  ws_sum, st), reps=2)
    chained = fine.d_chain is not None and mode == _lib.GS_LEXICOGRAPHIC
    t_pass_amortised = (sm3[0] - sm1[0]) / 4.0                   # one later pass of a symmetric sweep
    t_first = sm1[0] - t_pass_amortised                          # first pass of a call (helper + chain when chained)
    peak, peak_src = measured_peak()
    kern = {}
    ab_chain = N * (2 * b * b + 4 * b) * 8        # chain pass: 2 pre-multiplied blocks, c, d in; x, next c out
    ab_helper = N * (3 * b * b + 5 * b) * 8       # helper: Dinv + 2 blocks, rhs, x in; c, d (both streams) out
    rows = [("apply", k_apply, ab["apply"]), ("residual_norm", k_resid, ab["residual"]), ("gs_pass", k_gs, ab["gs_pass"])]
    if chained:
        rows += [("gs_chain_pass", (t_pass_amortised, 1), ab_chain),
                 ("gs_helper", (max(k_gs[0] - t_pass_amortised, 1e-6), 1), ab_helper)]
    else:
        rows += [("gs_pass_in_sweep", (t_pass_amortised, k_gs[1]), ab["gs_pass"])]
    for nm, (ms, nl), nbytes in rows:
        gbs = nbytes / (ms * 1e-3) / 1e9
        kern[nm] = {"ms": ms, "launches": nl, "algorithmic_bytes": nbytes, "GB/s": gbs, "frac": gbs / peak}
    kern["smoother_symmetric_1it_ms"] = sm1[0]
    kern["smoother_symmetric_3it_ms"] = sm3[0]
    # share of one V-cycle spent in each fine-level kernel family (reference schedule: 2 pre + 1 post symmetric
    # iterations = 6 passes in 2 smoother calls; 5 residual evaluations, the restriction reuses the smoother's last)
    n_res = 5 if args.check_residual else 1
    fam = {"residual_norm": n_res * kern["residual_norm"]["ms"]}
    if chained:
        fam["gs_chain_pass"] = 6 * kern["gs_chain_pass"]["ms"]
        fam["gs_helper"] = 2 * kern["gs_helper"]["ms"]
    else:
        fam["gs_pass"] = 6 * kern["gs_pass"]["ms"]
    for nm, ms in fam.items():
        kern[nm]["share_of_vcycle"] = ms / ms_per_step
    top = max(fam, key=fam.get)
    names = {"residual_norm": f"k_rows<{b}, residual> (r = rhs - A u and its norm, fine level)",
             "gs_chain_pass": f"k_gs_chain<{b}> (dependency chain of one lexicographic block-GS pass, fine level)",
             "gs_helper": f"k_gs_helper<{b}> (dependency-free part of a block-GS pass, fine level)",
             "gs_pass": f"block_gs_pass(fine level, b={b}, mode={args.gs_mode})"}
    # DRAM traffic of the dominant kernel from one `ncu --set full` capture of the same launch (profiles/), if there
    # is one for this workload and kernel
    traffic, traffic_src = None, None
    tfile = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles",
                         "r01_ncu_full_k_rows_b9_residual_summary.json")
    if top == "residual_norm" and n == 2048 and p == 2 and os.path.exists(tfile):
        with open(tfile) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["traffic_bytes_per_launch"], "profiles/" + os.path.basename(tfile)
    roofline = {"bound": "hbm", "kernel": names[top], "achieved": kern[top]["GB/s"], "peak": peak, "unit": "GB/s",
                "frac": kern[top]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "share_of_vcycle": kern[top]["share_of_vcycle"],
                "launches_per_call": kern[top]["launches"],
                "algorithmic_bytes_per_launch_group": kern[top]["algorithmic_bytes"],
                "other_kernels": {nm: {"frac": kern[nm]["frac"], "share_of_vcycle": kern[nm]["share_of_vcycle"]}
                                  for nm in fam if nm != top}}
    # V-cycle level traffic: reference schedule = 12 passes/level (+ transfers, ignored)
    vbytes = 0
    for g in d.grids[1:]:
        bb = g.d_data.shape[1]
        a_g = algorithmic_bytes(int(g.d_indices.numel()), g.Ni * g.Nj, bb)
        vbytes += 6 * a_g["gs_pass"] + (6 if args.check_residual else 1) * a_g["residual"]
    g0 = d.grids[0]
    a_0 = algorithmic_bytes(int(g0.d_indices.numel()), g0.Ni * g0.Nj, g0.d_data.shape[1])
    vbytes += 20 * a_0["gs_pass"] + (11 if args.check_residual else 0) * a_0["residual"]
    vcycle_gbs = vbytes / (ms_per_step * 1e-3) / 1e9

    # ---- CPU baseline (oracle), bounded sample ----------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        dt, sample_dofs, setup = cpu_reference_vcycle(args.cpu_sample, p, 2, 1)
        cpu_val = (1.0 / dt) * (sample_dofs / n_dof)
        cpu = {"value": cpu_val, "unit": "V-cycles/s", "cores": 1, "kind": "port",
               "sample": f"oracle V-cycle on Rectangle {args.cpu_sample}x{args.cpu_sample} p={p} ({sample_dofs} DOFs, "
                         f"{dt * 1e3:.0f} ms/cycle), scaled by DOFs to {n}x{n}; host has {os.cpu_count()} cores, "
                         f"the reference path is single-threaded"}

    line = {"metric": "multigrid_vcycles_per_s", "value": value, "unit": "V-cycles/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu,
            "kernels": kern,
            "vcycle": {"algorithmic_bytes": vbytes, "GB/s": vcycle_gbs, "frac_of_peak": vcycle_gbs / peak,
                       "normalised_residual_after_timed_cycles": res_after / res0,
                       "cycles_run": args.warmup + args.steps},
            "apply_dof_per_s": n_dof / (k_apply[0] * 1e-3),
            "vcycle_dof_per_s": n_dof * value,
            "setup_s": setup_s, "assemble_s": d.timings.get("assemble"),
            "assembly_elements_per_s": sum(g.Ni * g.Nj for g in d.grids) / d.timings["assemble"]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=2048, help="elements per direction (BASELINE configs[2]: 2048)")
    ap.add_argument("--p", type=int, default=2)
    ap.add_argument("--gs-mode", default="lexicographic", choices=["lexicographic", "redblack", "slab_lexicographic"])
    ap.add_argument("--check-residual", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exact-multi", action="store_true",
                    help="N>1: keep the exact global lexicographic order (slabs sweep one after the other)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
