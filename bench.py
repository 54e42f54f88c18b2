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
  roofline   the kernel with the largest measured share of the V-cycle (launches per cycle x CUDA-event time per
             launch, counted from the schedule the library runs), algorithmic bytes per launch as DESIGN.md
             section 4 states them, DRAM traffic from the tracked ncu summaries (profiles/r02_ncu_traffic.json)
  vcycle     bytes_min of SURVEY.md section 8d (6 passes + 1 residual per level) AND the bytes the library's
             kernels must move, each over the measured time -- never the reference schedule's 12 passes/level
  cpu_baseline  the oracle (CPU restatement of the reference: scipy-order BSR matvec + restated
             pyamg block-GS in C, single thread like the reference) on a bounded sample: one V-cycle of the
             same level structure on a --cpu-sample^2 grid.  `measured` is the un-scaled rate on that grid,
             `value` the same rate scaled by DOFs to the full grid (the per-DOF cost of a V-cycle is flat in n)
  parity     the GPU V-cycle on the CPU sample grid against the oracle's (same inputs), inside the bench

--config c4 / c5 run BASELINE.json configs[3] / configs[4] at full size (p=5 apply / block-Jacobi / block-GS
sweeps + assembly; Stokes assembly + apply) and print one line each in the same contract.
--impl reference times the CPU restatement alone (the reference is pure Python + pyamg/scipy native code; it
cannot run the 2048^2 case): a step = one oracle V-cycle on the sample grid.
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
    V-cycle, sample DOFs, setup seconds, u after the first V-cycle from u = 0)."""
    from dgoracle import multigrid, plot3d
    t0 = time.perf_counter()
    x, y = plot3d.rectangle_nodes(n_sample, n_sample, p)
    H = multigrid.Hierarchy(x, y, p, [1, p] if p > 1 else [1], h_factors(n_sample), exact_u=MMS_U,
                            rhs_all_levels=False)
    setup = time.perf_counter() - t0
    fine = H.levels[-1]
    sched = multigrid.Schedule()
    u = np.zeros_like(fine.RHS)
    u_first = None
    for _ in range(max(warmup, 1)):
        u = multigrid.v_cycle(H, sched, len(H.levels), fine.RHS, u)
        if u_first is None:
            u_first = u.copy()
    t0 = time.perf_counter()
    for _ in range(steps):
        u = multigrid.v_cycle(H, sched, len(H.levels), fine.RHS, u)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, fine.RHS.size, setup, u_first


def run_reference(args):
    """--impl reference: the CPU arm alone.  A step = one oracle V-cycle on the sample grid (same level structure as
    the workload); ms_per_step is what was measured, `value` the DOF-scaled full-grid equivalent (stated in
    cpu_baseline.sample / measured)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, p = args.size, args.p
    full_dofs = n * n * (p + 1) ** 2
    dt, sample_dofs, setup, _ = cpu_reference_vcycle(args.cpu_sample, p, args.steps, min(args.warmup, 1))
    value = (1.0 / dt) * (sample_dofs / full_dofs)
    cpu = cpu_baseline_block(args, dt, sample_dofs, setup, full_dofs)
    cfg = workload_config(args)
    cfg["cpu_sample"] = f"Rectangle {args.cpu_sample}x{args.cpu_sample} (timed), DOF-scaled to {n}x{n} (value)"
    line = {"impl": "reference", "metric": "multigrid_vcycles_per_s", "value": value, "unit": "V-cycles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "ms_per_step_is": "one oracle V-cycle on the sample grid (un-scaled)",
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "V-cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_block(args, dt, sample_dofs, setup, full_dofs):
    n, p, m = args.size, args.p, args.cpu_sample
    return {"value": (1.0 / dt) * (sample_dofs / full_dofs), "unit": "V-cycles/s", "cores": 1, "kind": "port",
            "measured": {"grid": f"{m}x{m}", "dofs": sample_dofs, "s_per_vcycle": dt, "vcycles_per_s": 1.0 / dt,
                         "dof_per_s": sample_dofs / dt, "setup_s_not_timed": setup},
            "sample": f"oracle (C/NumPy restatement of the reference, single thread like the reference's scipy/pyamg "
                      f"path) V-cycle on Rectangle {m}x{m} p={p}, same level structure: {dt:.3f} s per cycle "
                      f"measured; `value` scales that by DOFs ({sample_dofs} -> {full_dofs}) to {n}x{n}, which the "
                      f"CPU path cannot assemble in bench time; host has {os.cpu_count()} cores"}


def workload_config(args):
    n, p = args.size, args.p
    return {"workload": f"synthetic Rectangle {n}x{n} Plot3D grid, p={p} Poisson MMS multigrid V-cycle "
                        f"(BASELINE.json configs[2])",
            "elements": n * n, "fine_dofs": n * n * (p + 1) ** 2, "p_levels": [p, 1] if p > 1 else [1],
            "h_factors": h_factors(n), "smoother": "block_gauss_seidel_pyamg symmetric 2 pre / 1 post, 10 coarse",
            "gs_mode": args.gs_mode, "check_residual": bool(args.check_residual),
            "launch": "one CUDA graph per V-cycle (device-timed loop); plain launches through the host-buffer API (e2e)",
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
    replays0 = solver.graph_replays
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        solver._vcycle_device(nlev)
    e1.record()
    torch.cuda.synchronize()
    # kernels launched inside the timed region: direct launches + what the CUDA-graph replays launched
    launches = int(L.dgb_launch_count(0)) + (solver.graph_replays - replays0) * solver.graph_launches
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

    # ---- a full solve (Solver.solve_multigrid: residual test + V-cycle per iteration, from u = 0 to 1e-6) ----------
    solve = None
    if args.solve:
        import tempfile
        u0 = torch.zeros_like(u_k)
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:       # the loop pickles its residual history below the cwd (solver.py:128-138)
            os.chdir(tmp)
            try:
                for rep in range(2):         # the first run captures the graph of the cycle that follows the loop's residual
                    solver.residuals = []
                    solver.primed_cycles = 0
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    solver.solve_multigrid(nlev, fine.d_rhs, u0)
                    torch.cuda.synchronize()
                    dt_solve = time.perf_counter() - t0
            finally:
                os.chdir(cwd)
        ncyc = len(solver.residuals) - 1
        solve = {"cycles": ncyc, "s": dt_solve, "ms_per_iteration": 1e3 * dt_solve / max(ncyc, 1),
                 "final_normalised_residual": solver.residuals[-1], "tolerance": 1e-6,
                 "cycles_started_from_the_loops_residual": solver.primed_cycles,
                 "timing": "host wall clock around Solver.solve_multigrid (device-resident rhs/u, one host sync per cycle)"}
        rhs_k.copy_(fine.d_rhs)

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
    ctl0.zero_()
    chained = fine.d_chain is not None and mode == _lib.GS_LEXICOGRAPHIC
    t_pass_amortised = (sm3[0] - sm1[0]) / 4.0                   # one later pass of a symmetric sweep
    peak, peak_src = measured_peak()
    kern = {}
    # algorithmic bytes per launch
    ab_chain = N * (2 * b * b + 4 * b) * 8        # chain pass: 2 pre-multiplied blocks, c, d in; x, next c out
    ab_helper = N * (3 * b * b + 5 * b) * 8       # helper: Dinv + 2 blocks, rhs, x in; c, d (both streams) out
    ab_entry = ab["residual"] + N * (b * b + 3 * b) * 8   # fused entry residual + helper: 5 blocks + Dinv; + c, d, d out
    rows = [("apply", k_apply, ab["apply"]), ("residual_norm", k_resid, ab["residual"]), ("gs_pass", k_gs, ab["gs_pass"])]
    if chained:
        k_entry = timed(lambda: L.dgb_block_gs_entry_residual(__import__("ctypes").byref(op), _lib.ptr(fine.d_rhs),
                                                              _lib.ptr(xg), 1, _lib.ptr(y), _lib.ptr(ws_part),
                                                              _lib.ptr(ws_sum), st), reps=3)
        # residual right after a pass, from the records the pass left (k_residual_rec): after 1 backward pass
        _lib.call("dgb_block_gauss_seidel_pyamg", op, fine.d_rhs, xg, 0, 1, mode, 0, ctl0, ws_part, ws_sum, st)
        k_rr = timed(lambda: L.dgb_block_gs_residual_after_pass(__import__("ctypes").byref(op), _lib.ptr(xg), -1, _lib.ptr(y),
                                                                _lib.ptr(ws_part), _lib.ptr(ws_sum), None, st), reps=3)
        ab_rr = N * (3 * b * b + 5 * b) * 8          # 2 pre-multiplied blocks + c, d; diagonal block; x, r
        rows += [("residual_after_pass", k_rr, ab_rr)]
        rows += [("gs_chain_pass", (t_pass_amortised, 1), ab_chain),
                 ("gs_helper", (max(k_gs[0] - t_pass_amortised, 1e-6), 1), ab_helper),
                 ("gs_entry_residual", k_entry, ab_entry)]
    else:
        rows += [("gs_pass_in_sweep", (t_pass_amortised, k_gs[1]), ab["gs_pass"])]
    for nm, (ms, nl), nbytes in rows:
        gbs = nbytes / (ms * 1e-3) / 1e9
        kern[nm] = {"ms": ms, "launches": nl, "algorithmic_bytes": nbytes, "GB/s": gbs, "frac": gbs / peak}
    kern["smoother_symmetric_1it_ms"] = sm1[0]
    kern["smoother_symmetric_3it_ms"] = sm3[0]
    # launches per V-cycle on the fine level, from the schedule the library runs (dgb_vcycle.cu, gs_pyamg in
    # dgb_solve.cu): pre 2 + post 1 symmetric iterations = 6 passes in 2 smoother calls; with the residual tests
    # each call opens with the fused entry residual (chained) or a plain residual, every iteration closes with a
    # residual (3), and the restriction reuses the pre-smoother's last one
    chk = bool(args.check_residual)
    if chained:
        rec_res = b in (4, 9, 16)
        per_cycle = {"gs_chain_pass": 6, "residual_after_pass": 3 if (chk and rec_res) else 0,
                     "residual_norm": (0 if rec_res else 3) if chk else 1,
                     "gs_entry_residual": 2 if chk else 0, "gs_helper": 0 if chk else 2}
    else:
        per_cycle = {"gs_pass": 6, "residual_norm": 5 if chk else 1}
    fam = {nm: cnt * kern[nm]["ms"] for nm, cnt in per_cycle.items() if cnt}
    for nm, ms in fam.items():
        kern[nm]["launches_per_vcycle"] = per_cycle[nm]
        kern[nm]["share_of_vcycle"] = ms / ms_per_step
    top = max(fam, key=fam.get)
    names = {"residual_norm": f"k_rows<{b}, residual> (r = rhs - A u and its norm, fine level)",
             "gs_chain_pass": f"k_gs_chain<{b}> (dependency chain of one lexicographic block-GS pass, fine level)",
             "gs_helper": f"k_gs_helper<{b}, false> (dependency-free part of a block-GS pass, fine level)",
             "gs_entry_residual": f"k_gs_helper<{b}, true> (smoother entry residual fused with the dependency-free part)",
             "residual_after_pass": f"k_residual_rec<{b}> (residual test of a smoother iteration, from the pass's records)",
             "gs_pass": f"block_gs_pass(fine level, b={b}, mode={args.gs_mode})"}
    # DRAM traffic per launch from the tracked `ncu --set full` summaries of the same kernels (profiles/)
    traffic_tab = {}
    tfile = os.path.join(REPO, "profiles", "r02_ncu_traffic.json")
    if n == 2048 and p == 2 and os.path.exists(tfile):
        with open(tfile) as f:
            traffic_tab = json.load(f)
    tr = traffic_tab.get(top, {})
    roofline = {"bound": "hbm", "kernel": names[top], "achieved": kern[top]["GB/s"], "peak": peak, "unit": "GB/s",
                "frac": kern[top]["frac"], "traffic": tr.get("traffic_bytes_per_launch"),
                "traffic_source": tr.get("source"), "peak_source": peak_src,
                "share_of_vcycle": kern[top]["share_of_vcycle"], "launches_per_vcycle": per_cycle[top],
                "ms_per_launch": kern[top]["ms"],
                "algorithmic_bytes_per_launch": kern[top]["algorithmic_bytes"],
                "selection": "largest launches_per_vcycle x ms_per_launch among the fine-level kernels (measured live)",
                "other_kernels": {nm: {"frac": kern[nm]["frac"], "share_of_vcycle": kern[nm]["share_of_vcycle"],
                                       "launches_per_vcycle": per_cycle[nm],
                                       "traffic": traffic_tab.get(nm, {}).get("traffic_bytes_per_launch")}
                                  for nm in fam if nm != top}}
    # ---- V-cycle level traffic --------------------------------------------------------------
    # bytes_min: per non-coarsest level 6 passes + 1 residual, coarsest 20 passes, + transfers
    # bytes_moved: what this library's kernels have to move for the same cycle (chain passes stream 2 of the 5
    # blocks; the residual tests of the reference's schedule are kept)
    bytes_min = bytes_moved = 0
    for li, g in enumerate(d.grids):
        bb = g.d_data.shape[1]
        Ng = g.Ni * g.Nj
        a_g = algorithmic_bytes(int(g.d_indices.numel()), Ng, bb)
        ch = g.d_chain is not None and mode == _lib.GS_LEXICOGRAPHIC
        c_pass = Ng * (2 * bb * bb + 4 * bb) * 8 if ch else a_g["gs_pass"]
        c_entry = a_g["residual"] + (Ng * (bb * bb + 3 * bb) * 8 if ch else 0)
        c_res = Ng * (3 * bb * bb + 5 * bb) * 8 if (ch and bb in (4, 9, 16)) else a_g["residual"]
        if li == 0:
            bytes_min += 20 * a_g["gs_pass"]
            bytes_moved += 20 * c_pass + ((c_entry + 10 * c_res) if chk else (Ng * (3 * bb * bb + 5 * bb) * 8 if ch else 0))
        else:
            bytes_min += 6 * a_g["gs_pass"] + a_g["residual"]
            bytes_moved += 6 * c_pass + ((2 * c_entry + 3 * c_res) if chk
                                         else (a_g["residual"] + (2 * Ng * (3 * bb * bb + 5 * bb) * 8 if ch else 0)))
            cg = d.grids[li - 1]
            tb = 8 * (Ng * bb + cg.Ni * cg.Nj * cg.d_data.shape[1]) * 2 + 8 * Ng * bb      # restrict + prolong-add
            bytes_min += tb
            bytes_moved += tb
    vc_t = ms_per_step * 1e-3
    vcycle = {"bytes_min": bytes_min, "bytes_min_GBs": bytes_min / vc_t / 1e9, "frac_bytes_min": bytes_min / vc_t / 1e9 / peak,
              "bytes_moved": bytes_moved, "bytes_moved_GBs": bytes_moved / vc_t / 1e9,
              "frac_bytes_moved": bytes_moved / vc_t / 1e9 / peak,
              "frac_bytes_min_of_nominal_8TBs": bytes_min / vc_t / 1e9 / 8000.0,
              "normalised_residual_after_timed_cycles": res_after / res0,
              "cycles_run": args.warmup + args.steps}
    if solve is not None:
        # per iteration of the loop: the cycle's bytes_min plus the loop's own residual of the finest level
        it_bytes = bytes_min + ab["residual"]
        solve["bytes_min_per_iteration"] = it_bytes
        solve["frac_bytes_min"] = it_bytes * solve["cycles"] / solve["s"] / 1e9 / peak
        solve["dof_per_s"] = n_dof / solve["s"]
    # device memory the hierarchy holds (operator, inverse diagonal blocks, smoother streams, vectors)
    mem = {"data": 0, "dinv": 0, "gs_chain": 0, "gs_data": 0, "mailbox": 0, "vectors": 0}
    for g in d.grids:
        for key, t in (("data", g.d_data), ("dinv", g.d_dinv), ("gs_chain", g.d_chain), ("gs_data", g.d_gs),
                       ("mailbox", g.d_mailbox)):
            if t is not None:
                mem[key] += t.numel() * t.element_size()
        mem["vectors"] += 4 * g.Ni * g.Nj * g.d_data.shape[1] * 8
    mem["total_GB"] = sum(v for v in mem.values()) / 1e9
    mem["torch_allocated_GB"] = torch.cuda.memory_allocated() / 1e9

    # ---- CPU baseline (oracle) on a bounded sample + parity of the GPU cycle on that sample ------------
    cpu = parity = None
    if not args.no_cpu_baseline:
        dt, sample_dofs, setup, u_first = cpu_reference_vcycle(args.cpu_sample, p, 1, 1)
        cpu = cpu_baseline_block(args, dt, sample_dofs, setup, n_dof)
        m = args.cpu_sample
        s2 = Settings(make_params(m, p, args.gs_mode, bool(args.check_residual)))
        d2 = DGFEM(settings=s2, geometry=Geometry(None, s2, nodes=rectangle_nodes_file_order(m, p)),
                   solve_multigrid=True, write_results=False)
        f2 = d2.grids[-1]
        u_gpu = d2.solver.multigrid_V_cycle(len(d2.grids), f2.RHS, np.zeros_like(f2.RHS))
        err = float(np.abs(u_gpu - u_first).max() / np.abs(u_first).max())
        parity = {"grid": f"{m}x{m}", "check": "u after one V-cycle from u=0, GPU vs oracle, max rel err",
                  "rel_err": err, "tolerance": 1e-10, "ok": bool(err < 1e-10) if args.gs_mode == "lexicographic" else None}
        del d2

    line = {"metric": "multigrid_vcycles_per_s", "value": value, "unit": "V-cycles/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "graph_replays": solver.graph_replays,
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "solve": solve,
            "kernels": kern, "vcycle": vcycle, "memory": mem,
            "apply_dof_per_s": n_dof / (k_apply[0] * 1e-3),
            "vcycle_dof_per_s": n_dof * value,
            "setup_s": setup_s, "assemble_s": d.timings.get("assemble"),
            "smoother_setup_s": d.timings.get("smoother_setup"),
            "assembly_elements_per_s": sum(g.Ni * g.Nj for g in d.grids) / d.timings["assemble"]}
    if args.p5_apply and n >= 1024:
        # BASELINE's "operator-apply DOF/s (p=2,5)": the p=5 number at configs[3]'s size, in the same run
        del d, solver, H, xg, y, rhs_k, u_k, r_k
        torch.cuda.empty_cache()
        try:
            c4 = run_config_c4(1024, quick=True)
            line["apply_dof_per_s_p5"] = c4["apply_dof_per_s"]
            line["apply_p5"] = {k: c4[k] for k in ("config", "dofs", "b", "apply_ms", "apply_GBs", "apply_frac",
                                                   "assemble_s", "smoother_setup_s", "assembly_elements_per_s")}
        except Exception as e:                       # the headline line must not depend on the second workload
            line["apply_p5"] = {"error": repr(e)[:200]}
    print(json.dumps(line), flush=True)


def _lgl_line(edges, P):
    from dg_multigrid_solver_b200.tables import gauss_lobatto_nodes
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
    th = _lgl_line(-2.0 * np.pi * np.arange(n + 1) / n, P)
    rr = _lgl_line(redges, P)
    x = np.cos(th)[None, :] * rr[:, None]
    y = np.sin(th)[None, :] * rr[:, None]
    x[:, -1], y[:, -1] = x[:, 0], y[:, 0]          # close the O-grid exactly (grid.py:56-57)
    return np.ascontiguousarray(x), np.ascontiguousarray(y)


def _timed(fn, reps=3):
    import torch
    fn(); torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c) / reps


def run_config_c4(n=1024, quick=False):
    """BASELINE.json configs[3]: CircleInCircle n x n nPoly5 (O-grid, sigma-mult 2), p=5 single level: assembly,
    operator apply, block-Jacobi sweep, symmetric lexicographic block-GS iterations (36x36 blocks), and the
    reference's `-s --smoother block_gauss_seidel_pyamg` run (100 iterations with residual tests).
    quick: assembly + apply only."""
    import torch
    from dg_multigrid_solver_b200 import _lib
    from dg_multigrid_solver_b200.dgfem import DGFEM
    from dg_multigrid_solver_b200.grid import Geometry
    from dg_multigrid_solver_b200.settings import Settings
    peak, _ = measured_peak()
    st = _lib.stream_ptr()
    p = 5
    prm = make_params(n, p, "lexicographic", True)
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
    g.release_geometry()
    N, b, nnzb = g.Ni * g.Nj, g.b, int(g.d_indices.numel())
    ab = algorithmic_bytes(nnzb, N, b)
    x = torch.randn(N * b, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    op = g.operator()
    t_apply = _timed(lambda: _lib.call("dgb_bsr_apply", op, x, y, st))
    out = {"config": f"C4 CircleInCircle {n}x{n} p=5 O-grid (BASELINE.json configs[3])", "elements": N, "dofs": N * b,
           "b": b, "nnzb": nnzb, "operator_GB": nnzb * b * b * 8 / 1e9, "setup_s": setup,
           "assemble_s": d.timings.get("assemble"), "smoother_setup_s": d.timings.get("smoother_setup"),
           "assembly_elements_per_s": N / d.timings["assemble"],
           "apply_ms": t_apply, "apply_GBs": ab["apply"] / t_apply / 1e6, "apply_frac": ab["apply"] / t_apply / 1e6 / peak,
           "apply_dof_per_s": N * b / (t_apply * 1e-3)}
    if quick:
        return out
    t_jac = _timed(lambda: _lib.call("dgb_block_relax_sweep", op, g.d_rhs, x, y, 1.0, st))
    xg = torch.zeros_like(x)
    L = _lib.load()
    ctl = torch.zeros(32, dtype=torch.uint8, device="cuda")
    part = torch.zeros(L.dgb_partials_len(), dtype=torch.float64, device="cuda")
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")

    def smoother(iters):      # the smoother as Solver.solve_smoother calls it, without its residual tests
        _lib.call("dgb_block_gauss_seidel_pyamg", op, g.d_rhs, xg, 0, iters, 0, 0, ctl, part, ss, st)
    t1 = _timed(lambda: smoother(1), reps=2)
    t3 = _timed(lambda: smoother(3), reps=2)
    t_gs = (t3 - t1) / 2.0                       # one symmetric iteration inside a longer call = 2 chain passes
    # the reference's `-s --smoother block_gauss_seidel_pyamg` run: 100 symmetric iterations with residual tests
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d.solver.solve()
    torch.cuda.synchronize()
    t_s100 = time.perf_counter() - t0
    chained = g.d_chain is not None
    # bytes the pass kernels must move: a chained pass streams 2 pre-multiplied blocks + c, d, x, next c
    pass_bytes = N * (2 * b * b + 4 * b) * 8 if chained else ab["gs_pass"]
    out.update({"smoother_run_100_iterations_s": t_s100,
                "block_jacobi_sweep_ms": t_jac, "block_jacobi_GBs": ab["gs_pass"] / t_jac / 1e6,
                "block_jacobi_frac": ab["gs_pass"] / t_jac / 1e6 / peak,
                "block_jacobi_dof_per_s": N * b / (t_jac * 1e-3),
                "gs_first_symmetric_iteration_ms": t1, "gs_symmetric_iteration_ms": t_gs,
                "gs_pass_bytes_moved": pass_bytes, "gs_pass_GBs": 2 * pass_bytes / t_gs / 1e6,
                "gs_pass_frac": 2 * pass_bytes / t_gs / 1e6 / peak,
                "gs_pass_GBs_on_survey_8d_bytes": 2 * ab["gs_pass"] / t_gs / 1e6,
                "gs_sweep_dof_per_s": 2 * N * b / (t_gs * 1e-3), "device_error": L.dgb_device_error(1)})
    return out


def run_config_c5(n=1024):
    """BASELINE.json configs[4]: Rectangle n x n nPoly2, Stokes local ordering (p_u=2, p_p=1): assembly + apply."""
    import torch
    from dg_multigrid_solver_b200 import _lib
    from dg_multigrid_solver_b200.dgfem import DGFEM
    from dg_multigrid_solver_b200.grid import Geometry
    from dg_multigrid_solver_b200.settings import Settings
    peak, _ = measured_peak()
    st = _lib.stream_ptr()
    prm = make_params(n, 2, "lexicographic", True)
    prm["problem"]["type"] = "Stokes"
    prm["problem"]["include pressure BC"] = False
    prm["solution"]["p"]["polynomial degree"] = 1
    prm["solution"]["ordering"] = "local"
    s = Settings(prm)
    xn, yn = rectangle_nodes_file_order(n, 2)
    t0 = time.perf_counter()
    d = DGFEM(settings=s, geometry=Geometry(None, s, nodes=(xn, yn)), solve_direct=True, write_results=False)
    torch.cuda.synchronize()
    setup = time.perf_counter() - t0
    g = d.grids[-1]
    N, b, nnzb = g.Ni * g.Nj, int(g.d_data.shape[1]), int(g.d_indices.numel())
    ab = algorithmic_bytes(nnzb, N, b)
    x = torch.randn(N * b, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    op = g.operator()
    t_apply = _timed(lambda: _lib.call("dgb_bsr_apply", op, x, y, st))
    return {"config": f"C5 Rectangle {n}x{n} Stokes p_u=2 p_p=1 local order (BASELINE.json configs[4])", "elements": N,
            "dofs": N * b, "b": b, "nnzb": nnzb, "operator_GB": nnzb * b * b * 8 / 1e9, "setup_s": setup,
            "assemble_s": d.timings.get("assemble"), "assembly_elements_per_s": N / d.timings["assemble"],
            "apply_ms": t_apply, "apply_GBs": ab["apply"] / t_apply / 1e6, "apply_frac": ab["apply"] / t_apply / 1e6 / peak,
            "apply_dof_per_s": N * b / (t_apply * 1e-3)}


def run_other_config(args):
    """--config c4 | c5: one JSON line in the bench contract (metric = operator-apply DOF/s of that configuration,
    the configuration's other device-timed numbers beside it)."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    n = args.size if args.size != 2048 else 1024
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    out = run_config_c4(n) if args.config == "c4" else run_config_c5(n)
    clocks = sampler.stop()
    peak, peak_src = measured_peak()
    line = {"metric": "operator_apply_dof_per_s", "value": out["apply_dof_per_s"], "unit": "DOF/s", "n_gpus": 1,
            "steps": 3, "warmup": 1, "ms_per_step": out["apply_ms"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": out["config"], "l2_policy": f"inputs larger than L2 ({out['operator_GB']:.1f} GB operator)"},
            "clocks": clocks, "gpu_launches": 4,
            "roofline": {"bound": "hbm", "kernel": f"k_rows<{out['b']}, apply>", "achieved": out["apply_GBs"], "peak": peak,
                         "unit": "GB/s", "frac": out["apply_frac"], "traffic": None, "peak_source": peak_src},
            "e2e": None, "cpu_baseline": None, "details": out}
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
    ap.add_argument("--cpu-sample", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--solve", type=int, default=1, help="also time a full Solver.solve_multigrid run (u = 0 -> 1e-6)")
    ap.add_argument("--config", default="c3", choices=["c3", "c4", "c5"],
                    help="c3 = the V-cycle workload (BASELINE configs[2], default); c4 / c5 = configs[3] / configs[4]")
    ap.add_argument("--p5-apply", type=int, default=1,
                    help="also measure the p=5 operator apply (configs[3]'s operator, 1024^2) in the default run")
    ap.add_argument("--min-rows", type=int, default=0,
                    help="N>1: levels with fewer element rows per rank are replicated (0 = the library's default)")
    ap.add_argument("--exact-multi", action="store_true",
                    help="N>1: keep the exact global lexicographic order (slabs sweep one after the other)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.config != "c3":
        run_other_config(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
