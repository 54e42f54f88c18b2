#!/usr/bin/env python3
"""Timeline of one chained Gauss-Seidel pass (diagnostic build: DGB_LIB=.../libdgb200_trace.so, -DDGB_CHAIN_TRACE).
usage: chain_trace.py NI NJ P  -> per-band start / first-records / end times and wait shares, summarised per kind of
hand-over (shared-memory ring, cluster DSMEM, global mailbox)."""
import ctypes
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tools")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from probe_kernels import nodes  # noqa: E402
from dg_multigrid_solver_b200 import _lib  # noqa: E402
from dg_multigrid_solver_b200.discrete_system import DiscreteSystem  # noqa: E402
from dg_multigrid_solver_b200.grid import Geometry, Grid  # noqa: E402
from dg_multigrid_solver_b200.settings import Settings  # noqa: E402


def main():
    ni, nj, p = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    Pg = max(p, 1)
    prm = bench.make_params(max(ni, nj), Pg, "lexicographic", True)
    prm["solution"]["u"]["polynomial degree"] = p
    s = Settings(prm)
    s.update_setting("solver.method", "smoother")
    s.update_setting("solver.discretization", "dg")
    geo = Geometry(None, s, nodes=nodes(ni, nj, Pg))
    g = Grid(geo, ["u"]).initialize({"u": p}, None)
    DiscreteSystem(s).problem.assemble(g)
    g.release_geometry()
    L = _lib.load()
    st = _lib.stream_ptr()
    op = g.operator()
    b = g.b
    x = torch.randn(g.Ni * g.Nj * b, dtype=torch.float64, device="cuda")
    R = {4: 8, 9: 1, 16: 1, 25: 1}.get(b, 1)
    nb = (nj + R - 1) // R
    for _ in range(3):
        _lib.call("dgb_block_gs_pass", op, g.d_rhs, x, 1, 0, None, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.dgb_set_kernel_path(100 + 22)         # chain kernel alone
    e0.record()
    _lib.call("dgb_block_gs_pass", op, g.d_rhs, x, 1, 0, None, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out = np.zeros((min(nb, 8192), 16), dtype=np.int64)
    L.dgb_debug_chain_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rc = L.dgb_debug_chain_trace(out.ctypes.data, out.shape[0])
    assert rc == 0, rc
    t0 = out[:, 0].min()
    start, first, end = (out[:, 0] - t0) * 1e-3, (out[:, 1] - t0) * 1e-3, (out[:, 2] - t0) * 1e-3   # us
    nwait, cyc_wait, cyc_mbar, cyc_flow = out[:, 3], out[:, 4], out[:, 5], out[:, 6]
    info = out[:, 7]
    smid, pred, succ, w, crank = info & 0xffff, (info >> 16) & 0xf, (info >> 20) & 0xf, (info >> 24) & 0xff, info >> 32
    dur = end - first
    lag_end = np.diff(end)
    lag_first = np.diff(first)
    kind = np.where(pred[1:] == 2, "mailbox", np.where(w[1:] == 0, "dsmem", "ring"))
    clk = 1.9e3        # cycles per us (approx.)
    res = {"Ni": ni, "Nj": nj, "b": b, "bands": int(nb), "pass_ms": ms, "span_us": float(end.max()),
           "band_duration_us": {"median": float(np.median(dur)), "min": float(dur.min()), "max": float(dur.max())},
           "steps": ni + R - 1,
           "step_ns_of_median_band": float(np.median(dur)) * 1e3 / (ni + R - 1),
           "first_band_duration_us": float(dur[0]),
           "share_wait_up": float(np.median(cyc_wait / clk / np.maximum(dur, 1e-9))),
           "share_mbar": float(np.median(cyc_mbar / clk / np.maximum(dur, 1e-9))),
           "share_flow": float(np.median(cyc_flow / clk / np.maximum(dur, 1e-9))),
           "median_slow_path_steps": float(np.median(nwait))}
    for k in ("ring", "dsmem", "mailbox"):
        m = kind == k
        if m.any():
            res[f"lag_end_us[{k}]"] = {"n": int(m.sum()), "median": float(np.median(lag_end[m])), "mean": float(lag_end[m].mean())}
            res[f"lag_first_us[{k}]"] = {"median": float(np.median(lag_first[m])), "mean": float(lag_first[m].mean())}
    nch = np.maximum(out[:, 10], 1)
    res["cycles_per_chunk"] = {"mbar_wait": float(np.median(cyc_mbar / nch)), "flow_control": float(np.median(cyc_flow / nch)),
                               "steps": float(np.median(out[:, 8] / nch)), "epilogue": float(np.median(out[:, 9] / nch)),
                               "wait_up": float(np.median(cyc_wait / nch)), "band0": [float(out[0, k] / nch[0]) for k in (5, 6, 8, 9, 4)]}
    # distance (us) to the band above at the end of chunks 0, 1, 7, 63 and at the end of the row (ring hand-overs only)
    ring = np.concatenate([[False], kind == "ring"])
    for nm, col in (("chunk0", 11), ("chunk1", 12), ("chunk7", 13), ("chunk63", 14), ("end", 2)):
        tcol = (out[:, col] - t0) * 1e-3
        lag = tcol[1:] - tcol[:-1]
        ok = ring[1:] & (out[1:, col] > 0) & (out[:-1, col] > 0)
        if ok.any():
            res[f"lag_ring_us@{nm}"] = {"median": float(np.median(lag[ok])), "mean": float(lag[ok].mean())}
    res["sum_lag_end_us"] = {k: float(lag_end[kind == k].sum()) for k in ("ring", "dsmem", "mailbox")}
    # when do bands start relative to their predecessor's end (a band that starts after its SM freed up)
    res["bands_started_after_t0_us"] = [float(v) for v in np.percentile(start, [0, 25, 50, 75, 100])]
    res["device_error"] = L.dgb_device_error(1)
    print(json.dumps(res))
    np.save(os.path.join(REPO, "gpurun_out", f"chain_trace_{ni}x{nj}_b{b}.npy"), out)


if __name__ == "__main__":
    main()
