#!/usr/bin/env python3
"""Time the solve-phase kernels of one level in isolation (CUDA events), both kernel paths.
usage: probe_kernels.py NI NJ P [reps]   -- Rectangle NI x NJ elements, solution degree P."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from dg_multigrid_solver_b200 import _lib  # noqa: E402
from dg_multigrid_solver_b200.discrete_system import DiscreteSystem  # noqa: E402
from dg_multigrid_solver_b200.grid import Geometry, Grid  # noqa: E402
from dg_multigrid_solver_b200.settings import Settings  # noqa: E402


def nodes(ni, nj, P):
    from dg_multigrid_solver_b200.tables import gauss_lobatto_nodes
    xi = gauss_lobatto_nodes(P + 1)

    def line(n):
        e = np.linspace(-1.0, 1.0, n + 1)
        out = np.empty(n * P + 1)
        for k in range(n):
            out[k * P:(k + 1) * P + 1] = e[k] + (e[k + 1] - e[k]) * (xi + 1.0) / 2.0
        return out
    lx, ly = line(ni), line(nj)
    return np.repeat(lx[None, :], ly.size, axis=0), np.repeat(ly[:, None], lx.size, axis=1)


def main():
    ni, nj, p = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    only = sys.argv[5] if len(sys.argv) > 5 else ""
    Pg = max(p, 1)
    prm = bench.make_params(max(ni, nj), Pg, "lexicographic", True)
    prm["solution"]["u"]["polynomial degree"] = p
    s = Settings(prm)
    s.update_setting("solver.method", "smoother")
    s.update_setting("solver.discretization", "dg")
    if os.environ.get("DGB_GS_VARIANT"):                # tuning switches that also affect the assembly (51: DFMA kernel)
        _lib.load().dgb_set_kernel_path(100 + int(os.environ["DGB_GS_VARIANT"]))
    geo = Geometry(None, s, nodes=nodes(ni, nj, Pg))
    g = Grid(geo, ["u"]).initialize({"u": p}, None)
    DiscreteSystem(s).problem.assemble(g)
    asm_ms = None
    if os.environ.get("DGB_PROBE_ASSEMBLY"):            # re-run the operator assembly kernel alone, device-timed
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        _lib.call("dgb_assemble_poisson", g._h_tables, g.d_vol, g.d_face, g.d_area, g.Ni, g.Nj, 1.0, float(g.sigma),
                  g.flags, g.d_indptr, g.d_indices, g.d_data, g.d_minv, _lib.stream_ptr())
        ev1.record(); torch.cuda.synchronize()
        asm_ms = ev0.elapsed_time(ev1)
    g.release_geometry()
    L = _lib.load()
    st = _lib.stream_ptr()
    op = g.operator()
    N, b, nnzb = g.Ni * g.Nj, g.b, int(g.d_indices.numel())
    ab = bench.algorithmic_bytes(nnzb, N, b)
    x = torch.randn(N * b, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    part = torch.zeros(L.dgb_partials_len(), dtype=torch.float64, device="cuda")
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")

    def timed(fn):
        fn(); torch.cuda.synchronize()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        c.record(); torch.cuda.synchronize()
        return a.elapsed_time(c) / reps
    out = {"Ni": ni, "Nj": nj, "p": p, "b": b, "GB": {k: v / 1e9 for k, v in ab.items()}}
    if asm_ms is not None:
        out["assemble_poisson_ms"] = asm_ms
        out["assemble_elements_per_s"] = N / (asm_ms * 1e-3)
    calls = {
        "apply": (lambda: _lib.call("dgb_bsr_apply", op, x, y, st), ab["apply"]),
        "residual": (lambda: _lib.call("dgb_bsr_residual", op, g.d_rhs, x, None, part, ss, None, st), ab["residual"]),
        "gs_fwd": (lambda: _lib.call("dgb_block_gs_pass", op, g.d_rhs, x, 1, 0, None, st), ab["gs_pass"]),
        "gs_bwd": (lambda: _lib.call("dgb_block_gs_pass", op, g.d_rhs, x, -1, 0, None, st), ab["gs_pass"]),
        "entry_residual": (lambda: L.dgb_block_gs_entry_residual(__import__("ctypes").byref(op), _lib.ptr(g.d_rhs), _lib.ptr(x), 1,
                                                                 _lib.ptr(y), _lib.ptr(part), _lib.ptr(ss), st), ab["residual"]),
        "rec_residual": (lambda: (_lib.call("dgb_block_gs_pass", op, g.d_rhs, x, -1, 0, None, st),
                                  L.dgb_block_gs_residual_after_pass(__import__("ctypes").byref(op), _lib.ptr(x), -1, _lib.ptr(y),
                                                                     _lib.ptr(part), _lib.ptr(ss), None, st)), ab["residual"]),
        "redblack": (lambda: _lib.call("dgb_block_gs_pass", op, g.d_rhs, x, 1, 1, None, st), ab["gs_pass"]),
        "jacobi": (lambda: _lib.call("dgb_block_relax_sweep", op, g.d_rhs, x, y, 1.0, st), ab["gs_pass"]),
    }
    if os.environ.get("DGB_KSTREAM_MIN_B"):
        L.dgb_set_kernel_path(200 + int(os.environ["DGB_KSTREAM_MIN_B"]))
    if os.environ.get("DGB_CHAIN_CLUSTER"):
        L.dgb_set_kernel_path(400 + int(os.environ["DGB_CHAIN_CLUSTER"]))
    if os.environ.get("DGB_GS_VARIANT"):
        L.dgb_set_kernel_path(100 + int(os.environ["DGB_GS_VARIANT"]))
    for path, nm in ((0, "stream"), (1, "generic")):
        if only and nm != only.split(":")[0]:
            continue
        L.dgb_set_kernel_path(path)
        for k, (fn, nbytes) in calls.items():
            if only and ":" in only and k != only.split(":")[1]:
                continue
            if path == 1 and k.startswith("gs_") and max(ni, nj) > 600:
                continue
            ms = timed(fn)
            out[f"{nm}.{k}"] = {"ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1)}
    L.dgb_set_kernel_path(0)
    out["device_error"] = L.dgb_device_error(1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
