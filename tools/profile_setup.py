#!/usr/bin/env python3
"""Where the set-up of the bench workload spends its host time: cProfile of DGFEM(...) (grids, tables, metrics, assembly,
RHS, smoother streams) -- run 1 cold, run 2 warm (modules loaded, allocator warm), run 3 with CUDA_LAUNCH_BLOCKING-like
synchronisation after every library call (DGB_SYNC_CALLS=1) so that device time lands on the call that launched it.
usage: profile_setup.py [N=2048] [P=2]"""
import cProfile
import io
import os
import pstats
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO]
import torch  # noqa: E402

import bench  # noqa: E402
from dg_multigrid_solver_b200 import _lib  # noqa: E402
from dg_multigrid_solver_b200.dgfem import DGFEM  # noqa: E402
from dg_multigrid_solver_b200.grid import Geometry  # noqa: E402
from dg_multigrid_solver_b200.settings import Settings  # noqa: E402


def once(n, p, prof=None):
    s = Settings(bench.make_params(n, p, "lexicographic", True))
    nodes = bench.rectangle_nodes_file_order(n, p)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if prof:
        prof.enable()
    d = DGFEM(settings=s, geometry=Geometry(None, s, nodes=nodes), solve_multigrid=True, write_results=False)
    d.solver.hierarchy()
    torch.cuda.synchronize()
    if prof:
        prof.disable()
    dt = time.perf_counter() - t0
    tm = dict(d.timings)
    del d
    torch.cuda.empty_cache()
    return dt, tm


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    _lib.load()
    for label in ("cold", "warm", "warm, synchronised calls"):
        if "synchronised" in label:
            os.environ["DGB_SYNC_CALLS"] = "1"
        prof = cProfile.Profile()
        dt, tm = once(n, p, prof)
        print(f"=== {label}: {dt:.3f} s  timings {({k: round(v, 3) for k, v in tm.items()})}")
        if label != "cold":
            out = io.StringIO()
            pstats.Stats(prof, stream=out).sort_stats("cumulative").print_stats(38)
            txt = out.getvalue()
            print("\n".join(l[:170] for l in txt.splitlines()[4:]))


if __name__ == "__main__":
    main()
