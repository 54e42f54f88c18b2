#!/usr/bin/env python3
"""Time Solver.solve_multigrid (the `-m` path: residual test + V-cycle per iteration) on the bench workload, with the
loop's residual handed to the next cycle's pre-smoother (default) and without (DGB_SOLVE_PRIME=0).
usage: solve_time.py [N=2048] [P=2]"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from dg_multigrid_solver_b200.dgfem import DGFEM  # noqa: E402
from dg_multigrid_solver_b200.grid import Geometry  # noqa: E402
from dg_multigrid_solver_b200.settings import Settings  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    s = Settings(bench.make_params(n, p, "lexicographic", True))
    d = DGFEM(settings=s, geometry=Geometry(None, s, nodes=bench.rectangle_nodes_file_order(n, p)), solve_multigrid=True,
              write_results=False)
    for g in d.grids:
        g.release_geometry()
    fine = d.grids[-1]
    nlev = len(d.grids)
    out = {"N": n, "p": p, "dofs": int(fine.d_rhs.numel())}
    u0 = torch.zeros_like(fine.d_rhs)
    for rep in range(2):                      # the second round runs on the captured graphs
        for prime in ("0", "1"):
            os.environ["DGB_SOLVE_PRIME"] = prime
            d.solver.residuals = []
            d.solver.primed_cycles = 0
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            u = d.solver.solve_multigrid(nlev, fine.d_rhs, u0)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            cycles = len(d.solver.residuals) - 1
            out[f"prime{prime}_round{rep}"] = {"s": dt, "cycles": cycles, "ms_per_cycle": 1e3 * dt / max(cycles, 1),
                                               "primed_cycles": d.solver.primed_cycles,
                                               "final_residual": d.solver.residuals[-1], "u_norm": float(u.norm())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
