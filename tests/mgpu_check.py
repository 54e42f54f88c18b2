#!/usr/bin/env python3
"""Multi-GPU parity check, launched with torchrun (one rank per GPU):
  red-black across slabs      == single-GPU red-black          (same iteration; norms reduced in another order)
  exact lexicographic pipeline == single-GPU lexicographic      (the reference's sweep order)
  slab-lexicographic           == the oracle's restatement of that iteration (dgoracle.relax.slab_gs_pass:
                                  lexicographic inside a slab, halo frozen at the previous pass), same cycle count
Histories are compared at rtol 1e-10 + atol 1e-12 (the single-GPU tolerance, tests/test_gpu_parity.py).
usage: torchrun --nproc-per-node N tests/mgpu_check.py [N_elements]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "oracle")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    if os.environ.get("DGB_MGPU_ONE_DEVICE") == "1":
        # all ranks share cuda:0 and talk through gloo + host staging (parallel._host_staged): lets the round-end
        # single-GPU test box exercise the slab code with ghost rows
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from dg_multigrid_solver_b200.dgfem import DGFEM
    from dg_multigrid_solver_b200.grid import Geometry
    from dg_multigrid_solver_b200.parallel import build_distributed
    from dg_multigrid_solver_b200.settings import Settings
    xn, yn = bench.rectangle_nodes_file_order(n, 2)
    ok = True
    for mode, single_mode in (("redblack", "redblack"), ("lexicographic", "lexicographic"), ("slab_lexicographic", None)):
        s = Settings(bench.make_params(n, 2, mode, True))
        s.update_setting("solver.method", "multigrid")
        ds = build_distributed(s, xn, yn, world, rank, gs_mode=mode)
        transport = "native (peer memory, dgb_vcycle_slab)" if ds.native else "torch.distributed"
        hist = np.array(ds.solve(tol=1e-6, max_cycles=60))
        ds.check_native_error()
        ref = None
        if rank == 0 and single_mode is not None:
            s1 = Settings(bench.make_params(n, 2, single_mode, True))
            d = DGFEM(settings=s1, geometry=Geometry(None, s1, nodes=(xn, yn)), solve_multigrid=True, write_results=False)
            d.solver.solve()
            ref = np.array(d.solver.residuals)
        if rank == 0:
            msg = f"[mgpu_check] world={world} n={n} mode={mode} [{transport}]: {len(hist) - 1} cycles, final {hist[-1]:.3e}"
        if rank == 0 and single_mode is None:
            # the oracle's slab iteration on the same grid (CPU; the reference's assembly restated in NumPy)
            from dgoracle import multigrid, plot3d
            x, y = plot3d.rectangle_nodes(n, n, 2)
            H = multigrid.Hierarchy(x, y, 2, [2, 1], bench.h_factors(n))
            _, ref = multigrid.solve_multigrid(H, multigrid.Schedule(gs_mode="slab_lexicographic", world=world, min_rows=8))
        if rank == 0:
            if ref is not None:
                same = len(ref) == len(hist) and np.allclose(hist, ref, rtol=1e-10, atol=1e-12)
                msg += f" | {'single-GPU' if single_mode else 'oracle'} {len(ref) - 1} cycles, max rel diff " \
                       f"{np.max(np.abs(hist[:len(ref)] - ref[:len(hist)]) / ref[:len(hist)]):.2e} -> {'OK' if same else 'MISMATCH'}"
                ok &= bool(same)
            print(msg, flush=True)
        ds.close()
        del ds
        torch.cuda.empty_cache()
    flag = torch.tensor([1.0 if ok else 0.0], device="cpu" if dist.get_backend() == "gloo" else "cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("[mgpu_check] PASS" if ok else "[mgpu_check] FAIL", flush=True)
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
