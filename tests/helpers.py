"""Shared test helpers: fixture cases (the same table oracle/gen_golden.py generated the goldens
from), settings construction, oracle hierarchies."""
import copy
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, "tests", "golden")
GRIDS = os.path.join(GOLD, "grids")

from gen_golden import CASES  # noqa: E402  (oracle/gen_golden.py)


def golden(name):
    return np.load(os.path.join(GOLD, f"{name}.npz"))


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def base_params():
    from dg_multigrid_solver_b200.settings import load_params
    return load_params(os.path.join(REPO, "input", "paramfile.yml"))


def case_params(case, gs_mode="lexicographic", check_residual=True):
    p = copy.deepcopy(base_params())
    p["grid"]["folder"] = GRIDS
    p["grid"]["filename"] = case["grid"]
    p["grid"]["polynomial degree"] = case["pg"]
    p["grid"]["O grid"] = case["ogrid"]
    p["grid"]["circular"] = case["circ"]
    p["solution"]["u"]["polynomial degree"] = case["pu"]
    p["problem"]["SIP penalty parameter multiplier"] = case["sigmul"]
    p["solver"]["b200"]["gs mode"] = gs_mode
    p["solver"]["b200"]["check residual"] = check_residual
    mg = case.get("mg")
    if mg:
        pc = p["solver"]["multigrid"]["polynomial coarsening"]
        gc = p["solver"]["multigrid"]["geometric coarsening"]
        pc["levels"]["u"] = mg["levels_u"]
        gc["coarsening factors"] = mg["factors"]
        for blk in (pc, gc):                       # the same edits oracle/gen_golden.py applied to the reference
            if mg.get("smoother"):
                for s in ("pre smoother", "post smoother"):
                    blk[s]["smoother"] = mg["smoother"]
            if mg.get("post"):
                blk["post smoother"].update(mg["post"])
        if mg.get("coarse"):
            p["solver"]["multigrid"]["coarse grid solver"] = mg["coarse"]
    return p


def oracle_schedule(case, **kw):
    """dgoracle.multigrid.Schedule of a fixture case (coarse solver, independent post smoother)."""
    from dgoracle import multigrid
    mg = case.get("mg") or {}
    post = mg.get("post") or {}
    args = dict(coarse_solver=mg.get("coarse", "smoother"))
    if mg.get("smoother"):
        args["smoother"] = mg["smoother"]
    if post:
        args.update(post_smoother=post.get("smoother"), post_direction=post.get("direction"),
                    post_omega=post.get("relaxation factor"))
        if "iterations" in post:
            args["post"] = post["iterations"]
    args.update(kw)
    return multigrid.Schedule(**args)


def make_settings(case, **kw):
    from dg_multigrid_solver_b200.settings import Settings
    return Settings(case_params(case, **kw))


def grid_path(case):
    return os.path.join(GRIDS, case["grid"])


def oracle_hierarchy(case, x=None, y=None, **kw):
    from dgoracle import multigrid, plot3d
    if x is None:
        x, y, _, _ = plot3d.read_plot3d(grid_path(case), case["pg"])
    mg = case.get("mg") or {"levels_u": str(case["pu"]), "factors": ""}
    pl = [int(v) for v in str(mg["levels_u"]).split(",")]
    hf = [int(v) for v in str(mg["factors"]).split(",") if v != ""]
    return multigrid.Hierarchy(x, y, case["pg"], pl, hf, sigma_mult=case["sigmul"], O_grid=case["ogrid"], **kw)
