#!/usr/bin/env python3
"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE.  Runs only in the build container (the reference does not
travel to the GPU box); the .npz files it writes are committed under
tests/golden/ together with this script.

Recipe: the reference is executed in place from
/root/reference with an import-shim directory (oracle/shims) for the five
third-party modules missing offline; `pyamg` is restated (oracle/shims/pyamg),
everything else is the reference's own code on numpy 2.3 / scipy 1.18.

Usage:  python oracle/gen_golden.py            # all cases -> tests/golden/
        python oracle/gen_golden.py --case c1   # one case (worker mode)
"""
import argparse
import copy
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
GOLD = os.path.join(REPO, "tests", "golden")


def _mg(levels_u, factors, smoother="block_gauss_seidel_pyamg", coarse="smoother", post=None):
    """post: optional overrides of every coarsening's post smoother (smoother / direction / iterations /
    relaxation factor) -- the reference resolves pre and post independently (dgfem/solver.py:143-147,196)."""
    return {"levels_u": levels_u, "factors": factors, "smoother": smoother, "coarse": coarse, "post": post}


# name -> dict(grid file, P_grid, p_u, O-grid, circular, sigma multiplier, mode, multigrid settings)
CASES = {
    # BASELINE.json configs[0]
    "c1": dict(grid="Rectangle_4X4_nPoly2.xyz", pg=2, pu=2, ogrid=False, circ=False, sigmul=1.0,
               mode="multigrid", mg=_mg("2,1", 2), dump="full"),
    # BASELINE.json configs[1] (sigma multiplier 2: the shipped 1 diverges)
    "c2": dict(grid="CircleInCircle_8X8_nPoly5.xyz", pg=5, pu=5, ogrid=True, circ=True, sigmul=2.0,
               mode="multigrid", mg=_mg("5,3,1", 2), dump="full"),
    # the shipped paramfile as-is
    "shipped": dict(grid="Rectangle_8X8_nPoly5.xyz", pg=5, pu=5, ogrid=False, circ=False, sigmul=1.0,
                    mode="multigrid", mg=_mg("5,3,1", 2), dump="light"),
    # deeper h hierarchy (cf = 2,4) on a rectangle and on a curved O-grid
    "rect8_h24": dict(grid="Rectangle_8X8_nPoly2.xyz", pg=2, pu=2, ogrid=False, circ=False, sigmul=1.0,
                      mode="multigrid", mg=_mg("2,1", "2,4"), dump="full"),
    "circ8_h24": dict(grid="CircleInCircle_8X8_nPoly2.xyz", pg=2, pu=2, ogrid=True, circ=True, sigmul=2.0,
                      mode="multigrid", mg=_mg("2,1", "2,4"), dump="full"),
    # coarse grid solver: direct (dgfem/solver.py:56-59,199-200) on the same hierarchy
    "rect8_direct": dict(grid="Rectangle_8X8_nPoly2.xyz", pg=2, pu=2, ogrid=False, circ=False, sigmul=1.0,
                         mode="multigrid", mg=_mg("2,1", "2,4", coarse="direct"), dump="light"),
    # pre smoother != post smoother (different plugin, direction, iteration count and relaxation factor)
    "rect8_prepost": dict(grid="Rectangle_8X8_nPoly2.xyz", pg=2, pu=2, ogrid=False, circ=False, sigmul=1.0,
                          mode="multigrid",
                          mg=_mg("2,1", "2", post={"smoother": "block_gauss_seidel", "direction": "forward",
                                                   "iterations": 2, "relaxation factor": 0.9}), dump="light"),
    # p=1 geometry
    "rect4_p1": dict(grid="Rectangle_4X4_nPoly1.xyz", pg=1, pu=1, ogrid=False, circ=False, sigmul=1.0,
                     mode="multigrid", mg=_mg("1", 2), dump="full"),
    # smoother-only runs (python -m dgfem -s --smoother X), SURVEY 3.4 / App. C.4
    "smooth_circ4_p5": dict(grid="CircleInCircle_4X4_nPoly5.xyz", pg=5, pu=5, ogrid=True, circ=True, sigmul=2.0,
                            mode="smoother", dump="full"),
    "smooth_rect4_p2": dict(grid="Rectangle_4X4_nPoly2.xyz", pg=2, pu=2, ogrid=False, circ=False, sigmul=1.0,
                            mode="smoother", dump="full"),
    # Stokes local-order assembly + apply (BASELINE.json configs[4] at fixture size)
    "stokes_rect4": dict(grid="Rectangle_4X4_nPoly2.xyz", pg=2, pu=2, pp=1, ogrid=False, circ=False, sigmul=1.0,
                         mode="stokes", dump="full"),
    "stokes_circ4": dict(grid="CircleInCircle_4X4_nPoly2.xyz", pg=2, pu=2, pp=1, ogrid=True, circ=True, sigmul=2.0,
                         mode="stokes", dump="full"),
    # Stokes, global ordering + distributive Gauss-Seidel (lsq splitting), the `-s --smoother
    # distributive_gauss_seidel` run of dgfem/solver.py:61-66 (315 outer iterations)
    "stokes_dgs_rect4": dict(grid="Rectangle_4X4_nPoly2.xyz", pg=2, pu=2, pp=1, ogrid=False, circ=False, sigmul=1.0,
                             mode="stokes_dgs", dump="full"),
}


def make_workdir(case):
    w = tempfile.mkdtemp(prefix="dgref_")
    for d in ("input", "logs", "results", "cache/grid", "cache/discrete_system",
              "postprocessing/pickles/relaxation"):
        os.makedirs(os.path.join(w, d), exist_ok=True)
    shutil.copy(os.path.join(REF, "input", "paramfile.yml"), os.path.join(w, "input", "paramfile.yml"))
    shutil.copy(os.path.join(REF, "input", case["grid"]), os.path.join(w, "input", case["grid"]))
    return w


def worker(name):
    import numpy as np
    case = CASES[name]
    w = make_workdir(case)
    os.chdir(w)
    sys.path[:0] = [os.path.join(HERE, "shims"), REF]
    from input import params  # reference: input/__init__.py:4-7
    params = copy.deepcopy(params)
    params["grid"]["filename"] = case["grid"]
    params["grid"]["polynomial degree"] = case["pg"]
    params["grid"]["O grid"] = case["ogrid"]
    params["grid"]["circular"] = case["circ"]
    params["solution"]["u"]["polynomial degree"] = case["pu"]
    params["problem"]["SIP penalty parameter multiplier"] = case["sigmul"]
    params["visualization"]["automatically open paraview"] = False
    params["visualization"]["export"] = False
    params["logging"]["loglevel"] = "ERROR"
    out = {}
    from dgfem.settings import Settings
    from dgfem.dgfem import DGFEM

    def dump_grid(prefix, g, full=True):
        B = g.BSR
        out[prefix + "indptr"] = np.asarray(B.indptr, dtype=np.int32)
        out[prefix + "indices"] = np.asarray(B.indices, dtype=np.int32)
        out[prefix + "data_fro"] = np.sqrt(np.sum(np.asarray(B.data) ** 2))
        out[prefix + "data_sum"] = np.sum(np.asarray(B.data))
        out[prefix + "RHS"] = np.asarray(g.RHS)
        out[prefix + "meta"] = np.array([g.Ni, g.Nj, g.P_grid, g.P_sol["u"], g.N_int["u"], B.blocksize[0]], dtype=np.int64)
        out[prefix + "sigma"] = np.float64(g.sigma)
        if full:
            out[prefix + "data"] = np.asarray(B.data)
            A = np.array([[g.elements[i, j].A for j in range(g.Nj)] for i in range(g.Ni)])
            out[prefix + "area"] = A
            e = g.elements[0, 0]
            for key in ("J", "rx", "sx", "ry", "sy"):
                out[prefix + f"e00_{key}"] = np.asarray(e.gt[key]["e"]["u"])
                for f in ("imin", "imax", "jmin", "jmax"):
                    out[prefix + f"e00_{key}_{f}"] = np.asarray(e.gt[key][f]["u"])
            for f in ("imin", "imax", "jmin", "jmax"):
                out[prefix + f"e00_n_{f}"] = np.asarray(e.gt["n"][f]["u"])
            out[prefix + "e00_xint"] = np.asarray(e.xy_int["xy_int"][0]["u"])
            out[prefix + "e00_yint"] = np.asarray(e.xy_int["xy_int"][1]["u"])
            e = g.elements[g.Ni - 1, g.Nj - 1]
            for key in ("J", "rx", "sx", "ry", "sy"):
                out[prefix + f"eNN_{key}"] = np.asarray(e.gt[key]["e"]["u"])
                for f in ("imin", "imax", "jmin", "jmax"):
                    out[prefix + f"eNN_{key}_{f}"] = np.asarray(e.gt[key][f]["u"])
            out[prefix + "Minv_00"] = np.asarray(g.elements[0, 0].inv_mass_matrix)
            # one interior-ish face, all four blocks (face.py:115-127)
            fi = g.faces_i[min(1, g.Ni), 0]
            LL, LR, RL, RR = fi.compute_momentum_laplace_SIP_terms("Poisson")
            out[prefix + "face_i10"] = np.stack([LL, LR, RL, RR])
            out[prefix + "face_i10_hF"] = np.float64(fi.h_F)
            fj = g.faces_j[0, 0]
            LL, LR, RL, RR = fj.compute_momentum_laplace_SIP_terms("Poisson")
            out[prefix + "face_j00"] = np.stack([LL, LR, RL, RR])
            out[prefix + "K_00"] = np.asarray(g.elements[0, 0].compute_momentum_laplace_volume_integral("Poisson"))
            out[prefix + "M_00"] = np.asarray(g.elements[0, 0].compute_mass_matrix())

    if case["mode"] == "multigrid":
        mg = case["mg"]
        pc = params["solver"]["multigrid"]["polynomial coarsening"]
        gc = params["solver"]["multigrid"]["geometric coarsening"]
        pc["levels"]["u"] = mg["levels_u"]
        gc["coarsening factors"] = mg["factors"]
        for blk in (pc, gc):
            for s in ("pre smoother", "post smoother"):
                blk[s]["smoother"] = mg["smoother"]
            if mg.get("post"):
                blk["post smoother"].update(mg["post"])
        params["solver"]["multigrid"]["coarse grid solver"] = mg["coarse"]
        s = Settings(params)
        d = DGFEM(settings=s, solve_multigrid=True)
        out["nlevels"] = np.int64(len(d.grids))
        for k, g in enumerate(d.grids):
            dump_grid(f"L{k}_", g, full=(case["dump"] == "full"))
        for k, (R, P) in enumerate(zip(d.solver.restriction_operators, d.solver.prolongation_operators)):
            out[f"R{k}"] = np.asarray(R)
            out[f"P{k}"] = np.asarray(P)
        out["multigrid_type"] = np.array(d.solver.multigrid_type)
        # one V-cycle from u=0 (solver.py:141-207) and the full solve (solver.py:114-139)
        fine = d.grids[-1]
        u1 = d.solver.multigrid_V_cycle(k=len(d.grids), RHS=fine.RHS, u=np.zeros_like(fine.RHS))
        out["u_after_1_vcycle"] = np.asarray(u1)
        d.solver.residuals = []
        d.solve()
        out["residuals"] = np.asarray(d.solver.residuals)
        out["L1_error"] = np.float64(d.L1_error_u)
        out["L2_error"] = np.float64(d.L2_error_u)
        out["final_residual"] = np.float64(d.residual)
        # single smoother calls on the fine level, u0 = a deterministic non-trivial vector
        from dgfem.relaxation import Relaxation
        n = fine.RHS.size
        u0 = np.sin(0.37 * np.arange(n)) * 0.1
        out["smooth_u0"] = u0
        for direction in ("forward", "backward", "symmetric"):
            out[f"bgs_pyamg_{direction}_1"] = Relaxation.block_gauss_seidel_pyamg(
                grid=fine, RHS=fine.RHS, u=u0, direction=direction, max_iterations=1, omega=1.0)
        out["bgs_pyamg_symmetric_2"] = Relaxation.block_gauss_seidel_pyamg(
            grid=fine, RHS=fine.RHS, u=u0, direction="symmetric", max_iterations=2, omega=1.0)
        # coarsest level: 10 symmetric sweeps from zero on a synthetic rhs (may trigger the early exit)
        g0 = d.grids[0]
        rhs0 = np.cos(0.11 * np.arange(g0.RHS.size))
        out["coarse_rhs"] = rhs0
        out["coarse_bgs_10"] = Relaxation.block_gauss_seidel_pyamg(
            grid=g0, RHS=rhs0, u=np.zeros_like(rhs0), direction="symmetric", max_iterations=10, omega=1.0)
        out["A_u0_fine"] = fine.BSR @ u0
        # pyamg-free pin of the BACKWARD pass: the reference's own NumPy block_gauss_seidel (forward,
        # dgfem/relaxation.py:170-195) on the block-reversed system P A P^T equals a backward pass on A
        import types
        import scipy.sparse as sp
        B = fine.BSR
        b = B.blocksize[0]
        N = B.shape[0] // b
        perm = (np.arange(N)[::-1, None] * b + np.arange(b)[None, :]).ravel()      # block order reversed
        Arev = sp.bsr_array(sp.csr_array(B.tocsr())[perm][:, perm], blocksize=(b, b))
        Arev.sort_indices()
        stand_in = types.SimpleNamespace(BSR=Arev, BSR_D=None, BSR_E=None, BSR_F=None)
        ub = Relaxation.block_gauss_seidel(stand_in, fine.RHS[perm], u=u0[perm], max_iterations=1, omega=1)
        out["bgs_numpy_backward_1"] = ub[perm]
        fine.BSR_E = fine.BSR_D = fine.BSR_F = None
        uf = Relaxation.block_gauss_seidel(fine, fine.RHS, u=u0, max_iterations=1, omega=1)
        out["bgs_numpy_forward_1"] = uf
        stand_in = types.SimpleNamespace(BSR=Arev, BSR_D=None, BSR_E=None, BSR_F=None)
        out["bgs_numpy_symmetric_1"] = Relaxation.block_gauss_seidel(stand_in, fine.RHS[perm], u=uf[perm],
                                                                     max_iterations=1, omega=1)[perm]
    elif case["mode"] == "smoother":
        from dgfem.relaxation import Relaxation
        s = Settings(params)
        d = DGFEM(settings=s, solve_smoother=True, smoother="block_jacobi")
        g = d.grids[-1]
        dump_grid("L0_", g, full=True)
        out["nlevels"] = np.int64(1)
        for nm in ("block_jacobi", "block_gauss_seidel", "block_gauss_seidel_pyamg"):
            for its in (1, 2, 3, 100):
                g.BSR_E = g.BSR_D = g.BSR_F = None
                u = getattr(Relaxation, nm)(g, g.RHS, max_iterations=its, direction="symmetric")
                out[f"{nm}_{its}"] = np.asarray(u)
        # relaxation factor != 1 (block_jacobi / block_gauss_seidel honour omega, relaxation.py:148,194)
        for nm in ("block_jacobi", "block_gauss_seidel"):
            g.BSR_E = g.BSR_D = g.BSR_F = None
            out[f"{nm}_omega0p8_3"] = np.asarray(getattr(Relaxation, nm)(g, g.RHS, max_iterations=3, omega=0.8))
    elif case["mode"] == "stokes":
        params["problem"]["type"] = "Stokes"
        params["solution"]["p"]["polynomial degree"] = case["pp"]
        params["solution"]["ordering"] = "local"
        s = Settings(params)
        d = DGFEM(settings=s, solve_direct=True)
        g = d.grids[-1]
        B = g.BSR
        out["nlevels"] = np.int64(1)
        out["L0_indptr"] = np.asarray(B.indptr, dtype=np.int32)
        out["L0_indices"] = np.asarray(B.indices, dtype=np.int32)
        out["L0_data"] = np.asarray(B.data)
        out["L0_RHS"] = np.asarray(g.RHS)
        out["L0_meta"] = np.array([g.Ni, g.Nj, g.P_grid, g.P_sol["u"], g.P_sol["p"], g.N_int["u"], g.N_int["p"],
                                   B.blocksize[0]], dtype=np.int64)
        out["L0_sigma"] = np.float64(g.sigma)
        out["L0_gamma"] = np.float64(g.gamma)
        out["L0_Epsilon"] = np.float64(g.Epsilon)
        out["exact_p_mean"] = np.float64(d.exact_p_mean)
        n = g.RHS.size
        u0 = np.sin(0.37 * np.arange(n)) * 0.1
        out["smooth_u0"] = u0
        out["A_u0_fine"] = B @ u0
    elif case["mode"] == "stokes_dgs":
        import glob
        import pickle
        from dgfem.relaxation import Relaxation
        params["problem"]["type"] = "Stokes"
        params["solution"]["p"]["polynomial degree"] = case["pp"]
        params["solution"]["ordering"] = "global"
        s = Settings(params)
        d = DGFEM(settings=s, solve_smoother=True, smoother="distributive_gauss_seidel")
        g = d.grids[-1]
        out["nlevels"] = np.int64(1)
        for nm in ("A", "D", "G"):
            B = getattr(g, "BSR_block_" + nm)
            out[f"{nm}_indptr"], out[f"{nm}_indices"] = np.asarray(B.indptr, np.int32), np.asarray(B.indices, np.int32)
            out[f"{nm}_data"], out[f"{nm}_shape"] = np.asarray(B.data), np.array(B.shape + B.blocksize, np.int64)
        out["BSR_blocksize"] = np.array(g.BSR.blocksize, np.int64)
        out["RHS"] = np.asarray(g.RHS)
        # one and three outer iterations, then the full run (solver.py:63: max_iterations=1000000, splitting='lsq')
        for its in (1, 3):
            out[f"dgs_u_{its}"] = np.asarray(Relaxation.distributive_gauss_seidel(
                g, g.RHS, max_iterations=its, splitting="lsq", settings=s))
        DG = g.BSR_block_DG
        out["DG_indptr"], out["DG_indices"] = np.asarray(DG.indptr, np.int32), np.asarray(DG.indices, np.int32)
        out["DG_data"], out["DG_shape"] = np.asarray(DG.data), np.array(DG.shape + DG.blocksize, np.int64)
        u = d.solver.solve()
        out["dgs_u_final"] = np.asarray(u)
        f = glob.glob("postprocessing/pickles/relaxation/*.pkl")
        out["dgs_residuals"] = np.asarray(pickle.load(open(f[0], "rb")), dtype=np.float64)
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), **out)
    shutil.rmtree(w, ignore_errors=True)
    print(f"[gen_golden] {name}: wrote {len(out)} arrays")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--only", default=None, help="comma separated subset")
    a = ap.parse_args()
    if a.case:
        worker(a.case)
        return
    names = a.only.split(",") if a.only else list(CASES)
    for name in names:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name])
        if r.returncode != 0:
            print(f"[gen_golden] case {name} FAILED", file=sys.stderr)
            sys.exit(1)


if __name__ == "__main__":
    main()
