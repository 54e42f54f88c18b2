"""GPU tier: the CUDA path (through the C ABI / the reference-facing classes) against the golden
fixtures generated from the reference, and against the oracle on seeded/synthetic inputs."""
import os

import numpy as np
import pytest

from helpers import CASES, golden, grid_path, make_settings, oracle_hierarchy, rel_err

pytestmark = pytest.mark.gpu

MG_CASES = ["c1", "rect4_p1", "rect8_h24", "circ8_h24", "c2", "shipped", "rect8_direct", "rect8_prepost"]
HIST_RTOL, HIST_ATOL = 1e-10, 1e-12      # see tests/test_oracle_vs_golden.py


def build(case, **kw):
    from dg_multigrid_solver_b200.dgfem import DGFEM
    from dg_multigrid_solver_b200.grid import Geometry
    s = make_settings(case, **{k: v for k, v in kw.items() if k in ("gs_mode", "check_residual")})
    geo = Geometry(grid_path(case), s)
    mode = dict(solve_multigrid=True) if case["mode"] == "multigrid" else dict(solve_smoother=True, smoother="block_jacobi")
    return DGFEM(settings=s, geometry=geo, write_results=False, **mode)


@pytest.fixture(scope="module", params=MG_CASES)
def mg(request):
    name = request.param
    return name, golden(name), build(CASES[name])


def test_assembled_levels_match_reference(mg):
    name, g, d = mg
    assert len(d.grids) == int(g["nlevels"])
    assert list(d.solver.multigrid_type) == list(g["multigrid_type"])
    for k, grid in enumerate(d.grids):
        A = grid.BSR
        assert np.array_equal(A.indptr, g[f"L{k}_indptr"])            # bit-exact structure
        assert np.array_equal(A.indices, g[f"L{k}_indices"])
        assert A.blocksize == (int(g[f"L{k}_meta"][5]),) * 2
        if f"L{k}_data" in g.files:
            assert rel_err(A.data, g[f"L{k}_data"]) < 1e-12             # BASELINE.json: 1e-12 relative
            assert rel_err(grid.area_host().reshape(grid.Nj, grid.Ni).T, g[f"L{k}_area"]) < 1e-13
        else:
            assert abs(np.sqrt((A.data ** 2).sum()) - g[f"L{k}_data_fro"]) / g[f"L{k}_data_fro"] < 1e-12
        assert rel_err(grid.RHS, g[f"L{k}_RHS"]) < 1e-12
    for k, (R, P) in enumerate(zip(d.solver.restriction_operators, d.solver.prolongation_operators)):
        assert np.array_equal(R, g[f"R{k}"]) and np.array_equal(P, g[f"P{k}"])


def test_element_metrics_match_reference(mg):
    """a3 (dgfem/element.py:52-130,242-356): k_metrics' per-point arrays of the first and the last element of every
    level -- J, rx, sx, ry, sy and the point coordinates at the volume points, the face Jacobians and unit normals of
    the four faces -- against the values the reference's Element objects hold (h-coarsened levels included)."""
    name, g, d = mg
    if "L0_e00_J" not in g.files:
        pytest.skip("light fixture")
    for k, grid in enumerate(d.grids):
        vol = grid.d_vol.cpu().numpy()                 # [N][7][nq]: J, rx, sx, ry, sy, x, y
        face = grid.d_face.cpu().numpy()               # [N][4][8][nq1]: J_f, alpha, beta, x, y, nx, ny, -
        scale = max(np.abs(g[f"L{k}_e00_{key}"]).max() for key in ("rx", "sx", "ry", "sy"))
        for tag, m in (("e00", 0), ("eNN", grid.Ni * grid.Nj - 1)):
            for c, key in enumerate(("J", "rx", "sx", "ry", "sy")):
                ref = np.ravel(g[f"L{k}_{tag}_{key}"], order="F")      # point index r-fastest
                sc = np.abs(ref).max() if key == "J" else scale
                assert np.abs(vol[m, c] - ref).max() < 1e-12 * sc, (name, k, tag, key)
            for f, fname in enumerate(("imin", "imax", "jmin", "jmax")):
                ref = g[f"L{k}_{tag}_J_{fname}"]
                assert np.abs(face[m, f, 0] - ref).max() < 1e-12 * np.abs(ref).max(), (name, k, tag, fname)
        for f, fname in enumerate(("imin", "imax", "jmin", "jmax")):
            n = g[f"L{k}_e00_n_{fname}"]
            assert np.abs(face[0, f, 5] - n[:, 0]).max() < 1e-12 and np.abs(face[0, f, 6] - n[:, 1]).max() < 1e-12
        assert rel_err(vol[0, 5], np.ravel(g[f"L{k}_e00_xint"], order="F")) < 1e-13
        assert rel_err(vol[0, 6], np.ravel(g[f"L{k}_e00_yint"], order="F")) < 1e-13


def test_apply_and_smoother_calls(mg):
    from dg_multigrid_solver_b200.relaxation import Relaxation, bsr_apply
    name, g, d = mg
    fine = d.grids[-1]
    u0 = g["smooth_u0"]
    assert rel_err(bsr_apply(fine, u0), g["A_u0_fine"]) < 1e-13
    for direction in ("forward", "backward", "symmetric"):
        u = Relaxation.block_gauss_seidel_pyamg(grid=fine, RHS=fine.RHS, u=u0, direction=direction,
                                                max_iterations=1, omega=1.0)
        assert rel_err(u, g[f"bgs_pyamg_{direction}_1"]) < 1e-12
        assert rel_err(u, g[f"bgs_numpy_{direction}_1"]) < 1e-12        # pyamg-free pin (reference NumPy sweep)
        assert np.array_equal(u0, g["smooth_u0"])                      # inputs are never mutated
    u = Relaxation.block_gauss_seidel_pyamg(grid=fine, RHS=fine.RHS, u=u0, direction="symmetric", max_iterations=2)
    assert rel_err(u, g["bgs_pyamg_symmetric_2"]) < 1e-12
    # coarsest level, 10 symmetric sweeps from zero: exercises the device-side 1e-6 early exit
    u = Relaxation.block_gauss_seidel_pyamg(grid=d.grids[0], RHS=g["coarse_rhs"], u=None, direction="symmetric",
                                            max_iterations=10)
    assert rel_err(u, g["coarse_bgs_10"]) < 1e-11


def test_vcycle_and_residual_history(mg):
    name, g, d = mg
    fine = d.grids[-1]
    u1 = d.solver.multigrid_V_cycle(k=len(d.grids), RHS=fine.RHS, u=np.zeros_like(fine.RHS))
    assert rel_err(u1, g["u_after_1_vcycle"]) < 1e-11
    d.solver.residuals = []
    d.solve()
    hist = np.array(d.solver.residuals)
    assert len(hist) == len(g["residuals"])                              # identical V-cycle count
    assert np.allclose(hist, g["residuals"], rtol=HIST_RTOL, atol=HIST_ATOL)
    assert abs(d.L2_error_u - float(g["L2_error"])) < 1e-9 * max(1.0, float(g["L2_error"]))
    assert abs(d.L1_error_u - float(g["L1_error"])) < 1e-9 * max(1.0, float(g["L1_error"]))
    # the final residual sits ~1e-7 below the initial one: same rounding floor as the history's tail
    res0 = np.sqrt(np.mean(fine.RHS ** 2))
    assert abs(d.residual - float(g["final_residual"])) < HIST_ATOL * res0 + HIST_RTOL * float(g["final_residual"])


@pytest.mark.parametrize("name", ["smooth_rect4_p2", "smooth_circ4_p5"])
def test_smoother_only_runs(name):
    from dg_multigrid_solver_b200.relaxation import Relaxation
    g = golden(name)
    d = build(CASES[name])
    grid = d.grids[-1]
    assert rel_err(grid.BSR.data, g["L0_data"]) < 1e-12
    assert rel_err(grid.RHS, g["L0_RHS"]) < 1e-12
    for nm in ("block_jacobi", "block_gauss_seidel", "block_gauss_seidel_pyamg"):
        for its in (1, 2, 3, 100):
            u = getattr(Relaxation, nm)(grid, grid.RHS, max_iterations=its, direction="symmetric")
            assert rel_err(u, g[f"{nm}_{its}"]) < 1e-10, (nm, its)
    for nm in ("block_jacobi", "block_gauss_seidel"):
        u = getattr(Relaxation, nm)(grid, grid.RHS, max_iterations=3, omega=0.8)
        assert rel_err(u, g[f"{nm}_omega0p8_3"]) < 1e-12
    # `python -m ... -s --smoother X` path (Solver.solve_smoother: 100 symmetric iterations)
    d.settings.update_setting("solver.smoother", "block_gauss_seidel_pyamg")
    u = d.solver.solve()
    assert rel_err(u, g["block_gauss_seidel_pyamg_100"]) < 1e-10


@pytest.mark.parametrize("name", ["c1", "c2", "circ8_h24", "shipped"])
def test_streaming_kernels_match_generic_kernels(name):
    """Every level of the hierarchy (b = 4, 9, 16, 36; Dirichlet and O-grid stencils): the single-launch
    lexicographic kernels (k_gs_chain, k_gs_rows) against the generic per-anti-diagonal kernels, call by call."""
    import torch
    from dg_multigrid_solver_b200 import _lib
    from dg_multigrid_solver_b200.relaxation import Relaxation, bsr_apply, residual_norm
    d = build(CASES[name])
    L = _lib.load()
    rng = np.random.default_rng(11)
    try:
        for grid in d.grids:
            assert grid.stencil >= 0 and grid.d_gs is not None
            n = grid.d_rhs.numel()
            x = torch.from_numpy(rng.standard_normal(n)).cuda()
            rhs = grid.d_rhs
            out = {}
            for path in (1, 0):
                L.dgb_set_kernel_path(path)
                r = {}
                r["apply"] = bsr_apply(grid, x).clone()
                ss, res = residual_norm(grid, rhs, x, want_residual=True)
                r["resid"], r["sumsq"] = res.clone(), float(ss.item())
                for direction in ("forward", "backward", "symmetric"):
                    r["gs_" + direction] = Relaxation.block_gauss_seidel_pyamg(grid, rhs, x, direction, 1, 2).clone()
                    r["rb_" + direction] = Relaxation.block_gauss_seidel_pyamg(grid, rhs, x, direction, 1, 2,
                                                                             gs_mode="redblack").clone()
                r["jacobi1"] = Relaxation.block_jacobi(grid, rhs, x, None, 0.9, 1).clone()
                r["jacobi3"] = Relaxation.block_jacobi(grid, rhs, x, None, 0.9, 3).clone()
                r["bgs2"] = Relaxation.block_gauss_seidel(grid, rhs, x, None, 0.8, 2).clone()
                out[path] = r
            assert L.dgb_device_error(1) == 0
            for key, ref in out[1].items():
                got = out[0][key]
                if key == "sumsq":
                    assert abs(got - ref) <= 1e-12 * abs(ref)
                else:
                    scale = float(ref.abs().max())
                    assert float((got - ref).abs().max()) <= 1e-12 * scale, (name, grid.Ni, grid.b, key)
    finally:
        L.dgb_set_kernel_path(0)
        L.dgb_set_kernel_path(200 + 1000)


def test_block_diag_inverse_and_transfers():
    import torch
    from dg_multigrid_solver_b200 import _lib
    rng = np.random.default_rng(7)
    for b in (1, 4, 9, 16, 22, 25, 36):
        N = 37
        blocks = rng.standard_normal((N, b, b)) + 3 * np.eye(b)
        data = torch.from_numpy(blocks).cuda()
        indices = torch.arange(N, dtype=torch.int32, device="cuda")
        indptr = torch.arange(N + 1, dtype=torch.int32, device="cuda")
        dinv = torch.empty_like(data)
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.call("dgb_block_diag_inverse", data, indices, indptr, N, b, dinv, info, _lib.stream_ptr())
        assert int(info.item()) == 0
        ref = np.linalg.inv(blocks)
        err = np.abs(dinv.cpu().numpy() - ref).max(axis=(1, 2)) / np.abs(ref).max(axis=(1, 2))
        cond = np.linalg.cond(blocks)
        assert np.all(err < 1e-14 * cond + 1e-13)          # backward-stable up to the block's conditioning
    # singular block is flagged (the reference's pinv would silently pseudo-invert)
    blocks = rng.standard_normal((3, 4, 4)); blocks[1] = 0.0
    data = torch.from_numpy(blocks).cuda()
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.call("dgb_block_diag_inverse", data, torch.arange(3, dtype=torch.int32, device="cuda"),
              torch.arange(4, dtype=torch.int32, device="cuda"), 3, 4, torch.empty_like(data), info, _lib.stream_ptr())
    assert int(info.item()) == 2
    # transfers against the reference's literal reshape/transpose/einsum (solver.py:152-190),
    # square and non-square coarse grids
    from dgoracle.multigrid import h_restriction, p_restriction
    Rh, Ph = h_restriction()
    for (Nic, Njc) in ((2, 2), (5, 5), (3, 7), (8, 4)):
        fine = rng.standard_normal(4 * Nic * Njc * 4)
        ref = np.ravel(np.einsum("ij,kj->ki", Rh, fine.reshape((Nic, 2, Njc, 2, 4)).transpose((0, 2, 1, 3, 4)).reshape(-1, 16)))
        dR = torch.from_numpy(Rh.copy()).cuda(); dP = torch.from_numpy(np.ascontiguousarray(Ph)).cuda()
        df = torch.from_numpy(fine).cuda(); dc = torch.empty(Nic * Njc * 4, dtype=torch.float64, device="cuda")
        _lib.call("dgb_restrict", _lib.TRANSFER_H, dR, 4, 16, Nic, Njc, df, dc, _lib.stream_ptr())
        assert np.abs(dc.cpu().numpy() - ref).max() < 1e-14
        uc = rng.standard_normal(Nic * Njc * 4)
        v = np.einsum("ij,kj->ki", Ph, uc.reshape(-1, 4)).reshape((Nic, Njc, 2, 2, 4)).transpose((0, 2, 1, 3, 4))
        ref = fine + np.ravel(v)
        duc = torch.from_numpy(uc).cuda()
        _lib.call("dgb_prolong_add", _lib.TRANSFER_H, dP, 4, 16, Nic, Njc, duc, df, _lib.stream_ptr())
        assert np.abs(df.cpu().numpy() - ref).max() < 1e-14
    R = p_restriction(3, 5)
    fine = rng.standard_normal(11 * 36)
    dR = torch.from_numpy(R).cuda(); df = torch.from_numpy(fine).cuda()
    dc = torch.empty(11 * 16, dtype=torch.float64, device="cuda")
    _lib.call("dgb_restrict", _lib.TRANSFER_P, dR, 16, 36, 11, 1, df, dc, _lib.stream_ptr())
    assert np.array_equal(dc.cpu().numpy(), np.ravel(np.einsum("ij,kj->ki", R, fine.reshape(-1, 36))))


def test_redblack_mode_matches_its_oracle():
    """The 2-colour multicolour sweep has no counterpart in the reference; it is checked against
    the oracle's restatement of the same colouring, and as a whole solve."""
    from dgoracle import multigrid, relax
    from dg_multigrid_solver_b200.relaxation import Relaxation
    case = CASES["rect8_h24"]
    d = build(case, gs_mode="redblack")
    H = oracle_hierarchy(case)
    fine_o, fine = H.levels[-1], d.grids[-1]
    u0 = np.sin(0.37 * np.arange(fine.RHS.size)) * 0.1
    for direction in ("forward", "backward", "symmetric"):
        u = Relaxation.block_gauss_seidel_pyamg(grid=fine, RHS=fine.RHS, u=u0, direction=direction, max_iterations=2)
        uo = relax.red_black_gauss_seidel(fine_o.A, fine_o.RHS, fine_o.Ni, fine_o.Nj, u0, direction, 2)
        assert rel_err(u, uo) < 1e-12
    d.solve()
    uo, hist = multigrid.solve_multigrid(H, multigrid.Schedule(gs_mode="redblack"))
    assert len(d.solver.residuals) == len(hist)
    assert np.allclose(d.solver.residuals, hist, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", ["c1", "circ8_h24", "shipped"])
def test_colour_entry_is_residual_plus_first_colour(name):
    """dgb_block_gs_colour_entry == dgb_bsr_residual followed by dgb_block_gs_colour(first), on every level
    (b = 4, 9, 16, 36; Dirichlet and O-grid), either first colour, either row parity, with and without the freeze flag."""
    import torch
    from dg_multigrid_solver_b200 import _lib
    d = build(CASES[name])
    L = _lib.load()
    st = _lib.stream_ptr()
    rng = np.random.default_rng(5)
    part = torch.zeros(L.dgb_partials_len(), dtype=torch.float64, device="cuda")
    for grid in d.grids:
        op = grid.operator()
        n = grid.d_rhs.numel()
        x0 = torch.from_numpy(rng.standard_normal(n)).cuda()
        for first in (0, 1):
            for shift in (0, 1):
                ss_ref = torch.zeros(1, dtype=torch.float64, device="cuda")
                r_ref = torch.empty_like(x0)
                x_ref = x0.clone()
                _lib.call("dgb_bsr_residual", op, grid.d_rhs, x_ref, r_ref, part, ss_ref, None, st)
                _lib.call("dgb_block_gs_colour", op, grid.d_rhs, x_ref, first, shift, None, st)
                for frozen in (None, 0, 1):
                    flag = None if frozen is None else torch.full((1,), frozen, dtype=torch.int32, device="cuda")
                    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
                    r = torch.full_like(x0, float("nan"))
                    x = x0.clone()
                    _lib.call("dgb_block_gs_colour_entry", op, grid.d_rhs, x, r, first, shift, part, ss, flag, st)
                    assert float((r - r_ref).abs().max()) <= 1e-13 * float(r_ref.abs().max()), (name, grid.Ni, grid.b)
                    assert abs(float(ss.item()) - float(ss_ref.item())) <= 1e-13 * float(ss_ref.item())
                    assert torch.equal(x, x0 if frozen == 1 else x_ref), (name, grid.Ni, grid.b, first, shift, frozen)
                x = x0.clone()                                        # r is optional
                ss = torch.zeros(1, dtype=torch.float64, device="cuda")
                _lib.call("dgb_block_gs_colour_entry", op, grid.d_rhs, x, None, first, shift, part, ss, None, st)
                assert torch.equal(x, x_ref)
                assert abs(float(ss.item()) - float(ss_ref.item())) <= 1e-13 * float(ss_ref.item())
    assert L.dgb_device_error(1) == 0


def _synthetic(kind, Ni, Nj, P):
    from dgoracle import plot3d
    return plot3d.rectangle_nodes(Ni, Nj, P) if kind == "rect" else plot3d.circle_in_circle_nodes(Ni, Nj, P)


@pytest.mark.parametrize("kind,Ni,Nj,P,levels,factors,sigmul", [
    ("rect", 32, 32, 2, "2,1", "2,4,8", 1.0),       # deep h hierarchy incl. cf=8 (App. B.12 sampling)
    ("circ", 16, 16, 2, "2,1", "2,4", 2.0),          # curved O-grid
    ("rect", 24, 16, 2, "2,1", "2", 1.0),            # non-square: reproduces the reference's h-gather as is
    ("rect", 12, 12, 5, "5,3,1", "2", 1.0),          # high order
])
def test_synthetic_grids_against_oracle(kind, Ni, Nj, P, levels, factors, sigmul):
    """Seeded-free (the problem is deterministic) mid-size cases: GPU assembly + V-cycle vs oracle."""
    from dgoracle import multigrid
    from dg_multigrid_solver_b200.dgfem import DGFEM
    from dg_multigrid_solver_b200.grid import Geometry
    x, y = _synthetic(kind, Ni, Nj, P)
    case = dict(grid="synthetic.xyz", pg=P, pu=P, ogrid=(kind == "circ"), circ=(kind == "circ"), sigmul=sigmul,
                mode="multigrid", mg=dict(levels_u=levels, factors=factors))
    s = make_settings(case)
    geo = Geometry(None, s, nodes=(np.ascontiguousarray(x.T), np.ascontiguousarray(y.T)))
    d = DGFEM(settings=s, geometry=geo, solve_multigrid=True, write_results=False)
    H = oracle_hierarchy(case, x=x, y=y)
    assert len(d.grids) == len(H.levels)
    for grid, L in zip(d.grids, H.levels):
        A = grid.BSR
        assert np.array_equal(A.indptr, L.A.indptr) and np.array_equal(A.indices, L.A.indices)
        assert rel_err(A.data, L.A.data) < 1e-12
        assert rel_err(grid.RHS, L.RHS) < 1e-12
    fine = d.grids[-1]
    u1 = d.solver.multigrid_V_cycle(k=len(d.grids), RHS=fine.RHS, u=np.zeros_like(fine.RHS))
    uo = multigrid.v_cycle(H, multigrid.Schedule(), len(H.levels), H.levels[-1].RHS, np.zeros_like(fine.RHS))
    assert rel_err(u1, uo) < 1e-10
    if Ni != Nj:
        return      # the reference's h-gather scrambles non-square grids (App. A.7): no convergence claim
    d.solve()
    _, hist = multigrid.solve_multigrid(H, multigrid.Schedule())
    assert len(d.solver.residuals) == len(hist)
    assert np.allclose(d.solver.residuals, hist, rtol=1e-9, atol=1e-12)


def test_large_grid_properties():
    """512 x 512, p=2 (2.4 MDOF): size-independent properties -- linearity of the apply, the
    residual of the assembled system at the exact discrete solution of a manufactured vector, and a
    V-cycle that contracts the residual."""
    import torch
    from dg_multigrid_solver_b200 import _lib
    from dg_multigrid_solver_b200.dgfem import DGFEM
    from dg_multigrid_solver_b200.grid import Geometry
    from dg_multigrid_solver_b200.relaxation import bsr_apply, residual_norm
    x, y = _synthetic("rect", 512, 512, 2)
    case = dict(grid="synthetic.xyz", pg=2, pu=2, ogrid=False, circ=False, sigmul=1.0, mode="multigrid",
                mg=dict(levels_u="2,1", factors="2,4,8,16,32,64,128"))
    s = make_settings(case)
    geo = Geometry(None, s, nodes=(np.ascontiguousarray(x.T), np.ascontiguousarray(y.T)))
    d = DGFEM(settings=s, geometry=geo, solve_multigrid=True, write_results=False)
    fine = d.grids[-1]
    n = fine.d_rhs.numel()
    assert n == 512 * 512 * 9 and int(fine.d_indptr[-1].item()) == 512 * 512 + 2 * 511 * 512 * 2
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    lhs = bsr_apply(fine, 2.0 * a - 3.0 * b)
    rhs = 2.0 * bsr_apply(fine, a) - 3.0 * bsr_apply(fine, b)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 1e-13
    # r = (A a) - A a == 0 to rounding, and the fused norm agrees with torch's
    Aa = bsr_apply(fine, a)
    sumsq, r = residual_norm(fine, Aa, a, want_residual=True)
    assert float(r.abs().max() / Aa.abs().max()) < 1e-14
    sumsq2, r2 = residual_norm(fine, fine.d_rhs, a, want_residual=True)
    assert abs(float(sumsq2.item()) - float((r2 * r2).sum().item())) < 1e-12 * float(sumsq2.item())
    # uniform rectangle: all interior rows hold the same five blocks
    data = fine.d_data
    ip = fine.d_indptr.cpu().numpy().astype(np.int64)
    m0, m1 = 512 * 200 + 100, 512 * 300 + 317
    assert float((data[ip[m0]:ip[m0 + 1]] - data[ip[m1]:ip[m1 + 1]]).abs().max()) < 1e-9 * float(data[ip[m0]].abs().max())
    # V-cycles contract the residual
    d.solver.residuals = []
    u = d.solver.solve_multigrid(levels=len(d.grids), RHS=fine.d_rhs, u=torch.zeros_like(fine.d_rhs), tol=1e-6, max_cycles=40)
    hist = d.solver.residuals
    assert hist[-1] < 1e-6 and len(hist) <= 40
    assert all(h2 < h1 for h1, h2 in zip(hist[:-1], hist[1:]))
    _lib.require_cuda().cuda.synchronize()


@pytest.mark.parametrize("nproc", [1, 2])
def test_slab_partitioned_vcycle(nproc):
    """DistributedSolver (parallel.py) through torchrun: one rank (the slab code path with the gathered coarse
    hierarchy) and two ranks (ghost rows, halo exchange, the three sweep orders) -- over NCCL when two GPUs are
    visible, else both ranks on the one GPU with gloo and host-staged messages."""
    import subprocess
    import sys
    import torch
    from helpers import REPO
    env = dict(os.environ)
    if torch.cuda.device_count() < nproc:
        env["DGB_MGPU_ONE_DEVICE"] = "1"        # both ranks on cuda:0, gloo + host-staged messages
    port = 29500 + os.getpid() % 2000 + nproc
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(REPO, "tests", "mgpu_check.py"), "64"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "[mgpu_check] PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("name", ["stokes_rect4", "stokes_circ4"])
def test_stokes_local_order_assembly(name):
    """BASELINE.json configs[4] at fixture size: Stokes local-order operator + RHS + apply against the
    reference's output (rectangle and O-grid annulus, p_u=2, p_p=1, 22x22 blocks)."""
    import copy
    from helpers import case_params
    from dg_multigrid_solver_b200.dgfem import DGFEM
    from dg_multigrid_solver_b200.grid import Geometry
    from dg_multigrid_solver_b200.relaxation import Relaxation, bsr_apply
    from dg_multigrid_solver_b200.settings import Settings
    case = CASES[name]
    g = golden(name)
    prm = copy.deepcopy(case_params(case))
    prm["problem"]["type"] = "Stokes"
    prm["problem"]["include pressure BC"] = False
    prm["solution"]["p"]["polynomial degree"] = case["pp"]
    prm["solution"]["ordering"] = "local"
    s = Settings(prm)
    d = DGFEM(settings=s, geometry=Geometry(grid_path(case), s), solve_direct=True, write_results=False)
    grid = d.grids[-1]
    A = grid.BSR
    assert A.blocksize == (int(g["L0_meta"][7]),) * 2 == (22, 22)
    assert np.array_equal(A.indptr, g["L0_indptr"]) and np.array_equal(A.indices, g["L0_indices"])
    assert rel_err(A.data, g["L0_data"]) < 1e-12
    assert rel_err(grid.RHS, g["L0_RHS"]) < 1e-12
    assert rel_err(bsr_apply(grid, g["smooth_u0"]), g["A_u0_fine"]) < 1e-13


@pytest.mark.parametrize("Ni,Nj,P", [
    (3, 5, 1), (1, 7, 1), (7, 1, 2), (2, 2, 1), (17, 9, 2), (40, 33, 1), (33, 40, 2), (8, 8, 3), (6, 7, 4),
    (64, 64, 1), (130, 70, 2), (70, 130, 1), (5, 4, 5), (1, 3, 5), (24, 37, 5),
])
def test_chained_gauss_seidel_kernel(Ni, Nj, P):
    """k_gs_helper + k_gs_chain (dgb_chain.cu) on ragged grid shapes -- partial bands, single rows/columns,
    more bands than one CTA holds (mailbox hand-over) -- against the oracle's restatement of pyamg's
    block_gauss_seidel and against the anti-diagonal wavefront of the generic kernels."""
    import torch
    from dgoracle import native
    from dg_multigrid_solver_b200 import _lib
    from dg_multigrid_solver_b200.dgfem import DGFEM
    from dg_multigrid_solver_b200.grid import Geometry
    from dg_multigrid_solver_b200.relaxation import Relaxation
    x, y = _synthetic("rect", Ni, Nj, P)
    case = dict(grid="synthetic.xyz", pg=P, pu=P, ogrid=False, circ=False, sigmul=1.0)
    s = make_settings(case)
    geo = Geometry(None, s, nodes=(np.ascontiguousarray(x.T), np.ascontiguousarray(y.T)))
    L = _lib.load()
    L.dgb_set_kernel_path(300 + 31)         # chained kernel for every block size it supports
    try:
        d = DGFEM(settings=s, geometry=geo, solve_smoother=True, smoother="block_gauss_seidel_pyamg",
                  write_results=False)
        grid = d.grids[-1]
        assert grid.d_chain is not None, "chained kernel not selected"
        _chained_gs_checks(grid, Ni, Nj, L)
    finally:
        L.dgb_set_kernel_path(300 + CHAIN_MASK_DEFAULT)


CHAIN_MASK_DEFAULT = 31


def _chained_gs_checks(grid, Ni, Nj, L):
    import torch
    from dgoracle import native
    from dg_multigrid_solver_b200.relaxation import Relaxation
    A = grid.BSR
    b = A.blocksize[0]
    rng = np.random.default_rng(Ni * 1000 + Nj)
    rhs = rng.standard_normal(A.shape[0])
    x0 = rng.standard_normal(A.shape[0])
    Dinv = grid.d_dinv.cpu().numpy()
    # 3 symmetric iterations: passes 2..6 take their c from the previous pass's chain (no helper launch)
    for direction, sweeps in (("forward", [1]), ("backward", [-1]), ("symmetric", [1, -1]),
                              ("symmetric", [1, -1] * 3), ("forward", [1] * 2)):
        ref = x0.copy()
        N = Ni * Nj
        for sw in sweeps:
            span = (0, N, 1) if sw > 0 else (N - 1, -1, -1)
            native.block_gauss_seidel(A.indptr, A.indices, A.data, ref, rhs, Dinv, *span, b)
        got = {}
        try:
            for path in (0, 1):
                L.dgb_set_kernel_path(path)
                iters = len(sweeps) // (2 if direction == "symmetric" else 1)
                got[path] = Relaxation.block_gauss_seidel_pyamg(grid, rhs, x0, direction, 1, iters)
        finally:
            L.dgb_set_kernel_path(0)
        assert L.dgb_device_error(1) == 0
        scale = np.abs(ref).max()
        assert np.abs(got[1] - ref).max() <= 1e-12 * scale
        assert np.abs(got[0] - ref).max() <= 1e-12 * scale, (direction, np.abs(got[0] - ref).max() / scale)
    # the mailbox is all-sentinel again after the passes
    assert bool((grid.d_mailbox.view(torch.int64) == -1).all())
    # residual right after a pass, evaluated from the records the pass left (k_residual_rec), against rhs - A x
    from dg_multigrid_solver_b200 import _lib
    from dg_multigrid_solver_b200.relaxation import bsr_apply, residual_norm
    import ctypes
    op = grid.operator()
    st = _lib.stream_ptr()
    d_rhs = torch.from_numpy(rhs).cuda()
    part = torch.zeros(L.dgb_partials_len(), dtype=torch.float64, device="cuda")
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    for sweeps in ([1], [-1], [1, -1], [1, -1, 1]):
        d_x = torch.from_numpy(x0).cuda()
        prev = 0
        for sw in sweeps:
            _lib.call("dgb_block_gs_pass_seq", op, d_rhs, d_x, sw, prev, None, st)
            prev = sw
        r_rec = torch.full_like(d_x, float("nan"))
        rc = L.dgb_block_gs_residual_after_pass(ctypes.byref(op), _lib.ptr(d_x), prev, _lib.ptr(r_rec), _lib.ptr(part),
                                                _lib.ptr(ss), None, st)
        if rc == _lib.UNSUPPORTED:
            assert b not in (4, 9, 16)
            break
        assert rc == 0
        ss_rec = float(ss.item())
        ss_ref, r_ref = residual_norm(grid, d_rhs, d_x, want_residual=True)
        scale = float(torch.maximum(d_rhs.abs().max(), bsr_apply(grid, d_x).abs().max()))
        assert float((r_rec - r_ref).abs().max()) <= 1e-12 * scale, (sweeps, float((r_rec - r_ref).abs().max()) / scale)
        assert abs(ss_rec - float(ss_ref.item())) <= 1e-11 * max(float(ss_ref.item()), scale * scale * 1e-6)
    assert L.dgb_device_error(1) == 0


def test_vcycle_result_copied_out_on_second_stream():
    """Solver.multigrid_V_cycle(..., out=pinned buffer): u leaves on a second stream from the event dgb_vcycle
    records after the last sweep; same bits as the plain path."""
    import torch
    d = build(CASES["rect8_h24"])
    fine = d.grids[-1]
    k = len(d.grids)
    u_plain = d.solver.multigrid_V_cycle(k=k, RHS=fine.RHS, u=np.zeros_like(fine.RHS))
    out = torch.empty(fine.RHS.size, dtype=torch.float64).pin_memory()
    rhs_h = torch.from_numpy(fine.RHS.copy()).pin_memory()
    u_h = torch.zeros(fine.RHS.size, dtype=torch.float64).pin_memory()
    for _ in range(3):
        got = d.solver.multigrid_V_cycle(k=k, RHS=rhs_h, u=u_h, out=out)
        assert got is out
        assert np.array_equal(out.numpy(), u_plain)


@pytest.mark.parametrize("Ni,Nj,P", [(3, 4, 1), (8, 5, 1), (17, 9, 2), (12, 6, 3), (40, 30, 1), (64, 20, 2), (3, 2, 5),
                                     (20, 35, 5)])
def test_chained_gauss_seidel_kernel_ogrid(Ni, Nj, P):
    """The same checks on O-grids (periodic in i): the last element of every row also meets the row's first
    element across the wrap (side array of wrap blocks, fill/drain path of the chain kernel)."""
    from dg_multigrid_solver_b200 import _lib
    from dg_multigrid_solver_b200.dgfem import DGFEM
    from dg_multigrid_solver_b200.grid import Geometry
    x, y = _synthetic("circ", Ni, Nj, P)
    case = dict(grid="synthetic.xyz", pg=P, pu=P, ogrid=True, circ=True, sigmul=2.0)
    s = make_settings(case)
    geo = Geometry(None, s, nodes=(np.ascontiguousarray(x.T), np.ascontiguousarray(y.T)))
    L = _lib.load()
    L.dgb_set_kernel_path(300 + 31)
    try:
        d = DGFEM(settings=s, geometry=geo, solve_smoother=True, smoother="block_gauss_seidel_pyamg",
                  write_results=False)
        grid = d.grids[-1]
        assert grid.d_chain is not None, "chained kernel not selected"
        _chained_gs_checks(grid, Ni, Nj, L)
    finally:
        L.dgb_set_kernel_path(300 + CHAIN_MASK_DEFAULT)


def test_direct_solve_and_export():
    """`-d` (Solver.solve_directly, dgfem/solver.py:56-59: spsolve on the assembled system) through the dense
    inverse kernels, against scipy on the GPU-assembled matrix; the fused post-processing kernel
    (dgb_nodal_error) against the plain evaluation; the `.vts` export of DGFEM.solve."""
    import scipy.sparse.linalg as splin
    import torch
    from dg_multigrid_solver_b200.dgfem import DGFEM
    from dg_multigrid_solver_b200.grid import Geometry
    from dg_multigrid_solver_b200.mms import PoissonMMS
    case = CASES["rect8_h24"]
    s = make_settings(case)
    s.update_setting("visualization.export", True)
    d = DGFEM(settings=s, geometry=Geometry(grid_path(case), s), solve_direct=True, write_results=True)
    u = d.solve()
    fine = d.grids[-1]
    ref = splin.spsolve(fine.BSR.tocsr(), fine.RHS)
    assert rel_err(u, ref) < 1e-11
    assert d.residual_normalized < 1e-12
    # post-processing: V_DOF_grid @ u_e at the element nodes, exact solution there, L1 / L2 (dgfem.py:203-221)
    T = fine.tables
    un = u.reshape(-1, T.b) @ np.asarray(T.V_DOF_grid).T
    assert rel_err(d.u_nodal.cpu().numpy(), un) < 1e-13
    xn, yn = (t.cpu().numpy() for t in d.geometry.device_nodes())
    N1, Pg = fine.P_grid + 1, fine.P_grid
    ex = PoissonMMS(s).solution(torch.from_numpy(xn).cuda(), torch.from_numpy(yn).cuda()).cpu().numpy()
    delta = []
    for e in range(fine.Ni * fine.Nj):
        i, j = e % fine.Ni, e // fine.Ni
        blk = ex[j * Pg:j * Pg + N1, i * Pg:i * Pg + N1]          # [c, a]
        delta.append(un[e] - blk.reshape(-1))
    delta = np.concatenate(delta)
    assert abs(d.L1_error_u - np.abs(delta).mean()) < 1e-13
    assert abs(d.L2_error_u - np.sqrt((delta ** 2).mean())) < 1e-13
    vts = d.solution_visualization_filepath + ".vts"
    assert os.path.exists(vts) and os.path.getsize(vts) > fine.Ni * fine.Nj * N1 * N1 * 8 * 6


def test_midsize_vcycle_against_oracle():
    """256 x 256, p=2 (590 k DOFs, levels p 2,1 + h 2..64): the largest case the oracle finishes in seconds.  Operator
    apply (a16), one symmetric smoother call with its residual tests (a14) and one V-cycle (a13) against the oracle
    on the same inputs -- the bench compares the 512^2 V-cycle the same way (bench.py `parity`)."""
    import bench
    from dgoracle import multigrid, plot3d, relax
    from dg_multigrid_solver_b200.dgfem import DGFEM
    from dg_multigrid_solver_b200.grid import Geometry
    from dg_multigrid_solver_b200.relaxation import Relaxation, bsr_apply
    from dg_multigrid_solver_b200.settings import Settings
    n, p = 256, 2
    s = Settings(bench.make_params(n, p, "lexicographic", True))
    d = DGFEM(settings=s, geometry=Geometry(None, s, nodes=bench.rectangle_nodes_file_order(n, p)), solve_multigrid=True,
              write_results=False)
    x, y = plot3d.rectangle_nodes(n, n, p)
    H = multigrid.Hierarchy(x, y, p, [1, p], bench.h_factors(n), exact_u=bench.MMS_U, rhs_all_levels=False)
    fine, fo = d.grids[-1], H.levels[-1]
    assert np.array_equal(fine.BSR.indices, fo.A.indices)
    assert rel_err(fine.RHS, fo.RHS) < 1e-12
    u0 = np.sin(0.37 * np.arange(fo.RHS.size)) * 0.1
    assert rel_err(bsr_apply(fine, u0), fo.A @ u0) < 1e-12
    u_s = Relaxation.block_gauss_seidel_pyamg(grid=fine, RHS=fine.RHS, u=u0, direction="symmetric", max_iterations=2)
    assert rel_err(u_s, relax.block_gauss_seidel_pyamg(fo.A, fo.RHS, u0, "symmetric", 1, 2)) < 1e-11
    u1 = d.solver.multigrid_V_cycle(k=len(d.grids), RHS=fine.RHS, u=np.zeros_like(fine.RHS))
    uo = multigrid.v_cycle(H, multigrid.Schedule(), len(H.levels), fo.RHS, np.zeros_like(fo.RHS))
    assert rel_err(u1, uo) < 1e-10


@pytest.mark.parametrize("name,gs_mode", [("rect8_h24", "lexicographic"), ("c2", "lexicographic"), ("rect8_h24", "redblack")])
def test_solve_loop_hands_its_residual_to_the_next_cycle(name, gs_mode, monkeypatch):
    """solve_multigrid's residual after a cycle (dgfem/solver.py:119) is the next pre-smoother's entry residual
    (relaxation.py:202): with the chained kernel the loop evaluates it once (dgb_block_gs_entry_residual +
    dgb_vcycle_ex(DGB_VCYCLE_ENTRY_PRIMED)).  Same history and iterate as the plain loop; the 2-colour mode, whose
    smoother opens differently, falls back to the plain cycle."""
    out = {}
    for prime in ("0", "1"):
        monkeypatch.setenv("DGB_SOLVE_PRIME", prime)
        d = build(CASES[name], gs_mode=gs_mode)
        fine = d.grids[-1]
        u = d.solver.solve_multigrid(len(d.grids), fine.RHS, np.zeros_like(fine.RHS))
        out[prime] = (np.array(d.solver.residuals), u, d.solver.primed_cycles)
    h0, u0, n0 = out["0"]
    h1, u1, n1 = out["1"]
    assert n0 == 0 and n1 == (len(h1) - 1 if gs_mode == "lexicographic" else 0)
    assert len(h0) == len(h1)
    assert np.allclose(h1, h0, rtol=1e-12, atol=1e-14)
    assert rel_err(u1, u0) < 1e-12
    if gs_mode == "lexicographic":
        assert np.allclose(h1, golden(name)["residuals"], rtol=HIST_RTOL, atol=HIST_ATOL)


def test_stokes_global_order_and_distributive_gauss_seidel():
    """global-order Stokes blocks (dgfem/discrete_system.py:416-745) and
    Relaxation.distributive_gauss_seidel with the `lsq` splitting (dgfem/relaxation.py:221-283) -- the
    `-s --smoother distributive_gauss_seidel` run -- against the reference's output (315 outer
    iterations, first residuals 0.76194 / 0.46150 / 0.26073)."""
    import copy
    from helpers import case_params
    from dg_multigrid_solver_b200.dgfem import DGFEM
    from dg_multigrid_solver_b200.grid import Geometry
    from dg_multigrid_solver_b200.relaxation import Relaxation
    from dg_multigrid_solver_b200.settings import Settings
    from dg_multigrid_solver_b200.stokes import Stokes
    case = CASES["stokes_dgs_rect4"]
    g = golden("stokes_dgs_rect4")
    prm = copy.deepcopy(case_params(case))
    prm["problem"]["type"] = "Stokes"
    prm["problem"]["include pressure BC"] = False
    prm["solution"]["p"]["polynomial degree"] = case["pp"]
    prm["solution"]["ordering"] = "global"
    s = Settings(prm)
    d = DGFEM(settings=s, geometry=Geometry(grid_path(case), s), solve_smoother=True,
              smoother="distributive_gauss_seidel", write_results=False)
    grid = d.grids[-1]
    # the regrouped blocks: SciPy's block sizes, stored order and values as the reference has them
    for nm in ("A", "D", "G"):
        B = getattr(grid, "BSR_block_" + nm).to_scipy()
        assert tuple(B.shape) + tuple(B.blocksize) == tuple(int(v) for v in g[f"{nm}_shape"])
        assert np.array_equal(B.indptr, g[f"{nm}_indptr"]) and np.array_equal(B.indices, g[f"{nm}_indices"])
        assert rel_err(B.data, g[f"{nm}_data"]) < 1e-12
    assert tuple(grid.BSR.blocksize) == tuple(int(v) for v in g["BSR_blocksize"])
    assert rel_err(grid.RHS, g["RHS"]) < 1e-12
    DG = Stokes(s).block_DG(grid).to_scipy()
    assert np.array_equal(DG.indptr, g["DG_indptr"]) and np.array_equal(DG.indices, g["DG_indices"])
    assert rel_err(DG.data, g["DG_data"]) < 1e-12
    for its in (1, 3):
        u = Relaxation.distributive_gauss_seidel(grid, grid.RHS, max_iterations=its, splitting="lsq", settings=s)
        assert rel_err(u, g[f"dgs_u_{its}"]) < 1e-11
    u = d.solver.solve()                                    # Solver.solve_smoother -> 1e6 iterations, lsq
    hist = np.array(Relaxation.last_residuals)
    ref = g["dgs_residuals"]
    assert len(hist) == len(ref) == 316                     # 315 outer iterations + the converged one
    assert np.allclose(hist, ref, rtol=1e-8, atol=1e-12)    # 315 accumulated iterations: 1.5e-10 already oracle vs reference
    assert rel_err(u, g["dgs_u_final"]) < 1e-10


def test_cli_entry_points(tmp_path, monkeypatch):
    """`python -m dg_multigrid_solver_b200 -m | -s --smoother X | -d` from a directory laid out like the reference's
    (input/paramfile.yml + grid folder): the shipped paramfile as-is reproduces the reference's 8 V-cycles and error
    norms, the smoother and direct runs reach the reference's residuals."""
    import shutil
    from helpers import REPO
    from dg_multigrid_solver_b200.__main__ import main
    (tmp_path / "input").mkdir()
    shutil.copy(os.path.join(REPO, "input", "paramfile.yml"), tmp_path / "input" / "paramfile.yml")
    shutil.copy(os.path.join(REPO, "input", "Rectangle_8X8_nPoly5.xyz"), tmp_path / "input" / "Rectangle_8X8_nPoly5.xyz")
    monkeypatch.chdir(tmp_path)
    g = golden("shipped")
    d = main(["-m", "--silent"])
    assert d is not None
    assert len(d.solver.residuals) == len(g["residuals"]) == 9
    assert np.allclose(d.solver.residuals, g["residuals"], rtol=HIST_RTOL, atol=HIST_ATOL)
    assert abs(d.L2_error_u - float(g["L2_error"])) < 1e-9
    assert os.path.exists(os.path.join(d.results_dir, "summary.txt"))
    d = main(["-s", "--smoother", "block_gauss_seidel_pyamg", "--silent"])
    assert d is not None and 0.0 < d.residual_normalized < 1.0
    d = main(["-d", "--silent"])
    assert d is not None and d.residual_normalized < 1e-11
