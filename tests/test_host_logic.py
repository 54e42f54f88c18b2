"""CPU tier: host-side logic of the product package (settings, tables, Plot3D, CLI, hierarchy
bookkeeping) -- everything that does not need a device."""
import numpy as np
import pytest

from helpers import CASES, GRIDS, base_params, golden, grid_path, make_settings


def test_settings_attribute_view_and_updates():
    from dg_multigrid_solver_b200.settings import Settings
    s = Settings(base_params())
    assert s.grid.polynomial_degree == 5 and s.grid.O_grid is False
    assert s.solver.multigrid.polynomial_coarsening.pre_smoother.smoother == "block_gauss_seidel_pyamg"
    assert s.solution.u.integration_polynomial_degree_factor == 3
    s.update_settings({"grid_file": "x.xyz", "p_grid": 2, "smoother": "block_jacobi"})
    assert s.grid.filename == "x.xyz" and s.grid.polynomial_degree == 2
    assert s.solver.smoother == "block_jacobi" and s.solver.discretization == "dg"
    s.update_setting("solver.method", "multigrid")
    assert s.solver.method == "multigrid"
    assert s.get("solver.b200.gs_mode") == "lexicographic"
    assert s.get("solver.nothing.here", 7) == 7


def test_tables_match_oracle_tables():
    from dg_multigrid_solver_b200.tables import Tables, h_restriction, p_restriction
    from dgoracle import multigrid as om
    from dgoracle import tables as ot
    for Pg, p in [(1, 1), (2, 1), (2, 2), (5, 3), (5, 5)]:
        T, O = Tables(Pg, p), ot.LevelTables(Pg, p)
        assert T.nq1 == O.N_int and T.b == O.b
        assert np.array_equal(T.V, O.V_DOF_int) and np.array_equal(T.Vr, O.Vr_DOF_int)
        assert np.array_equal(T.Vs, O.Vs_DOF_int) and np.array_equal(T.w2, np.ravel(O.w_int_2D, order="F"))
        for k, nm in enumerate(("iL", "iR", "jL", "jR")):
            assert np.array_equal(T.Vf[k], O.V_face[nm]) and np.array_equal(T.Vrf[k], O.Vr_face[nm])
        assert np.allclose(T.GX, O.L_int, rtol=0, atol=1e-14)
        assert np.allclose(T.GR, O.Dr_int @ O.L_gg, rtol=0, atol=1e-13)
        for k, f in enumerate(("imin", "imax", "jmin", "jmax")):
            assert np.allclose(T.FS[k], O.Ds_face[f] @ O.L_gg, rtol=0, atol=1e-13)
        assert np.array_equal(T.V_DOF_grid, O.V_DOF_grid)
    for pc, pf in [(1, 2), (1, 3), (3, 5), (0, 1), (2, 5)]:
        assert np.array_equal(p_restriction(pc, pf), om.p_restriction(pc, pf))
    assert np.array_equal(h_restriction()[0], om.h_restriction()[0])
    g = golden("c2")
    assert np.array_equal(h_restriction()[0], g["R0"]) and np.array_equal(p_restriction(3, 5), g["R2"])


def test_mms_fields_in_slices_equal_whole_evaluation():
    """Device evaluation of the manufactured solution runs in slices along the element axis (mms.Field.CHUNK): same
    values entry by entry as one evaluation, and as the NumPy evaluation (dgfem/dgfem.py:410-483) to rounding."""
    import torch
    from dg_multigrid_solver_b200 import mms
    f = mms.Field("-2*sin(pi*x)**2*sin(pi*y)*cos(pi*y)")
    g = torch.Generator().manual_seed(3)
    vol = torch.rand(301, 7, 9, dtype=torch.float64, generator=g) * 2 - 1
    X, Y = vol[:, 5, :], vol[:, 6, :]                      # strided views, as assemble_RHS_Poisson passes them
    whole = f(X, Y)
    old = mms.Field.CHUNK
    try:
        mms.Field.CHUNK = 64
        sliced = f(X, Y)
        face = torch.rand(50, 4, 8, 3, dtype=torch.float64, generator=g)
        sliced4 = f(face[:, :, 3, :], face[:, :, 4, :])
    finally:
        mms.Field.CHUNK = old
    assert sliced.shape == whole.shape and sliced.is_contiguous() and torch.equal(sliced, whole)
    assert np.abs(whole.numpy() - f(X.numpy(), Y.numpy())).max() < 1e-14
    assert torch.equal(sliced4, f(face[:, :, 3, :], face[:, :, 4, :]))
    assert torch.equal(mms.Field("3.5")(X, Y), torch.full_like(X, 3.5))


def test_coarse_tables_follow_reference_point_location():
    """cf = 2, 4, 8: sub-element offsets and local coordinates (element.py:273-310, App. B.12)."""
    from dg_multigrid_solver_b200.tables import Tables
    from dgoracle import geometry as og
    from dgoracle import tables as ot
    for cf in (2, 4, 8, 16, 64, 512):           # 512: the coarsest level of the 2048^2 bench hierarchy
        T, O = Tables(2, 1, cf=cf), ot.LevelTables(2, 1)
        for (iR, iS, m, n, r, s) in og.coarse_point_map(O, cf):
            q = iR + T.nq1 * iS
            assert tuple(T.sub_vol[q]) == (m, n)
            L, Dr, Ds = O.point_ops(r, s)
            assert np.allclose(T.GX[q], L[0], atol=1e-14)
            assert np.allclose(T.GR[q], cf * (Dr @ O.L_gg)[0], atol=1e-12)
        # B.12: for cf >= 8 the "imin" face is sampled from an interior fine element -- the one that holds the
        # first volume point
        assert T.sub_face[0, 0, 0] == int((T.r_int[0] + 1.0) / (2.0 / cf)) and (T.sub_face[0, 0, 0] > 0) == (cf >= 8)


def test_plot3d_reader_matches_oracle_reader():
    from dg_multigrid_solver_b200.grid import Geometry
    from dgoracle import plot3d
    for name in ("c1", "c2"):
        case = CASES[name]
        geo = Geometry(grid_path(case), make_settings(case))
        x, y, Ni, Nj = plot3d.read_plot3d(grid_path(case), case["pg"])
        assert (geo.Ni, geo.Nj) == (Ni, Nj)
        assert np.array_equal(geo.x, x) and np.array_equal(geo.y, y)
        assert geo.xn.shape == (geo.jl, geo.il) and geo.xn[1, 2] == x[2, 1]


def test_plot3d_rejects_open_o_grid():
    from dg_multigrid_solver_b200.grid import Geometry
    case = dict(CASES["c1"], ogrid=True)
    with pytest.raises(ValueError):
        Geometry(grid_path(CASES["c1"]), make_settings(case))


def test_synthetic_grid_rules_reproduce_shipped_files():
    from dgoracle import plot3d
    x, y, Ni, Nj = plot3d.read_plot3d(f"{GRIDS}/Rectangle_8X8_nPoly2.xyz", 2)
    xs, ys = plot3d.rectangle_nodes(Ni, Nj, 2)
    assert np.array_equal(x, xs) and np.array_equal(y, ys)          # byte-identical rule
    x, y, Ni, Nj = plot3d.read_plot3d(f"{GRIDS}/CircleInCircle_8X8_nPoly5.xyz", 5)
    xs, ys = plot3d.circle_in_circle_nodes(Ni, Nj, 5)
    assert np.abs(x - xs).max() < 5e-16 and np.abs(y - ys).max() < 5e-16


def test_cli_parser_matches_reference_flags():
    from dg_multigrid_solver_b200.__main__ import build_parser
    p = build_parser()
    a = p.parse_args(["-m"])
    assert a.solve_multigrid and not a.solve_smoother
    a = p.parse_args(["-s", "--smoother", "block_jacobi", "-f", "g.xyz", "--p-grid", "2", "-v"])
    assert a.solve_smoother and a.smoother == "block_jacobi" and a.grid_file == "g.xyz" and a.p_grid == 2
    with pytest.raises(SystemExit):
        p.parse_args(["-m", "-d"])                                # mutually exclusive solver flags
    with pytest.raises(SystemExit):
        p.parse_args([])                                          # one solver flag is required


def test_mms_fields_host():
    from dg_multigrid_solver_b200.mms import PoissonMMS
    from dgoracle.mms import PoissonMMS as OMMS
    s = make_settings(CASES["c1"])
    m, o = PoissonMMS(s), OMMS(s.problem.exact_solution.u, 1.0)
    X, Y = np.meshgrid(np.linspace(-1, 1, 7), np.linspace(-1, 1, 5))
    assert np.array_equal(m.solution(X, Y), o.solution(X, Y))
    assert np.array_equal(m.source(X, Y), o.source(X, Y))


def test_full_size_config_generators_follow_the_shipped_grid_rule():
    """bench.py generates the C3/C4 grids with its own code (the product side may not import the oracle); they
    must agree with the oracle's restatement of the shipped grids' rule."""
    from dgoracle import plot3d
    import bench
    x, y = plot3d.rectangle_nodes(6, 6, 2)
    xn, yn = bench.rectangle_nodes_file_order(6, 2)
    assert np.array_equal(xn, x.T) and np.array_equal(yn, y.T)
    xc, yc = plot3d.circle_in_circle_nodes(8, 8, 3)
    xg, yg = bench.circle_nodes_file_order(8, 3)
    assert np.allclose(xg, xc.T, rtol=0, atol=1e-15) and np.allclose(yg, yc.T, rtol=0, atol=1e-15)


def test_plot3d_writer_round_trip(tmp_path):
    """visualization.write_plot3d writes what Geometry.read (dgfem/grid.py:26-63) reads; a shipped grid is
    reproduced byte for byte."""
    from dg_multigrid_solver_b200.grid import Geometry
    from dg_multigrid_solver_b200.visualization import write_plot3d
    for name in ("c1", "c2"):
        case = CASES[name]
        geo = Geometry(grid_path(case), make_settings(case))
        out = write_plot3d(str(tmp_path / f"{name}.xyz"), geo.xn, geo.yn)
        assert open(out, "rb").read() == open(grid_path(case), "rb").read()
        geo2 = Geometry(out, make_settings(case))
        assert np.array_equal(geo2.xn, geo.xn) and np.array_equal(geo2.yn, geo.yn)


def test_vts_export_layout(tmp_path, monkeypatch):
    """elements_to_vtk: the reference's point layout (dgfem/visualization.py:67-117: element nodes side by side,
    Ni*N1 x Nj*N1 points, i fastest) in a VTK XML StructuredGrid with raw appended data."""
    import re
    import struct
    from dg_multigrid_solver_b200.visualization import element_node_arrays, elements_to_vtk, nodal_to_elements
    monkeypatch.chdir(tmp_path)
    Ni, Nj, Pg = 3, 2, 2
    N1 = Pg + 1
    xl = np.linspace(0.0, 3.0, Ni * Pg + 1)
    yl = np.linspace(0.0, 1.0, Nj * Pg + 1)
    xn, yn = np.meshgrid(xl, yl)                        # file order [jl][il]
    x_el, y_el = element_node_arrays(xn, yn, Ni, Nj, Pg)
    assert x_el.shape == (Ni, Nj, N1, N1)
    assert x_el[2, 1, 1, 0] == xl[2 * Pg + 1] and y_el[2, 1, 1, 2] == yl[1 * Pg + 2]
    u_nodal = (np.arange(Ni * Nj)[:, None] * 100 + np.arange(N1 * N1)[None, :]).astype(float)   # [N, ng]
    phi = nodal_to_elements(u_nodal, Ni, Nj, Pg)
    assert phi[1, 1, 2, 0] == (1 * Ni + 1) * 100 + 2 and phi[0, 1, 0, 1] == (1 * Ni) * 100 + N1
    path = elements_to_vtk("sol", x_el, y_el, "Poisson", {"phi": phi, "abs_error_phi": np.abs(phi)})
    raw = open(path, "rb").read()
    head, tail = raw.split(b'<AppendedData encoding="raw">\n_')
    nx, ny = Ni * N1, Nj * N1
    assert f'WholeExtent="0 {nx - 1} 0 {ny - 1} 0 0"'.encode() in head
    offs = [int(v) for v in re.findall(rb'offset="(\d+)"', head)]
    names = re.findall(rb'Name="(\w+)"', head)
    assert names == [b"phi", b"abs_error_phi", b"points"] and sorted(offs)[0] == 0
    npts = struct.unpack("<Q", tail[:8])[0]
    assert npts == nx * ny * 3 * 8
    pts = np.frombuffer(tail[8:8 + npts], dtype="<f8").reshape(ny, nx, 3)
    assert pts[0, 1, 0] == xl[1] and pts[N1, 0, 1] == yl[Pg] and np.all(pts[:, :, 2] == 0)
    o = offs[names.index(b"phi")]
    nb = struct.unpack("<Q", tail[o:o + 8])[0]
    v = np.frombuffer(tail[o + 8:o + 8 + nb], dtype="<f8").reshape(ny, nx)
    assert v[1 * N1 + 0, 1 * N1 + 2] == phi[1, 1, 2, 0]
