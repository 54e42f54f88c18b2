"""Geometry / Grid / CoarseGrid: per-level containers, struct-of-arrays on the device.

Host-side mirror of dgfem/grid.py (Geometry :14-63, Grid :65-270, CoarseGrid :272-360).  The
reference keeps one Python object per element and per face; here a level is a handful of
device arrays (element-major, m = j*Ni + i):

    d_vol  [N][7][nq]      J, rx, sx, ry, sy, x, y at the volume quadrature points
    d_face [N][4][8][nq1]  per face imin,imax,jmin,jmax: J_f, alpha, beta, x, y, nx, ny, -
    d_area [N]             A = sum J w                                  (element.py:30)
    d_data/d_indices/d_indptr   the BSR operator (scipy layout)          (discrete_system.py:145)
    d_minv [N][b][b], d_dinv [N][b][b], d_rhs [N*b]

`grid.BSR` / `grid.RHS` materialise host copies lazily for callers that read them the way the
reference's post-processing does; `grid.elements[i, j]` is a light view (x, y, A, V_DOF_grid).
"""
import os

import numpy as np

from . import _lib
from .tables import Tables


def _read_fortran_records(path):
    raw = np.fromfile(path, dtype=np.uint8)
    recs, off = [], 0
    while off < raw.size:
        n = int(raw[off:off + 4].view("<u4")[0])
        payload = raw[off + 4:off + 4 + n]
        if int(raw[off + 4 + n:off + 8 + n].view("<u4")[0]) != n:
            raise ValueError("corrupt Fortran unformatted record")
        recs.append(payload)
        off += n + 8
    return recs


class Geometry:
    """Plot3D single-block 2-D grid (dgfem/grid.py:26-63)."""

    def __init__(self, filepath, settings, nodes=None):
        self.settings = settings
        self.filepath = filepath
        self.P_grid = settings.grid.polynomial_degree
        self.N_grid = self.P_grid + 1
        self.N_DOF_grid = self.N_grid ** 2
        self.O_grid = settings.grid.O_grid
        self.fully_periodic_boundaries = settings.grid.fully_periodic_boundaries
        self._dev = None
        # element-slab partitioning (parallel.py): fine element rows of halo available below / above this
        # slab's own rows inside the local node arrays (0 on a physical boundary or without partitioning)
        self.halo_lo = self.halo_hi = 0
        if nodes is not None:
            self._from_nodes(*nodes)
        else:
            self.read()

    def read(self):
        recs = _read_fortran_records(self.filepath)
        nblocks = recs[0].view("<i4")
        if nblocks.size != 1 or recs[0].size != 4:
            raise ValueError(f"Size of the record nblocks is {recs[0].size} instead of 4")
        if nblocks[0] != 1:
            raise ValueError(f"Number of blocks is {nblocks[0]} instead of 1")
        if recs[1].size != 12:
            raise ValueError(f"Size of the record dims is {recs[1].size} instead of 12")
        il, jl, kl = (int(v) for v in recs[1].view("<i4"))
        if kl != 1:
            raise ValueError("More than one point in third dimension")
        coords = recs[2].view("<f8")
        # file order is i-fastest: node (i, j) at j*il + i (grid.py:49-54)
        xn = np.ascontiguousarray(coords[:il * jl]).reshape(jl, il)
        yn = np.ascontiguousarray(coords[il * jl:2 * il * jl]).reshape(jl, il)
        self._from_nodes(xn, yn)

    def _from_nodes(self, xn, yn):
        """xn, yn: [jl][il] (file order).  self.x / self.y are the reference's [i, j] views."""
        self.xn, self.yn = np.ascontiguousarray(xn, dtype=np.float64), np.ascontiguousarray(yn, dtype=np.float64)
        jl, il = self.xn.shape
        self.il, self.jl = il, jl
        self.x, self.y = self.xn.T, self.yn.T
        if self.O_grid:
            if not np.all(abs(self.x[0, :] - self.x[-1, :]) < 1e-15) or not np.all(abs(self.y[0, :] - self.y[-1, :]) < 1e-15):
                raise ValueError("O-grid is not closed")
        self.Ni = (il - 1) // self.P_grid
        self.Nj = (jl - 1) // self.P_grid
        self.N = self.Ni * self.Nj

    def device_nodes(self):
        if self._dev is None:
            torch = _lib.require_cuda()
            self._dev = (torch.from_numpy(self.xn).cuda(), torch.from_numpy(self.yn).cuda())
        return self._dev


def padded_blocks(nblocks, b):
    """[nblocks, b, b] fp64 device tensor with 16 bytes of readable slack behind it (the streaming
    kernels issue 16-byte aligned bulk copies that may over-read by up to 8 bytes)."""
    torch = _lib.require_cuda()
    store = torch.zeros(nblocks * b * b + 2, dtype=torch.float64, device="cuda")
    view = store[:nblocks * b * b].view(nblocks, b, b)
    view._dgb_store = store
    return view


def upload_tables(T):
    """dgb_tables_create for a host Tables object; returns the opaque device handle."""
    import ctypes
    L = _lib.load()
    keep = {}

    def hp(name, arr, dtype=np.float64):
        a = np.ascontiguousarray(arr, dtype=dtype)
        keep[name] = a
        return a.ctypes.data
    d = _lib.TablesDesc(Pg=T.Pg, p=T.p, nq1=T.nq1, cf=T.cf,
                        h_V=hp("V", T.V), h_Vr=hp("Vr", T.Vr), h_Vs=hp("Vs", T.Vs),
                        h_w2=hp("w2", T.w2), h_w1=hp("w1", T.w1),
                        h_Vf=hp("Vf", T.Vf), h_Vrf=hp("Vrf", T.Vrf), h_Vsf=hp("Vsf", T.Vsf),
                        h_GX=hp("GX", T.GX), h_GR=hp("GR", T.GR), h_GS=hp("GS", T.GS),
                        h_FX=hp("FX", T.FX), h_FR=hp("FR", T.FR), h_FS=hp("FS", T.FS),
                        h_sub_vol=hp("sv", T.sub_vol, np.int32), h_sub_face=hp("sf", T.sub_face, np.int32))
    h = ctypes.c_void_p()
    _lib.check(L.dgb_tables_create(ctypes.byref(d), ctypes.byref(h)), "dgb_tables_create")
    return h


class _ElementView:
    """What callers of the reference read from grid.elements[i, j]."""

    def __init__(self, grid, i, j):
        g = grid
        s = g._node_stride
        ia = (i * g.P_grid + np.arange(g.P_grid + 1)) * s
        ja = (j * g.P_grid + np.arange(g.P_grid + 1)) * s
        self.x = g.x[np.ix_(ia, ja)]
        self.y = g.y[np.ix_(ia, ja)]
        self.V_DOF_grid = {"u": {"u": g.tables.V_DOF_grid}}
        self._grid, self._m = g, j * g.Ni + i

    @property
    def A(self):
        return float(self._grid.area_host()[self._m])

    @property
    def inv_mass_matrix(self):
        g = self._grid
        return None if g.d_minv is None else g.d_minv[self._m].cpu().numpy()


class _ElementArray:
    def __init__(self, grid):
        self._g = grid
        self.shape = (grid.Ni, grid.Nj)

    def __getitem__(self, ij):
        i, j = ij
        g = self._g
        return _ElementView(g, i % g.Ni, j % g.Nj)


class Grid:
    def __init__(self, geometry, vars, discretization="dg"):
        self.geometry = geometry
        for k in ("settings", "filepath", "P_grid", "N_grid", "N_DOF_grid", "O_grid",
                  "fully_periodic_boundaries", "x", "y", "il", "jl", "Ni", "Nj", "N"):
            setattr(self, k, getattr(geometry, k))
        self.coarsening_factor = None
        self._node_stride = 1
        # slab partitioning: one ghost element row below / above (the neighbour slab's edge row at this
        # level's resolution); the level's node window starts `_node_row0` fine element rows into the
        # local node arrays
        self.ghost_lo = 1 if geometry.halo_lo > 0 else 0
        self.ghost_hi = 1 if geometry.halo_hi > 0 else 0
        self._node_row0 = geometry.halo_lo - self.ghost_lo
        self.Nj = geometry.Nj - geometry.halo_lo - geometry.halo_hi + self.ghost_lo + self.ghost_hi
        self.N = self.Ni * self.Nj
        self.vars = vars
        self.discretization = discretization
        self.tables = None
        self._h_tables = None
        self.d_vol = self.d_face = self.d_area = None
        self.d_data = self.d_indices = self.d_indptr = None
        self.d_minv = self.d_dinv = self.d_rhs = self.d_gs = self.d_mailbox = self.d_chain = None
        self._BSR = self._RHS = self._area_host = None
        self.stencil = -1          # >= 0 once the BSR structure is verified to be the 5-point DG stencil
        self.BSR_E = self.BSR_D = self.BSR_F = None
        self.Epsilon = None

    # -- reference-facing attributes ---------------------------------------------------------
    @property
    def elements(self):
        return _ElementArray(self)

    @property
    def BSR(self):
        """scipy.sparse.bsr_array copy of the device operator (discrete_system.py:145)."""
        if self._BSR is None and getattr(self, "ordering", "local") == "global":
            self._BSR = self.BSR_global.to_scipy()          # Stokes global ordering (discrete_system.py:745)
        if self._BSR is None and self.d_data is not None:
            import scipy.sparse as sp
            n = self.N * self.N_DOF_sol_tot
            self._BSR = sp.bsr_array((self.d_data.cpu().numpy(), self.d_indices.cpu().numpy(),
                                      self.d_indptr.cpu().numpy()), shape=(n, n))
        return self._BSR

    @BSR.setter
    def BSR(self, value):
        """Accept a host matrix (e.g. assembled elsewhere) and mirror it on the device."""
        self._BSR = value
        if value is not None:
            torch = _lib.require_cuda()
            data = np.ascontiguousarray(value.data, dtype=np.float64)
            self.d_data = padded_blocks(data.shape[0], data.shape[1])
            self.d_data.copy_(torch.from_numpy(data))
            self.d_indices = torch.from_numpy(np.ascontiguousarray(value.indices, dtype=np.int32)).cuda()
            self.d_indptr = torch.from_numpy(np.ascontiguousarray(value.indptr, dtype=np.int32)).cuda()
            self.d_dinv = self.d_gs = self.d_mailbox = self.d_chain = None
            self.stencil = -1

    def operator(self):
        """dgb_operator view of this level's device arrays (include/dgb200.h)."""
        b = int(self.d_data.shape[1])
        return _lib.Operator(Ni=int(self.Ni), Nj=int(self.Nj), b=b, nnzb=int(self.d_indices.numel()),
                             stencil=int(self.stencil), reserved=0,
                             data=self.d_data.data_ptr(), indices=self.d_indices.data_ptr(),
                             indptr=self.d_indptr.data_ptr(),
                             dinv=self.d_dinv.data_ptr() if self.d_dinv is not None else None,
                             gs_data=self.d_gs.data_ptr() if self.d_gs is not None else None,
                             gs_mailbox=self.d_mailbox.data_ptr() if self.d_mailbox is not None else None,
                             gs_chain=self.d_chain.data_ptr() if getattr(self, "d_chain", None) is not None else None)

    @property
    def RHS(self):
        if self._RHS is None and self.d_rhs is not None:
            self._RHS = self.d_rhs.cpu().numpy()
        return self._RHS

    @RHS.setter
    def RHS(self, value):
        self._RHS = value
        if value is not None:
            torch = _lib.require_cuda()
            self.d_rhs = torch.from_numpy(np.ascontiguousarray(value, dtype=np.float64)).cuda()

    def area_host(self):
        if self._area_host is None:
            self._area_host = self.d_area.cpu().numpy()
        return self._area_host

    @property
    def b(self):
        return self.N_DOF_sol_tot

    # -- initialisation (grid.py:95-149) -----------------------------------------------------
    def _set_solution_space(self, P_sol, sigma, gamma):
        s = self.settings
        self.P_sol = dict(P_sol)
        self.N_sol = {v: self.P_sol[v] + 1 for v in self.vars}
        self.N_DOF_sol = {v: self.N_sol[v] ** 2 for v in self.vars}
        self.N_DOF_sol_tot = self.N_DOF_sol["u"] if self.vars == ["u"] else \
            sum(n * 2 if v == "u" else n for v, n in self.N_DOF_sol.items())          # grid.py:106
        self.N_int = {v: getattr(getattr(s.solution, v), "integration_polynomial_degree_factor") * self.P_sol[v] // 2 + 1
                      for v in self.vars}                                              # grid.py:107
        self.sigma = sigma
        if not self.sigma:
            self.sigma = s.problem.SIP_penalty_parameter if s.problem.SIP_penalty_parameter else \
                (self.P_sol["u"] + 1) ** 2 * s.problem.SIP_penalty_parameter_multiplier  # grid.py:110
        self.gamma = gamma if gamma else s.problem.velocity_penalty_parameter

    def _make_tables(self, cf):
        factor = self.settings.solution.u.integration_polynomial_degree_factor
        self.tables = Tables(self.P_grid, self.P_sol["u"], factor=factor, cf=cf)
        self._h_tables = upload_tables(self.tables)

    def _run_metrics(self):
        torch = _lib.require_cuda()
        T = self.tables
        xn, yn = self.geometry.device_nodes()
        if self._node_row0:
            xn, yn = xn[self._node_row0 * self.P_grid:], yn[self._node_row0 * self.P_grid:]
        N = self.Ni * self.Nj
        self.d_vol = torch.empty((N, 7, T.nq), dtype=torch.float64, device="cuda")
        self.d_face = torch.empty((N, 4, 8, T.nq1), dtype=torch.float64, device="cuda")
        self.d_area = torch.empty((N,), dtype=torch.float64, device="cuda")
        _lib.call("dgb_metrics", self._h_tables, xn, yn, self.il, self.Ni, self.Nj,
                  self.d_vol, self.d_face, self.d_area, _lib.stream_ptr())

    def initialize(self, P_sol, sigma=None, gamma=None):
        self._set_solution_space(P_sol, sigma, gamma)
        self._make_tables(cf=1)
        self._run_metrics()
        if self.O_grid:
            self._check_closed()
        return self

    def _check_closed(self):
        # grid.py:223-225: first and last element columns must share their i-face nodes
        s = self._node_stride
        last = self.Ni * self.P_grid * s
        if not np.all(abs(self.x[0, :] - self.x[last, :]) < 1e-15) or not np.all(abs(self.y[0, :] - self.y[last, :]) < 1e-15):
            raise ValueError("Element does not close O-grid with neighbouring element")

    def release_geometry(self):
        """Free the per-point metric arrays once the operator and RHS exist."""
        self.d_vol = self.d_face = None

    def __del__(self):
        try:
            if self._h_tables is not None:
                _lib.load().dgb_tables_destroy(self._h_tables)
                self._h_tables = None
        except Exception:
            pass


class CoarseGrid(Grid):
    """h-coarsened level (dgfem/grid.py:272-360): every cf-th node, metrics sampled from the
    containing fine elements."""

    def __init__(self, geometry, fine_grid, vars, discretization="dg"):
        super().__init__(geometry, vars, discretization)
        self._fine = fine_grid

    def initialize(self, coarsening_factor):
        f = self._fine
        self.coarsening_factor = coarsening_factor
        self._node_stride = coarsening_factor
        self.Ni_fine, self.Nj_fine = f.Ni, f.Nj
        own_rows = f.Nj - f.ghost_lo - f.ghost_hi                      # active fine rows of this slab
        if own_rows % coarsening_factor != 0:
            raise ValueError(f"a slab of {own_rows} element rows cannot be coarsened by {coarsening_factor}")
        g = self.geometry
        if (self.ghost_lo and g.halo_lo < coarsening_factor) or (self.ghost_hi and g.halo_hi < coarsening_factor):
            raise ValueError("slab halo too thin for this coarsening factor")
        self._node_row0 = g.halo_lo - coarsening_factor * self.ghost_lo
        self.Ni = f.Ni // coarsening_factor
        self.Nj = own_rows // coarsening_factor + self.ghost_lo + self.ghost_hi
        self.N = self.Ni * self.Nj
        if self.Ni == 0 or self.Nj == 0:
            raise ValueError(f"The number of original elements (Ni,Nj)={(f.Ni, f.Nj)} cannot be divided by a factor {coarsening_factor}")
        self._set_solution_space(f.P_sol, f.sigma, f.gamma)
        self._make_tables(cf=coarsening_factor)
        self._run_metrics()
        if self.O_grid:
            self._check_closed()
        return self
