"""Stokes (pressure-robust DG): assembly of the operator and right-hand side on the device.  Same surface as the
reference's `Stokes` problem class (dgfem/discrete_system.py:405-414, 416-1028):
`DiscreteSystem(settings).problem.assemble(grid)` fills grid.BSR / grid.RHS.

  local ordering   one (2 b_u + b_p)^2 block per element pair (discrete_system.py:812-965): dgb_assemble_stokes
  global ordering  [u of all elements | v | p] (discrete_system.py:416-745): the same integrals, regrouped into
                   grid.BSR_block_A / _D / _G (+ the full matrix) with exactly the SciPy calls the reference makes on
                   its per-component matrices -- run here on index-coded data, so the block sizes SciPy picks
                   (App. B.6) and the stored order are the reference's and the values are a device gather out of
                   the locally ordered blocks.  Feeds Relaxation.distributive_gauss_seidel.

The reference has no Stokes multigrid (README "future work"; settings.py:33-36), so this path ends at the
assembled system, `grid.BSR @ u`, the single-level block smoothers and distributive Gauss-Seidel."""
import numpy as np
import sympy as sym

from . import _lib
from .discrete_system import prepare_smoother_data, stencil_flags
from .grid import padded_blocks, upload_tables
from .mms import Field
from .tables import Tables


class StokesMMS:
    """dgfem/dgfem.py:410-483 for problem == 'Stokes' (momentum source = -div(nu grad u) + grad p)."""

    def __init__(self, settings):
        x, y = sym.symbols("x y")
        nu = settings.problem.kinematic_viscosity
        ex = settings.problem.exact_solution
        u, v, p = sym.sympify(ex.u), sym.sympify(ex.v), sym.sympify(ex.p)
        f_cont = sym.diff(u, x) + sym.diff(v, y)
        if settings.solution.manufactured_solution and not sym.simplify(f_cont).is_zero:
            raise ValueError("Manufactured solution is not divergence-free")          # dgfem.py:427-429
        lap = lambda w: -(sym.diff(nu * sym.diff(w, x), x) + sym.diff(nu * sym.diff(w, y), y))   # noqa: E731
        self.u, self.v = Field(u), Field(v)
        self.fx, self.fy = Field(lap(u) + sym.diff(p, x)), Field(lap(v) + sym.diff(p, y))   # dgfem.py:466-469
        self.fc = Field(f_cont)


class GlobalBlock:
    """One block (A, D, G, D G or the whole matrix) of the global-order Stokes system: a generic BSR operator on
    the device (dgb_operator with stencil = -1), square b x b blocks, any shape."""

    def __init__(self, shape, b, indptr, indices, d_data):
        torch = _lib.require_cuda()
        self.shape, self.b = (int(shape[0]), int(shape[1])), int(b)
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.d_indptr = torch.from_numpy(self.indptr).cuda()
        self.d_indices = torch.from_numpy(self.indices).cuda()
        self.d_data = d_data                       # [nnzb][b][b] (+ 16 bytes of slack)
        self.d_dinv = self.d_gs_rows = None
        self._h_off = None
        self.nlev = (0, 0)

    @classmethod
    def from_coded(cls, coded, src_flat):
        """coded: scipy BSR whose data holds 1 + (flat index into src_flat), 0 = structural padding."""
        torch = _lib.require_cuda()
        b = int(coded.blocksize[0])
        assert coded.blocksize[0] == coded.blocksize[1]
        nnzb = coded.data.shape[0]
        d_data = padded_blocks(nnzb, b)
        if nnzb:
            code = torch.from_numpy(np.ascontiguousarray(coded.data, dtype=np.float64)).cuda().to(torch.int64).view(-1)
            vals = torch.where(code > 0, src_flat[(code - 1).clamp_(min=0)], torch.zeros((), dtype=torch.float64, device="cuda"))
            d_data.view(-1)[:vals.numel()].copy_(vals)
        return cls(coded.shape, b, coded.indptr, coded.indices, d_data)

    @property
    def n_brow(self):
        return self.shape[0] // self.b

    def operator(self):
        return _lib.Operator(Ni=self.n_brow, Nj=1, b=self.b, nnzb=int(self.indices.size), stencil=-1, reserved=0,
                             data=self.d_data.data_ptr(), indices=self.d_indices.data_ptr(),
                             indptr=self.d_indptr.data_ptr(),
                             dinv=self.d_dinv.data_ptr() if self.d_dinv is not None else None,
                             gs_data=None, gs_mailbox=None, gs_chain=None,
                             gs_rows=self.d_gs_rows.data_ptr() if self.d_gs_rows is not None else None,
                             h_gs_offsets=self._h_off.ctypes.data if self._h_off is not None else None,
                             gs_nlevels_fwd=self.nlev[0], gs_nlevels_bwd=self.nlev[1])

    def apply(self, x):
        torch = _lib.require_cuda()
        y = torch.empty(self.shape[0], dtype=torch.float64, device="cuda")
        _lib.call("dgb_bsr_apply", self.operator(), x, y, _lib.stream_ptr())
        return y

    def prepare_gauss_seidel(self):
        """Inverse diagonal blocks (get_block_diag(A, blocksize, inv_flag=True), pyamg_relaxation.py:230-231) and the
        level schedule of the lexicographic sweep: level(k) = 1 + max level(j) over the stored j < k of row k
        (j > k for the backward sweep); rows of one level do not couple (the pattern is structurally symmetric)."""
        torch = _lib.require_cuda()
        if self.d_dinv is not None:
            return
        assert self.shape[0] == self.shape[1]
        n = self.n_brow
        self.d_dinv = padded_blocks(n, self.b)
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.call("dgb_block_diag_inverse", self.d_data, self.d_indices, self.d_indptr, n, self.b, self.d_dinv, info,
                  _lib.stream_ptr())
        self._dinv_info = info
        rows = torch.repeat_interleave(torch.arange(n, device="cuda"), torch.from_numpy(np.diff(self.indptr)).cuda())
        cols = self.d_indices.to(torch.int64)
        lists, offs = [], []
        for lower in (True, False):
            m = cols < rows if lower else cols > rows
            r, c = rows[m], cols[m]
            level = torch.zeros(n, dtype=torch.int64, device="cuda")
            for _ in range(n + 1):
                new = torch.zeros_like(level).scatter_reduce_(0, r, level[c] + 1, "amax", include_self=True)
                if torch.equal(new, level):
                    break
                level = new
            order = torch.argsort(level * n + torch.arange(n, device="cuda"))
            lists.append(order.to(torch.int32))
            counts = torch.bincount(level, minlength=int(level.max().item()) + 1).cpu().numpy()
            offs.append(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32))
        self.d_gs_rows = torch.cat(lists).contiguous()
        self._h_off = np.ascontiguousarray(np.concatenate(offs), dtype=np.int32)
        self.nlev = (len(offs[0]) - 1, len(offs[1]) - 1)

    def to_scipy(self):
        import scipy.sparse as sp
        nnzb = int(self.indices.size)
        data = self.d_data.view(-1)[:nnzb * self.b * self.b].view(nnzb, self.b, self.b).cpu().numpy()
        return sp.bsr_array((data, self.indices, self.indptr), shape=self.shape)


class Stokes:
    def __init__(self, settings):
        self.settings = settings

    def assemble(self, grid):
        order = self.settings.solution.ordering.lower()
        self.assemble_BSR_Stokes_local_order(grid)
        self.assemble_RHS_Stokes(grid)
        if order == "global":
            self.assemble_BSR_Stokes_global_order(grid)
        elif order != "local":
            raise ValueError("solution ordering must be local|global")

    def assemble_BSR_Stokes_global_order(self, grid):
        """discrete_system.py:416-745 from the locally ordered device blocks (bitwise the same integrals: the
        reference's two assembly routines differ only in where they put them)."""
        import scipy.sparse as sp
        torch = _lib.require_cuda()
        S = self._setup(grid)
        bu, bp = S["bu"], S["bp"]
        bt = 2 * bu + bp
        N = grid.Ni * grid.Nj
        indptr = grid.d_indptr.cpu().numpy()
        indices = grid.d_indices.cpu().numpy()
        nnzb = indices.size
        src = grid.d_data.view(-1)[:nnzb * bt * bt]
        code = 1.0 + np.arange(nnzb * bt * bt, dtype=np.float64).reshape(nnzb, bt, bt)      # exact below 2^53
        rows = {"xmom": slice(0, bu), "ymom": slice(bu, 2 * bu), "cont": slice(2 * bu, bt)}
        cols = {"u": slice(0, bu), "v": slice(bu, 2 * bu), "p": slice(2 * bu, bt)}

        def comp(r, c, shape):
            return sp.bsr_array((np.ascontiguousarray(code[:, rows[r], cols[c]]), indices, indptr), shape=shape)
        nu_, np_ = N * bu, N * bp
        Au_x, Av_x = comp("xmom", "u", (nu_, nu_)), comp("xmom", "v", (nu_, nu_))
        Au_y, Av_y = comp("ymom", "u", (nu_, nu_)), comp("ymom", "v", (nu_, nu_))
        # the reference's own calls (discrete_system.py:731-745): SciPy picks the block sizes
        cA = sp.bsr_array(sp.vstack([sp.hstack([Au_x, Av_x]), sp.hstack([Au_y, Av_y])], format="bsr"))
        cD = sp.bsr_array(sp.hstack([comp("cont", "u", (np_, nu_)), comp("cont", "v", (np_, nu_))], format="bsr"))
        cG = sp.bsr_array(sp.vstack([comp("xmom", "p", (nu_, np_)), comp("ymom", "p", (nu_, np_))], format="bsr"))
        c0 = sp.bsr_array(np.zeros((np_, np_)))
        if self.settings.get("solver.method") == "direct":
            z = np.zeros((np_, np_)); z[0, 0] = -1.0            # coded below as "the pinned entry" (value 1.0)
            c0 = sp.bsr_array(z)
        cK = sp.bsr_array(sp.vstack([sp.hstack([cA, cG]), sp.hstack([cD, c0])], format="bsr"))
        grid.BSR_block_A = GlobalBlock.from_coded(cA, src)
        grid.BSR_block_D = GlobalBlock.from_coded(cD, src)
        grid.BSR_block_G = GlobalBlock.from_coded(cG, src)
        # whole matrix: the pinned entry (code -1) carries the value 1 (discrete_system.py:741)
        pin = cK.data < 0
        cK.data[pin] = 0.0
        grid.BSR_global = GlobalBlock.from_coded(cK, src)
        if pin.any():
            flat = torch.from_numpy(np.flatnonzero(pin.ravel())).cuda()
            grid.BSR_global.d_data.view(-1)[flat] = 1.0
        grid.BSR_block_DG = None
        # right-hand side and the permutation local -> global ordering
        e = np.arange(N)
        perm = np.concatenate([(e[:, None] * bt + c * bu + np.arange(bu)[None, :]).ravel() for c in (0, 1)] +
                              [(e[:, None] * bt + 2 * bu + np.arange(bp)[None, :]).ravel()])
        grid.d_perm_global = torch.from_numpy(perm).cuda()
        grid.d_rhs_local = grid.d_rhs
        grid.d_rhs = grid.d_rhs_local[grid.d_perm_global].contiguous()
        grid._RHS = None
        grid._BSR = None
        grid.ordering = "global"

    def block_DG(self, grid):
        """grid.BSR_block_DG = grid.BSR_block_D @ grid.BSR_block_G (dgfem/relaxation.py:240): SciPy's structure of
        the product, values by dgb_bsr_spgemm."""
        import scipy.sparse as sp
        if grid.BSR_block_DG is not None:
            return grid.BSR_block_DG
        D, G = grid.BSR_block_D, grid.BSR_block_G
        assert D.b == G.b

        def ones(B):
            return sp.bsr_array((np.ones((B.indices.size, B.b, B.b)), B.indices, B.indptr), shape=B.shape)
        C = ones(D) @ ones(G)                    # structure (and stored order) as the reference's product has it
        C = sp.bsr_array(C)
        nnzb = C.indices.size
        d_data = padded_blocks(nnzb, D.b)
        DG = GlobalBlock(C.shape, D.b, C.indptr, C.indices, d_data)
        _lib.call("dgb_bsr_spgemm", D.b, D.n_brow, D.d_indptr, D.d_indices, D.d_data, G.d_indptr, G.d_indices, G.d_data,
                  DG.d_indptr, DG.d_indices, DG.d_data, _lib.stream_ptr())
        grid.BSR_block_DG = DG
        return DG

    def _setup(self, grid):
        """Tables and metrics at the velocity and at the pressure quadrature points."""
        torch = _lib.require_cuda()
        if getattr(grid, "_stokes", None) is not None:
            return grid._stokes
        s = self.settings
        pu, pp = grid.P_sol["u"], grid.P_sol["p"]
        fu = s.solution.u.integration_polynomial_degree_factor
        fp = s.solution.p.integration_polynomial_degree_factor
        n1u, n1p = fu * pu // 2 + 1, fp * pp // 2 + 1                                  # grid.py:107
        T = dict(uu=Tables(grid.P_grid, pu, nq1=n1u), pu=Tables(grid.P_grid, pp, nq1=n1u),
                 up=Tables(grid.P_grid, pu, nq1=n1p), pp=Tables(grid.P_grid, pp, nq1=n1p))
        H = {k: upload_tables(v) for k, v in T.items()}
        xn, yn = grid.geometry.device_nodes()
        N = grid.Ni * grid.Nj
        st = _lib.stream_ptr()
        geo = {}
        for key, tab in (("u", "uu"), ("p", "up")):
            nq1 = T[tab].nq1
            vol = torch.empty((N, 7, nq1 * nq1), dtype=torch.float64, device="cuda")
            face = torch.empty((N, 4, 8, nq1), dtype=torch.float64, device="cuda")
            area = torch.empty((N,), dtype=torch.float64, device="cuda")
            _lib.call("dgb_metrics", H[tab], xn, yn, grid.il, grid.Ni, grid.Nj, vol, face, area, st)
            geo[key] = (vol, face, area)
        grid._stokes = dict(T=T, H=H, geo=geo, bu=T["uu"].b, bp=T["pp"].b)
        return grid._stokes

    def assemble_BSR_Stokes_local_order(self, grid):
        torch = _lib.require_cuda()
        L = _lib.load()
        S = self._setup(grid)
        s = self.settings
        flags = stencil_flags(grid, s) & ~_lib.FLAG_MINV        # no inverse-mass scaling (discrete_system.py:941)
        if grid.fully_periodic_boundaries:
            raise NotImplementedError("the Stokes assembly of the reference has no fully periodic branch")
        bt = 2 * S["bu"] + S["bp"]
        N = grid.Ni * grid.Nj
        nnzb = int(L.dgb_poisson_nnzb(grid.Ni, grid.Nj, flags))
        grid.d_data = padded_blocks(nnzb, bt)
        grid.d_indices = torch.empty(nnzb, dtype=torch.int32, device="cuda")
        grid.d_indptr = torch.empty(N + 1, dtype=torch.int32, device="cuda")
        (vu, fu, area), (vp, fp, _) = S["geo"]["u"], S["geo"]["p"]
        H = S["H"]
        pin = 1 if s.get("solver.method") == "direct" else 0                     # discrete_system.py:946
        _lib.call("dgb_assemble_stokes", H["uu"], H["pu"], H["up"], H["pp"], vu, fu, vp, fp, area, grid.Ni, grid.Nj,
                  float(s.problem.kinematic_viscosity), float(grid.sigma), float(grid.gamma), flags, pin,
                  grid.d_indptr, grid.d_indices, grid.d_data, _lib.stream_ptr())
        grid.d_area = area
        grid.flags, grid.nnzb, grid._BSR = flags, nnzb, None
        grid.stencil = flags
        if not pin:
            # block smoothers need the inverse diagonal blocks; the pinned direct-solve matrix does not
            prepare_smoother_data(grid)

    def assemble_RHS_Stokes(self, grid):
        torch = _lib.require_cuda()
        S = self._setup(grid)
        s = self.settings
        if s.problem.include_pressure_BC:
            raise NotImplementedError("`include pressure BC: True` is not accelerated")
        mms = StokesMMS(s)
        (vu, fu, area), (vp, fp, _) = S["geo"]["u"], S["geo"]["p"]
        xu, yu = vu[:, 5, :], vu[:, 6, :]
        f_mom = torch.stack([mms.fx(xu, yu), mms.fy(xu, yu)], dim=1).contiguous()                 # [N,2,nqu]
        f_cont = mms.fc(vp[:, 5, :], vp[:, 6, :]).contiguous()
        g_u = torch.stack([mms.u(fu[:, :, 3, :], fu[:, :, 4, :]), mms.v(fu[:, :, 3, :], fu[:, :, 4, :])], dim=2).contiguous()
        g_p = torch.stack([mms.u(fp[:, :, 3, :], fp[:, :, 4, :]), mms.v(fp[:, :, 3, :], fp[:, :, 4, :])], dim=2).contiguous()
        bt = 2 * S["bu"] + S["bp"]
        grid.d_rhs = torch.empty(grid.Ni * grid.Nj * bt, dtype=torch.float64, device="cuda")
        H = S["H"]
        _lib.call("dgb_assemble_rhs_stokes", H["uu"], H["pu"], H["up"], H["pp"], vu, fu, vp, fp, area, f_mom, f_cont,
                  g_u, g_p, grid.Ni, grid.Nj, float(s.problem.kinematic_viscosity), float(grid.sigma),
                  float(grid.gamma), grid.flags, grid.d_rhs, _lib.stream_ptr())
        grid._RHS = None
