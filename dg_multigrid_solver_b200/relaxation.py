"""Relaxation: the smoother plugins, resolved by name with getattr(Relaxation, name) exactly as
the reference's Solver does (dgfem/solver.py:65,147,196,204).

Same call surface as dgfem/relaxation.py: classmethods
    Relaxation.<name>(grid, RHS, u=None, direction=..., omega=1, max_iterations=...) -> new vector
that never mutate RHS/u.  RHS/u may be NumPy arrays (host; copied to the device and back) or
CUDA torch tensors (stay on the device).  All arithmetic runs in libdgb200.so.
"""
import numpy as np

from . import _lib
from .discrete_system import check_dinv, prepare_smoother_data

_DIRECTION = {"symmetric": 0, "forward": 1, "backward": -1}


class _Workspace:
    """Per-process scratch: partial sums, the norm scalar and one smoother control block."""
    _inst = None

    def __init__(self):
        torch = _lib.require_cuda()
        self.partials = torch.zeros(_lib.load().dgb_partials_len(), dtype=torch.float64, device="cuda")
        self.sumsq = torch.zeros(1, dtype=torch.float64, device="cuda")
        self.ctl = torch.zeros(32, dtype=torch.uint8, device="cuda")      # sizeof(dgb_smoother_ctl)

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def read_ctl(self):
        raw = self.ctl.cpu().numpy().tobytes()
        return _lib.SmootherCtl.from_buffer_copy(raw)


def _to_device(v, like=None):
    torch = _lib.require_cuda()
    if v is None:
        return None, False
    if isinstance(v, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).cuda(), True
    return v, False


def _gs_settings(grid):
    s = getattr(grid, "settings", None)
    mode = s.get("solver.b200.gs_mode", "lexicographic") if s is not None and hasattr(s, "get") else "lexicographic"
    chk = s.get("solver.b200.check_residual", True) if s is not None and hasattr(s, "get") else True
    return (_lib.GS_REDBLACK if mode == "redblack" else _lib.GS_LEXICOGRAPHIC), (1 if chk else 0)


def _ensure_dinv(grid):
    if grid.d_dinv is None:
        prepare_smoother_data(grid)
    return grid.d_dinv


class Relaxation:
    @classmethod
    def _prep(cls, grid, RHS, u):
        torch = _lib.require_cuda()
        d_rhs, host = _to_device(RHS)
        if isinstance(u, np.ndarray):
            d_u = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64)).cuda()
        elif u is None or not hasattr(u, "data_ptr"):
            d_u = torch.zeros_like(d_rhs)                      # relaxation.py:199
        else:
            d_u = u.clone()                                     # relaxation.py:200 (never mutate the input)
        return d_rhs, d_u, host

    @classmethod
    def block_gauss_seidel_pyamg(cls, grid, RHS, u=None, direction="symmetric", omega=1, max_iterations=1e3,
                                 gs_mode=None, check_residual=None):
        """dgfem/relaxation.py:198-218 (omega accepted and ignored, as there)."""
        d_rhs, d_u, host = cls._prep(grid, RHS, u)
        ws = _Workspace.get()
        mode, chk = _gs_settings(grid)
        if gs_mode is not None:
            mode = _lib.GS_REDBLACK if gs_mode in ("redblack", _lib.GS_REDBLACK) else _lib.GS_LEXICOGRAPHIC
        if check_residual is not None:
            chk = 1 if check_residual else 0
        _ensure_dinv(grid)
        ws.ctl.zero_()                 # `diverged` is sticky on the device: every call starts from a clean block
        _lib.call("dgb_block_gauss_seidel_pyamg", grid.operator(), d_rhs, d_u, _DIRECTION[direction],
                  int(max_iterations), mode, chk, ws.ctl, ws.partials, ws.sumsq, _lib.stream_ptr())
        if host:
            cls.finish(grid, chk)
            return d_u.cpu().numpy()
        # device tensors in, device tensor out: no host synchronisation here -- the caller runs
        # Relaxation.finish(grid) where it synchronises (divergence / singular-block / kernel-error checks)
        return d_u

    @classmethod
    def finish(cls, grid, check_residual=True):
        """The host-side checks of a smoother call (dgfem/relaxation.py:211-216 and the inverse of a singular
        diagonal block raising in pyamg): divergence -> SystemExit, early exit -> the reference's message;
        also the error flag of the asynchronous kernels.  Synchronises."""
        ws = _Workspace.get()
        check_dinv(grid)
        _lib.check_device_error([getattr(grid, "d_mailbox", None)])
        if check_residual:
            ctl = ws.read_ctl()
            cls.last_info = {"iterations": ctl.iters, "ratio": ctl.ratio, "early_exit": bool(ctl.skip and not ctl.diverged)}
            if ctl.diverged:
                print(f"diverging, residual={ctl.ratio:.6e}")      # relaxation.py:214-216
                raise SystemExit()
            if ctl.skip:
                print(f"Residual reduced by 6 orders in {ctl.iters} sweeps")   # relaxation.py:212

    @classmethod
    def block_jacobi(cls, grid, RHS, u=None, direction=None, omega=1, max_iterations=1e3):
        """dgfem/relaxation.py:123-150.  The reference's `u = u_new` aliases the two buffers after the
        first pass, so iteration 1 is block-Jacobi and every later one is an in-place forward
        block-GS pass; reproduced as such."""
        torch = _lib.require_cuda()
        d_rhs, d_u, host = cls._prep(grid, RHS, u)
        st = _lib.stream_ptr()
        _ensure_dinv(grid)
        args = (grid.operator(), d_rhs)
        if int(max_iterations) > 0:
            d_new = torch.empty_like(d_u)
            _lib.call("dgb_block_relax_sweep", *args, d_u, d_new, float(omega), st)
            for _ in range(int(max_iterations) - 1):
                _lib.call("dgb_block_relax_sweep", *args, d_new, d_new, float(omega), st)
            d_u = d_new
        if host:
            check_dinv(grid)
            _lib.check_device_error([getattr(grid, "d_mailbox", None)])
            return d_u.cpu().numpy()
        return d_u

    @classmethod
    def block_gauss_seidel(cls, grid, RHS, u=None, direction="forward", omega=1, max_iterations=1e3):
        """dgfem/relaxation.py:170-195 (forward lexicographic order; `direction` is ignored there)."""
        d_rhs, d_u, host = cls._prep(grid, RHS, u)
        st = _lib.stream_ptr()
        _ensure_dinv(grid)
        op = grid.operator()
        for _ in range(int(max_iterations)):
            _lib.call("dgb_block_relax_sweep", op, d_rhs, d_u, d_u, float(omega), st)
        if host:
            check_dinv(grid)
            _lib.check_device_error([getattr(grid, "d_mailbox", None)])
            return d_u.cpu().numpy()
        return d_u

    @classmethod
    def _bgs_block(cls, blk, d_rhs, sweeps):
        """Relaxation.block_gauss_seidel_pyamg(grid, rhs, u=0, direction='symmetric', max_iterations=sweeps) with
        grid.BSR = one global-order block (dgfem/relaxation.py:252,256,263): the level-scheduled lexicographic
        sweep of the generic kernels, block size = the block's (SciPy-chosen) blocksize."""
        torch = _lib.require_cuda()
        ws = _Workspace.get()
        blk.prepare_gauss_seidel()
        x = torch.zeros(blk.shape[0], dtype=torch.float64, device="cuda")
        ws.ctl.zero_()
        _lib.call("dgb_block_gauss_seidel_pyamg", blk.operator(), d_rhs, x, 0, int(sweeps), _lib.GS_LEXICOGRAPHIC, 1,
                  ws.ctl, ws.partials, ws.sumsq, _lib.stream_ptr())
        return x

    @classmethod
    def distributive_gauss_seidel(cls, grid, RHS, u=None, inner_smoother="block_gauss_seidel_pyamg",
                                  splitting="classical_exact", omega=1, max_iterations=1e3, settings=None):
        """dgfem/relaxation.py:221-283, `lsq` splitting (the one Solver.solve_smoother uses, solver.py:63): per outer
        iteration three symmetric block-GS iterations (on A, on D G, on D G) and eight block mat-vecs, all on the
        device; one scalar per outer iteration comes back for the reference's convergence test."""
        import os
        import pickle
        torch = _lib.require_cuda()
        if settings.problem.type != "Stokes":
            raise ValueError("Distributive Gauss-Seidel is only possible for the Stokes equations")
        if settings.solution.ordering != "global":
            raise ValueError("The solution ordering must be global in order to use distributive Gauss-Seidel")
        if splitting != "lsq":
            raise NotImplementedError(f"distributive_gauss_seidel splitting '{splitting}' is outside the accelerated "
                                      "path (the reference's -s run uses 'lsq', dgfem/solver.py:63)")
        from .stokes import Stokes
        A, D, G = grid.BSR_block_A, grid.BSR_block_D, grid.BSR_block_G
        DG = Stokes(settings).block_DG(grid)
        d_rhs, host = _to_device(RHS)
        if isinstance(u, np.ndarray):
            d_u = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64)).cuda()
        elif u is None or not hasattr(u, "data_ptr"):
            d_u = torch.zeros_like(d_rhs)
        else:
            d_u = u.clone()
        n_u = A.shape[0]
        f_mom, f_cont = d_rhs[:n_u], d_rhs[n_u:]

        def full_residual_rms(v):
            r1 = f_mom - A.apply(v[:n_u]) - G.apply(v[n_u:])
            r2 = f_cont - D.apply(v[:n_u])
            return float(torch.sqrt(((r1 * r1).sum() + (r2 * r2).sum()) / d_rhs.numel()).item())
        residual_0 = full_residual_rms(d_u)
        residuals = []
        sweeps, n = 1, 0
        with np.errstate(divide="ignore", invalid="ignore"):
            while n < max_iterations:
                u_k, p_k = d_u[:n_u], d_u[n_u:]
                rhs_mom = f_mom - A.apply(u_k) - G.apply(p_k)
                du_star = cls._bgs_block(A, rhs_mom, sweeps)
                rhs_cont = f_cont - D.apply(u_k + du_star)
                dp_star = cls._bgs_block(DG, rhs_cont, sweeps)
                du = du_star + G.apply(dp_star)
                rhs_dg = -D.apply(A.apply(G.apply(dp_star)))
                dp = cls._bgs_block(DG, rhs_dg, sweeps)
                d_u[:n_u] += du
                d_u[n_u:] += dp
                residual = np.float64(full_residual_rms(d_u)) / np.float64(residual_0)
                residuals.append(float(residual))
                if residual < 1e-6:
                    print(f"Residual reduced by 6 orders in {n} sweeps")         # relaxation.py:272
                    break
                elif residual > 1e10:
                    print(f"diverging, residual={residual:.6e}")
                    raise SystemExit()
                n += 1
        _lib.check_device_error([])
        cls.last_residuals = residuals
        try:                                                                       # relaxation.py:229-234,282-283
            path = os.path.join(os.getcwd(), "postprocessing", "pickles", "relaxation")
            os.makedirs(path, exist_ok=True)
            name = f"residuals_{settings.problem.type}_{grid.Ni}X{grid.Nj}_nPoly{grid.P_grid}_Pu{grid.P_sol.get('u')}" \
                   f"_Pp{grid.P_sol.get('p')}_{splitting}" + ("_circle" if settings.grid.circular else "_rectangle") + ".pkl"
            with open(os.path.join(path, name), "wb") as f:
                pickle.dump(residuals, f)
        except OSError:
            pass
        return d_u.cpu().numpy() if host else d_u

    # names the reference resolves but that are outside the accelerated path
    @classmethod
    def _out_of_scope(cls, *a, **k):
        raise NotImplementedError("this smoother is outside the B200 hot path (SURVEY.md section 2.1 row 8)")

    jacobi = jacobi_pyamg = gauss_seidel = gauss_seidel_pyamg = _out_of_scope
    calculate_amplification = _out_of_scope


def bsr_apply(grid, x):
    """grid.BSR @ x on the device (scipy bsr_matvec; dgfem/solver.py:117,119,150)."""
    torch = _lib.require_cuda()
    d_x, host = _to_device(x)
    y = torch.empty_like(d_x)
    _lib.call("dgb_bsr_apply", grid.operator(), d_x, y, _lib.stream_ptr())
    return y.cpu().numpy() if host else y


def residual_norm(grid, rhs, x, want_residual=False):
    """(sum((rhs - A x)^2), residual or None), device tensors."""
    torch = _lib.require_cuda()
    ws = _Workspace.get()
    r = torch.empty_like(rhs) if want_residual else None
    _lib.call("dgb_bsr_residual", grid.operator(), rhs, x, r, ws.partials, ws.sumsq, None, _lib.stream_ptr())
    return ws.sumsq, r
