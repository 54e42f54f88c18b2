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

    # names the reference resolves but that are outside the accelerated path
    @classmethod
    def _out_of_scope(cls, *a, **k):
        raise NotImplementedError("this smoother is outside the B200 hot path (SURVEY.md section 2.1 row 8)")

    jacobi = jacobi_pyamg = gauss_seidel = gauss_seidel_pyamg = distributive_gauss_seidel = _out_of_scope
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
