"""Solver: multigrid / smoother drivers on the device.

Same surface as dgfem/solver.py:13-207 -- attributes grids, restriction_operators,
prolongation_operators, multigrid_type, residuals; methods solve(), solve_multigrid(levels,
RHS, u, tol, max_cycles), multigrid_V_cycle(k, RHS, u), solve_smoother(grid, RHS).

One V-cycle is a single call into libdgb200 (dgb_vcycle): every kernel of the cycle is enqueued
on the current stream with no host synchronisation; the smoother's early-exit state lives in
device memory.  The only host round trip of solve_multigrid is the one scalar per cycle the
convergence test needs (dgfem/solver.py:119-123).
"""
import ctypes
import os
import pickle

import numpy as np

from . import _lib
from .relaxation import Relaxation, _to_device
from .timer import Timer


def compute_Lp_norm(delta, p):
    """utils/helpers.py:16-26 (an RMS-type norm: (sum|d|^p / n)^(1/p))."""
    return (np.sum(abs(delta) ** p) / delta.size) ** (1 / p)


class Solver:
    def __init__(self, method, settings):
        self.settings = settings
        self.method = method
        self.grids = []
        self.restriction_operators = []
        self.prolongation_operators = []
        self.multigrid_type = []
        self.residuals = []
        self._hier = None
        self.timings = {}
        self.graph_launches = self.graph_replays = 0
        self.primed_cycles = 0           # V-cycles of solve_multigrid that started from the loop's own residual

    # ------------------------------------------------------------------------------------
    def solve(self):
        reference_grid = self.grids[-1]
        with Timer() as timer:
            if self.method == "smoother":
                u = self.solve_smoother(reference_grid, reference_grid.RHS)
            elif self.method == "direct":
                u = self.solve_directly(reference_grid, reference_grid.RHS)          # dgfem/solver.py:33-35
            elif self.method == "multigrid":
                RHS_0 = reference_grid.RHS
                u_0 = np.zeros_like(RHS_0)
                mg = self.settings.solver.multigrid
                u = self.solve_multigrid(levels=len(self.grids), RHS=RHS_0, u=u_0, tol=mg.tolerance,
                                         max_cycles=mg.max_cycles)
            else:
                raise NotImplementedError(
                    f"solver method '{self.method}' is outside the B200 hot path (SURVEY.md section 2.1 row 10); "
                    "use -m, -s or -d")
        self.timings["solve"] = timer.elapsed()
        return u

    def solve_smoother(self, grid, RHS):
        """dgfem/solver.py:61-66."""
        name = self.settings.solver.smoother
        if name == "distributive_gauss_seidel":                                   # dgfem/solver.py:62-63
            return Relaxation.distributive_gauss_seidel(grid, RHS, max_iterations=1000000, splitting="lsq",
                                                        settings=self.settings)
        return getattr(Relaxation, name)(grid, RHS, max_iterations=100, direction="symmetric")

    # ------------------------------------------------------------------------------------
    def _smoother_block(self, kind):
        return getattr(self.settings.solver.multigrid, f"{kind}_coarsening")

    def _build_hierarchy(self):
        """Device level descriptors (dgb_level) for the current grids / transfer operators."""
        torch = _lib.require_cuda()
        n = len(self.grids)
        levels = (_lib.Level * n)()
        vecs, ops = [], []
        s = self.settings
        gs_mode = s.get("solver.b200.gs_mode", "lexicographic")
        chk = s.get("solver.b200.check_residual", True)
        for k, g in enumerate(self.grids):
            b = g.d_data.shape[1]
            N = g.Ni * g.Nj
            if g.d_dinv is None:
                from .discrete_system import prepare_smoother_data
                prepare_smoother_data(g)
            rhs = torch.zeros(N * b, dtype=torch.float64, device="cuda")
            u = torch.zeros(N * b, dtype=torch.float64, device="cuda")
            r = torch.zeros(N * b, dtype=torch.float64, device="cuda")
            vecs.append((rhs, u, r))
            L = levels[k]
            L.op = g.operator()
            L.rhs, L.u, L.r = rhs.data_ptr(), u.data_ptr(), r.data_ptr()
            # smoother settings come from the coarsening that links this level to the next coarser
            # one; the coarsest level uses the first link's (dgfem/solver.py:143,202)
            kind = self.multigrid_type[k - 1] if k > 0 else (self.multigrid_type[0] if self.multigrid_type else "polynomial")
            blk = self._smoother_block(kind)
            # pre- and post-smoother are resolved independently (dgfem/solver.py:143-147,196)
            pre, post = blk.pre_smoother, blk.post_smoother
            for sm in (pre, post):
                if sm.smoother not in _lib.SMOOTHER_IDS:
                    raise AttributeError(f"Relaxation has no accelerated smoother '{sm.smoother}'")
            directions = {"symmetric": 0, "forward": 1, "backward": -1}
            L.smoother = _lib.SMOOTHER_IDS[pre.smoother]
            L.direction = directions[pre.direction]
            L.pre_iterations, L.post_iterations = int(pre.iterations), int(post.iterations)
            L.omega = float(pre.relaxation_factor)
            L.post_smoother = _lib.SMOOTHER_IDS[post.smoother]
            L.post_direction = directions[post.direction]
            L.post_omega = float(post.relaxation_factor)
            if k < n - 1:
                R = torch.from_numpy(np.ascontiguousarray(self.restriction_operators[k], dtype=np.float64)).cuda()
                P = torch.from_numpy(np.ascontiguousarray(self.prolongation_operators[k], dtype=np.float64)).cuda()
                ops.append((R, P))
                L.R, L.P = R.data_ptr(), P.data_ptr()
                L.nc, L.nf = int(R.shape[0]), int(R.shape[1])
                kind_up = self.multigrid_type[k]
                if kind_up == "geometric":
                    if s.solver.multigrid.geometric_coarsening.use_FVM:
                        raise NotImplementedError("FVM coarse levels are out of scope (SURVEY.md section 2.1 row 19)")
                    L.transfer_kind = _lib.TRANSFER_H
                elif kind_up == "polynomial":
                    L.transfer_kind = _lib.TRANSFER_P
                else:
                    raise NotImplementedError(f"multigrid type '{kind_up}' is out of scope")
        opts = _lib.VcycleOpts(gs_mode=_lib.GS_REDBLACK if gs_mode == "redblack" else _lib.GS_LEXICOGRAPHIC,
                               check_residual=1 if chk else 0, coarse_iterations=10,
                               coarse_solver=_lib.COARSE_SMOOTHER)
        coarse_inverse = None
        coarse = s.solver.multigrid.coarse_grid_solver
        if coarse == "direct":
            # dgfem/solver.py:199-200: spsolve on the coarsest level -> dense inverse, once per hierarchy
            coarse_inverse = self.coarse_direct_inverse(self.grids[0])
            opts.coarse_solver = _lib.COARSE_DIRECT
            opts.coarse_inverse = coarse_inverse.data_ptr()
        elif coarse != "smoother":
            raise NotImplementedError(f"coarse grid solver '{coarse}' is outside the B200 hot path "
                                      "(paramfile.yml:23: 'smoother' and 'direct' run on the device)")
        ctl = torch.zeros(32 * n, dtype=torch.uint8, device="cuda")
        partials = torch.zeros(_lib.load().dgb_partials_len(), dtype=torch.float64, device="cuda")
        sumsq = torch.zeros(1, dtype=torch.float64, device="cuda")
        self._hier = dict(levels=levels, n=n, vecs=vecs, ops=ops, opts=opts, ctl=ctl, partials=partials, sumsq=sumsq,
                          grids=list(self.grids), coarse_inverse=coarse_inverse)
        return self._hier

    @staticmethod
    def coarse_direct_inverse(grid):
        """Dense inverse of a level's operator (dgb_dense_inverse): the device-side `solve_directly`."""
        torch = _lib.require_cuda()
        if getattr(grid, "ordering", "local") == "global":       # Stokes, global ordering: the regrouped matrix
            grid = grid.BSR_global
        b = grid.d_data.shape[1]
        N = grid.d_indptr.numel() - 1
        inv = torch.empty((N * b, N * b), dtype=torch.float64, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.call("dgb_dense_inverse", grid.d_data, grid.d_indices, grid.d_indptr, N, b, inv, info, _lib.stream_ptr())
        bad = int(info.item())
        if bad:
            raise np.linalg.LinAlgError(f"coarse-grid operator is singular (zero pivot in column {bad - 1})")
        return inv

    def solve_directly(self, grid, RHS):
        """dgfem/solver.py:56-59 (spsolve(grid.BSR.tocsr(), RHS)) on the device; meant for coarse levels."""
        torch = _lib.require_cuda()
        d_rhs, host = _to_device(RHS)
        inv = self.coarse_direct_inverse(grid)
        u = torch.empty_like(d_rhs)
        _lib.call("dgb_dense_solve", inv, int(d_rhs.numel()), d_rhs, u, _lib.stream_ptr())
        return u.cpu().numpy() if host else u

    def hierarchy(self):
        if self._hier is None or self._hier["grids"] != list(self.grids):
            self._build_hierarchy()
        return self._hier

    def _launch_vcycle(self, H, k, primed=False):
        rc = _lib.load().dgb_vcycle_ex(H["levels"], k, ctypes.byref(H["opts"]), H["ctl"].data_ptr(),
                                       H["partials"].data_ptr(), H["sumsq"].data_ptr(), _lib.stream_ptr(),
                                       _lib.VCYCLE_ENTRY_PRIMED if primed else 0)
        if primed and rc == _lib.UNSUPPORTED:       # nothing was launched
            return False
        _lib.check(rc, "dgb_vcycle_ex")
        return True

    def _vcycle_device(self, k, primed=False):
        """One V-cycle (dgb_vcycle).  The launch sequence of a full cycle is fixed (the smoother's early exit is a
        device-side flag), so from the third call on it is replayed as one CUDA graph: the ~150 kernels of the
        coarse levels are shorter than their launch latency otherwise (`solver.b200.cuda graph: False` or
        DGB_VCYCLE_GRAPH=0 keeps plain launches).  primed: the caller has just run dgb_block_gs_entry_residual on the
        finest level (solve_multigrid's residual after the previous cycle), the cycle starts with its first pass;
        returns False (nothing launched) when the library does not support that for this hierarchy."""
        H = self.hierarchy()
        use_graph = k == H["n"] and not H["opts"].u_final_event and H.get("graph_ok", True) and \
            os.environ.get("DGB_VCYCLE_GRAPH", "1") != "0" and self.settings.get("solver.b200.cuda_graph", True)
        if not use_graph:
            return self._launch_vcycle(H, k, primed)
        gkey, ckey = ("graph_primed", "vcalls_primed") if primed else ("graph", "vcalls")
        if H.get(gkey) is not None:
            H[gkey].replay()
            self.graph_replays += 1
            return True
        H[ckey] = H.get(ckey, 0) + 1
        if H[ckey] < 3:
            return self._launch_vcycle(H, k, primed)
        torch = _lib.require_cuda()
        try:
            # capture_begin / capture_end on a side stream instead of `with torch.cuda.graph(g)`: that context manager
            # synchronises the device, runs the Python garbage collector and empties the allocator cache on entry
            # (~25 ms in a 0.4 s solve); nothing in the cycle allocates through torch
            g = torch.cuda.CUDAGraph()
            cur = torch.cuda.current_stream()
            side = H.get("capture_stream")
            if side is None:
                side = H["capture_stream"] = torch.cuda.Stream()
            side.wait_stream(cur)
            before = _lib.load().dgb_launch_count(0)
            with torch.cuda.stream(side):
                g.capture_begin()
                try:
                    self._launch_vcycle(H, k, primed)
                finally:
                    g.capture_end()
            cur.wait_stream(side)
            if not primed:
                self.graph_launches = int(_lib.load().dgb_launch_count(0) - before)  # kernels one replay launches
            H[gkey] = g
        except Exception:                    # capture is an optimisation: fall back to plain launches
            H["graph_ok"] = False
            torch.cuda.synchronize()
            return self._launch_vcycle(H, k, primed)
        H[gkey].replay()
        self.graph_replays += 1
        return True

    def _check_divergence(self):
        """Host-side view of the device state, wherever the host synchronises anyway: the sticky `diverged`
        flags of the smoother control blocks (dgfem/relaxation.py:214-216 exits the process there) and the
        error flag of the asynchronous kernels."""
        H = self.hierarchy()
        raw = H["ctl"].cpu().numpy().tobytes()
        _lib.check_device_error([g.d_mailbox for g in H["grids"]])
        for k in range(H["n"]):
            c = _lib.SmootherCtl.from_buffer_copy(raw[32 * k:32 * (k + 1)])
            if c.diverged:
                H["ctl"].zero_()                                     # the flag is sticky on the device
                print(f"diverging, residual={c.ratio:.6e}")          # dgfem/relaxation.py:214-216
                raise SystemExit()

    @staticmethod
    def _load(dst, src):
        """Copy a host (NumPy / CPU tensor, ideally pinned) or device vector into a level buffer;
        returns True when the source lives on the host."""
        torch = _lib.require_cuda()
        if isinstance(src, np.ndarray):
            src = torch.from_numpy(np.ascontiguousarray(src, dtype=np.float64))
        dst.copy_(src, non_blocking=True)
        return not src.is_cuda

    def multigrid_V_cycle(self, k, RHS, u, out=None):
        """dgfem/solver.py:141-207.  RHS/u: host vectors (NumPy arrays or CPU tensors; copied to the
        device and the result copied back) or CUDA tensors (everything stays on the device).
        `out`: optional host buffer (pinned CPU tensor / NumPy array) that receives the result."""
        torch = _lib.require_cuda()
        H = self.hierarchy()
        rhs_k, u_k, _ = H["vecs"][k - 1]
        host = self._load(rhs_k, RHS)
        self._load(u_k, u)
        if host and out is not None and k == H["n"]:
            # the result leaves on a second stream as soon as the last sweep has written it, while the
            # post-smoother's closing residual test (which only reads u) still runs on the main stream
            out_t = torch.from_numpy(out) if isinstance(out, np.ndarray) else out
            if "u_event" not in H:
                H["u_event"] = torch.cuda.Event()
                H["u_event"].record()                     # materialises the CUDA event handle
                H["copy_stream"] = torch.cuda.Stream()
            H["opts"].u_final_event = H["u_event"].cuda_event
            try:
                self._vcycle_device(k)
            finally:
                H["opts"].u_final_event = None
            main = torch.cuda.current_stream()
            with torch.cuda.stream(H["copy_stream"]):
                H["copy_stream"].wait_event(H["u_event"])
                out_t.copy_(u_k, non_blocking=True)
            main.wait_stream(H["copy_stream"])
            main.synchronize()
            self._check_divergence()
            return out
        self._vcycle_device(k)
        if host:
            if out is not None:
                out_t = torch.from_numpy(out) if isinstance(out, np.ndarray) else out
                out_t.copy_(u_k, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                self._check_divergence()
                return out
            res = u_k.cpu().numpy()
            self._check_divergence()
            return res
        return u_k.clone()

    def solve_multigrid(self, levels, RHS, u, tol=1e-6, max_cycles=100):
        """dgfem/solver.py:114-139."""
        torch = _lib.require_cuda()
        H = self.hierarchy()
        fine = self.grids[-1]
        rhs_k, u_k, r_k = H["vecs"][levels - 1]
        host = self._load(rhs_k, RHS)
        self._load(u_k, u)
        fine_rhs = fine.d_rhs
        fine_op = fine.operator()
        b = fine.d_data.shape[1]
        nrow = fine.d_indptr.numel() - 1
        n_dof = float(nrow * b)
        st = _lib.stream_ptr()

        # The residual the loop evaluates after a cycle (solver.py:119) is the entry residual the next cycle's
        # pre-smoother opens with (relaxation.py:202: same operator, rhs and u).  Where that smoother call starts with
        # the fused kernel (chained lexicographic sweep) the loop runs it itself and hands the cycle its outcome
        # (dgb_vcycle_ex, DGB_VCYCLE_ENTRY_PRIMED): one evaluation of A u on the finest level per cycle instead of two.
        L = _lib.load()
        top = H["levels"][levels - 1]
        first_dir = 1 if top.direction >= 0 else -1
        state = {"primed": levels == H["n"] and levels >= 2 and bool(torch.equal(rhs_k, fine_rhs)) and
                 os.environ.get("DGB_SOLVE_PRIME", "1") != "0"}

        def rms():
            if state["primed"]:
                rc = L.dgb_block_gs_entry_residual(ctypes.byref(fine_op), rhs_k.data_ptr(), u_k.data_ptr(), first_dir,
                                                   r_k.data_ptr(), H["partials"].data_ptr(), H["sumsq"].data_ptr(), st)
                if rc == _lib.UNSUPPORTED:
                    state["primed"] = False
                else:
                    _lib.check(rc, "dgb_block_gs_entry_residual")
            if not state["primed"]:
                _lib.call("dgb_bsr_residual", fine_op, fine_rhs, u_k, None, H["partials"], H["sumsq"], None, st)
            return float(np.sqrt(H["sumsq"].item() / n_dof))

        def cycle():
            if state["primed"]:
                if self._vcycle_device(levels, primed=True):
                    self.primed_cycles += 1
                    return
                state["primed"] = False            # this hierarchy's pre-smoother opens differently: the plain cycle
            self._vcycle_device(levels)
        n = 0
        residual_0 = rms()
        norm = residual_0                          # solver.py:117,119 evaluate the same vector twice before cycle 1
        with np.errstate(divide="ignore", invalid="ignore"):
            while n < max_cycles:
                residual = np.float64(norm) / np.float64(residual_0)
                self.residuals.append(float(residual))
                if residual < tol or np.isnan(residual) or np.isinf(residual):
                    break
                cycle()
                n += 1
                self._check_divergence()          # the reference leaves at the first diverging smoother call
                norm = rms()
        self._pickle_residuals(fine)
        return u_k.cpu().numpy() if host else u_k.clone()

    def _pickle_residuals(self, grid):
        """dgfem/solver.py:128-138 (the residual history is an observable output)."""
        try:
            path = os.path.join(os.getcwd(), "postprocessing", "pickles", "multigrid")
            os.makedirs(path, exist_ok=True)
            name = f"residuals_{self.settings.problem.type}_{grid.Ni}X{grid.Nj}_nPoly{grid.P_grid}"
            name += "_" + "_".join(sorted(set(self.multigrid_type)))
            name += "_circle" if self.settings.grid.circular else "_rectangle"
            with open(os.path.join(path, name + ".pkl"), "wb") as f:
                pickle.dump(self.residuals, f)
        except OSError:
            pass
