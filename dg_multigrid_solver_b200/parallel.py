"""Element-slab data parallelism: one process per GPU, torch.distributed (NCCL over NVLink; gloo for the
CPU tests of the host logic).

Partitioning: element numbering m = j*Ni + i makes contiguous index ranges slabs
of whole j-rows; rank g owns rows [g*Nj/G, (g+1)*Nj/G).  The 5-point block stencil couples a slab only to
the edge rows of its neighbours, so each level carries ONE ghost element row per interior slab edge
(`DGB_FLAG_GHOST_LO/HI`): ghost rows have vector entries (the halo) but no matrix rows.

  per operator pass     halo exchange of one element row per edge (Ni*b*8 bytes) -- NCCL send/recv
  norms                 all-reduce of one fp64 (sum of squares over owned rows)
  p-transfer            local;  h-transfer: local (slabs hold whole 2x2 child groups)
  coarse levels         levels with fewer than `min_rows` rows per rank are gathered to rank 0, which runs
                        the rest of the V-cycle with the single-GPU driver (dgb_vcycle) and scatters the
                        correction back (north_star: "coarsest level gathered to one GPU")

Two transports for the same schedule:
  native (default on one GPU per rank)   the whole V-cycle is ONE call into libdgb200 (dgb_vcycle_slab): halo
                        exchange, norm reduction and the coarse gather are kernels that store straight into the
                        peers' memory over NVLink (CUDA IPC mapped arenas, dgb_comm_*); the levels below the
                        distributed ones are replicated on every rank (all-gather instead of gather + scatter).
                        Only `redblack` and `slab_lexicographic` run natively.
  torch.distributed     NCCL send/recv + all_reduce sequenced from Python (also the gloo path of the tests, and the
                        exact `lexicographic` pipeline)

Smoother orderings across slabs (`solver.b200.gs mode`):
  lexicographic        exact global lexicographic order: slab g sweeps after it received slab g-1's edge
                       row (a pipeline across ranks: exact, no parallel speed-up of the sweep itself)
  slab_lexicographic   lexicographic inside each slab, halo from the neighbour's previous pass
                       (block-Jacobi coupling between slabs; not the reference's iteration)
  redblack             2-colour sweep, one halo exchange per colour
"""
import ctypes
import os

import numpy as np

from . import _lib


class SlabPartition:
    """Rows [j0, j1) of an Nj-row element grid owned by `rank` of `world`."""

    def __init__(self, Nj, world, rank):
        if Nj % world != 0:
            raise ValueError(f"{Nj} element rows cannot be split evenly over {world} ranks")
        self.Nj, self.world, self.rank = Nj, world, rank
        self.rows = Nj // world
        self.j0, self.j1 = rank * self.rows, (rank + 1) * self.rows
        self.has_lo, self.has_hi = rank > 0, rank < world - 1


def default_min_rows(world):
    """Levels with fewer element rows per rank than this are replicated instead of distributed (measured on 8 B200,
    profiles/r02_bench_2048_n8_min_rows.md: a distributed level costs ~36 collectives of a few microseconds per
    cycle whatever its size, a replicated 128^2 level less than that)."""
    return 64


def distributed_levels(Nj, world, h_factors, min_rows=8):
    """Which h-coarsening factors stay distributed: every distributed level needs >= min_rows rows per
    rank and an even row count per rank on the next finer level (2x2 children stay inside a slab)."""
    keep = []
    for cf in sorted(h_factors):
        rows = Nj // world // cf
        if Nj % (world * cf) == 0 and rows >= min_rows:
            keep.append(cf)
        else:
            break
    return keep


def slab_nodes(xn, yn, Pg, part, halo):
    """Node rows of a slab with `halo` fine element rows of overlap per interior edge.
    xn, yn: full grid in Plot3D file order [jl][il].  Returns (xn_loc, yn_loc, halo_lo, halo_hi)."""
    lo = halo if part.has_lo else 0
    hi = halo if part.has_hi else 0
    r0, r1 = (part.j0 - lo) * Pg, (part.j1 + hi) * Pg + 1
    return np.ascontiguousarray(xn[r0:r1]), np.ascontiguousarray(yn[r0:r1]), lo, hi


def _host_staged(group=None):
    """True when the process group cannot move device tensors itself (gloo): the slab code then stages its
    messages through host memory.  Used by the tests to run two ranks on ONE GPU; production runs use NCCL."""
    import torch.distributed as dist
    return dist.get_backend(group) == "gloo"


def _p2p(sends, recvs, group=None):
    """sends: [(device tensor, dst)], recvs: [(device tensor, src)] -- one batch of point-to-point messages."""
    import torch.distributed as dist
    if not sends and not recvs:
        return
    if _host_staged(group):
        hs = [(t.cpu(), dst) for t, dst in sends]
        hr = [(t.new_empty(t.shape, device="cpu"), src, t) for t, src in recvs]
        reqs = [dist.isend(h, dst, group=group) for h, dst in hs] + [dist.irecv(h, src, group=group) for h, src, _ in hr]
        for q in reqs:
            q.wait()
        for h, _, t in hr:
            t.copy_(h)
        return
    ops = [dist.P2POp(dist.isend, t, dst, group) for t, dst in sends] + \
          [dist.P2POp(dist.irecv, t, src, group) for t, src in recvs]
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def _all_reduce(t, group=None, op=None):
    import torch.distributed as dist
    op = op or dist.ReduceOp.SUM
    if _host_staged(group):
        h = t.cpu()
        dist.all_reduce(h, op=op, group=group)
        t.copy_(h)
    else:
        dist.all_reduce(t, op=op, group=group)


def exchange_halo(vec, Ni, b, ghost_lo, ghost_hi, rank, world, group=None, upward=True, downward=True):
    """Fill the ghost rows of `vec` ([Nj_ext*Ni*b], ghost rows first/last) from the neighbour slabs.
    upward:   my last owned row  -> rank+1's lower ghost row (and I receive rank-1's into my lower ghost)
    downward: my first owned row -> rank-1's upper ghost row (and I receive rank+1's into my upper ghost)"""
    import torch.distributed as dist
    row = Ni * b
    n = vec.numel() // row
    sends, recvs = [], []
    if upward:
        if ghost_hi:
            sends.append((vec[(n - 2) * row:(n - 1) * row], rank + 1))
        if ghost_lo:
            recvs.append((vec[0:row], rank - 1))
    if downward:
        if ghost_lo:
            sends.append((vec[row:2 * row], rank - 1))
        if ghost_hi:
            recvs.append((vec[(n - 1) * row:n * row], rank + 1))
    _p2p(sends, recvs, group)


def gather_rows(local_owned, world, rank, group=None):
    """Concatenate the ranks' owned chunks (equal sizes) on rank 0 -> global vector in element order."""
    import torch
    import torch.distributed as dist
    if _host_staged(group):
        h = local_owned.cpu()
        out = [torch.empty_like(h) for _ in range(world)] if rank == 0 else None
        dist.gather(h, out, dst=0, group=group)
        return torch.cat(out).to(local_owned.device) if rank == 0 else None
    out = [torch.empty_like(local_owned) for _ in range(world)] if rank == 0 else None
    dist.gather(local_owned, out, dst=0, group=group)
    return torch.cat(out) if rank == 0 else None


def scatter_rows(global_vec, like, world, rank, group=None):
    import torch
    import torch.distributed as dist
    if _host_staged(group):
        out = torch.empty(like.shape, dtype=like.dtype, device="cpu")
        chunks = [c.contiguous() for c in global_vec.cpu().chunk(world)] if rank == 0 else None
        dist.scatter(out, chunks, src=0, group=group)
        return out.to(like.device)
    out = torch.empty_like(like)
    chunks = list(global_vec.chunk(world)) if rank == 0 else None
    dist.scatter(out, chunks, src=0, group=group)
    return out


class DistributedSolver:
    """Multigrid V-cycle over element slabs.  `local` is this rank's DGFEM-like object holding the
    distributed levels (grids with ghost rows); `coarse` (rank 0 only) the single-GPU Solver of the
    gathered levels."""

    def __init__(self, settings, grids, R_ops, P_ops, types, coarse_solver, part, gs_mode="redblack",
                 check_residual=True, group=None):
        torch = _lib.require_cuda()
        self.settings, self.grids, self.R, self.P, self.types = settings, grids, R_ops, P_ops, types
        self.coarse, self.part, self.group = coarse_solver, part, group
        self.gs_mode, self.check = gs_mode, bool(check_residual)
        self.rank, self.world = part.rank, part.world
        L = _lib.load()
        self.L = L
        self.partials = torch.zeros(L.dgb_partials_len(), dtype=torch.float64, device="cuda")
        self.sumsq = torch.zeros(1, dtype=torch.float64, device="cuda")
        self.ctl = torch.zeros(32 * (len(grids) + 1), dtype=torch.uint8, device="cuda")
        self.vec = []
        self.ops = []
        for g in grids:
            n = g.Ni * g.Nj * g.b
            self.vec.append(tuple(torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(3)))
            self.ops.append(g.operator())
        self.dR = [torch.from_numpy(np.ascontiguousarray(R)).cuda() for R in R_ops]
        self.dP = [torch.from_numpy(np.ascontiguousarray(P)).cuda() for P in P_ops]
        mg = settings.solver.multigrid
        self.sched = []
        for k in range(len(grids)):
            kind = types[k]                       # coarsening that links level k to the next coarser one
            blk = getattr(mg, f"{kind}_coarsening")
            self.sched.append((blk.pre_smoother, blk.post_smoother))
        self.n_owned = [g.Ni * (g.Nj - g.ghost_lo - g.ghost_hi) * g.b for g in grids]
        # vector that carries the restricted residual of the coarsest distributed level to rank 0
        g0 = grids[0]
        R0 = R_ops[0]
        self.bc = R0.shape[0]
        self.kind0 = _lib.TRANSFER_H if types[0] == "geometric" else _lib.TRANSFER_P
        rows0 = g0.Nj - g0.ghost_lo - g0.ghost_hi
        self.c_rows = rows0 // 2 if self.kind0 == _lib.TRANSFER_H else rows0
        self.c_Ni = g0.Ni // 2 if self.kind0 == _lib.TRANSFER_H else g0.Ni
        self.c_rhs = torch.zeros(self.c_rows * self.c_Ni * self.bc, dtype=torch.float64, device="cuda")
        self.native, self.comm = False, None
        self._graph, self._graph_failed, self._native_calls = None, False, 0
        self.graph_launches = self.graph_replays = 0

    # ---- native transport (peer memory, dgb_comm_* / dgb_vcycle_slab) ---------------------------------
    @staticmethod
    def _arena_tensor(ptr, n):
        """float64 CUDA tensor view of n doubles at device address ptr (arena memory owned by libdgb200)."""
        torch = _lib.require_cuda()

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(raw, device="cuda")

    def enable_native(self):
        """Move the iterates of the distributed levels into a peer-mapped arena and describe the hierarchy to
        dgb_vcycle_slab.  Needs the replicated coarse hierarchy on EVERY rank (build_distributed(replicate=True))."""
        import torch
        import torch.distributed as dist
        if self.gs_mode not in ("redblack", "slab_lexicographic"):
            raise NotImplementedError("the native slab V-cycle runs `redblack` and `slab_lexicographic`")
        if self.coarse is None:
            raise RuntimeError("native slab V-cycle: this rank has no replicated coarse hierarchy")
        L = _lib.load()
        al = lambda nbytes: (int(nbytes) + 255) & ~255      # noqa: E731
        blocks, off = [], 0
        for g in self.grids:
            rows = g.Nj - g.ghost_lo - g.ghost_hi
            nb = (rows + 2) * g.Ni * g.b * 8
            blocks.append((off, rows))
            off += al(nb)
        chunk = self.c_rows * self.c_Ni * self.bc
        gather_off = off
        off += al(self.world * chunk * 8)
        comm = ctypes.c_void_p()
        _lib.check(L.dgb_comm_create(self.rank, self.world, off, ctypes.byref(comm)), "dgb_comm_create")
        hb = L.dgb_comm_handle_bytes()
        mine = (ctypes.c_ubyte * hb)()
        _lib.check(L.dgb_comm_export(comm, mine), "dgb_comm_export")
        h = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device="cuda")
        allh = [torch.empty_like(h) for _ in range(self.world)]
        if self.world > 1:
            dist.all_gather(allh, h, group=self.group)
        else:
            allh = [h]
        blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
        _lib.check(L.dgb_comm_connect(comm, blob), "dgb_comm_connect")
        nbytes = ctypes.c_int64()
        base = L.dgb_comm_arena(comm, ctypes.byref(nbytes))
        self.comm, self._arena_base = comm, base
        # iterates -> arena blocks [ghost below | owned | ghost above]
        n = len(self.grids)
        levels = (_lib.SlabLevel * n)()
        self._keep = []
        for k, g in enumerate(self.grids):
            boff, rows = blocks[k]
            row = g.Ni * g.b
            blk = base + boff
            u_new = self._arena_tensor(blk + (1 - g.ghost_lo) * row * 8, g.Ni * g.Nj * g.b)
            u_new.copy_(self.vec[k][1])
            rhs, _, r = self.vec[k]
            self.vec[k] = (rhs, u_new, r)
            SL = levels[k]
            lev = SL.lev
            lev.op = self.ops[k]
            lev.rhs, lev.u, lev.r = rhs.data_ptr(), u_new.data_ptr(), r.data_ptr()
            pre, post = self.sched[k]
            for sm in (pre, post):
                if sm.smoother != "block_gauss_seidel_pyamg":
                    raise NotImplementedError("the native slab V-cycle smooths with block_gauss_seidel_pyamg")
            directions = {"symmetric": 0, "forward": 1, "backward": -1}
            lev.smoother = lev.post_smoother = _lib.SMOOTHER_IDS["block_gauss_seidel_pyamg"]
            lev.direction, lev.post_direction = directions[pre.direction], directions[post.direction]
            lev.pre_iterations, lev.post_iterations = int(pre.iterations), int(post.iterations)
            lev.omega, lev.post_omega = float(pre.relaxation_factor), float(post.relaxation_factor)
            if k < n - 1:               # transfer between this level (coarse side) and level k + 1: operators index k + 1
                R, P = self.dR[k + 1], self.dP[k + 1]
                lev.R, lev.P = R.data_ptr(), P.data_ptr()
                lev.nc, lev.nf = int(R.shape[0]), int(R.shape[1])
                lev.transfer_kind = _lib.TRANSFER_H if self.types[k + 1] == "geometric" else _lib.TRANSFER_P
            SL.u_block = blk
            SL.ghost_lo, SL.ghost_hi = int(g.ghost_lo), int(g.ghost_hi)
            SL.colour_shift = int(self._colour_shift(g)) & 1
            SL.n_global = int(self.n_owned[k] * self.world)
        # replicated hierarchy: its top level's rhs is the all-gather destination
        H = self.coarse.hierarchy()
        top = H["n"] - 1
        gather_ptr = base + gather_off
        H["levels"][top].rhs = gather_ptr
        self._coarse_rhs = self._arena_tensor(gather_ptr, self.world * chunk)
        opts = _lib.SlabOpts()
        opts.gs_mode = _lib.GS_REDBLACK if self.gs_mode == "redblack" else _lib.GS_SLAB_LEXICOGRAPHIC
        opts.check_residual = 1 if self.check else 0
        opts.link_kind = self.kind0
        opts.link_nc, opts.link_nf = int(self.dR[0].shape[0]), int(self.dR[0].shape[1])
        opts.link_Ni_c, opts.link_rows_c = int(self.c_Ni), int(self.c_rows)
        opts.n_coarse = H["n"]
        opts.link_R, opts.link_P = self.dR[0].data_ptr(), self.dP[0].data_ptr()
        opts.link_rhs_local = self.c_rhs.data_ptr()
        opts.coarse_levels = ctypes.cast(H["levels"], ctypes.POINTER(_lib.Level))
        opts.coarse_ctl = H["ctl"].data_ptr()
        opts.coarse_opts = H["opts"]
        self._slab_levels, self._slab_opts, self._coarse_H = levels, opts, H
        self.native = True

    def _launch_native(self):
        rc = _lib.load().dgb_vcycle_slab(self.comm, self._slab_levels, len(self.grids), ctypes.byref(self._slab_opts),
                                         self.ctl.data_ptr(), self.partials.data_ptr(), self.sumsq.data_ptr(), self._st())
        _lib.check(rc, "dgb_vcycle_slab")

    def _native_cycle(self):
        """One V-cycle through dgb_vcycle_slab.  The launch sequence of a cycle is fixed (the smoother's early exit is
        a device-side flag, the collectives count their sequence numbers on the device), so from the third call on it
        is replayed as ONE CUDA graph: at 8 GPUs a rank's ~400 short launches per cycle are launch-bound otherwise."""
        import torch
        if self._graph is not None:
            self._graph.replay()
            self.graph_replays += 1
            return
        self._native_calls += 1
        if self._native_calls < 3 or os.environ.get("DGB_MGPU_GRAPH", "1") == "0" or self._graph_failed:
            self._launch_native()
            return
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            L = _lib.load()
            before = L.dgb_launch_count(0)
            with torch.cuda.graph(g):
                self._launch_native()
            self.graph_launches = int(L.dgb_launch_count(0) - before)     # kernels one replay launches
            self._graph = g
        except Exception:                       # capture is an optimisation: fall back to plain launches
            self._graph_failed = True
            torch.cuda.synchronize()
            self._launch_native()
            return
        self._graph.replay()
        self.graph_replays += 1

    def close(self):
        self._graph = None
        if getattr(self, "comm", None) is not None:
            _lib.require_cuda().cuda.synchronize()
            self.vec = [(rhs, None, r) for rhs, _, r in self.vec]      # views into the arena die with it
            self._coarse_rhs = None
            _lib.load().dgb_comm_destroy(self.comm)
            self.comm = None
            self.native = False

    def check_native_error(self):
        if getattr(self, "comm", None) is not None:
            err = _lib.load().dgb_comm_error(self.comm, 1)
            if err:
                raise _lib.DgbError(f"a peer did not arrive at a collective within the time limit (dgb_comm_error={err})")
            _lib.check_device_error([g.d_mailbox for g in self.grids])

    # ---- building blocks -------------------------------------------------------------------
    def _st(self):
        return _lib.stream_ptr()

    def _halo(self, k, v, upward=True, downward=True):
        g = self.grids[k]
        exchange_halo(v, g.Ni, g.b, g.ghost_lo, g.ghost_hi, self.rank, self.world, self.group, upward, downward)

    def residual_sumsq(self, k, rhs, u, r=None, skip=None, first_direction=0):
        """Global sum of squares of rhs - A u over owned rows (device scalar self.sumsq).
        first_direction != 0: this is the entry residual of a lexicographic smoother call whose first pass runs in
        that direction -- it shares the launch of that pass's dependency-free part; returns True when it did."""
        import torch.distributed as dist
        if getattr(self, "native", False) and u.data_ptr() == self.vec[k][1].data_ptr() and first_direction == 0:
            g = self.grids[k]
            L = _lib.load()
            _lib.check(L.dgb_halo_exchange(self.comm, self._slab_levels[k].u_block, g.Ni * g.b,
                                           g.Nj - g.ghost_lo - g.ghost_hi, self._st()), "dgb_halo_exchange")
            _lib.call("dgb_bsr_residual", self.ops[k], rhs, u, r, self.partials, self.sumsq, skip, self._st())
            _lib.check(L.dgb_allreduce_sum(self.comm, self.sumsq.data_ptr(), 0, None, 0, self._st()), "dgb_allreduce_sum")
            self._entry_fused = False
            return self.sumsq
        self._halo(k, u)
        fused = False
        if first_direction != 0 and skip is None:
            L = _lib.load()
            rc = L.dgb_block_gs_entry_residual(ctypes.byref(self.ops[k]), _lib.ptr(rhs), _lib.ptr(u), first_direction,
                                               _lib.ptr(r), _lib.ptr(self.partials), _lib.ptr(self.sumsq), self._st())
            if rc != _lib.UNSUPPORTED:
                _lib.check(rc, "dgb_block_gs_entry_residual")
                fused = True
        if not fused:
            _lib.call("dgb_bsr_residual", self.ops[k], rhs, u, r, self.partials, self.sumsq, skip, self._st())
        if self.world > 1:
            _all_reduce(self.sumsq, self.group)
        self._entry_fused = fused
        return self.sumsq

    def _ctl_ptr(self, k):
        return self.ctl.data_ptr() + 32 * k

    def _gs_pass(self, k, rhs, u, direction, skip, prev=0):
        """One directional pass.  prev: direction of the previous pass of the same smoother call (0 = first):
        an opposite previous pass lets the chained kernel skip its dependency-free launch."""
        import torch.distributed as dist
        g, op, st = self.grids[k], self.ops[k], self._st()
        if self.gs_mode == "redblack":
            for c in (0, 1):
                self._halo(k, u)
                _lib.call("dgb_block_gs_colour", op, rhs, u, c if direction > 0 else 1 - c, self._colour_shift(g),
                          skip, st)
            return
        if self.gs_mode == "slab_lexicographic" or self.world == 1:
            self._halo(k, u)
            _lib.call("dgb_block_gs_pass_seq", op, rhs, u, direction, prev, skip, st)
            return
        # exact global lexicographic order: pipeline across ranks
        self._halo(k, u, upward=(direction < 0), downward=(direction > 0))     # old values of the slab ahead
        row = g.Ni * g.b
        n = u.numel() // row
        first, last = (self.rank == 0, self.rank == self.world - 1) if direction > 0 else \
                      (self.rank == self.world - 1, self.rank == 0)
        src = self.rank - 1 if direction > 0 else self.rank + 1
        dst = self.rank + 1 if direction > 0 else self.rank - 1
        if not first:
            ghost = u[0:row] if direction > 0 else u[(n - 1) * row:n * row]
            _p2p([], [(ghost, src)], self.group)
        _lib.call("dgb_block_gs_pass_seq", op, rhs, u, direction, prev, skip, st)
        if not last:
            edge = u[(n - 2) * row:(n - 1) * row] if direction > 0 else u[row:2 * row]
            _p2p([(edge, dst)], [], self.group)

    def _colour_shift(self, g):
        """Global colour (i+j_global)&1 from the local row index: j_global = j_local - ghost_lo + j0/cf."""
        cf = g.coarsening_factor or 1
        return (self.part.j0 // cf - g.ghost_lo) & 1

    def smooth(self, k, rhs, u, spec, iterations, r_keep=None):
        """One smoother call.  r_keep: every residual test also stores the residual vector there; returns True
        when r_keep holds rhs - A u of the returned u (the caller then skips its own residual evaluation)."""
        name = spec.smoother
        direction = {"symmetric": 0, "forward": 1, "backward": -1}[spec.direction]
        g, st = self.grids[k], self._st()
        n_global = self.n_owned[k] * self.world
        if name == "block_gauss_seidel_pyamg":
            skip = None
            prev = 0
            if self.check:
                first = +1 if direction >= 0 else -1
                lexi = self.gs_mode != "redblack" and int(iterations) > 0
                self.residual_sumsq(k, rhs, u, r=r_keep, first_direction=first if lexi else 0)
                if self._entry_fused:
                    prev = -first                             # the first pass finds its right-hand sides in place
                _lib.call("dgb_smoother_begin", self._ctl_ptr(k), self.sumsq, n_global, st)
                skip = self._ctl_ptr(k) + 16                 # &ctl->skip
            for _ in range(int(iterations)):
                if direction >= 0:
                    self._gs_pass(k, rhs, u, +1, skip, prev)
                    prev = +1
                if direction <= 0:
                    self._gs_pass(k, rhs, u, -1, skip, prev)
                    prev = -1
                if self.check:
                    self.residual_sumsq(k, rhs, u, r=r_keep, skip=skip)
                    _lib.call("dgb_smoother_check", self._ctl_ptr(k), self.sumsq, n_global, st)
            return bool(self.check and r_keep is not None and int(iterations) > 0)
        elif name == "block_jacobi":
            tmp = self.vec[k][2]
            self._halo(k, u)
            _lib.call("dgb_block_relax_sweep", self.ops[k], rhs, u, tmp, float(spec.relaxation_factor), st)
            u.copy_(tmp)
            if int(iterations) > 1:
                raise NotImplementedError("block_jacobi with > 1 iteration turns into forward GS in the reference "
                                          "(App. B.1); use block_gauss_seidel_pyamg across slabs")
        else:
            raise NotImplementedError(f"smoother {name} is not available across slabs")

    # ---- the cycle -------------------------------------------------------------------------
    def vcycle(self, k=None):
        """One V-cycle on distributed level k (default: finest).  vec[k] = (rhs, u, r)."""
        import torch.distributed as dist
        if getattr(self, "native", False) and (k is None or k == len(self.grids) - 1):
            self._native_cycle()
            return
        k = len(self.grids) - 1 if k is None else k
        g, st = self.grids[k], self._st()
        rhs, u, r = self.vec[k]
        pre, post = self.sched[k]
        # the pre-smoother's last residual test already evaluated the vector the restriction needs (solver.py:150)
        if not self.smooth(k, rhs, u, pre, pre.iterations, r_keep=r):
            self.residual_sumsq(k, rhs, u, r=r)
        kind = _lib.TRANSFER_H if self.types[k] == "geometric" else _lib.TRANSFER_P
        if k > 0:
            c = self.grids[k - 1]
            crhs, cu, _ = self.vec[k - 1]
            _lib.call("dgb_restrict_slab", kind, self.dR[k], self.dR[k].shape[0], self.dR[k].shape[1], c.Ni, c.Nj,
                      c.ghost_lo, c.ghost_hi, g.ghost_lo, r, crhs, st)
            cu.zero_()
            self.vcycle(k - 1)
            _lib.call("dgb_prolong_add_slab", kind, self.dP[k], self.dP[k].shape[1], self.dP[k].shape[0], c.Ni, c.Nj,
                      c.ghost_lo, c.ghost_hi, g.ghost_lo, cu, u, st)
        else:
            # restrict into an un-ghosted local vector, gather on rank 0, finish the cycle there
            _lib.call("dgb_restrict_slab", kind, self.dR[0], self.dR[0].shape[0], self.dR[0].shape[1], self.c_Ni,
                      self.c_rows, 0, 0, g.ghost_lo, r, self.c_rhs, st)
            full = gather_rows(self.c_rhs, self.world, self.rank, self.group) if self.world > 1 else self.c_rhs
            cu_full = None
            if self.rank == 0:
                cs = self.coarse
                nlev = len(cs.grids)
                cu_full = cs.multigrid_V_cycle(nlev, full, full.new_zeros(full.numel()))
            cu = scatter_rows(cu_full, self.c_rhs, self.world, self.rank, self.group) if self.world > 1 else cu_full
            _lib.call("dgb_prolong_add_slab", kind, self.dP[0], self.dP[0].shape[1], self.dP[0].shape[0], self.c_Ni,
                      self.c_rows, 0, 0, g.ghost_lo, cu, u, st)
        self.smooth(k, rhs, u, post, post.iterations)

    def residual_rms(self):
        k = len(self.grids) - 1
        rhs, u, _ = self.vec[k]
        ss = self.residual_sumsq(k, self.grids[k].d_rhs, u)
        return float(np.sqrt(ss.item() / (self.n_owned[k] * self.world)))

    def solve(self, tol=1e-6, max_cycles=100):
        """Solver.solve_multigrid (dgfem/solver.py:114-126) across slabs; returns the residual history."""
        k = len(self.grids) - 1
        rhs, u, _ = self.vec[k]
        rhs.copy_(self.grids[k].d_rhs)
        u.zero_()
        hist = []
        r0 = self.residual_rms()
        n = 0
        while n < max_cycles:
            res = self.residual_rms() / r0
            hist.append(res)
            if res < tol or not np.isfinite(res):
                break
            self.vcycle()
            n += 1
        return hist


def native_transport_possible(group=None):
    """One GPU per rank (kernels of different ranks wait for each other: they must not share a device) and a
    device-capable process group."""
    import torch
    import torch.distributed as dist
    if os.environ.get("DGB_MGPU_NATIVE", "1") == "0" or os.environ.get("DGB_MGPU_ONE_DEVICE") == "1":
        return False
    if not dist.is_initialized() or _host_staged(group):
        return False
    world = dist.get_world_size(group)
    dev = torch.tensor([torch.cuda.current_device()], device="cuda")
    devs = [torch.empty_like(dev) for _ in range(world)]
    dist.all_gather(devs, dev, group=group)
    ids = [int(d.item()) for d in devs]
    return len(set(ids)) == world and torch.cuda.device_count() >= world


def build_distributed(settings, xn, yn, world, rank, min_rows=8, group=None, gs_mode=None, native=None):
    """Build this rank's slab hierarchy (+ the gathered coarse hierarchy on rank 0).
    xn, yn: the FULL grid's nodes in Plot3D file order (every rank reads / generates them)."""
    from .dgfem import DGFEM, _int_list
    from .discrete_system import DiscreteSystem
    from .grid import CoarseGrid, Geometry, Grid
    from .settings import Settings
    from .tables import h_restriction, p_restriction
    torch = _lib.require_cuda()
    Pg = settings.grid.polynomial_degree
    Nj = (xn.shape[0] - 1) // Pg
    part = SlabPartition(Nj, world, rank)
    mg = settings.solver.multigrid
    p_levels = sorted(_int_list(mg.polynomial_coarsening.levels.u))
    factors = sorted(_int_list(mg.geometric_coarsening.coarsening_factors)) if mg.geometric_coarsening.enabled else []
    dist_f = distributed_levels(Nj, world, factors, min_rows)
    coarse_f = [f for f in factors if f not in dist_f]
    halo = max(dist_f) if dist_f else 1
    lx, ly, hlo, hhi = slab_nodes(xn, yn, Pg, part, halo)
    geo = Geometry(None, settings, nodes=(lx, ly))
    geo.halo_lo, geo.halo_hi = hlo, hhi
    mult = settings.problem.SIP_penalty_parameter_multiplier
    grids, R_ops, P_ops, types = [], [], [], []
    for p in p_levels:
        grids.append(Grid(geo, ["u"]).initialize({"u": p}, (p + 1) ** 2 * mult))
    for k in range(len(p_levels) - 1):
        R = p_restriction(p_levels[k], p_levels[k + 1])
        R_ops.append(R); P_ops.append(R.T.copy()); types.append("polynomial")
    base = grids[0]
    Rh, Ph = h_restriction()
    hl = [CoarseGrid(geo, base, ["u"]).initialize(coarsening_factor=cf) for cf in sorted(dist_f, reverse=True)]
    grids[0:0] = hl
    R_ops[0:0] = [Rh for _ in hl]; P_ops[0:0] = [Ph for _ in hl]; types[0:0] = ["geometric" for _ in hl]
    # the link from the coarsest distributed level down to the gathered hierarchy
    if coarse_f:
        R_ops.insert(0, Rh); P_ops.insert(0, Ph); types.insert(0, "geometric")
    else:
        raise NotImplementedError("at least one gathered (coarse) level is required")
    ds = DiscreteSystem(settings)
    for g in grids:
        ds.problem.assemble(g)
        g.release_geometry()
    mode = gs_mode or settings.get("solver.b200.gs_mode", "redblack")
    if native is None:
        # one rank: the same C++ driver and collective kernels with nobody to wait for (what a 1-GPU box can test)
        native = mode in ("redblack", "slab_lexicographic") and os.environ.get("DGB_MGPU_NATIVE", "1") != "0" and \
            (world == 1 or native_transport_possible(group))
    coarse_solver = None
    if rank == 0 or native:
        full_geo = Geometry(None, settings, nodes=(xn, yn))
        from .solver import Solver
        cgrids, cR, cP, ctypes_ = [], [], [], []
        base_full = Grid(full_geo, ["u"])       # only its solution space / sigma are read by CoarseGrid
        base_full._set_solution_space({"u": p_levels[0]}, (p_levels[0] + 1) ** 2 * mult, None)
        for cf in sorted(coarse_f, reverse=True):
            cgrids.append(CoarseGrid(full_geo, base_full, ["u"]).initialize(coarsening_factor=cf))
        for _ in cgrids[:-1]:
            cR.append(Rh); cP.append(Ph); ctypes_.append("geometric")
        for g in cgrids:
            ds.problem.assemble(g)
            g.release_geometry()
        del base_full
        coarse_solver = Solver("multigrid", settings)
        coarse_solver.grids = cgrids
        coarse_solver.restriction_operators, coarse_solver.prolongation_operators = cR, cP
        coarse_solver.multigrid_type = ctypes_ if ctypes_ else ["geometric"]
        full_geo._dev = None
        torch.cuda.empty_cache()
    ds = DistributedSolver(settings, grids, R_ops, P_ops, types, coarse_solver, part, gs_mode=mode,
                           check_residual=settings.get("solver.b200.check_residual", True), group=group)
    if native:
        ds.enable_native()
    return ds


def run_bench_multi_gpu(args, bench):
    """bench.py for N > 1 (torchrun, one rank per GPU): the same 2048^2 p=2 V-cycle, element slabs across
    ranks (strong scaling).  Timing: barrier + synchronize on both sides, max over ranks; rank 0 prints."""
    import json
    import time
    import torch
    import torch.distributed as dist
    from .settings import Settings
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, p = args.size, args.p
    # N > 1 default: the 2-colour sweep (bandwidth-bound, the same iteration on any number of GPUs; pinned by the
    # oracle's red_black_gauss_seidel).  The exact lexicographic order is a pipeline across slabs (--exact-multi),
    # slab_lexicographic keeps the wavefront latency of Ni steps per pass on every rank.
    mode = args.gs_mode if args.gs_mode != "lexicographic" or args.exact_multi else "redblack"
    settings = Settings(bench.make_params(n, p, mode, bool(args.check_residual)))
    settings.update_setting("solver.method", "multigrid")
    xn, yn = bench.rectangle_nodes_file_order(n, p)
    t0 = time.perf_counter()
    mr = getattr(args, "min_rows", 0) or default_min_rows(world)
    ds = build_distributed(settings, xn, yn, world, rank, min_rows=mr, gs_mode=mode)
    del xn, yn
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    k = len(ds.grids) - 1
    rhs, u, _ = ds.vec[k]
    rhs.copy_(ds.grids[k].d_rhs)
    u.zero_()
    r0 = ds.residual_rms()
    for _ in range(args.warmup):
        ds.vcycle()
    dist.barrier()
    torch.cuda.synchronize()
    L = _lib.load()
    L.dgb_launch_count(1)
    replays0 = ds.graph_replays
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = bench.ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0.record()
    for _ in range(args.steps):
        ds.vcycle()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # kernels launched inside the timed region: direct launches + what the graph replays launched
    launches = torch.tensor([float(L.dgb_launch_count(0) + (ds.graph_replays - replays0) * ds.graph_launches)],
                            dtype=torch.float64, device="cuda")
    dist.all_reduce(launches)
    clocks = sampler.stop() if sampler else None
    res = ds.residual_rms() / r0
    ds.check_native_error()
    ms_per_step = float(ms.item()) / args.steps
    # end to end: host RHS/u chunks in, u chunk out, every step
    n_own = ds.n_owned[k]
    g = ds.grids[k]
    off = g.ghost_lo * g.Ni * g.b
    h_rhs = torch.empty(n_own, dtype=torch.float64, pin_memory=True)
    h_u = torch.zeros(n_own, dtype=torch.float64, pin_memory=True)
    h_out = torch.empty(n_own, dtype=torch.float64, pin_memory=True)
    h_rhs.copy_(g.d_rhs[off:off + n_own])
    e2e_steps = max(2, min(args.steps, 5))
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(e2e_steps):
        rhs[off:off + n_own].copy_(h_rhs, non_blocking=True)
        u[off:off + n_own].copy_(h_u, non_blocking=True)
        ds.vcycle()
        h_out.copy_(u[off:off + n_own], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2.item()) / e2e_steps
    if rank == 0:
        cfg = bench.workload_config(args)
        cfg["gs_mode"] = mode
        cfg["transport"] = ("native: dgb_vcycle_slab, halo/all-reduce/all-gather kernels over peer memory (NVLink)" +
                            (", one CUDA graph per cycle" if ds._graph is not None else "")) \
            if ds.native else "torch.distributed (NCCL send/recv, all_reduce) sequenced from Python"
        cfg["partition"] = (f"{world} slabs of {n // world} element rows, halo exchange per pass, levels with < {mr} "
                            f"rows/rank {'replicated on every rank (all-gather)' if ds.native else 'gathered to rank 0'}")
        line = {"metric": "multigrid_vcycles_per_s", "value": 1e3 / ms_per_step, "unit": "V-cycles/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "clocks": clocks, "gpu_launches": int(launches.item()),
                "e2e": {"value": 1e3 / e2e_ms, "unit": "V-cycles/s", "h2d_bytes_per_step": 2 * n_own * 8 * world,
                        "d2h_bytes_per_step": n_own * 8 * world, "ms_per_step": e2e_ms, "steps": e2e_steps,
                        "api": "DistributedSolver.vcycle with pinned host RHS/u slabs in, u slab out, per rank"},
                "roofline": None, "cpu_baseline": None,
                "vcycle": {"normalised_residual_after_timed_cycles": res, "cycles_run": args.warmup + args.steps},
                "vcycle_dof_per_s": n * n * (p + 1) ** 2 * 1e3 / ms_per_step, "setup_s": setup_s}
        print(json.dumps(line), flush=True)
    ds.check_native_error()
    ds.close()
    dist.destroy_process_group()
