"""DiscreteSystem / Poisson: assembly of the DG operator and right-hand side on the device.

Same surface as dgfem/discrete_system.py:11-52: `DiscreteSystem(settings).problem.assemble(grid)`
fills grid.BSR (lazily materialised scipy view of the device BSR arrays) and grid.RHS.
The work is done by dgb_assemble_poisson / dgb_assemble_rhs / dgb_block_diag_inverse
(include/dgb200.h).
"""
import ctypes
import os

import numpy as np

from . import _lib
from .mms import PoissonMMS


class DiscreteSystem:
    def __init__(self, settings):
        self.problem = self.select_problem(settings)

    def select_problem(self, settings):
        kind = settings.problem.type.lower()
        if kind == "poisson":
            return Poisson(settings)
        if kind == "stokes":
            from .stokes import Stokes
            return Stokes(settings)
        raise NotImplementedError(f"There exists no implementation for the {settings.problem.type} equation(s)")


def stencil_flags(grid, settings):
    f = 0
    if grid.O_grid or grid.fully_periodic_boundaries:
        f |= _lib.FLAG_PERIODIC_I                       # discrete_system.py:84,95
    if grid.fully_periodic_boundaries:
        f |= _lib.FLAG_PERIODIC_J                       # discrete_system.py:106,117
    if settings.problem.multiply_inverse_mass_matrix:
        f |= _lib.FLAG_MINV                             # discrete_system.py:139
    if getattr(grid, "ghost_lo", 0):
        f |= _lib.FLAG_GHOST_LO                         # slab partitioning (parallel.py)
    if getattr(grid, "ghost_hi", 0):
        f |= _lib.FLAG_GHOST_HI
    return f


class Poisson:
    def __init__(self, settings):
        self.settings = settings
        self._mms = None

    def mms(self):
        if self._mms is None:
            self._mms = PoissonMMS(self.settings)
        return self._mms

    def assemble(self, grid):
        if grid.discretization != "dg":
            raise NotImplementedError("the FVM discretisation is out of scope (SURVEY.md section 2.1 row 19)")
        self.assemble_BSR_Poisson(grid)
        self.assemble_RHS_Poisson(grid)

    def assemble_BSR_Poisson(self, grid):
        if self.settings.problem.type != "Poisson":
            raise ValueError("The governing equation(s) field in the paramfile is not set to Poisson")
        torch = _lib.require_cuda()
        L = _lib.load()
        flags = stencil_flags(grid, self.settings)
        b = grid.N_DOF_sol["u"]
        N = grid.Ni * grid.Nj
        nnzb = int(L.dgb_poisson_nnzb(grid.Ni, grid.Nj, flags))
        from .grid import padded_blocks
        grid.d_data = padded_blocks(nnzb, b)
        grid.d_indices = torch.empty(nnzb, dtype=torch.int32, device="cuda")
        grid.d_indptr = torch.empty(N + 1, dtype=torch.int32, device="cuda")
        grid.d_minv = torch.empty((N, b, b), dtype=torch.float64, device="cuda")
        nu = float(self.settings.problem.kinematic_viscosity)
        _lib.call("dgb_assemble_poisson", grid._h_tables, grid.d_vol, grid.d_face, grid.d_area, grid.Ni, grid.Nj,
                  nu, float(grid.sigma), flags, grid.d_indptr, grid.d_indices, grid.d_data, grid.d_minv,
                  _lib.stream_ptr())
        grid.flags = flags
        grid.nnzb = nnzb
        grid._BSR = None
        grid.stencil = flags & ~_lib.FLAG_MINV                                 # structure is ours by construction
        prepare_smoother_data(grid)

    def assemble_RHS_Poisson(self, grid):
        torch = _lib.require_cuda()
        T = grid.tables
        b = grid.N_DOF_sol["u"]
        N = grid.Ni * grid.Nj
        mms = self.mms()
        f_vol = mms.source(grid.d_vol[:, 5, :], grid.d_vol[:, 6, :]).contiguous()           # discrete_system.py:375
        g_face = mms.solution(grid.d_face[:, :, 3, :], grid.d_face[:, :, 4, :]).contiguous()  # :381-394
        grid.d_rhs = torch.empty(N * b, dtype=torch.float64, device="cuda")
        nu = float(self.settings.problem.kinematic_viscosity)
        _lib.call("dgb_assemble_rhs", grid._h_tables, grid.d_vol, grid.d_face, grid.d_area, grid.d_minv,
                  f_vol, g_face, grid.Ni, grid.Nj, nu, float(grid.sigma), grid.flags, grid.d_rhs,
                  _lib.stream_ptr())
        grid._RHS = None
        assert T.b == b


# seconds spent in prepare_smoother_data since the last reset (the reference has no such phase: pyamg rebuilds the
# inverse diagonal blocks inside every smoother call); DGFEM.initialize reports it beside the assembly time
SETUP_TIMINGS = {"smoother_setup": 0.0}


def prepare_smoother_data(grid):
    """Inverse diagonal blocks and the smoother stream, once per level (the reference recomputes
    the block inverses on every smoother call: pyamg_relaxation.py:230-231)."""
    import time
    torch = _lib.require_cuda()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        return _prepare_smoother_data(grid)
    finally:
        torch.cuda.synchronize()
        SETUP_TIMINGS["smoother_setup"] += time.perf_counter() - t0


def _prepare_smoother_data(grid):
    torch = _lib.require_cuda()
    from .grid import padded_blocks
    b = grid.d_data.shape[1]
    N = grid.d_indptr.numel() - 1
    st = _lib.stream_ptr()
    grid.d_dinv = padded_blocks(N, b)              # 16 bytes of slack: rows are read in aligned 16-byte pieces
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.call("dgb_block_diag_inverse", grid.d_data, grid.d_indices, grid.d_indptr, N, b, grid.d_dinv, info, st)
    grid._dinv_info = info
    if grid.stencil >= 0:
        # belt and braces: the streaming kernels rely on the closed-form 5-point structure
        mism = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.call("dgb_check_stencil", grid.d_indices, grid.d_indptr, grid.Ni, grid.Nj, grid.stencil, mism, st)
        if int(mism.item()) != 0:
            grid.stencil = -1
    grid.d_gs = grid.d_mailbox = grid.d_chain = None
    if grid.stencil >= 0:
        # row hand-over mailbox of the lexicographic GS kernels: all-ones (sentinel NaN) outside a pass
        grid.d_mailbox = torch.full((N * b,), -1, dtype=torch.int64, device="cuda").view(torch.float64)
        # chained kernel (two launches per pass: dependency-free part + chain) where the operator allows it,
        # else the row-pipelined kernel on a copy of the matrix with inverted diagonal blocks
        chain_len = int(_lib.load().dgb_gs_chain_len(b, int(grid.Ni), int(grid.Nj), int(grid.stencil)))
        if chain_len > 0:
            grid.d_chain = torch.empty(chain_len, dtype=torch.float64, device="cuda")
            op = grid.operator()
            _lib.call("dgb_build_gs_chain", op, st)
        nbytes = int(grid.d_indices.numel()) * b * b * 8
        if chain_len == 0 or os.environ.get("DGB_GS_STREAM") == "1" or nbytes < (64 << 20):
            grid.d_gs = padded_blocks(int(grid.d_indices.numel()), b)
            _lib.call("dgb_build_gs_stream", grid.d_data, grid.d_indices, grid.d_indptr, grid.d_dinv, N, b,
                      grid.d_gs, st)
    return grid.d_dinv


def check_dinv(grid):
    bad = int(grid._dinv_info.item())
    if bad:
        raise np.linalg.LinAlgError(f"singular diagonal block in block row {bad - 1}")
