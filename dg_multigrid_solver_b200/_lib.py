"""ctypes binding of libdgb200.so (include/dgb200.h).  PyTorch only owns the device buffers;
every call passes raw device pointers (tensor.data_ptr()) and the current CUDA stream.

There is NO CPU fallback: if the CUDA extension is missing or no GPU is visible, the
product path raises."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DGB_LIB", os.path.join(_HERE, "libdgb200.so"))      # DGB_LIB: a diagnostic build
_lib = None

c_i32, c_i64, c_f64, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p


class DgbError(RuntimeError):
    pass


class SmootherCtl(ctypes.Structure):
    """dgb_smoother_ctl (include/dgb200.h)."""
    _fields_ = [("res0", c_f64), ("ratio", c_f64), ("skip", c_i32), ("diverged", c_i32),
                ("iters", c_i32), ("calls", c_i32)]


class Operator(ctypes.Structure):
    """dgb_operator (include/dgb200.h): host struct of device pointers."""
    _fields_ = [("Ni", c_i32), ("Nj", c_i32), ("b", c_i32), ("nnzb", c_i32),
                ("stencil", c_i32), ("reserved", c_i32),
                ("data", c_vp), ("indices", c_vp), ("indptr", c_vp), ("dinv", c_vp), ("gs_data", c_vp),
                ("gs_mailbox", c_vp), ("gs_chain", c_vp),
                ("gs_rows", c_vp), ("h_gs_offsets", c_vp), ("gs_nlevels_fwd", c_i32), ("gs_nlevels_bwd", c_i32)]


class Level(ctypes.Structure):
    """dgb_level (include/dgb200.h)."""
    _fields_ = [("op", Operator),
                ("rhs", c_vp), ("u", c_vp), ("r", c_vp),
                ("transfer_kind", c_i32), ("nc", c_i32), ("nf", c_i32), ("pad0", c_i32),
                ("R", c_vp), ("P", c_vp),
                ("smoother", c_i32), ("direction", c_i32),
                ("pre_iterations", c_i32), ("post_iterations", c_i32), ("omega", c_f64),
                ("post_smoother", c_i32), ("post_direction", c_i32), ("post_omega", c_f64)]


class VcycleOpts(ctypes.Structure):
    _fields_ = [("gs_mode", c_i32), ("check_residual", c_i32), ("coarse_iterations", c_i32),
                ("coarse_solver", c_i32), ("u_final_event", c_vp), ("coarse_inverse", c_vp)]


class SlabLevel(ctypes.Structure):
    """dgb_slab_level (include/dgb200.h)."""
    _fields_ = [("lev", Level), ("u_block", c_vp), ("ghost_lo", c_i32), ("ghost_hi", c_i32),
                ("colour_shift", c_i32), ("pad", c_i32), ("n_global", c_i64)]


class SlabOpts(ctypes.Structure):
    """dgb_slab_opts (include/dgb200.h)."""
    _fields_ = [("gs_mode", c_i32), ("check_residual", c_i32), ("link_kind", c_i32), ("link_nc", c_i32),
                ("link_nf", c_i32), ("link_Ni_c", c_i32), ("link_rows_c", c_i32), ("n_coarse", c_i32),
                ("link_R", c_vp), ("link_P", c_vp), ("link_rhs_local", c_vp),
                ("coarse_levels", ctypes.POINTER(Level)), ("coarse_ctl", c_vp), ("coarse_opts", VcycleOpts)]


class TablesDesc(ctypes.Structure):
    _fields_ = [("Pg", c_i32), ("p", c_i32), ("nq1", c_i32), ("cf", c_i32)] + \
        [(n, c_vp) for n in ("h_V", "h_Vr", "h_Vs", "h_w2", "h_w1", "h_Vf", "h_Vrf", "h_Vsf",
                             "h_GX", "h_GR", "h_GS", "h_FX", "h_FR", "h_FS", "h_sub_vol", "h_sub_face")]


GS_LEXICOGRAPHIC, GS_REDBLACK, GS_SLAB_LEXICOGRAPHIC = 0, 1, 2
TRANSFER_P, TRANSFER_H = 1, 2
SMOOTHER_IDS = {"block_gauss_seidel_pyamg": 0, "block_jacobi": 1, "block_gauss_seidel": 2}
FLAG_PERIODIC_I, FLAG_PERIODIC_J, FLAG_MINV, FLAG_GHOST_LO, FLAG_GHOST_HI = 1, 2, 4, 8, 16
UNSUPPORTED = 100        # DGB_UNSUPPORTED
COARSE_SMOOTHER, COARSE_DIRECT = 0, 1
VCYCLE_ENTRY_PRIMED = 1  # DGB_VCYCLE_ENTRY_PRIMED

# name -> (restype, argtypes); every symbol include/dgb200.h declares
OP = ctypes.POINTER(Operator)
SIGNATURES = {
    "dgb_abi_version": (c_i32, []),
    "dgb_last_error": (ctypes.c_char_p, []),
    "dgb_sm_count": (c_i32, []),
    "dgb_partials_len": (c_i32, []),
    "dgb_launch_count": (ctypes.c_longlong, [c_i32]),
    "dgb_set_kernel_path": (c_i32, [c_i32]),
    "dgb_device_error": (c_i32, [c_i32]),
    "dgb_fill_sentinel": (c_i32, [c_vp, c_i64, c_vp]),
    "dgb_dense_inverse": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dgb_dense_solve": (c_i32, [c_vp, c_i32, c_vp, c_vp, c_vp]),
    "dgb_comm_create": (c_i32, [c_i32, c_i32, c_i64, ctypes.POINTER(c_vp)]),
    "dgb_comm_handle_bytes": (c_i32, []),
    "dgb_comm_export": (c_i32, [c_vp, c_vp]),
    "dgb_comm_connect": (c_i32, [c_vp, c_vp]),
    "dgb_comm_arena": (c_vp, [c_vp, ctypes.POINTER(c_i64)]),
    "dgb_comm_error": (c_i32, [c_vp, c_i32]),
    "dgb_comm_destroy": (None, [c_vp]),
    "dgb_halo_exchange": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp]),
    "dgb_allreduce_sum": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_i64, c_vp]),
    "dgb_allgather": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "dgb_vcycle_slab": (c_i32, [c_vp, ctypes.POINTER(SlabLevel), c_i32, ctypes.POINTER(SlabOpts), c_vp, c_vp, c_vp, c_vp]),
    "dgb_nodal_error": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dgb_bsr_apply": (c_i32, [OP, c_vp, c_vp, c_vp]),
    "dgb_bsr_residual": (c_i32, [OP, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dgb_bsr_residual_colour": (c_i32, [OP, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "dgb_bsr_spgemm": (c_i32, [c_i32, c_i32] + [c_vp] * 10),
    "dgb_sumsq": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "dgb_block_diag_inverse": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dgb_build_gs_stream": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "dgb_gs_chain_len": (c_i64, [c_i32, c_i32, c_i32, c_i32]),
    "dgb_build_gs_chain": (c_i32, [OP, c_vp]),
    "dgb_check_stencil": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "dgb_block_gs_pass": (c_i32, [OP, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "dgb_block_gs_pass_seq": (c_i32, [OP, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "dgb_block_gs_entry_residual": (c_i32, [OP, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "dgb_block_gs_residual_after_pass": (c_i32, [OP, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dgb_block_gs_colour": (c_i32, [OP, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "dgb_block_gs_colour_entry": (c_i32, [OP, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "dgb_block_relax_sweep": (c_i32, [OP, c_vp, c_vp, c_vp, c_f64, c_vp]),
    "dgb_smoother_begin": (c_i32, [c_vp, c_vp, c_i64, c_vp]),
    "dgb_smoother_check": (c_i32, [c_vp, c_vp, c_i64, c_vp]),
    "dgb_block_gauss_seidel_pyamg": (c_i32, [OP, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "dgb_restrict": (c_i32, [c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dgb_prolong_add": (c_i32, [c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dgb_restrict_slab": (c_i32, [c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dgb_prolong_add_slab": (c_i32, [c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dgb_vcycle": (c_i32, [ctypes.POINTER(Level), c_i32, ctypes.POINTER(VcycleOpts), c_vp, c_vp, c_vp, c_vp]),
    "dgb_vcycle_ex": (c_i32, [ctypes.POINTER(Level), c_i32, ctypes.POINTER(VcycleOpts), c_vp, c_vp, c_vp, c_vp, c_i32]),
    "dgb_tables_create": (c_i32, [ctypes.POINTER(TablesDesc), ctypes.POINTER(c_vp)]),
    "dgb_tables_destroy": (None, [c_vp]),
    "dgb_metrics": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "dgb_poisson_nnzb": (c_i64, [c_i32, c_i32, c_i32]),
    "dgb_assemble_poisson": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_f64, c_f64, c_i32,
                                     c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dgb_assemble_stokes": (c_i32, [c_vp] * 9 + [c_i32, c_i32, c_f64, c_f64, c_f64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "dgb_assemble_rhs_stokes": (c_i32, [c_vp] * 13 + [c_i32, c_i32, c_f64, c_f64, c_f64, c_i32, c_vp, c_vp]),
    "dgb_assemble_rhs": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_f64, c_f64,
                                 c_i32, c_vp, c_vp]),
}


def load(path=None):
    """dlopen libdgb200.so and bind every declared symbol.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise DgbError(f"{path} not found: build it with `python -m dg_multigrid_solver_b200.build` "
                       "(there is no CPU fallback)")
    L = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.dgb_abi_version() != 6:
        raise DgbError("libdgb200.so ABI version mismatch")
    if os.environ.get("DGB_KERNELS", "auto") == "generic":
        L.dgb_set_kernel_path(1)
    if os.environ.get("DGB_CHAIN_MASK"):        # tuning: block sizes of the chained Gauss-Seidel kernel
        L.dgb_set_kernel_path(300 + int(os.environ["DGB_CHAIN_MASK"]))
    if os.environ.get("DGB_GS_VARIANT"):        # tuning: A/B switches of the smoother kernels (dgb_stream.cu)
        L.dgb_set_kernel_path(100 + int(os.environ["DGB_GS_VARIANT"]))
    _lib = L
    return L


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    if isinstance(t, int):
        return t
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def check(rc, what=""):
    if rc != 0:
        msg = load().dgb_last_error().decode(errors="replace")
        raise DgbError(f"{what} failed (rc={rc}): {msg}")


def call(name, *args):
    L = load()
    conv = []
    for a in args:
        if isinstance(a, Operator):
            conv.append(ctypes.byref(a))
        elif hasattr(a, "data_ptr") or isinstance(a, np.ndarray):
            conv.append(ptr(a))
        else:
            conv.append(a)
    rc = getattr(L, name)(*conv)
    check(rc, name)
    if os.environ.get("DGB_SYNC_CALLS") == "1":      # diagnostics (tools/profile_setup.py): device time lands on its call
        require_cuda().cuda.synchronize()
    return rc


def check_device_error(mailboxes=()):
    """Read (and reset) the error flag of the asynchronous smoother kernels wherever the host already
    synchronises.  A non-zero flag means a bounded wait timed out and a pass returned early: the mailboxes
    are refilled with the sentinel so that a later pass starts clean, and the solve is reported as failed."""
    L = load()
    err = L.dgb_device_error(1)
    if err:
        for m in mailboxes:
            if m is not None:
                call("dgb_fill_sentinel", m, int(m.numel()), stream_ptr())
        raise DgbError(f"a lexicographic Gauss-Seidel kernel timed out on the device (dgb_device_error={err}: "
                       "1 = TMA/mbarrier wait, 2 = row hand-over wait); the iterate is partly updated -- "
                       "results of this solve are invalid (GPU time-sliced or stalled under a profiler?)")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise DgbError("no CUDA device visible: dg_multigrid_solver_b200 has no CPU path")
    load()
    return torch
