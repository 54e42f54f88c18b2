"""ctypes loader for oracle/csrc/dgoracle.c (restated pyamg / scipy native routines)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_ROOT, "_build", "libdgoracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_ROOT, "csrc", "dgoracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _ROOT, "-s", "_build/libdgoracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        L.orc_bsr_matvec.argtypes = [ctypes.c_int32, ctypes.c_int32, i32p, i32p, f64p, f64p, f64p]
        L.orc_bsr_matvec.restype = None
        L.orc_bsr_residual.argtypes = [ctypes.c_int32, ctypes.c_int32, i32p, i32p, f64p, f64p, f64p, f64p]
        L.orc_bsr_residual.restype = ctypes.c_double
        L.orc_block_gauss_seidel.argtypes = [i32p, i32p, f64p, f64p, f64p, f64p,
                                             ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
        L.orc_block_gauss_seidel.restype = None
        L.orc_block_gauss_seidel_colour.argtypes = [i32p, i32p, f64p, f64p, f64p, f64p,
                                                    ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                                    ctypes.c_int32, ctypes.c_int32]
        L.orc_block_gauss_seidel_colour.restype = None
        _lib = L
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def block_gauss_seidel(Ap, Aj, Ax, x, b, Dinv, row_start, row_stop, row_step, blocksize):
    """pyamg.amg_core.block_gauss_seidel (pyamg 5.0.1 relaxation.h), x updated in place.
    Reference call site: dgfem/pyamg_relaxation.py:252-255."""
    assert x.dtype == np.float64 and x.flags.c_contiguous
    lib().orc_block_gauss_seidel(_i32(Ap), _i32(Aj), _f64(np.ravel(Ax)), x, _f64(b), _f64(np.ravel(Dinv)),
                                 int(row_start), int(row_stop), int(row_step), int(blocksize))


def block_gauss_seidel_colour(Ap, Aj, Ax, x, b, Dinv, Ni, Nj, ncolours, colour, blocksize):
    assert x.dtype == np.float64 and x.flags.c_contiguous
    lib().orc_block_gauss_seidel_colour(_i32(Ap), _i32(Aj), _f64(np.ravel(Ax)), x, _f64(b), _f64(np.ravel(Dinv)),
                                        int(Ni), int(Nj), int(ncolours), int(colour), int(blocksize))


def bsr_matvec(Ap, Aj, Ax, x, blocksize):
    """scipy _sparsetools.bsr_matvec restated (y = A x)."""
    n_brow = len(Ap) - 1
    y = np.zeros(n_brow * blocksize)
    lib().orc_bsr_matvec(n_brow, int(blocksize), _i32(Ap), _i32(Aj), _f64(np.ravel(Ax)), _f64(x), y)
    return y


def bsr_residual(Ap, Aj, Ax, x, b, blocksize):
    n_brow = len(Ap) - 1
    r = np.zeros(n_brow * blocksize)
    s = lib().orc_bsr_residual(n_brow, int(blocksize), _i32(Ap), _i32(Aj), _f64(np.ravel(Ax)), _f64(x), _f64(b), r)
    return r, s
