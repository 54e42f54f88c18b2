"""Restated pyamg.amg_core entry points (the C lives in oracle/csrc/dgoracle.c)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dgoracle import native as _native  # noqa: E402


def block_gauss_seidel(Ap, Aj, Ax, x, b, Tx, row_start, row_stop, row_step, blocksize):
    _native.block_gauss_seidel(Ap, Aj, Ax, x, b, Tx, row_start, row_stop, row_step, blocksize)


def _out_of_scope(*a, **k):
    raise NotImplementedError("point smoothers are out of scope (SURVEY.md section 2.2)")


bsr_jacobi = bsr_gauss_seidel = gauss_seidel = _out_of_scope
