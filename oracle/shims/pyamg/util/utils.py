"""Restated pyamg.util.utils.get_block_diag (pyamg 5.0.1): extract the diagonal
bs x bs blocks of a BSR matrix and (inv_flag) pseudo-invert each.  pyamg uses
amg_core.pinv_array (bs<7) or LAPACK gelss (bs>=7) with a relative singular-value
cutoff; numpy.linalg.pinv is the same SVD pseudo-inverse and equals the true
inverse for the nonsingular blocks met here."""
import numpy as np
from scipy import sparse


def get_block_diag(A, blocksize, inv_flag=True):
    A = sparse.bsr_matrix(A, blocksize=(blocksize, blocksize)) if not sparse.isspmatrix_bsr(A) else A
    if A.blocksize != (blocksize, blocksize):
        A = A.tobsr(blocksize=(blocksize, blocksize))
    N = A.shape[0] // blocksize
    block_diag = np.zeros((N, blocksize, blocksize), dtype=A.dtype)
    indptr, indices, data = A.indptr, A.indices, A.data
    rows = np.repeat(np.arange(N), np.diff(indptr))
    sel = np.nonzero(indices == rows)[0]
    # duplicates (if any) are summed, as a BSR->dense diagonal extraction would
    np.add.at(block_diag, rows[sel], data[sel])
    if inv_flag:
        block_diag = np.linalg.pinv(block_diag)
    return block_diag


def type_prep(*a, **k):
    raise NotImplementedError


def get_diagonal(*a, **k):
    raise NotImplementedError
