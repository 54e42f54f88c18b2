import numpy as np


def set_tol(dtype):
    return {np.dtype('float32'): 1e-7}.get(np.dtype(dtype), 1e-15)
