import numpy as np


def norm(x, pnorm='2'):
    return float(np.sqrt(np.inner(np.ravel(x).conj(), np.ravel(x)).real))
