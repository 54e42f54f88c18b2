"""Plot3D reader + synthetic grid rules (TEST INFRASTRUCTURE).

Restates Geometry.read (dgfem/grid.py:26-63) and the rules the shipped input/*.xyz
files follow."""
import numpy as np

from .tables import lgl


def read_plot3d(path, P_grid):
    """grid.py:29-62.  Returns x[il, jl], y[il, jl] (first axis = i), Ni, Nj."""
    raw = np.fromfile(path, dtype=np.uint8)
    off = 0
    recs = []
    while off < raw.size:
        n = int(raw[off:off + 4].view("<u4")[0])
        recs.append(raw[off + 4:off + 4 + n])
        tail = int(raw[off + 4 + n:off + 8 + n].view("<u4")[0])
        if tail != n:
            raise ValueError("corrupt Fortran record")
        off += 8 + n
    nblocks = recs[0].view("<i4")[0]
    if nblocks != 1:
        raise ValueError(f"Number of blocks is {nblocks} instead of 1")
    il, jl, kl = (int(v) for v in recs[1].view("<i4"))
    if kl != 1:
        raise ValueError("More than one point in third dimension")
    coords = recs[2].view("<f8")
    x = coords[:il * jl].reshape((jl, il)).T
    y = coords[il * jl:2 * il * jl].reshape((jl, il)).T
    return np.ascontiguousarray(x), np.ascontiguousarray(y), (il - 1) // P_grid, (jl - 1) // P_grid


def write_plot3d(path, x, y):
    il, jl = x.shape
    def rec(b):
        n = np.array([len(b)], dtype="<u4").tobytes()
        return n + b + n
    coords = np.concatenate([x.T.ravel(), y.T.ravel(), np.zeros(il * jl)]).astype("<f8")
    with open(path, "wb") as f:
        f.write(rec(np.array([1], dtype="<i4").tobytes()))
        f.write(rec(np.array([il, jl, 1], dtype="<i4").tobytes()))
        f.write(rec(coords.tobytes()))


def _lgl_line(edges, P):
    """Nodes of a 1-D mesh with element edges `edges`, LGL interior points per element."""
    xi = lgl(P + 1)
    N = len(edges) - 1
    out = np.empty(N * P + 1)
    for e in range(N):
        a, b = edges[e], edges[e + 1]
        out[e * P:(e + 1) * P + 1] = a + (b - a) * (xi + 1.0) / 2.0
    return out


def rectangle_nodes(Ni, Nj, P, lo=-1.0, hi=1.0):
    """Rectangle_{N}X{N}_nPoly{P}: uniform elements on [-1,1]^2, LGL interior nodes."""
    xe = np.linspace(lo, hi, Ni + 1)
    ye = np.linspace(lo, hi, Nj + 1)
    xl = _lgl_line(xe, P)
    yl = _lgl_line(ye, P)
    x = np.repeat(xl[:, None], len(yl), axis=1)
    y = np.repeat(yl[None, :], len(xl), axis=0)
    return x, y


def circle_in_circle_nodes(Ni, Nj, P, r_in=0.1, r_out=1.0):
    """CircleInCircle_{N}X{N}_nPoly{P} (App. A.9): i = angle, clockwise from 0; j = radius with
    element widths in geometric progression of ratio 10^(1/(Nj-1)); LGL interior nodes in both."""
    q = 10.0 ** (1.0 / (Nj - 1)) if Nj > 1 else 1.0
    if Nj > 1:
        w0 = (r_out - r_in) * (q - 1.0) / (q ** Nj - 1.0)
        widths = w0 * q ** np.arange(Nj)
    else:
        widths = np.array([r_out - r_in])
    redges = r_in + np.concatenate([[0.0], np.cumsum(widths)])
    redges[-1] = r_out
    tedges = -2.0 * np.pi * np.arange(Ni + 1) / Ni
    th = _lgl_line(tedges, P)
    rr = _lgl_line(redges, P)
    x = np.cos(th)[:, None] * rr[None, :]
    y = np.sin(th)[:, None] * rr[None, :]
    # close the O-grid exactly (grid.py:56-57 requires |x[0]-x[-1]| < 1e-15)
    x[-1, :] = x[0, :]
    y[-1, :] = y[0, :]
    return x, y
