"""Host-side generation of the (tiny) per-level tables the CUDA kernels consume.

Mirrors the reference's Interpolation class and Grid.initialize_interpolation
(dgfem/interpolation.py:29-170, dgfem/grid.py:178-213): orthonormal Legendre modal basis,
mode n = j_s*(p+1)+i_r, point index i_r + N_int*i_s (r fastest), Gauss-Legendre volume/face
quadrature with N_int = factor*p//2 + 1 points, LGL geometry nodes.

On top of the reference's tables this module folds the geometry mapping into three operator
tables per point set (node values -> x, x_r, x_s at the points), which is what the metrics
kernel (dgb_metrics) multiplies with the raw Plot3D nodes:
    x   = L  @ nodes            (Element.metric_xy_rs,          dgfem/element.py:115-130)
    x_r = Dr @ (L_gg @ nodes)   (Element.compute_geometric_terms, dgfem/element.py:54,76-80)
For an h-coarsened level every coarse quadrature point carries the offset (m, n) of the fine
sub-element that contains it and operator rows evaluated at its local coordinates, scaled by
the coarsening factor (CoarseElement._init_coarse_element, dgfem/element.py:273-310).
"""
from math import factorial

import numpy as np
from scipy.special import eval_jacobi, gamma, roots_jacobi

FACES = ("imin", "imax", "jmin", "jmax")
TRACES = ("iL", "iR", "jL", "jR")


def _jacobi_on(x, a, b, n):
    """Orthonormal Jacobi polynomial P_n^(a,b) (dgfem/interpolation.py:42-44)."""
    x = np.asarray(x, dtype=np.float64)
    h = 2.0 ** (a + b + 1) * gamma(n + a + 1) * gamma(n + b + 1) / ((2 * n + a + b + 1) * gamma(n + a + b + 1) * factorial(n))
    return eval_jacobi(n, a, b, x) / np.sqrt(h)


def _phi(x, n):
    return _jacobi_on(x, 0, 0, n)


def _dphi(x, n):
    """d/dx of the orthonormal Legendre polynomial (dgfem/interpolation.py:52-57)."""
    x = np.asarray(x, dtype=np.float64)
    if n == 0:
        return np.zeros_like(x)
    return np.sqrt(n * (n + 1.0)) * _jacobi_on(x, 1, 1, n - 1)


def gauss_legendre(n):
    return roots_jacobi(n, 0, 0)


def gauss_lobatto_nodes(n):
    if n < 2:
        raise ValueError("The polynomial order P must be a positive integer")
    xi = np.empty(n)
    xi[0], xi[-1] = -1.0, 1.0
    if n > 2:
        xi[1:-1] = roots_jacobi(n - 2, 1, 1)[0]
    return xi


def tensor_basis(n1, r, s, dr=False, ds=False):
    """Rows: points (r fastest); columns: modes (i_r fastest).  dr/ds select the derivative."""
    r = np.atleast_1d(np.asarray(r, dtype=np.float64))
    s = np.atleast_1d(np.asarray(s, dtype=np.float64))
    out = np.empty((r.size * s.size, n1 * n1))
    for js in range(n1):
        fs = _dphi(s, js) if ds else _phi(s, js)
        for ir in range(n1):
            fr = _dphi(r, ir) if dr else _phi(r, ir)
            out[:, js * n1 + ir] = np.outer(fr, fs).ravel(order="F")
    return out


class Tables:
    """All host tables of one level: geometry degree Pg, solution degree p, coarsening factor cf."""

    def __init__(self, Pg, p, factor=3, cf=1, nq1=None):
        """nq1 overrides the number of quadrature points per direction: the Stokes system evaluates the
        velocity basis at the pressure's points and vice versa (grid.py:185-210 builds every combination)."""
        self.Pg, self.p, self.cf = int(Pg), int(p), int(cf)
        self.ng1, self.n1 = self.Pg + 1, self.p + 1
        self.ng, self.b = self.ng1 ** 2, self.n1 ** 2
        self.nq1 = factor * self.p // 2 + 1 if nq1 is None else int(nq1)   # dgfem/grid.py:107
        self.nq = self.nq1 ** 2
        self.r_grid = gauss_lobatto_nodes(self.ng1)
        self.r_int, self.w1 = gauss_legendre(self.nq1)
        self.w2 = np.outer(self.w1, self.w1).ravel(order="F")
        ri, n1 = self.r_int, self.n1
        self.V = tensor_basis(n1, ri, ri)
        self.Vr = tensor_basis(n1, ri, ri, dr=True)
        self.Vs = tensor_basis(n1, ri, ri, ds=True)
        tr = {"iL": ([1.0], ri), "iR": ([-1.0], ri), "jL": (ri, [1.0]), "jR": (ri, [-1.0])}   # grid.py:203-210
        self.Vf = np.stack([tensor_basis(n1, *tr[k]) for k in TRACES])
        self.Vrf = np.stack([tensor_basis(n1, *tr[k], dr=True) for k in TRACES])
        self.Vsf = np.stack([tensor_basis(n1, *tr[k], ds=True) for k in TRACES])
        self.V_DOF_grid = tensor_basis(n1, self.r_grid, self.r_grid)                          # grid.py:213
        self._geometry_ops()

    # -- geometry operators ------------------------------------------------------------------
    def _ops_at(self, r, s):
        """(L, Dr.L_gg, Ds.L_gg) rows at the tensor points r x s of the geometry basis."""
        Vgg = tensor_basis(self.ng1, self.r_grid, self.r_grid)
        Vinv = np.linalg.inv(Vgg)
        L_gg = Vgg @ np.linalg.inv(Vgg.T).T            # (inv(V^T) V^T)^T, element.py:122 at the nodes
        L = tensor_basis(self.ng1, r, s) @ np.linalg.inv(Vgg.T).T
        Dr = tensor_basis(self.ng1, r, s, dr=True) @ Vinv   # (inv(V)^T Vr^T)^T, element.py:76
        Ds = tensor_basis(self.ng1, r, s, ds=True) @ Vinv
        return L, Dr @ L_gg, Ds @ L_gg

    def _geometry_ops(self):
        ri, nq1, cf = self.r_int, self.nq1, self.cf
        if cf == 1:
            self.GX, self.GR, self.GS = self._ops_at(ri, ri)
            fp = {"imin": ([-1.0], ri), "imax": ([1.0], ri), "jmin": (ri, [-1.0]), "jmax": (ri, [1.0])}
            ops = [self._ops_at(*fp[f]) for f in FACES]
            self.FX = np.stack([o[0] for o in ops])
            self.FR = np.stack([o[1] for o in ops])
            self.FS = np.stack([o[2] for o in ops])
            self.sub_vol = np.zeros((self.nq, 2), dtype=np.int32)
            self.sub_face = np.zeros((4, nq1, 2), dtype=np.int32)
            return
        if nq1 < 2:
            raise ValueError("h-coarsening needs at least 2 quadrature points per direction")
        delta = 2.0 / cf
        GX = np.zeros((self.nq, self.ng)); GR = np.zeros_like(GX); GS = np.zeros_like(GX)
        FX = np.zeros((4, nq1, self.ng)); FR = np.zeros_like(FX); FS = np.zeros_like(FX)
        sub_vol = np.zeros((self.nq, 2), dtype=np.int32)
        sub_face = np.zeros((4, nq1, 2), dtype=np.int32)

        k2 = 1 + 2 * np.arange(cf)

        def locate(R, S):
            # first containing fine sub-element, n outer / m inner (element.py:278-287).  r depends on m only and s on
            # n only, so the first match of the double loop is (first n whose s fits, first m whose r fits); the two
            # expressions are the reference's, evaluated for all m (n) at once with the same operations
            r_all = (2 * R + 2 - delta * k2) / delta
            s_all = (2 * S + 2 - delta * k2) / delta
            ok_r = (-1 <= r_all) & (r_all <= 1)
            ok_s = (-1 <= s_all) & (s_all <= 1)
            if not (ok_r.any() and ok_s.any()):
                raise RuntimeError("coarse quadrature point not located in any fine element")
            m, n = int(np.argmax(ok_r)), int(np.argmax(ok_s))
            return m, n, float(r_all[m]), float(s_all[n])

        def put_face(f, idx, m, n, r, s):
            L, Dr, Ds = self._ops_at([r], [s])
            FX[f, idx], FR[f, idx], FS[f, idx] = L[0], cf * Dr[0], cf * Ds[0]
            sub_face[f, idx] = (m, n)

        for iS, S in enumerate(ri):
            for iR, R in enumerate(ri):
                m, n, r, s = locate(R, S)
                q = iR + nq1 * iS
                L, Dr, Ds = self._ops_at([r], [s])
                GX[q], GR[q], GS[q] = L[0], cf * Dr[0], cf * Ds[0]      # element.py:81-85
                sub_vol[q] = (m, n)
                # face samples come from the element that holds the first/last VOLUME point
                # (element.py:295-310, including its if/elif pairing)
                if iR == 0:
                    put_face(0, iS, m, n, -1.0, s)
                elif iR == nq1 - 1:
                    put_face(1, iS, m, n, 1.0, s)
                if iS == 0:
                    put_face(2, iR, m, n, r, -1.0)
                elif iS == nq1 - 1:
                    put_face(3, iR, m, n, r, 1.0)
        self.GX, self.GR, self.GS, self.FX, self.FR, self.FS = GX, GR, GS, FX, FR, FS
        self.sub_vol, self.sub_face = sub_vol, sub_face


def p_restriction(p_coarse, p_fine):
    """Zero-padded identity selecting the modes (i <= p_c, j <= p_c) of the hierarchical basis;
    same matrix as dgfem/dgfem.py:306-317 builds with np.insert/np.append."""
    nc1, nf1 = p_coarse + 1, p_fine + 1
    R = np.zeros((nc1 * nc1, nf1 * nf1))
    for j in range(nc1):
        for i in range(nc1):
            R[j * nc1 + i, j * nf1 + i] = 1.0
    return R


def h_restriction():
    """L2 projection of the four p=1 children onto their parent (dgfem/dgfem.py:362-367),
    children ordered (a_j, a_i), 4 modes each; prolongation = 4 R^T."""
    s3 = np.sqrt(3)
    R = np.array([
        np.array([1., 0., 0., 0., 1., 0., 0., 0., 1., 0., 0., 0., 1., 0., 0., 0.]) / 4.,
        np.array([-s3, 1., 0., 0., s3, 1., 0., 0., -s3, 1., 0., 0., s3, 1., 0., 0.]) / 8.,
        np.array([-s3, 0., 1., 0., -s3, 0., 1., 0., s3, 0., 1., 0., s3, 0., 1., 0.]) / 8.,
        np.array([3., -s3, -s3, 1., -3., -s3, s3, 1., -3., s3, -s3, 1., 3., s3, s3, 1.]) / 16.])
    return R, R.T * 4.
