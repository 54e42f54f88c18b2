"""Vectorised element geometry (TEST INFRASTRUCTURE).

Restates Element.compute_geometric_terms / metric_xy_rs (dgfem/element.py:52-130),
Element.A (element.py:30) and the coarse-element sampling
CoarseElement._init_coarse_element (element.py:242-356) + CoarseGrid.initialize
(dgfem/grid.py:288-360), for all elements of a level at once.

Array conventions: element arrays are indexed [i, j, ...]; volume point arrays are
[i, j, q] with q = i_r + N_int*i_s (r fastest, the reference's order='F' ravel);
face point arrays are [i, j, N_int]."""
import numpy as np

FACES = ("imin", "imax", "jmin", "jmax")


class LevelGeometry:
    pass


def element_nodes(x, y, Ni, Nj, Pg, stride=1):
    """grid.py:89-93 / 282-286: nodes of element (i,j), F-order flattened -> [Ni,Nj,(Pg+1)^2]."""
    ia = (np.arange(Ni)[:, None] * Pg * stride + np.arange(Pg + 1)[None, :] * stride)     # [Ni, Pg+1]
    ja = (np.arange(Nj)[:, None] * Pg * stride + np.arange(Pg + 1)[None, :] * stride)
    xe = x[ia[:, None, :, None], ja[None, :, None, :]]          # [Ni,Nj,a,c]
    ye = y[ia[:, None, :, None], ja[None, :, None, :]]
    # F-order ravel: index = a + (Pg+1)*c
    xe = xe.transpose(0, 1, 3, 2).reshape(Ni, Nj, -1)
    ye = ye.transpose(0, 1, 3, 2).reshape(Ni, Nj, -1)
    return xe, ye


def _metrics(xe, ye, L_gg, L, Dr, Ds, scale=1.0):
    """element.py:54,76-95 for a set of points given operator rows [npts, ng]."""
    x_rs = xe @ L_gg.T
    y_rs = ye @ L_gg.T
    xr, xs = x_rs @ Dr.T, x_rs @ Ds.T
    yr, ys = y_rs @ Dr.T, y_rs @ Ds.T
    if scale != 1.0:
        xr, xs, yr, ys = xr * scale, xs * scale, yr * scale, ys * scale
    J = xr * ys - yr * xs
    out = dict(J=J, rx=ys / J, sx=-yr / J, ry=-xs / J, sy=xr / J, xr=xr, xs=xs, yr=yr, ys=ys)
    if L is not None:
        out["x"] = xe @ L.T
        out["y"] = ye @ L.T
    return out


def _face_finish(m, face):
    """element.py:96-102."""
    if face in ("imin", "imax"):
        Jf = np.sqrt(m["xs"] ** 2 + m["ys"] ** 2)
        nrm = np.sqrt(m["rx"] ** 2 + m["ry"] ** 2)
        n = np.stack([m["rx"] / nrm, m["ry"] / nrm], axis=-1)
    else:
        Jf = np.sqrt(m["xr"] ** 2 + m["yr"] ** 2)
        nrm = np.sqrt(m["sx"] ** 2 + m["sy"] ** 2)
        n = np.stack([m["sx"] / nrm, m["sy"] / nrm], axis=-1)
    return Jf, n


def fine_geometry(x, y, Ni, Nj, T):
    """All-element version of Element.__init__ (element.py:16-30) for tables T (LevelTables)."""
    g = LevelGeometry()
    g.Ni, g.Nj = Ni, Nj
    xe, ye = element_nodes(x, y, Ni, Nj, T.Pg)
    vol = _metrics(xe, ye, T.L_gg, T.L_int, T.Dr_int, T.Ds_int)
    g.vol = {k: vol[k] for k in ("J", "rx", "sx", "ry", "sy", "x", "y")}
    g.face = {}
    for f in FACES:
        m = _metrics(xe, ye, T.L_gg, T.L_face[f], T.Dr_face[f], T.Ds_face[f])
        Jf, n = _face_finish(m, f)
        g.face[f] = dict(J=Jf, rx=m["rx"], sx=m["sx"], ry=m["ry"], sy=m["sy"], n=n, x=m["x"], y=m["y"])
    g.A = np.einsum("ijq,q->ij", g.vol["J"], np.ravel(T.w_int_2D, order="F"))
    return g


def coarse_point_map(T, cf):
    """element.py:273-287: for every coarse quadrature point (iR, iS) the fine sub-element (m, n)
    that contains it (first match, n outer / m inner) and the local coordinates (r, s)."""
    delta = 2.0 / cf
    R_int = T.r_int
    pts = []
    for iS, S in enumerate(R_int):
        for iR, R in enumerate(R_int):
            found = None
            for n in range(cf):
                if found:
                    break
                for m in range(cf):
                    r = (2 * R + 2 - delta * (1 + m * 2)) / delta
                    s = (2 * S + 2 - delta * (1 + n * 2)) / delta
                    if -1 <= r <= 1 and -1 <= s <= 1:
                        found = (m, n, r, s)
                        break
            if found is None:
                raise RuntimeError("coarse quadrature point not located (element.py:284)")
            pts.append((iR, iS) + found)
    return pts


def coarse_geometry(x, y, Ni_f, Nj_f, T, cf):
    """Coarse level by a factor cf: metrics sampled from the containing fine element,
    derivatives scaled by cf (element.py:81-85,292-310).  Includes the reference's choice of
    source element for face terms."""
    Ni, Nj = Ni_f // cf, Nj_f // cf
    if Ni == 0 or Nj == 0:
        raise ValueError("grid cannot be divided by the coarsening factor (grid.py:325)")
    g = LevelGeometry()
    g.Ni, g.Nj = Ni, Nj
    xe_f, ye_f = element_nodes(x, y, Ni_f, Nj_f, T.Pg)           # fine elements
    nq = T.N_int
    vol = {k: np.zeros((Ni, Nj, nq * nq)) for k in ("J", "rx", "sx", "ry", "sy", "x", "y")}
    face = {f: {k: np.zeros((Ni, Nj, nq)) for k in ("J", "rx", "sx", "ry", "sy", "x", "y")} for f in FACES}
    for f in FACES:
        face[f]["n"] = np.zeros((Ni, Nj, nq, 2))
    I = np.arange(Ni)[:, None]
    Jc = np.arange(Nj)[None, :]
    for (iR, iS, m, n, r, s) in coarse_point_map(T, cf):
        q = iR + nq * iS
        xe = xe_f[I * cf + m, Jc * cf + n]                     # [Ni,Nj,ng]
        ye = ye_f[I * cf + m, Jc * cf + n]
        L, Dr, Ds = T.point_ops(r, s)
        mm = _metrics(xe, ye, T.L_gg, L, Dr, Ds, scale=float(cf))
        for k in vol:
            vol[k][:, :, q] = mm[k][..., 0]
        def put(fname, rr, ss, idx):
            Lf, Drf, Dsf = T.point_ops(rr, ss)
            mf = _metrics(xe, ye, T.L_gg, Lf, Drf, Dsf, scale=float(cf))
            Jf, nn = _face_finish(mf, fname)
            for k in ("rx", "sx", "ry", "sy", "x", "y"):
                face[fname][k][:, :, idx] = mf[k][..., 0]
            face[fname]["J"][:, :, idx] = Jf[..., 0]
            face[fname]["n"][:, :, idx, :] = nn[..., 0, :]
        # element.py:295-310 (note the if/elif pairs)
        if iR == 0:
            put("imin", -1.0, s, iS)
        elif iR == nq - 1:
            put("imax", 1.0, s, iS)
        if iS == 0:
            put("jmin", r, -1.0, iR)
        elif iS == nq - 1:
            put("jmax", r, 1.0, iR)
    g.vol, g.face = vol, face
    g.A = np.einsum("ijq,q->ij", vol["J"], np.ravel(T.w_int_2D, order="F"))
    return g
