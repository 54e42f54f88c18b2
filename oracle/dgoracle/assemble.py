"""Vectorised Poisson DG assembly (TEST INFRASTRUCTURE).

Restates, for all elements/faces of a level at once:
  Element.compute_momentum_laplace_volume_integral   dgfem/element.py:181-199
  Element.compute_mass_matrix                        dgfem/element.py:132-133
  Element.compute_source_momentum_volume_integral    dgfem/element.py:161-167
  Face.__init__ (h_F, face Jacobian choice)          dgfem/face.py:13-35
  Face.compute_momentum_laplace_SIP_{flux,penalty,symmetrizing}_term  dgfem/face.py:129-280
  Poisson.assemble_BSR_Poisson                       dgfem/discrete_system.py:54-145
  Poisson.assemble_RHS_Poisson                       dgfem/discrete_system.py:355-403
"""
import numpy as np


def volume_blocks(G, T, nu):
    """K[i,j,k,l] (row k = test, col l = trial) and M (element.py:181-199,132-133)."""
    w = np.ravel(T.w_int_2D, order="F")
    v = G.vol
    Vr, Vs, V = T.Vr_DOF_int, T.Vs_DOF_int, T.V_DOF_int
    dx = Vr[None, None] * v["rx"][..., None] + Vs[None, None] * v["sx"][..., None]     # [Ni,Nj,q,b]
    dy = Vr[None, None] * v["ry"][..., None] + Vs[None, None] * v["sy"][..., None]
    wJ = v["J"] * w[None, None, :]
    K = nu * (np.einsum("ijqk,ijq,ijql->ijkl", dx, wJ, dx) + np.einsum("ijqk,ijq,ijql->ijkl", dy, wJ, dy))
    M = np.einsum("qk,ijq,ql->ijkl", V, wJ, V)
    return K, M


def _dn(T, side, fgeo):
    """Normal derivative of the trace basis using that side's own face metrics (face.py:131-135)."""
    Vr, Vs = T.Vr_face[side], T.Vs_face[side]
    ux = Vr[None, None] * fgeo["rx"][..., None] + Vs[None, None] * fgeo["sx"][..., None]
    uy = Vr[None, None] * fgeo["ry"][..., None] + Vs[None, None] * fgeo["sy"][..., None]
    return fgeo["n"][..., 0, None] * ux + fgeo["n"][..., 1, None] * uy                    # [Ni,Nj,pts,b]


def face_blocks(G, T, nu, sigma, direction, periodic):
    """All faces of one direction.  Returns dict with LL, LR, RL, RR of shape [NiF, NjF, b, b]
    where face (i,j) of direction 'i' sits between elements (i-1,j) [L] and (i,j) [R]
    (grid.py:151-176); boundary faces have the missing side's blocks zero.
    Closed form, term by term as face.py:129-280."""
    Ni, Nj = G.Ni, G.Nj
    d = direction
    fmax, fmin = f"{d}max", f"{d}min"
    VL, VR = T.V_face[f"{d}L"], T.V_face[f"{d}R"]
    dnL_all = _dn(T, f"{d}L", G.face[fmax])      # element as L side: its max face
    dnR_all = _dn(T, f"{d}R", G.face[fmin])      # element as R side: its min face
    w = T.w_int
    if d == "i":
        nF = Ni + 1
        Lidx = np.arange(-1, Ni)                  # L element index along i for faces 0..Ni
        Ridx = np.arange(0, Ni + 1)
        if periodic:
            Lidx[0] = Ni - 1
            Ridx[Ni] = 0
            hasL = np.ones(nF, bool)
            hasR = np.ones(nF, bool)
        else:
            hasL = Lidx >= 0
            hasR = Ridx < Ni
        Lc, Rc = np.clip(Lidx, 0, Ni - 1), np.clip(Ridx, 0, Ni - 1)
        take = lambda a, idx: a[idx]              # noqa: E731
        bshape = (nF, 1)
    else:
        nF = Nj + 1
        Lidx = np.arange(-1, Nj)
        Ridx = np.arange(0, Nj + 1)
        if periodic:
            Lidx[0] = Nj - 1
            Ridx[Nj] = 0
            hasL = np.ones(nF, bool)
            hasR = np.ones(nF, bool)
        else:
            hasL = Lidx >= 0
            hasR = Ridx < Nj
        Lc, Rc = np.clip(Lidx, 0, Nj - 1), np.clip(Ridx, 0, Nj - 1)
        take = lambda a, idx: a[:, idx]           # noqa: E731
        bshape = (1, nF)
    hasL_b = hasL.reshape(bshape)
    hasR_b = hasR.reshape(bshape)
    A_L, A_R = take(G.A, Lc), take(G.A, Rc)
    # face.py:13-35: h_F and which side's face Jacobian is used
    hF = np.where(hasL_b & hasR_b, (np.sqrt(A_L) + np.sqrt(A_R)) / 2,
                  np.where(hasL_b, np.sqrt(A_L), np.sqrt(A_R)))
    JL = take(G.face[fmax]["J"], Lc)
    JR = take(G.face[fmin]["J"], Rc)
    Jf = np.where(hasL_b[..., None], JL, JR)
    W = Jf * w[None, None, :]                                     # [.,.,pts]
    c = np.where(hasL_b & hasR_b, 0.5, 1.0)[..., None, None]
    mL = hasL_b[..., None, None].astype(float)
    mR = hasR_b[..., None, None].astype(float)
    dnL = take(dnL_all, Lc)
    dnR = take(dnR_all, Rc)
    pen = (sigma * nu / hF)[..., None, None]
    e = np.einsum
    # rows = test side, cols = trial side
    VWV_LL = e("pk,ijp,pl->ijkl", VL, W, VL)
    VWV_LR = e("pk,ijp,pl->ijkl", VL, W, VR)
    VWV_RL = e("pk,ijp,pl->ijkl", VR, W, VL)
    VWV_RR = e("pk,ijp,pl->ijkl", VR, W, VR)
    flux_LL = -c * nu * e("pk,ijp,ijpl->ijkl", VL, W, dnL)
    flux_LR = -c * nu * e("pk,ijp,ijpl->ijkl", VL, W, dnR)
    flux_RL = +c * nu * e("pk,ijp,ijpl->ijkl", VR, W, dnL)
    flux_RR = +c * nu * e("pk,ijp,ijpl->ijkl", VR, W, dnR)
    sym_LL = -c * nu * e("ijpk,ijp,pl->ijkl", dnL, W, VL)
    sym_LR = +c * nu * e("ijpk,ijp,pl->ijkl", dnL, W, VR)
    sym_RL = -c * nu * e("ijpk,ijp,pl->ijkl", dnR, W, VL)
    sym_RR = +c * nu * e("ijpk,ijp,pl->ijkl", dnR, W, VR)
    out = dict(
        LL=(flux_LL + pen * VWV_LL + sym_LL) * mL,
        LR=(flux_LR - pen * VWV_LR + sym_LR) * mL * mR,
        RL=(flux_RL - pen * VWV_RL + sym_RL) * mL * mR,
        RR=(flux_RR + pen * VWV_RR + sym_RR) * mR,
        hF=hF, W=W, hasL=hasL, hasR=hasR, dnL=dnL, dnR=dnR)
    return out


def assemble_bsr(G, T, nu, sigma, O_grid=False, fully_periodic=False, multiply_inverse_mass=True):
    """discrete_system.py:54-145.  Returns data[nnzb,b,b], indices, indptr, Minv[Ni,Nj,b,b]."""
    Ni, Nj, b = G.Ni, G.Nj, T.b
    K, M = volume_blocks(G, T, nu)
    Minv = np.linalg.inv(M)
    per_i = O_grid or fully_periodic
    per_j = fully_periodic
    Fi = face_blocks(G, T, nu, sigma, "i", per_i)
    Fj = face_blocks(G, T, nu, sigma, "j", per_j)
    diag = K + Fi["RR"][:Ni] + Fi["LL"][1:] + Fj["RR"][:, :Nj] + Fj["LL"][:, 1:]
    off = {"iL": Fi["RL"][:Ni], "iR": Fi["LR"][1:], "jL": Fj["RL"][:, :Nj], "jR": Fj["LR"][:, 1:]}
    if multiply_inverse_mass:
        diag = Minv @ diag
        off = {k: Minv @ v for k, v in off.items()}
    I, J = np.meshgrid(np.arange(Ni), np.arange(Nj), indexing="ij")
    m = J * Ni + I
    NONE = -1
    cols = {
        "m": m,
        "iL": np.where(I > 0, m - 1, (J * Ni + Ni - 1) if per_i else NONE),
        "iR": np.where(I < Ni - 1, m + 1, (J * Ni) if per_i else NONE),
        "jL": np.where(J > 0, m - Ni, ((Nj - 1) * Ni + I) if per_j else NONE),
        "jR": np.where(J < Nj - 1, m + Ni, I if per_j else NONE),
    }
    order = ["m", "iL", "iR", "jL", "jR"]            # discrete_system.py:136
    blocks = {"m": diag, **off}
    data, indices, indptr = [], [], [0]
    colarr = np.stack([cols[k] for k in order], axis=-1)            # [Ni,Nj,5]
    for j in range(Nj):
        for i in range(Ni):
            cr = colarr[i, j]
            present = [k for k in range(5) if cr[k] != NONE]
            srt = sorted(present, key=lambda k: cr[k])                # stable, as sorted() in :138
            for k in srt:
                indices.append(int(cr[k]))
                data.append(blocks[order[k]][i, j])
            indptr.append(indptr[-1] + len(present))
    return (np.array(data).reshape(-1, b, b), np.array(indices, dtype=np.int32),
            np.array(indptr, dtype=np.int32), Minv)


def assemble_bsr_fast(G, T, nu, sigma, O_grid=False, fully_periodic=False, multiply_inverse_mass=True):
    """Same result as assemble_bsr, without the Python double loop (for the CPU-baseline sizes)."""
    Ni, Nj, b = G.Ni, G.Nj, T.b
    K, M = volume_blocks(G, T, nu)
    Minv = np.linalg.inv(M)
    per_i = O_grid or fully_periodic
    per_j = fully_periodic
    Fi = face_blocks(G, T, nu, sigma, "i", per_i)
    Fj = face_blocks(G, T, nu, sigma, "j", per_j)
    diag = K + Fi["RR"][:Ni] + Fi["LL"][1:] + Fj["RR"][:, :Nj] + Fj["LL"][:, 1:]
    blk = [diag, Fi["RL"][:Ni], Fi["LR"][1:], Fj["RL"][:, :Nj], Fj["LR"][:, 1:]]
    if multiply_inverse_mass:
        blk = [Minv @ v for v in blk]
    I, J = np.meshgrid(np.arange(Ni), np.arange(Nj), indexing="ij")
    m = J * Ni + I
    BIG = np.iinfo(np.int64).max
    cols = np.stack([
        m,
        np.where(I > 0, m - 1, (J * Ni + Ni - 1) if per_i else BIG),
        np.where(I < Ni - 1, m + 1, (J * Ni) if per_i else BIG),
        np.where(J > 0, m - Ni, ((Nj - 1) * Ni + I) if per_j else BIG),
        np.where(J < Nj - 1, m + Ni, I if per_j else BIG)], axis=-1).astype(np.int64)   # [Ni,Nj,5]
    blocks = np.stack(blk, axis=2)                                    # [Ni,Nj,5,b,b]
    # element-major order m = j*Ni + i
    cols = cols.transpose(1, 0, 2).reshape(Ni * Nj, 5)
    blocks = blocks.transpose(1, 0, 2, 3, 4).reshape(Ni * Nj, 5, b, b)
    srt = np.argsort(cols, axis=1, kind="stable")
    cols_s = np.take_along_axis(cols, srt, axis=1)
    blocks_s = np.take_along_axis(blocks, srt[:, :, None, None], axis=1)
    present = cols_s != BIG
    counts = present.sum(axis=1)
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    indices = cols_s[present].astype(np.int32)
    data = blocks_s[present]
    return data, indices, indptr, Minv.transpose(1, 0, 2, 3).reshape(Ni * Nj, b, b)


def assemble_rhs(G, T, nu, sigma, f_vol, g_face, Minv, O_grid=False, fully_periodic=False,
                 multiply_inverse_mass=True):
    """discrete_system.py:355-403.  f_vol[i,j,q]: source at volume points; g_face[f][i,j,pts]:
    Dirichlet data at the four face point sets of every element (only boundary ones are used).
    Minv[Ni,Nj,b,b]."""
    Ni, Nj, b = G.Ni, G.Nj, T.b
    w2 = np.ravel(T.w_int_2D, order="F")
    F = np.einsum("qk,ijq,ijq->ijk", T.V_DOF_int, G.vol["J"] * w2[None, None], f_vol)     # element.py:161-164
    if not fully_periodic:
        def bterm(face, side, elem_sel, sign):
            # face.py:183 / 194 (penalty) and :230 / :246 (symmetrising), boundary face of element
            fg = {k: v[elem_sel] for k, v in G.face[face].items()}
            V = T.V_face[side]
            Vr, Vs = T.Vr_face[side], T.Vs_face[side]
            psix = Vr[None] * fg["rx"][..., None] + Vs[None] * fg["sx"][..., None]
            psiy = Vr[None] * fg["ry"][..., None] + Vs[None] * fg["sy"][..., None]
            psin = fg["n"][..., 0, None] * psix + fg["n"][..., 1, None] * psiy            # [n,pts,b]
            hF = np.sqrt(G.A[elem_sel])
            g = g_face[face][elem_sel]
            Wg = g * T.w_int[None] * fg["J"]
            pen = (sigma * nu / hF)[:, None] * np.einsum("pk,np->nk", V, Wg)
            sym = sign * nu * np.einsum("npk,np->nk", psin, Wg)
            return pen + sym
        if not O_grid:
            F[0, :] += bterm("imin", "iR", (0, slice(None)), +1.0)
            F[Ni - 1, :] += bterm("imax", "iL", (Ni - 1, slice(None)), -1.0)
        F[:, 0] += bterm("jmin", "jR", (slice(None), 0), +1.0)
        F[:, Nj - 1] += bterm("jmax", "jL", (slice(None), Nj - 1), -1.0)
    if multiply_inverse_mass:
        F = np.einsum("ijkl,ijl->ijk", Minv, F)
    return np.ascontiguousarray(F.transpose(1, 0, 2)).reshape(-1)          # m = j*Ni + i
