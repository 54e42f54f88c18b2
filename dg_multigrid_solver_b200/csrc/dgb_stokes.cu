// dgb_stokes.cu -- pressure-robust DG Stokes system, local ordering: one (2 b_u + b_p)^2 block per
// element pair, rows (x-momentum, y-momentum, continuity), columns (u, v, p); no inverse-mass scaling.
//
// Reference semantics restated here (closed form validated in SURVEY.md App. A.11):
//   Stokes.assemble_BSR_Stokes_local_order                      dgfem/discrete_system.py:812-965
//   Stokes.assemble_RHS_Stokes                                  dgfem/discrete_system.py:967-1028
//   Element.compute_continuity_volume_integral                  dgfem/element.py:169-179
//   Element.compute_momentum_pressure_volume_integral           dgfem/element.py:201-211
//   Element.compute_momentum_velocity_penalty_volume_integral   dgfem/element.py:213-231
//   Face.compute_continuity_surface_integral                    dgfem/face.py:79-113
//   Face.compute_momentum_pressure_surface_integral             dgfem/face.py:282-320
//   Face.compute_momentum_velocity_penalty_surface_integral     dgfem/face.py:322-371
// Velocity and pressure use different quadrature sets: 'u' points for the momentum rows, 'p' points
// for the continuity row (dgfem/grid.py:107,185-187), so geometry exists at both point sets.
#include "dgb_async.cuh"
#include "dgb_tables.cuh"

namespace dgb {

struct StokesTabs {
    TabView uu;   // velocity basis @ velocity points (+ geometry operators of the velocity points)
    TabView pu;   // pressure basis @ velocity points
    TabView up;   // velocity basis @ pressure points (+ geometry operators of the pressure points)
    TabView pp;   // pressure basis @ pressure points
};

struct FaceInfoS {
    double c;      // 1/2 interior, 1 boundary
    double pen;    // sigma nu / h_F
    double gpen;   // gamma / h_F
    double st;     // +1 if this element is the L side (max faces), -1 on min faces
    int nbr;
};

// one CTA per element
__global__ void __launch_bounds__(256)
k_assemble_stokes(StokesTabs T, const double *__restrict__ vol_u, const double *__restrict__ face_u,
                  const double *__restrict__ vol_p, const double *__restrict__ face_p,
                  const double *__restrict__ area, Stencil S, double nu, double sigma, double gamma,
                  int pin_pressure, int32_t *__restrict__ indptr, int32_t *__restrict__ indices,
                  double *__restrict__ data) {
    extern __shared__ double sm[];
    const int bu = T.uu.b, bp = T.pp.b, bt = 2 * bu + bp, bb = bt * bt;
    const int nqu = T.uu.nq, nqp = T.pp.nq, n1u = T.uu.nq1, n1p = T.pp.nq1;
    double *blk = sm;                          // [5][bt*bt]
    double *Dxu = blk + 5 * bb;                // [nqu][bu]
    double *Dyu = Dxu + nqu * bu;
    double *Dxp = Dyu + nqu * bu;              // [nqp][bu]  velocity-basis derivatives at the pressure points
    double *Dyp = Dxp + nqp * bu;
    double *wJu = Dyp + nqp * bu;              // [nqu]
    double *wJp = wJu + nqu;                   // [nqp]
    double *Wu = wJp + nqp;                    // [4][n1u]
    double *Wp = Wu + 4 * n1u;                 // [4][n1p]
    double *dnO = Wp + 4 * n1p;                // [4][n1u][bu]
    double *dnN = dnO + 4 * n1u * bu;          // [4][n1u][bu]
    double *nOu = dnN + 4 * n1u * bu;          // [4][n1u][2]  own / neighbour unit normals, velocity face points
    double *nNu = nOu + 4 * n1u * 2;
    double *nOp = nNu + 4 * n1u * 2;           // [4][n1p][2]  pressure face points
    double *nNp = nOp + 4 * n1p * 2;
    __shared__ FaceInfoS fi[4];
    __shared__ int s_cols[5], s_rank[5];
    __shared__ long long s_row0;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int N = S.Ni * S.Nj;
    for (int e = blockIdx.x; e < N; e += gridDim.x) {
        const int i = e % S.Ni, j = e / S.Ni;
        if (!S.active(j)) {
            if (tid == 0) {
                indptr[e] = (int32_t)S.row_start(i, j);
                if (e == N - 1) indptr[N] = (int32_t)S.row_start(0, S.Nj);
            }
            continue;
        }
        if (tid == 0) {
            int c[5], rk[5];
            S.cols(i, j, c);
            slot_ranks(c, rk);
            for (int s = 0; s < 5; ++s) { s_cols[s] = c[s]; s_rank[s] = rk[s]; }
            s_row0 = S.row_start(i, j);
            const double Ae = area[e];
            for (int f = 0; f < 4; ++f) {
                const int nb = c[1 + f];
                FaceInfoS q;
                q.nbr = nb;
                q.st = (f & 1) ? 1.0 : -1.0;
                const double hF = nb >= 0 ? 0.5 * (sqrt(Ae) + sqrt(area[nb])) : sqrt(Ae);   // face.py:14,21,28
                q.pen = sigma * nu / hF;
                q.gpen = gamma / hF;
                q.c = nb >= 0 ? 0.5 : 1.0;
                fi[f] = q;
            }
        }
        __syncthreads();
        const double *vu = vol_u + (size_t)e * VOL_NC * nqu;
        const double *vp = vol_p + (size_t)e * VOL_NC * nqp;
        for (int t = tid; t < nqu * bu; t += nt) {
            const int q = t / bu;
            Dxu[t] = T.uu.Vr[t] * vu[nqu + q] + T.uu.Vs[t] * vu[2 * nqu + q];
            Dyu[t] = T.uu.Vr[t] * vu[3 * nqu + q] + T.uu.Vs[t] * vu[4 * nqu + q];
        }
        for (int t = tid; t < nqp * bu; t += nt) {
            const int q = t / bu;
            Dxp[t] = T.up.Vr[t] * vp[nqp + q] + T.up.Vs[t] * vp[2 * nqp + q];
            Dyp[t] = T.up.Vr[t] * vp[3 * nqp + q] + T.up.Vs[t] * vp[4 * nqp + q];
        }
        for (int q = tid; q < nqu; q += nt) wJu[q] = vu[q] * T.uu.w2[q];
        for (int q = tid; q < nqp; q += nt) wJp[q] = vp[q] * T.pp.w2[q];
        // face weights and normals at both point sets (face Jacobian of the L side, face.py:15,22,30)
        for (int t = tid; t < 4 * n1u; t += nt) {
            const int f = t / n1u, k = t - f * n1u;
            const int nb = fi[f].nbr;
            const double *fo = face_u + ((size_t)e * 4 + f) * FACE_NC * n1u;
            double Jf = fo[k];
            nOu[2 * t] = fo[5 * n1u + k]; nOu[2 * t + 1] = fo[6 * n1u + k];
            nNu[2 * t] = 0.0; nNu[2 * t + 1] = 0.0;
            if (nb >= 0) {
                const double *fn = face_u + ((size_t)nb * 4 + opp_face(f)) * FACE_NC * n1u;
                if (!(f & 1)) Jf = fn[k];
                nNu[2 * t] = fn[5 * n1u + k]; nNu[2 * t + 1] = fn[6 * n1u + k];
            }
            Wu[t] = Jf * T.uu.w1[k];
        }
        for (int t = tid; t < 4 * n1p; t += nt) {
            const int f = t / n1p, k = t - f * n1p;
            const int nb = fi[f].nbr;
            const double *fo = face_p + ((size_t)e * 4 + f) * FACE_NC * n1p;
            double Jf = fo[k];
            nOp[2 * t] = fo[5 * n1p + k]; nOp[2 * t + 1] = fo[6 * n1p + k];
            nNp[2 * t] = 0.0; nNp[2 * t + 1] = 0.0;
            if (nb >= 0) {
                const double *fn = face_p + ((size_t)nb * 4 + opp_face(f)) * FACE_NC * n1p;
                if (!(f & 1)) Jf = fn[k];
                nNp[2 * t] = fn[5 * n1p + k]; nNp[2 * t + 1] = fn[6 * n1p + k];
            }
            Wp[t] = Jf * T.pp.w1[k];
        }
        for (int t = tid; t < 4 * n1u * bu; t += nt) {
            const int f = t / (n1u * bu), rem = t - f * n1u * bu;
            const int k = rem / bu, l = rem - k * bu;
            const int nb = fi[f].nbr;
            const double *fo = face_u + ((size_t)e * 4 + f) * FACE_NC * n1u;
            const int trO = own_trace(f);
            dnO[t] = T.uu.Vrf[((size_t)trO * n1u + k) * bu + l] * fo[n1u + k] +
                     T.uu.Vsf[((size_t)trO * n1u + k) * bu + l] * fo[2 * n1u + k];
            double v = 0.0;
            if (nb >= 0) {
                const double *fn = face_u + ((size_t)nb * 4 + opp_face(f)) * FACE_NC * n1u;
                v = T.uu.Vrf[((size_t)f * n1u + k) * bu + l] * fn[n1u + k] +
                    T.uu.Vsf[((size_t)f * n1u + k) * bu + l] * fn[2 * n1u + k];
            }
            dnN[t] = v;
        }
        __syncthreads();
        // ---- block entries: item = (slot s, row R, col C) ----
        for (int t = tid; t < 5 * bb; t += nt) {
            const int s = t / bb, rc = t - s * bb;
            const int R = rc / bt, Cc = rc - R * bt;
            const int rtype = R < bu ? 0 : (R < 2 * bu ? 1 : 2);        // x-mom, y-mom, continuity
            const int ctype = Cc < bu ? 0 : (Cc < 2 * bu ? 1 : 2);      // u, v, p
            const int k = rtype == 0 ? R : (rtype == 1 ? R - bu : R - 2 * bu);
            const int l = ctype == 0 ? Cc : (ctype == 1 ? Cc - bu : Cc - 2 * bu);
            double acc = 0.0;
            if (s == 0) {
                // ---------------- diagonal block ----------------
                if (rtype < 2 && ctype < 2) {
                    const double *Da = rtype == 0 ? Dxu : Dyu;
                    const double *Db = ctype == 0 ? Dxu : Dyu;
                    double lap = 0.0, gd = 0.0;
                    for (int q = 0; q < nqu; ++q) {
                        if (rtype == ctype)
                            lap = fma(wJu[q], Dxu[q * bu + k] * Dxu[q * bu + l] + Dyu[q * bu + k] * Dyu[q * bu + l], lap);
                        gd = fma(wJu[q] * Da[q * bu + k], Db[q * bu + l], gd);
                    }
                    acc = nu * lap + gamma * gd;             // element.py:181-199, 213-231
                    for (int f = 0; f < 4; ++f) {
                        const int tr = own_trace(f);
                        const double *Vt = T.uu.Vf + (size_t)tr * n1u * bu;
                        const double *dn = dnO + (size_t)f * n1u * bu;
                        double flux = 0.0, pen = 0.0, sym = 0.0, vp2 = 0.0;
                        for (int q = 0; q < n1u; ++q) {
                            const double w = Wu[f * n1u + q];
                            const double vk = Vt[q * bu + k], vl = Vt[q * bu + l];
                            if (rtype == ctype) {
                                flux = fma(vk * w, dn[q * bu + l], flux);
                                pen = fma(vk * w, vl, pen);
                                sym = fma(dn[q * bu + k] * w, vl, sym);
                            }
                            const double na = nOu[2 * (f * n1u + q) + rtype], nb2 = nOu[2 * (f * n1u + q) + ctype];
                            vp2 = fma(vk * w * na * nb2, vl, vp2);
                        }
                        const double cn = fi[f].c * nu;
                        acc += -cn * fi[f].st * flux + fi[f].pen * pen - cn * fi[f].st * sym + fi[f].gpen * vp2;
                    }
                } else if (rtype < 2 && ctype == 2) {
                    // G: -int p div(v)  +  int {p} [v.n]            element.py:201-211, face.py:282-320
                    const double *Da = rtype == 0 ? Dxu : Dyu;
                    for (int q = 0; q < nqu; ++q) acc = fma(-wJu[q] * Da[q * bu + k], T.pu.V[q * bp + l], acc);
                    for (int f = 0; f < 4; ++f) {
                        const int tr = own_trace(f);
                        const double *Vt = T.uu.Vf + (size_t)tr * n1u * bu;
                        const double *Pt = T.pu.Vf + (size_t)tr * n1u * bp;
                        double g = 0.0;
                        for (int q = 0; q < n1u; ++q)
                            g = fma(Vt[q * bu + k] * Wu[f * n1u + q] * nOu[2 * (f * n1u + q) + rtype], Pt[q * bp + l], g);
                        acc += fi[f].c * fi[f].st * g;
                    }
                } else if (rtype == 2 && ctype < 2) {
                    // D: -int q div(u)  +  int [u.n] {q}   at the pressure points   element.py:169-179, face.py:79-113
                    const double *Db = ctype == 0 ? Dxp : Dyp;
                    for (int q = 0; q < nqp; ++q) acc = fma(-wJp[q] * T.pp.V[q * bp + k], Db[q * bu + l], acc);
                    for (int f = 0; f < 4; ++f) {
                        const int tr = own_trace(f);
                        const double *Pt = T.pp.Vf + (size_t)tr * n1p * bp;
                        const double *Vt = T.up.Vf + (size_t)tr * n1p * bu;
                        double g = 0.0;
                        for (int q = 0; q < n1p; ++q)
                            g = fma(Pt[q * bp + k] * Wp[f * n1p + q] * nOp[2 * (f * n1p + q) + ctype], Vt[q * bu + l], g);
                        acc += fi[f].c * fi[f].st * g;          // s_u = s_own for the diagonal block
                    }
                }
            } else {
                // ---------------- off-diagonal block of face f: trial side = neighbour ----------------
                const int f = s - 1;
                if (fi[f].nbr >= 0) {
                    const int tr = own_trace(f);
                    if (rtype < 2 && ctype < 2) {
                        const double *Vt = T.uu.Vf + (size_t)tr * n1u * bu;
                        const double *Vn = T.uu.Vf + (size_t)f * n1u * bu;
                        const double *dO = dnO + (size_t)f * n1u * bu;
                        const double *dN = dnN + (size_t)f * n1u * bu;
                        double flux = 0.0, pen = 0.0, sym = 0.0, vp2 = 0.0;
                        for (int q = 0; q < n1u; ++q) {
                            const double w = Wu[f * n1u + q];
                            const double vk = Vt[q * bu + k], vl = Vn[q * bu + l];
                            if (rtype == ctype) {
                                flux = fma(vk * w, dN[q * bu + l], flux);
                                pen = fma(vk * w, vl, pen);
                                sym = fma(dO[q * bu + k] * w, vl, sym);
                            }
                            const double na = nNu[2 * (f * n1u + q) + rtype], nb2 = nNu[2 * (f * n1u + q) + ctype];
                            vp2 = fma(vk * w * na * nb2, vl, vp2);
                        }
                        const double cn = fi[f].c * nu;
                        acc = -cn * fi[f].st * flux - fi[f].pen * pen + cn * fi[f].st * sym - fi[f].gpen * vp2;
                    } else if (rtype < 2 && ctype == 2) {
                        const double *Vt = T.uu.Vf + (size_t)tr * n1u * bu;
                        const double *Pn = T.pu.Vf + (size_t)f * n1u * bp;
                        double g = 0.0;
                        for (int q = 0; q < n1u; ++q)
                            g = fma(Vt[q * bu + k] * Wu[f * n1u + q] * nNu[2 * (f * n1u + q) + rtype], Pn[q * bp + l], g);
                        acc = fi[f].c * fi[f].st * g;
                    } else if (rtype == 2 && ctype < 2) {
                        const double *Pt = T.pp.Vf + (size_t)tr * n1p * bp;
                        const double *Vn = T.up.Vf + (size_t)f * n1p * bu;
                        double g = 0.0;
                        for (int q = 0; q < n1p; ++q)
                            g = fma(Pt[q * bp + k] * Wp[f * n1p + q] * nNp[2 * (f * n1p + q) + ctype], Vn[q * bu + l], g);
                        acc = -fi[f].c * fi[f].st * g;          // s_u = -s_own
                    }
                }
            }
            blk[t] = acc;
        }
        __syncthreads();
        // pressure pin of the direct solver (discrete_system.py:946): first stored block of element 0
        for (int t = tid; t < 5 * bb; t += nt) {
            const int s = t / bb, rc = t - s * bb;
            const int rk = s_rank[s];
            if (rk < 0) continue;
            double v = blk[t];
            if (pin_pressure && e == 0 && rk == 0 && rc == (2 * bu) * bt + 2 * bu) v = 1.0;
            data[((size_t)s_row0 + rk) * bb + rc] = v;
        }
        if (tid < 5 && s_rank[tid] >= 0) indices[s_row0 + s_rank[tid]] = s_cols[tid];
        if (tid == 0) {
            indptr[e] = (int32_t)s_row0;
            if (e == N - 1) indptr[N] = (int32_t)S.row_start(0, S.Nj);
        }
        __syncthreads();
    }
}

// RHS: one CTA of 64 threads per element (b_tot <= 64)
__global__ void __launch_bounds__(64)
k_assemble_rhs_stokes(StokesTabs T, const double *__restrict__ vol_u, const double *__restrict__ face_u,
                      const double *__restrict__ vol_p, const double *__restrict__ face_p,
                      const double *__restrict__ area, const double *__restrict__ f_mom /* [N][2][nqu] */,
                      const double *__restrict__ f_cont /* [N][nqp] */, const double *__restrict__ g_u /* [N][4][2][n1u] */,
                      const double *__restrict__ g_p /* [N][4][2][n1p] */, Stencil S, double nu, double sigma,
                      double gamma, double *__restrict__ rhs) {
    const int bu = T.uu.b, bp = T.pp.b, bt = 2 * bu + bp;
    const int nqu = T.uu.nq, nqp = T.pp.nq, n1u = T.uu.nq1, n1p = T.pp.nq1;
    const int N = S.Ni * S.Nj;
    const int R = threadIdx.x;
    for (int e = blockIdx.x; e < N; e += gridDim.x) {
        const int i = e % S.Ni, j = e / S.Ni;
        if (R >= bt) continue;
        if (!S.active(j)) { rhs[(size_t)e * bt + R] = 0.0; continue; }
        int c[5];
        S.cols(i, j, c);
        const int rtype = R < bu ? 0 : (R < 2 * bu ? 1 : 2);
        const int k = rtype == 0 ? R : (rtype == 1 ? R - bu : R - 2 * bu);
        const double hF = sqrt(area[e]);
        double acc = 0.0;
        if (rtype < 2) {
            const double *vu = vol_u + (size_t)e * VOL_NC * nqu;
            const double *fv = f_mom + ((size_t)e * 2 + rtype) * nqu;
            for (int q = 0; q < nqu; ++q) acc = fma(T.uu.V[q * bu + k] * (vu[q] * T.uu.w2[q]), fv[q], acc);   // element.py:161-167
            for (int f = 0; f < 4; ++f) {
                if (c[1 + f] >= 0) continue;
                const double *fo = face_u + ((size_t)e * 4 + f) * FACE_NC * n1u;
                const double *ga = g_u + (((size_t)e * 4 + f) * 2 + rtype) * n1u;
                const double *gx = g_u + (((size_t)e * 4 + f) * 2 + 0) * n1u;
                const double *gy = g_u + (((size_t)e * 4 + f) * 2 + 1) * n1u;
                const int tr = own_trace(f);
                const double st = (f & 1) ? 1.0 : -1.0;
                double pen = 0.0, sym = 0.0, vpn = 0.0;
                for (int q = 0; q < n1u; ++q) {
                    const double w = T.uu.w1[q] * fo[q];
                    const double vk = T.uu.Vf[((size_t)tr * n1u + q) * bu + k];
                    const double dn = T.uu.Vrf[((size_t)tr * n1u + q) * bu + k] * fo[n1u + q] +
                                      T.uu.Vsf[((size_t)tr * n1u + q) * bu + k] * fo[2 * n1u + q];
                    const double nx = fo[5 * n1u + q], ny = fo[6 * n1u + q];
                    pen = fma(vk, ga[q] * w, pen);
                    sym = fma(dn, ga[q] * w, sym);
                    vpn = fma(vk * (rtype == 0 ? nx : ny), (gx[q] * nx + gy[q] * ny) * w, vpn);   // face.py:325-341
                }
                acc += sigma * nu / hF * pen - st * nu * sym + gamma / hF * vpn;
            }
        } else {
            const double *vp = vol_p + (size_t)e * VOL_NC * nqp;
            const double *fc = f_cont + (size_t)e * nqp;
            for (int q = 0; q < nqp; ++q) acc = fma(-T.pp.V[q * bp + k] * (vp[q] * T.pp.w2[q]), fc[q], acc);   // element.py:158-159
            for (int f = 0; f < 4; ++f) {
                if (c[1 + f] >= 0) continue;
                const double *fo = face_p + ((size_t)e * 4 + f) * FACE_NC * n1p;
                const double *gx = g_p + (((size_t)e * 4 + f) * 2 + 0) * n1p;
                const double *gy = g_p + (((size_t)e * 4 + f) * 2 + 1) * n1p;
                const int tr = own_trace(f);
                const double st = (f & 1) ? 1.0 : -1.0;
                double s = 0.0;
                for (int q = 0; q < n1p; ++q) {
                    const double w = T.pp.w1[q] * fo[q];
                    s = fma(T.pp.Vf[((size_t)tr * n1p + q) * bp + k], (gx[q] * fo[5 * n1p + q] + gy[q] * fo[6 * n1p + q]) * w, s);
                }
                acc += st * s;                                  // face.py:82-83 (-), :91-92 (+)
            }
        }
        rhs[(size_t)e * bt + R] = acc;
    }
}

}  // namespace dgb

using namespace dgb;

extern "C" {

int dgb_assemble_stokes(const dgb_tables *t_uu, const dgb_tables *t_pu, const dgb_tables *t_up,
                        const dgb_tables *t_pp, const double *vol_u, const double *face_u, const double *vol_p,
                        const double *face_p, const double *area, int32_t Ni, int32_t Nj, double nu, double sigma,
                        double gamma, int32_t flags, int32_t pin_pressure, int32_t *indptr, int32_t *indices,
                        double *data, void *stream) {
    DGB_ARG(t_uu && t_pu && t_up && t_pp && vol_u && face_u && vol_p && face_p && area && indptr && indices && data);
    DGB_ARG(t_uu->b == t_up->b && t_pu->b == t_pp->b && t_uu->nq1 == t_pu->nq1 && t_up->nq1 == t_pp->nq1);
    StokesTabs T{view(t_uu), view(t_pu), view(t_up), view(t_pp)};
    Stencil S = make_stencil(Ni, Nj, flags);
    const size_t bu = T.uu.b, bp = T.pp.b, bt = 2 * bu + bp;
    const size_t nqu = T.uu.nq, nqp = T.pp.nq, n1u = T.uu.nq1, n1p = T.pp.nq1;
    const size_t smem = sizeof(double) * (5 * bt * bt + 2 * nqu * bu + 2 * nqp * bu + nqu + nqp + 4 * n1u + 4 * n1p +
                                          2 * 4 * n1u * bu + 2 * 8 * n1u + 2 * 8 * n1p);
    DGB_ARG(smem <= 227 * 1024);
    cudaStream_t st = (cudaStream_t)stream;
    DGB_CUDA_OK(cudaFuncSetAttribute(k_assemble_stokes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t g = (int64_t)Ni * Nj;
    if (g > sm_count() * 16) g = sm_count() * 16;
    k_assemble_stokes<<<(int)g, 256, smem, st>>>(T, vol_u, face_u, vol_p, face_p, area, S, nu, sigma, gamma,
                                               pin_pressure, indptr, indices, data);
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_assemble_rhs_stokes(const dgb_tables *t_uu, const dgb_tables *t_pu, const dgb_tables *t_up,
                            const dgb_tables *t_pp, const double *vol_u, const double *face_u, const double *vol_p,
                            const double *face_p, const double *area, const double *f_mom, const double *f_cont,
                            const double *g_u, const double *g_p, int32_t Ni, int32_t Nj, double nu, double sigma,
                            double gamma, int32_t flags, double *rhs, void *stream) {
    DGB_ARG(t_uu && t_pu && t_up && t_pp && vol_u && face_u && vol_p && face_p && area && f_mom && f_cont && g_u &&
            g_p && rhs);
    StokesTabs T{view(t_uu), view(t_pu), view(t_up), view(t_pp)};
    DGB_ARG(2 * T.uu.b + T.pp.b <= 64);
    Stencil S = make_stencil(Ni, Nj, flags);
    int64_t g = (int64_t)Ni * Nj;
    if (g > sm_count() * 32) g = sm_count() * 32;
    k_assemble_rhs_stokes<<<(int)g, 64, 0, (cudaStream_t)stream>>>(T, vol_u, face_u, vol_p, face_p, area, f_mom, f_cont,
                                                                 g_u, g_p, S, nu, sigma, gamma, rhs);
    DGB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
