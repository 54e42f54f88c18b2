// dgb_assemble.cu -- K0..K3: batched DG assembly of the SIPG Poisson operator on sm_100a.
//
// Reference semantics restated here (closed form validated in SURVEY.md App. A.6a):
//   K0  Element.compute_geometric_terms / metric_xy_rs        dgfem/element.py:52-130
//       CoarseElement._init_coarse_element                    dgfem/element.py:242-356
//   K1  Element.compute_momentum_laplace_volume_integral      dgfem/element.py:181-199
//       Element.compute_mass_matrix + np.linalg.inv           dgfem/element.py:132-133,
//                                                             dgfem/discrete_system.py:127-128
//   K2  Face.compute_momentum_laplace_SIP_terms               dgfem/face.py:115-280
//       Poisson.assemble_BSR_Poisson (row layout, sorting)    dgfem/discrete_system.py:54-145
//   K3  Poisson.assemble_RHS_Poisson                          dgfem/discrete_system.py:355-403
//
// Every interior face is evaluated once per adjacent element (test side = that element),
// i.e. only the two blocks that element's row needs -- the reference evaluates all four
// blocks twice and discards half (discrete_system.py:74-77).
#include "dgb_async.cuh"
#include "dgb_common.cuh"

#include "dgb_tables.cuh"
#include "dgb_mma.cuh"

namespace dgb {

extern int g_gs_variant;

// ---------------------------------------------------------------------------------------
// K0: one thread per (element, point); point < nq: volume point, else face point.
__global__ void __launch_bounds__(128)
k_metrics(TabView T, const double *__restrict__ xn, const double *__restrict__ yn, int il, int Ni,
          int Nj, double *__restrict__ vol, double *__restrict__ face) {
    const int npt = T.nq + 4 * T.nq1;
    const int64_t total = (int64_t)Ni * Nj * npt;
    const int N1 = T.Pg + 1;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
         it += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(it / npt);
        const int pt = (int)(it - (int64_t)e * npt);
        const int i = e % Ni, j = e / Ni;
        const double *gx, *gr, *gs;
        int sm, sn, f = -1, t = pt;
        if (pt < T.nq) {
            gx = T.GX + (size_t)pt * T.ng; gr = T.GR + (size_t)pt * T.ng; gs = T.GS + (size_t)pt * T.ng;
            sm = T.sub_vol[2 * pt]; sn = T.sub_vol[2 * pt + 1];
        } else {
            const int q = pt - T.nq;
            f = q / T.nq1; t = q - f * T.nq1;
            gx = T.FX + (size_t)q * T.ng; gr = T.FR + (size_t)q * T.ng; gs = T.FS + (size_t)q * T.ng;
            sm = T.sub_face[2 * q]; sn = T.sub_face[2 * q + 1];
        }
        // source (fine) element and its first node in the Plot3D array [jl][il]
        const int64_t fi = (int64_t)i * T.cf + sm, fj = (int64_t)j * T.cf + sn;
        const double *xe = xn + (fj * T.Pg) * il + fi * T.Pg;
        const double *ye = yn + (fj * T.Pg) * il + fi * T.Pg;
        double x = 0, y = 0, xr = 0, xs = 0, yr = 0, ys = 0;
        for (int c = 0; c < N1; ++c) {
            for (int a = 0; a < N1; ++a) {
                const int n = a + N1 * c;                 // F-order node index (element.py:79)
                const double xv = xe[(int64_t)c * il + a], yv = ye[(int64_t)c * il + a];
                const double lx = gx[n], lr = gr[n], ls = gs[n];
                x = fma(lx, xv, x);   y = fma(lx, yv, y);
                xr = fma(lr, xv, xr); yr = fma(lr, yv, yr);
                xs = fma(ls, xv, xs); ys = fma(ls, yv, ys);
            }
        }
        const double J = xr * ys - yr * xs;               // element.py:93
        const double rx = ys / J, sx = -yr / J, ry = -xs / J, sy = xr / J;   // :94-95
        if (f < 0) {
            double *o = vol + (size_t)e * VOL_NC * T.nq + t;
            o[0] = J; o[T.nq] = rx; o[2 * T.nq] = sx; o[3 * T.nq] = ry; o[4 * T.nq] = sy;
            o[5 * T.nq] = x; o[6 * T.nq] = y;
        } else {
            double Jf, nx, ny;
            if (f < 2) {                                  // i-faces (element.py:97-99)
                Jf = sqrt(xs * xs + ys * ys);
                const double nn = sqrt(rx * rx + ry * ry);
                nx = rx / nn; ny = ry / nn;
            } else {                                      // j-faces (element.py:100-102)
                Jf = sqrt(xr * xr + yr * yr);
                const double nn = sqrt(sx * sx + sy * sy);
                nx = sx / nn; ny = sy / nn;
            }
            double *o = face + ((size_t)e * 4 + f) * FACE_NC * T.nq1 + t;
            o[0] = Jf;
            o[T.nq1] = nx * rx + ny * ry;                 // alpha: d_n phi = Vr*alpha + Vs*beta
            o[2 * T.nq1] = nx * sx + ny * sy;             // beta
            o[3 * T.nq1] = x; o[4 * T.nq1] = y; o[5 * T.nq1] = nx; o[6 * T.nq1] = ny; o[7 * T.nq1] = 0.0;
        }
    }
}

__global__ void __launch_bounds__(256)
k_area(TabView T, const double *__restrict__ vol, int64_t N, double *__restrict__ area) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N; e += (int64_t)gridDim.x * blockDim.x) {
        const double *J = vol + (size_t)e * VOL_NC * T.nq;
        double a = 0.0;
        for (int q = 0; q < T.nq; ++q) a = fma(J[q], T.w2[q], a);   // element.py:30
        area[e] = a;
    }
}

// in-place Gauss-Jordan inverse with partial pivoting by one warp (matrix in smem)
__device__ __forceinline__ bool warp_invert(double *a, int *piv, int B, int lane) {
    for (int k = 0; k < B; ++k) {
        double best = -1.0;
        int bi = k;
        for (int i = k + lane; i < B; i += 32) {
            const double v = fabs(a[i * B + k]);
            if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) piv[k] = bi;
        if (best == 0.0 || !(best == best)) return false;
        if (bi != k)
            for (int c = lane; c < B; c += 32) {
                const double t = a[k * B + c];
                a[k * B + c] = a[bi * B + c];
                a[bi * B + c] = t;
            }
        __syncwarp();
        const double pinv = 1.0 / a[k * B + k];
        __syncwarp();
        for (int c = lane; c < B; c += 32) a[k * B + c] = (c == k) ? pinv : a[k * B + c] * pinv;
        __syncwarp();
        for (int t = lane; t < B * B; t += 32) {
            const int i = t / B, c = t - i * B;
            if (i == k || c == k) continue;
            a[t] = fma(-a[i * B + k], a[k * B + c], a[t]);
        }
        __syncwarp();
        for (int i = lane; i < B; i += 32)
            if (i != k) a[i * B + k] = -a[i * B + k] * pinv;
        __syncwarp();
    }
    for (int k = B - 1; k >= 0; --k) {
        const int p = piv[k];
        if (p != k)
            for (int i = lane; i < B; i += 32) {
                const double t = a[i * B + k];
                a[i * B + k] = a[i * B + p];
                a[i * B + p] = t;
            }
        __syncwarp();
    }
    return true;
}

// ---------------------------------------------------------------------------------------
// K1+K2: one CTA per element.
// smem: blk[6][b*b] (diag, iL, iR, jL, jR, M), Dx[nq][b], Dy[nq][b], wJ[nq],
//       per face f: W[f][nq1], dnO[f][nq1][b], dnN[f][nq1][b], scalars
struct FaceInfo {
    double c_nu;   // c * nu (c = 1/2 interior, 1 boundary)
    double pen;    // sigma nu / h_F
    double st;     // +1 if this element is the L side of the face (max faces), -1 on min faces
    int nbr;       // neighbour element or -1
};

// BT, QT: block size and quadrature points per direction as compile-time constants (0 = read them from the
// tables): all the index arithmetic of the item loops then folds into multiplications
template <int BT, int QT>
__global__ void __launch_bounds__(256)
k_assemble_poisson(TabView T, const double *__restrict__ vol, const double *__restrict__ face,
                   const double *__restrict__ area, Stencil S, double nu, double sigma, int use_minv,
                   int32_t *__restrict__ indptr, int32_t *__restrict__ indices,
                   double *__restrict__ data, double *__restrict__ minv_out) {
    extern __shared__ double sm[];
    const int b = BT > 0 ? BT : T.b, nq1 = QT > 0 ? QT : T.nq1, nq = nq1 * nq1, bb = b * b;
    double *blk = sm;                       // [6][bb]
    double *Dx = blk + 6 * bb;              // [nq][b]
    double *Dy = Dx + nq * b;               // [nq][b]
    double *wJ = Dy + nq * b;               // [nq]
    double *W = wJ + nq;                    // [4][nq1]
    double *dnO = W + 4 * nq1;              // [4][nq1][b]
    double *dnN = dnO + 4 * nq1 * b;        // [4][nq1][b]
    __shared__ FaceInfo fi[4];
    __shared__ int s_cols[5], s_rank[5], s_piv[64];
    __shared__ int64_t s_row0;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int N = S.Ni * S.Nj;
    for (int e = blockIdx.x; e < N; e += gridDim.x) {
        const int i = e % S.Ni, j = e / S.Ni;
        if (!S.active(j)) {       // ghost row of a slab: an empty matrix row
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
                FaceInfo q;
                q.nbr = nb;
                q.st = (f & 1) ? 1.0 : -1.0;
                const double hF = nb >= 0 ? 0.5 * (sqrt(Ae) + sqrt(area[nb])) : sqrt(Ae);   // face.py:14,21,28
                q.pen = sigma * nu / hF;
                q.c_nu = (nb >= 0 ? 0.5 : 1.0) * nu;
                fi[f] = q;
            }
        }
        __syncthreads();
        // volume derivative tables and weights (element.py:182-193)
        const double *ve = vol + (size_t)e * VOL_NC * nq;
        for (int t = tid; t < nq * b; t += nt) {
            const int q = t / b, l = t - q * b;
            const double vr = T.Vr[t], vs = T.Vs[t];
            Dx[t] = vr * ve[nq + q] + vs * ve[2 * nq + q];
            Dy[t] = vr * ve[3 * nq + q] + vs * ve[4 * nq + q];
            (void)l;
        }
        for (int q = tid; q < nq; q += nt) wJ[q] = ve[q] * T.w2[q];
        // face tables: W = w1 * Jf (L side's max-face J if L exists, face.py:15,22,30)
        for (int t = tid; t < 4 * nq1; t += nt) {
            const int f = t / nq1, k = t - f * nq1;
            const int nb = fi[f].nbr;
            const double *fo = face + ((size_t)e * 4 + f) * FACE_NC * nq1;
            double Jf = fo[k];
            if (!(f & 1) && nb >= 0)       // min face with an L neighbour: use L's max-face J
                Jf = face[((size_t)nb * 4 + opp_face(f)) * FACE_NC * nq1 + k];
            W[t] = Jf * T.w1[k];
        }
        for (int t = tid; t < 4 * nq1 * b; t += nt) {
            const int f = t / (nq1 * b), rem = t - f * nq1 * b;
            const int k = rem / b, l = rem - k * b;
            const int nb = fi[f].nbr;
            const double *fo = face + ((size_t)e * 4 + f) * FACE_NC * nq1;
            const int trO = own_trace(f);
            dnO[t] = T.Vrf[((size_t)trO * nq1 + k) * b + l] * fo[nq1 + k] +
                     T.Vsf[((size_t)trO * nq1 + k) * b + l] * fo[2 * nq1 + k];
            double v = 0.0;
            if (nb >= 0) {
                const double *fn = face + ((size_t)nb * 4 + opp_face(f)) * FACE_NC * nq1;
                const int trN = f;          // neighbour's own trace on its opposite face = own_trace(f^1) = f
                v = T.Vrf[((size_t)trN * nq1 + k) * b + l] * fn[nq1 + k] +
                    T.Vsf[((size_t)trN * nq1 + k) * b + l] * fn[2 * nq1 + k];
            }
            dnN[t] = v;
        }
        __syncthreads();
        // the mass matrix first (slot 5) ...
        for (int kl = tid; kl < bb; kl += nt) {
            const int k = kl / b, l = kl - k * b;
            double acc = 0.0;
            for (int q = 0; q < nq; ++q) acc = fma(T.V[q * b + k] * wJ[q], T.V[q * b + l], acc);
            blk[5 * bb + kl] = acc;
        }
        __syncthreads();
        // ... then warp 0 inverts it in place while the other warps compute the five operator blocks
        // (item = (slot s in 0..4, k, l))
        if (tid < 32) {
            warp_invert(blk + 5 * bb, s_piv, b, tid);
        }
        const int first = nt > 32 ? 32 : 0, nw = nt > 32 ? nt - 32 : nt;
        for (int t = tid - first; t >= 0 && t < 5 * bb; t += nw) {
            const int s = t / bb, kl = t - s * bb;
            const int k = kl / b, l = kl - k * b;
            double acc = 0.0;
            if (s == 0) {
                double kv = 0.0;
                for (int q = 0; q < nq; ++q)
                    kv = fma(wJ[q], Dx[q * b + k] * Dx[q * b + l] + Dy[q * b + k] * Dy[q * b + l], kv);
                acc = nu * kv;
                for (int f = 0; f < 4; ++f) {
                    const int tr = own_trace(f);
                    const double *Vt = T.Vf + (size_t)tr * nq1 * b;
                    const double *dn = dnO + (size_t)f * nq1 * b;
                    double flux = 0.0, pen = 0.0, sym = 0.0;
                    for (int q = 0; q < nq1; ++q) {
                        const double w = W[f * nq1 + q];
                        const double vk = Vt[q * b + k], vl = Vt[q * b + l];
                        flux = fma(vk * w, dn[q * b + l], flux);
                        pen = fma(vk * w, vl, pen);
                        sym = fma(dn[q * b + k] * w, vl, sym);
                    }
                    acc += -fi[f].c_nu * fi[f].st * flux + fi[f].pen * pen - fi[f].c_nu * fi[f].st * sym;
                }
            } else {
                const int f = s - 1;
                if (fi[f].nbr >= 0) {
                    const int tr = own_trace(f);
                    const double *Vt = T.Vf + (size_t)tr * nq1 * b;
                    const double *Vn = T.Vf + (size_t)f * nq1 * b;      // neighbour's trace
                    const double *dO = dnO + (size_t)f * nq1 * b;
                    const double *dN = dnN + (size_t)f * nq1 * b;
                    double flux = 0.0, pen = 0.0, sym = 0.0;
                    for (int q = 0; q < nq1; ++q) {
                        const double w = W[f * nq1 + q];
                        const double vk = Vt[q * b + k];
                        flux = fma(vk * w, dN[q * b + l], flux);
                        pen = fma(vk * w, Vn[q * b + l], pen);
                        sym = fma(dO[q * b + k] * w, Vn[q * b + l], sym);
                    }
                    // s_u = -s_t
                    acc = -fi[f].c_nu * fi[f].st * flux - fi[f].pen * pen + fi[f].c_nu * fi[f].st * sym;
                }
            }
            blk[t] = acc;
        }
        __syncthreads();
        const double *Mi = blk + 5 * bb;
        for (int t = tid; t < bb; t += nt) minv_out[(size_t)e * bb + t] = Mi[t];
        // out = Minv * blk (or blk), written in ascending-column order
        for (int t = tid; t < 5 * bb; t += nt) {
            const int s = t / bb, kl = t - s * bb;
            const int rk = s_rank[s];
            if (rk < 0) continue;
            const int k = kl / b, l = kl - k * b;
            double v;
            if (use_minv) {
                v = 0.0;
                const double *src = blk + s * bb;
                for (int m = 0; m < b; ++m) v = fma(Mi[k * b + m], src[m * b + l], v);
            } else {
                v = blk[t];
            }
            data[((size_t)s_row0 + rk) * bb + kl] = v;
        }
        if (tid < 5 && s_rank[tid] >= 0) indices[s_row0 + s_rank[tid]] = s_cols[tid];
        if (tid == 0) {
            indptr[e] = (int32_t)s_row0;
            if (e == N - 1) indptr[N] = (int32_t)S.row_start(0, S.Nj);
        }
        __syncthreads();
    }
}

// =========================================================================================
// K1+K2 on the FP64 tensor cores (b >= 16, i.e. p >= 3).
//
// Every block of an element row is a sum over "points" t of outer products, C[k][l] = sum_t L[t][k] R[t][l]:
//   mass      L = w J V,                 R = V                                    (nq points)
//   diagonal  L = nu w J Dx | nu w J Dy | W Vt      | -c st W dnO                 (2 nq + 8 nq1 points)
//             R = Dx        | Dy        | pen Vt - c st dnO | Vt
//   neighbour L = W Vt                  | c st W dnO                              (2 nq1 points per face)
//             R = -c st dnN - pen Vn    | Vn
//   out       = Minv blk  with L = Minv (symmetric), R = blk                      (b points)
// (the same terms k_assemble_poisson sums entry by entry; face.py:115-280, element.py:181-199).  The point tables
// are built in shared memory by the whole CTA, the products run as DMMA m8n8k4 (`mma.sync ... f64`): tile (I, J) of
// C takes its A fragment from L[t0 + lane%4][8I + lane/4] and its B fragment from R[t0 + lane%4][8J + lane/4].
// Measured on the p=5 contraction: 24.4 TFLOP/s against 5.0 for the per-entry loop and 10.0 for 4 x 4 register
// tiles (profiles/r02_dmma_vs_dfma.md).  The mass matrix is inverted by the whole CTA (Gauss-Jordan, partial
// pivoting), not by one warp.
// in-place Gauss-Jordan inverse with partial pivoting by the whole CTA (256 threads), matrix a[b][stride] in smem
template <int BT>
__device__ void cta_invert(double *a, int stride, double *col, int *piv, int *s_p) {
    const int tid = threadIdx.x;
    for (int k = 0; k < BT; ++k) {
        if (tid < 32) {                        // first maximum of |a[i][k]|, i >= k
            double best = -1.0;
            int bi = k;
            for (int i = k + tid; i < BT; i += 32) {
                const double v = fabs(a[i * stride + k]);
                if (v > best) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (tid == 0) { *s_p = bi; piv[k] = bi; }
        }
        __syncthreads();
        const int p = *s_p;
        if (p != k && tid < BT) {
            const double t = a[k * stride + tid];
            a[k * stride + tid] = a[p * stride + tid];
            a[p * stride + tid] = t;
        }
        __syncthreads();
        if (tid < BT) col[tid] = a[tid * stride + k];
        __syncthreads();
        const double pinv = 1.0 / col[k];
        if (tid < BT) a[k * stride + tid] = (tid == k) ? pinv : a[k * stride + tid] * pinv;
        __syncthreads();
        for (int t = tid; t < BT * BT; t += 256) {
            const int i = t / BT, c = t - i * BT;
            if (i == k) continue;
            a[i * stride + c] = (c == k) ? -col[i] * pinv : fma(-col[i], a[k * stride + c], a[i * stride + c]);
        }
        __syncthreads();
    }
    for (int k = BT - 1; k >= 0; --k) {        // undo the row interchanges as column interchanges
        const int p = piv[k];
        if (p != k && tid < BT) {
            const double t = a[tid * stride + k];
            a[tid * stride + k] = a[tid * stride + p];
            a[tid * stride + p] = t;
        }
        __syncthreads();
    }
}

template <int BT, int QT>
__global__ void __launch_bounds__(256)
k_assemble_poisson_mma(TabView T, const double *__restrict__ vol, const double *__restrict__ face,
                       const double *__restrict__ area, Stencil S, double nu, double sigma, int use_minv,
                       int32_t *__restrict__ indptr, int32_t *__restrict__ indices, double *__restrict__ data,
                       double *__restrict__ minv_out) {
    using C = MmaCfg<BT>;
    constexpr int b = BT, nq1 = QT, nq = QT * QT, bb = BT * BT, BS = C::BS, B4 = C::B4;
    constexpr int NQ4 = (nq + 3) & ~3;                  // volume points rounded up to the k-step
    constexpr int TCF = (2 * nq1 + 3) & ~3;             // rows per face in the neighbour tables
    constexpr int TC = NQ4 > 8 * nq1 ? (NQ4 > 4 * TCF ? NQ4 : 4 * TCF) : (8 * nq1 > 4 * TCF ? 8 * nq1 : 4 * TCF);
    extern __shared__ __align__(16) double sm[];
    double *Lt = sm;                          // [TC][BS] + 8
    double *Rt = Lt + TC * BS + 8;            // [TC][BS] + 8
    double *Mi = Rt + TC * BS + 8;            // [B4][BS] + 8: mass matrix, then its inverse
    double *Bk = Mi + B4 * BS + 8;            // [B4][BS] + 8: the block being assembled
    double *colv = Bk + B4 * BS + 8;          // [B4]
    __shared__ FaceInfo fi[4];
    __shared__ int s_cols[5], s_rank[5], s_piv[64], s_p;
    __shared__ int64_t s_row0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = S.Ni * S.Nj;
    // zero everything once: padding rows / columns of the tables stay zero for the whole kernel
    for (int t = tid; t < 2 * (TC * BS + 8) + 2 * (B4 * BS + 8) + B4; t += 256) sm[t] = 0.0;
    __syncthreads();
    for (int e = blockIdx.x; e < N; e += gridDim.x) {
        const int i = e % S.Ni, j = e / S.Ni;
        if (!S.active(j)) {       // ghost row of a slab: an empty matrix row
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
                FaceInfo q;
                q.nbr = nb;
                q.st = (f & 1) ? 1.0 : -1.0;
                const double hF = nb >= 0 ? 0.5 * (sqrt(Ae) + sqrt(area[nb])) : sqrt(Ae);   // face.py:14,21,28
                q.pen = sigma * nu / hF;
                q.c_nu = (nb >= 0 ? 0.5 : 1.0) * nu;
                fi[f] = q;
            }
        }
        const double *ve = vol + (size_t)e * VOL_NC * nq;
        double acc[C::MAXT][2];
        auto zero_acc = [&]() {
#pragma unroll
            for (int s = 0; s < C::MAXT; ++s) acc[s][0] = acc[s][1] = 0.0;
        };
        // ---- mass matrix: L = w J V, R = V ----
        for (int t = tid; t < nq * b; t += 256) {
            const int q = t / b, k = t - q * b;
            const double v = T.V[t];
            Lt[q * BS + k] = v * (ve[q] * T.w2[q]);
            Rt[q * BS + k] = v;
        }
        for (int t = tid; t < (NQ4 - nq) * BS; t += 256) Lt[nq * BS + t] = Rt[nq * BS + t] = 0.0;
        __syncthreads();
        zero_acc();
        mma_lr<BT>(acc, Lt, Rt, NQ4, warp, lane);
        mma_store<BT>(acc, Mi, BS, warp, lane);
        __syncthreads();
        cta_invert<BT>(Mi, BS, colv, s_piv, &s_p);
        for (int t = tid; t < bb; t += 256) minv_out[(size_t)e * bb + t] = Mi[(t / b) * BS + (t % b)];
        // out = Minv * Bk (or Bk), written as the block of sorted rank rk of this row
        auto emit = [&](int rk) {
            double *dst = data + ((size_t)s_row0 + rk) * bb;
            if (use_minv) {
                zero_acc();
                mma_lr<BT>(acc, Mi, Bk, B4, warp, lane);
                mma_store<BT>(acc, dst, b, warp, lane);
            } else {
                for (int t = tid; t < bb; t += 256) dst[t] = Bk[(t / b) * BS + (t % b)];
            }
        };
        // ---- diagonal block: volume part in two chunks (x-, y-derivative), then the four faces ----
        zero_acc();
        for (int part = 0; part < 2; ++part) {
            __syncthreads();
            for (int t = tid; t < nq * b; t += 256) {
                const int q = t / b, k = t - q * b;
                const double d = T.Vr[t] * ve[(1 + 2 * part) * nq + q] + T.Vs[t] * ve[(2 + 2 * part) * nq + q];   // element.py:182-193
                Lt[q * BS + k] = nu * (ve[q] * T.w2[q]) * d;
                Rt[q * BS + k] = d;
            }
            __syncthreads();
            mma_lr<BT>(acc, Lt, Rt, NQ4, warp, lane);
        }
        __syncthreads();
        for (int t = tid; t < 4 * nq1 * b; t += 256) {
            const int f = t / (nq1 * b), rem = t - f * nq1 * b;
            const int q = rem / b, k = rem - q * b;
            const int nb = fi[f].nbr;
            const double *fo = face + ((size_t)e * 4 + f) * FACE_NC * nq1;
            double Jf = fo[q];
            if (!(f & 1) && nb >= 0)       // min face with an L neighbour: use L's max-face J (face.py:15,22,30)
                Jf = face[((size_t)nb * 4 + opp_face(f)) * FACE_NC * nq1 + q];
            const double W = Jf * T.w1[q];
            const int tr = own_trace(f);
            const size_t ti = ((size_t)tr * nq1 + q) * b + k;
            const double vt = T.Vf[ti];
            const double dn = T.Vrf[ti] * fo[nq1 + q] + T.Vsf[ti] * fo[2 * nq1 + q];
            const double cst = fi[f].c_nu * fi[f].st;
            const int r0 = (f * 2) * nq1 + q, r1 = (f * 2 + 1) * nq1 + q;
            Lt[r0 * BS + k] = W * vt;           Rt[r0 * BS + k] = fi[f].pen * vt - cst * dn;
            Lt[r1 * BS + k] = -cst * W * dn;    Rt[r1 * BS + k] = vt;
        }
        __syncthreads();
        mma_lr<BT>(acc, Lt, Rt, 8 * nq1, warp, lane);
        mma_store<BT>(acc, Bk, BS, warp, lane);
        __syncthreads();
        emit(s_rank[0]);
        __syncthreads();
        // ---- neighbour blocks: tables of all four faces at once, one product per existing neighbour ----
        for (int t = tid; t < 4 * TCF * BS; t += 256) Lt[t] = Rt[t] = 0.0;
        __syncthreads();
        for (int t = tid; t < 4 * nq1 * b; t += 256) {
            const int f = t / (nq1 * b), rem = t - f * nq1 * b;
            const int q = rem / b, k = rem - q * b;
            const int nb = fi[f].nbr;
            if (nb < 0) continue;
            const double *fo = face + ((size_t)e * 4 + f) * FACE_NC * nq1;
            const double *fn = face + ((size_t)nb * 4 + opp_face(f)) * FACE_NC * nq1;
            const double Jf = (f & 1) ? fo[q] : fn[q];
            const double W = Jf * T.w1[q];
            const size_t to = ((size_t)own_trace(f) * nq1 + q) * b + k, tn = ((size_t)f * nq1 + q) * b + k;
            const double vt = T.Vf[to], vn = T.Vf[tn];
            const double dO = T.Vrf[to] * fo[nq1 + q] + T.Vsf[to] * fo[2 * nq1 + q];
            const double dN = T.Vrf[tn] * fn[nq1 + q] + T.Vsf[tn] * fn[2 * nq1 + q];
            const double cst = fi[f].c_nu * fi[f].st;
            const int r0 = f * TCF + q, r1 = f * TCF + nq1 + q;
            Lt[r0 * BS + k] = W * vt;           Rt[r0 * BS + k] = -cst * dN - fi[f].pen * vn;
            Lt[r1 * BS + k] = cst * W * dO;     Rt[r1 * BS + k] = vn;
        }
        __syncthreads();
        for (int f = 0; f < 4; ++f) {
            if (fi[f].nbr < 0) continue;          // uniform over the CTA
            zero_acc();
            mma_lr<BT>(acc, Lt + f * TCF * BS, Rt + f * TCF * BS, TCF, warp, lane);
            mma_store<BT>(acc, Bk, BS, warp, lane);
            __syncthreads();
            emit(s_rank[1 + f]);
            __syncthreads();
        }
        if (tid < 5 && s_rank[tid] >= 0) indices[s_row0 + s_rank[tid]] = s_cols[tid];
        if (tid == 0) {
            indptr[e] = (int32_t)s_row0;
            if (e == N - 1) indptr[N] = (int32_t)S.row_start(0, S.Nj);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// K3: one warp-group of b threads per element would under-fill; use one CTA (64 thr) / element
__global__ void __launch_bounds__(64)
k_assemble_rhs(TabView T, const double *__restrict__ vol, const double *__restrict__ face,
               const double *__restrict__ area, const double *__restrict__ minv,
               const double *__restrict__ f_vol, const double *__restrict__ g_face, Stencil S,
               double nu, double sigma, int use_minv, double *__restrict__ rhs) {
    __shared__ double F[64];
    const int b = T.b, nq = T.nq, nq1 = T.nq1;
    const int N = S.Ni * S.Nj;
    const int k = threadIdx.x;
    for (int e = blockIdx.x; e < N; e += gridDim.x) {
        const int i = e % S.Ni, j = e / S.Ni;
        if (!S.active(j)) {
            if (k < b) rhs[(size_t)e * b + k] = 0.0;
            continue;
        }
        int c[5];
        S.cols(i, j, c);
        if (k < b) {
            const double *ve = vol + (size_t)e * VOL_NC * nq;
            const double *fv = f_vol + (size_t)e * nq;
            double acc = 0.0;
            for (int q = 0; q < nq; ++q) acc = fma(T.V[q * b + k] * (ve[q] * T.w2[q]), fv[q], acc);   // element.py:162
            const double hF = sqrt(area[e]);
            for (int f = 0; f < 4; ++f) {
                if (c[1 + f] >= 0) continue;                 // Dirichlet faces only (discrete_system.py:378-396)
                const double *fo = face + ((size_t)e * 4 + f) * FACE_NC * nq1;
                const double *g = g_face + ((size_t)e * 4 + f) * nq1;
                const int tr = own_trace(f);
                const double st = (f & 1) ? 1.0 : -1.0;
                double pen = 0.0, sym = 0.0;
                for (int q = 0; q < nq1; ++q) {
                    const double wg = g[q] * T.w1[q] * fo[q];
                    const double vk = T.Vf[((size_t)tr * nq1 + q) * b + k];
                    const double dn = T.Vrf[((size_t)tr * nq1 + q) * b + k] * fo[nq1 + q] +
                                      T.Vsf[((size_t)tr * nq1 + q) * b + k] * fo[2 * nq1 + q];
                    pen = fma(vk, wg, pen);
                    sym = fma(dn, wg, sym);
                }
                acc += sigma * nu / hF * pen - st * nu * sym;   // face.py:183,194,230,246
            }
            F[k] = acc;
        }
        __syncthreads();
        if (k < b) {
            double v = F[k];
            if (use_minv) {
                v = 0.0;
                const double *Mi = minv + (size_t)e * b * b + (size_t)k * b;
                for (int m = 0; m < b; ++m) v = fma(Mi[m], F[m], v);
            }
            rhs[(size_t)e * b + k] = v;
        }
        __syncthreads();
    }
}

}  // namespace dgb

using namespace dgb;

extern "C" {

int dgb_tables_create(const dgb_tables_desc *d, dgb_tables **out) {
    DGB_ARG(d && out);
    DGB_ARG(d->Pg >= 1 && d->Pg <= 5 && d->p >= 0 && d->p <= 5 && d->nq1 >= 1 && d->nq1 <= 8 && d->cf >= 1);
    dgb_tables *t = new dgb_tables();
    t->Pg = d->Pg; t->p = d->p; t->nq1 = d->nq1; t->cf = d->cf;
    t->ng = (d->Pg + 1) * (d->Pg + 1);
    t->b = (d->p + 1) * (d->p + 1);
    t->nq = d->nq1 * d->nq1;
    const size_t nq = t->nq, nq1 = t->nq1, b = t->b, ng = t->ng;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n + 1) & ~(size_t)1; return o; };
    t->oV = take(nq * b); t->oVr = take(nq * b); t->oVs = take(nq * b);
    t->oW2 = take(nq); t->oW1 = take(nq1);
    t->oVf = take(4 * nq1 * b); t->oVrf = take(4 * nq1 * b); t->oVsf = take(4 * nq1 * b);
    t->oGX = take(nq * ng); t->oGR = take(nq * ng); t->oGS = take(nq * ng);
    t->oFX = take(4 * nq1 * ng); t->oFR = take(4 * nq1 * ng); t->oFS = take(4 * nq1 * ng);
    double *h = (double *)calloc(off, sizeof(double));
    auto put = [&](size_t o, const double *src, size_t n) { memcpy(h + o, src, n * sizeof(double)); };
    put(t->oV, d->h_V, nq * b); put(t->oVr, d->h_Vr, nq * b); put(t->oVs, d->h_Vs, nq * b);
    put(t->oW2, d->h_w2, nq); put(t->oW1, d->h_w1, nq1);
    put(t->oVf, d->h_Vf, 4 * nq1 * b); put(t->oVrf, d->h_Vrf, 4 * nq1 * b); put(t->oVsf, d->h_Vsf, 4 * nq1 * b);
    put(t->oGX, d->h_GX, nq * ng); put(t->oGR, d->h_GR, nq * ng); put(t->oGS, d->h_GS, nq * ng);
    put(t->oFX, d->h_FX, 4 * nq1 * ng); put(t->oFR, d->h_FR, 4 * nq1 * ng); put(t->oFS, d->h_FS, 4 * nq1 * ng);
    cudaError_t e1 = cudaMalloc(&t->d_buf, off * sizeof(double));
    cudaError_t e2 = cudaMalloc(&t->d_sub, (2 * nq + 8 * nq1) * sizeof(int32_t));
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        set_error("dgb_tables_create: cudaMalloc failed");
        free(h);
        delete t;
        return -1;
    }
    cudaMemcpy(t->d_buf, h, off * sizeof(double), cudaMemcpyHostToDevice);
    free(h);
    int32_t *hs = (int32_t *)calloc(2 * nq + 8 * nq1, sizeof(int32_t));
    if (d->h_sub_vol) memcpy(hs, d->h_sub_vol, 2 * nq * sizeof(int32_t));
    if (d->h_sub_face) memcpy(hs + 2 * nq, d->h_sub_face, 8 * nq1 * sizeof(int32_t));
    cudaMemcpy(t->d_sub, hs, (2 * nq + 8 * nq1) * sizeof(int32_t), cudaMemcpyHostToDevice);
    free(hs);
    DGB_CUDA_OK(cudaGetLastError());
    *out = t;
    return 0;
}

void dgb_tables_destroy(dgb_tables *t) {
    if (!t) return;
    cudaFree(t->d_buf);
    cudaFree(t->d_sub);
    delete t;
}

int dgb_metrics(const dgb_tables *t, const double *xn, const double *yn, int32_t il, int32_t Ni,
                int32_t Nj, double *vol, double *face, double *area, void *stream) {
    DGB_ARG(t && xn && yn && vol && face && area && Ni > 0 && Nj > 0);
    DGB_ARG(il >= Ni * t->cf * t->Pg + 1);
    cudaStream_t st = (cudaStream_t)stream;
    TabView T = view(t);
    const int64_t total = (int64_t)Ni * Nj * (T.nq + 4 * T.nq1);
    int64_t g = (total + 127) / 128;
    if (g > sm_count() * 32) g = sm_count() * 32;
    k_metrics<<<(int)g, 128, 0, st>>>(T, xn, yn, il, Ni, Nj, vol, face);
    DGB_LAUNCH_OK();
    int64_t g2 = ((int64_t)Ni * Nj + 255) / 256;
    if (g2 > sm_count() * 16) g2 = sm_count() * 16;
    k_area<<<(int)g2, 256, 0, st>>>(T, vol, (int64_t)Ni * Nj, area);
    DGB_LAUNCH_OK();
    return 0;
}

int64_t dgb_poisson_nnzb(int32_t Ni, int32_t Nj, int32_t flags) {
    Stencil S = make_stencil(Ni, Nj, flags);
    return S.row_start(0, Nj);
}

int dgb_assemble_poisson(const dgb_tables *t, const double *vol, const double *face,
                         const double *area, int32_t Ni, int32_t Nj, double nu, double sigma,
                         int32_t flags, int32_t *indptr, int32_t *indices, double *data,
                         double *minv, void *stream) {
    DGB_ARG(t && vol && face && area && indptr && indices && data && minv && Ni > 0 && Nj > 0);
    DGB_ARG(!((flags & DGB_FLAG_PERIODIC_I) && Ni < 2) && !((flags & DGB_FLAG_PERIODIC_J) && Nj < 2));
    DGB_ARG(dgb_poisson_nnzb(Ni, Nj, flags) < 2147483647LL);
    cudaStream_t st = (cudaStream_t)stream;
    TabView T = view(t);
    Stencil S = make_stencil(Ni, Nj, flags);
    const size_t bb = (size_t)T.b * T.b;
    const size_t smem = sizeof(double) * (6 * bb + 2 * (size_t)T.nq * T.b + T.nq + 4 * T.nq1 +
                                          2 * 4 * (size_t)T.nq1 * T.b);
    int64_t g = (int64_t)Ni * Nj;
    if (g > sm_count() * 16) g = sm_count() * 16;
    const int nt = T.b >= 16 ? 256 : 128;
    const int um = (flags & DGB_FLAG_MINV) ? 1 : 0;
#define DGB_ASM_LAUNCH(BT, QT)                                                                                  \
    do {                                                                                                        \
        DGB_CUDA_OK(cudaFuncSetAttribute(k_assemble_poisson<BT, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                           \
        k_assemble_poisson<BT, QT><<<(int)g, nt, smem, st>>>(T, vol, face, area, S, nu, sigma, um, indptr,     \
                                                             indices, data, minv);                              \
    } while (0)
    // (p+1)^2 blocks with the reference's N_int = 3p/2 + 1 points per direction (grid.py:107); anything else
    // (e.g. another integration-order factor) takes the runtime-sized instance
#define DGB_ASM_LAUNCH_MMA(BT, QT)                                                                              \
    do {                                                                                                        \
        constexpr int NQ4 = (QT * QT + 3) & ~3, TCF = (2 * QT + 3) & ~3, BS = MmaCfg<BT>::BS, B4 = MmaCfg<BT>::B4;  \
        constexpr int TC = NQ4 > 8 * QT ? (NQ4 > 4 * TCF ? NQ4 : 4 * TCF) : (8 * QT > 4 * TCF ? 8 * QT : 4 * TCF);  \
        const size_t sm2 = sizeof(double) * (2 * (TC * BS + 8) + 2 * (B4 * BS + 8) + B4);                       \
        DGB_CUDA_OK(cudaFuncSetAttribute(k_assemble_poisson_mma<BT, QT>,                                        \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));               \
        k_assemble_poisson_mma<BT, QT><<<(int)g, 256, sm2, st>>>(T, vol, face, area, S, nu, sigma, um, indptr, \
                                                                 indices, data, minv);                          \
    } while (0)
    // b >= 16 with the reference's quadrature: the contractions run on the FP64 tensor cores (DMMA); tuning value
    // dgb_set_kernel_path(100 + 51) keeps the per-entry DFMA kernel (for the comparison in profiles/)
    const bool mma = dgb::g_gs_variant != 51;
    if (mma && T.b == 16 && T.nq1 == 5) DGB_ASM_LAUNCH_MMA(16, 5);
    else if (mma && T.b == 25 && T.nq1 == 7) DGB_ASM_LAUNCH_MMA(25, 7);
    else if (mma && T.b == 36 && T.nq1 == 8) DGB_ASM_LAUNCH_MMA(36, 8);
    else if (T.b == 4 && T.nq1 == 2) DGB_ASM_LAUNCH(4, 2);
    else if (T.b == 9 && T.nq1 == 4) DGB_ASM_LAUNCH(9, 4);
    else if (T.b == 16 && T.nq1 == 5) DGB_ASM_LAUNCH(16, 5);
    else if (T.b == 25 && T.nq1 == 7) DGB_ASM_LAUNCH(25, 7);
    else if (T.b == 36 && T.nq1 == 8) DGB_ASM_LAUNCH(36, 8);
    else DGB_ASM_LAUNCH(0, 0);
#undef DGB_ASM_LAUNCH
#undef DGB_ASM_LAUNCH_MMA
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_assemble_rhs(const dgb_tables *t, const double *vol, const double *face,
                     const double *area, const double *minv, const double *f_vol,
                     const double *g_face, int32_t Ni, int32_t Nj, double nu, double sigma,
                     int32_t flags, double *rhs, void *stream) {
    DGB_ARG(t && vol && face && area && minv && f_vol && g_face && rhs && Ni > 0 && Nj > 0);
    DGB_ARG(t->b <= 64);
    cudaStream_t st = (cudaStream_t)stream;
    TabView T = view(t);
    Stencil S = make_stencil(Ni, Nj, flags);
    int64_t g = (int64_t)Ni * Nj;
    if (g > sm_count() * 32) g = sm_count() * 32;
    k_assemble_rhs<<<(int)g, 64, 0, st>>>(T, vol, face, area, minv, f_vol, g_face, S, nu, sigma,
                                         (flags & DGB_FLAG_MINV) ? 1 : 0, rhs);
    DGB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
