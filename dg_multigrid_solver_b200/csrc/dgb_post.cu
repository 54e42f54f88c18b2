// dgb_post.cu -- post-processing of a solve: modal -> nodal evaluation and the L1 / L2 error sums in one pass.
//
// Reference: DGFEM.solve (dgfem/dgfem.py:188-232): per element u_nodal = V_DOF_grid @ u_e at the element's
// (Pg+1)^2 geometry nodes, the exact solution at those nodes, L1 = mean |u_nodal - u_exact|,
// L2 = sqrt(mean (u_nodal - u_exact)^2) over all element nodes (nodes shared by neighbouring elements count once
// per element, as in the reference's per-element loop).
// One thread per (element, node): the b-term dot product against the staged evaluation matrix, the gather of the
// exact value from the grid's node array (evaluated once per grid node by the MMS expression on the device), and
// the two running sums; per-CTA partials are reduced in a fixed order (bitwise reproducible norms).
#include "dgb_common.cuh"

namespace dgb {

__global__ void __launch_bounds__(256)
k_nodal_error(const double *__restrict__ Vg, int ng, int b, int Pg, int Ni, int il, long long N,
              const double *__restrict__ u, const double *__restrict__ exact_nodes, double *__restrict__ u_nodal,
              double *__restrict__ partials) {
    extern __shared__ double s_V[];         // [ng][b]
    __shared__ double s_red[32];
    for (int t = threadIdx.x; t < ng * b; t += 256) s_V[t] = Vg[t];
    __syncthreads();
    const int N1 = Pg + 1;
    const long long total = N * ng;
    double s1 = 0.0, s2 = 0.0;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
        const long long e = t / ng;
        const int a = (int)(t - e * ng);              // node a_i + N1 * a_j of the element (i fastest)
        const int aj = a / N1, ai = a - aj * N1;
        const long long j = e / Ni, i = e - j * Ni;
        const double *ue = u + e * b;
        const double *v = s_V + a * b;
        double acc = 0.0;
        for (int c = 0; c < b; ++c) acc = fma(v[c], ue[c], acc);
        const double ex = exact_nodes[(j * Pg + aj) * (long long)il + i * Pg + ai];
        if (u_nodal != nullptr) u_nodal[t] = acc;
        const double dlt = acc - ex;
        s1 += fabs(dlt);
        s2 = fma(dlt, dlt, s2);
    }
    const double t1 = block_sum<256>(s1, s_red);
    const double t2 = block_sum<256>(s2, s_red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = t1;
        partials[gridDim.x + blockIdx.x] = t2;
    }
}

__global__ void __launch_bounds__(1024)
k_sum_two(const double *__restrict__ partials, int n, double *out) {
    __shared__ double s_red[32];
    double a = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) {
        a += partials[i];
        c += partials[n + i];
    }
    const double ta = block_sum<1024>(a, s_red);
    const double tc = block_sum<1024>(c, s_red);
    if (threadIdx.x == 0) {
        out[0] = ta;
        out[1] = tc;
    }
}

}  // namespace dgb

using namespace dgb;

extern "C" int dgb_nodal_error(const double *V_grid, int32_t ng, int32_t b, int32_t Pg, int32_t Ni, int32_t Nj,
                               int32_t il, const double *u, const double *exact_nodes, double *u_nodal,
                               double *partials, double *sums, void *stream) {
    DGB_ARG(V_grid && u && exact_nodes && partials && sums);
    DGB_ARG(ng == (Pg + 1) * (Pg + 1) && b > 0 && Ni > 0 && Nj > 0 && il == Ni * Pg + 1);
    cudaStream_t st = (cudaStream_t)stream;
    const long long N = (long long)Ni * Nj;
    long long g = (N * ng + 255) / 256;
    const int cap = kMaxPartials / 2 < sm_count() * 8 ? kMaxPartials / 2 : sm_count() * 8;
    if (g > cap) g = cap;
    k_nodal_error<<<(int)g, 256, sizeof(double) * ng * b, st>>>(V_grid, ng, b, Pg, Ni, il, N, u, exact_nodes, u_nodal, partials);
    DGB_LAUNCH_OK();
    k_sum_two<<<1, 1024, 0, st>>>(partials, (int)g, sums);
    DGB_LAUNCH_OK();
    return 0;
}
