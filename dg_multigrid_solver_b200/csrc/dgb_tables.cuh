// dgb_tables.cuh -- device-resident basis / quadrature / geometry-operator tables of one level and the
// layout of the per-point metric arrays (shared by the Poisson and Stokes assembly kernels).
#pragma once
#include "dgb_common.cuh"

struct dgb_tables {
    int Pg, p, nq1, cf;
    int ng, b, nq;
    double *d_buf;     // all double tables, one allocation
    int32_t *d_sub;    // sub_vol[nq][2] then sub_face[4][nq1][2]
    // offsets (in doubles) into d_buf
    size_t oV, oVr, oVs, oW2, oW1, oVf, oVrf, oVsf, oGX, oGR, oGS, oFX, oFR, oFS;
};

namespace dgb {

struct TabView {
    int Pg, p, nq1, cf, ng, b, nq;
    const double *V, *Vr, *Vs, *w2, *w1, *Vf, *Vrf, *Vsf, *GX, *GR, *GS, *FX, *FR, *FS;
    const int32_t *sub_vol, *sub_face;
};

static inline TabView view(const dgb_tables *t) {
    TabView v;
    v.Pg = t->Pg; v.p = t->p; v.nq1 = t->nq1; v.cf = t->cf; v.ng = t->ng; v.b = t->b; v.nq = t->nq;
    const double *d = t->d_buf;
    v.V = d + t->oV; v.Vr = d + t->oVr; v.Vs = d + t->oVs; v.w2 = d + t->oW2; v.w1 = d + t->oW1;
    v.Vf = d + t->oVf; v.Vrf = d + t->oVrf; v.Vsf = d + t->oVsf;
    v.GX = d + t->oGX; v.GR = d + t->oGR; v.GS = d + t->oGS;
    v.FX = d + t->oFX; v.FR = d + t->oFR; v.FS = d + t->oFS;
    v.sub_vol = t->d_sub;
    v.sub_face = t->d_sub + 2 * t->nq;
    return v;
}

// face order everywhere: 0 imin, 1 imax, 2 jmin, 3 jmax
// trace-table order (Vf/Vrf/Vsf): 0 iL (r=+1), 1 iR (r=-1), 2 jL (s=+1), 3 jR (s=-1)
// the element's own trace on face f: imin -> iR, imax -> iL, jmin -> jR, jmax -> jL
__device__ __forceinline__ int own_trace(int f) { return f ^ 1; }
// the neighbour across face f shows its opposite face / the opposite trace
__device__ __forceinline__ int opp_face(int f) { return f ^ 1; }

constexpr int VOL_NC = 7;    // J, rx, sx, ry, sy, x, y
constexpr int FACE_NC = 8;   // Jf, alpha, beta, x, y, nx, ny, pad

}  // namespace dgb
