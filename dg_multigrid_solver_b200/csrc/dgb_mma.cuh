// dgb_mma.cuh -- small dense products C[b x b] = sum_t L[t][.]^T R[t][.] on the FP64 tensor cores (DMMA m8n8k4,
// `mma.sync.aligned.m8n8k4.row.col.f64`), operands in shared memory; shared by the assembly kernels and the
// builder of the smoother records.  Eight warps per CTA, tile (I, J) of C belongs to warp (I * NTL + J) % 8.
#pragma once

namespace dgb {

template <int BT>
struct MmaCfg {
    static constexpr int BS = BT == 36 ? 36 : BT == 25 ? 28 : 20;       // row stride of the tables: rows 32 / 96 bytes apart (mod 128)
    static constexpr int NTL = (BT + 7) / 8;                             // 8 x 8 tiles per direction
    static constexpr int NT2 = NTL * NTL;
    static constexpr int MAXT = (NT2 + 7) / 8;                           // tiles per warp (8 warps)
    static constexpr int B4 = (BT + 3) & ~3;                             // b rounded up to the DMMA k-step
};

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int BT>
__device__ __forceinline__ void mma_lr(double (&acc)[MmaCfg<BT>::MAXT][2], const double *L, const double *R, int tcount,
                                       int warp, int lane) {
    using C = MmaCfg<BT>;
    const int m = lane >> 2, kk = lane & 3;
    for (int t0 = 0; t0 < tcount; t0 += 4) {
        const double *l = L + (t0 + kk) * C::BS + m, *r = R + (t0 + kk) * C::BS + m;
#pragma unroll
        for (int s = 0; s < C::MAXT; ++s) {
            const int tile = warp + 8 * s;
            if (tile < C::NT2) dmma884(acc[s], l[8 * (tile / C::NTL)], r[8 * (tile % C::NTL)]);
        }
    }
}
// C tiles -> dst[row * stride + col] (rows, cols < b); lane holds C[m][2 kk], C[m][2 kk + 1]
template <int BT>
__device__ __forceinline__ void mma_store(const double (&acc)[MmaCfg<BT>::MAXT][2], double *dst, int stride, int warp,
                                          int lane) {
    using C = MmaCfg<BT>;
    const int m = lane >> 2, kk = lane & 3;
#pragma unroll
    for (int s = 0; s < C::MAXT; ++s) {
        const int tile = warp + 8 * s;
        if (tile >= C::NT2) continue;
        const int row = 8 * (tile / C::NTL) + m, col = 8 * (tile % C::NTL) + 2 * kk;
        if (row < BT && col < BT) dst[row * stride + col] = acc[s][0];
        if (row < BT && col + 1 < BT) dst[row * stride + col + 1] = acc[s][1];
    }
}

}  // namespace dgb
