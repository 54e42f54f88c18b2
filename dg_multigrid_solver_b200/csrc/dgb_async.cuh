// dgb_async.cuh -- sm_100a async-copy / mbarrier primitives (inline PTX) and the closed-form
// structure of the DG 5-point block stencil.
#pragma once
#include <stdint.h>

namespace dgb {

// ---- bounded spinning: a wait that never completes must not hang the GPU ---------------------
constexpr int kSpinLimit = 1 << 22;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking probe: the result arrives ~150 cycles later, nothing waits for it until it is used
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// returns false on timeout (and records it in *err)
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int *err) {
    for (int spin = 0; spin < kSpinLimit; ++spin) {
        if (mbar_try_wait(bar, parity)) return true;
        if ((spin & 1023) == 1023 && *(volatile int *)err != 0) return false;
    }
    atomicExch(err, 1);
    return false;
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int ID, int NTHREADS>
__device__ __forceinline__ void named_bar_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NTHREADS) : "memory");
}

// ---- BSR structure of the DG 5-point block stencil, closed form ------------------------------
// (Poisson.assemble_BSR_Poisson, dgfem/discrete_system.py:83-144: slots {m, iL, iR, jL, jR},
//  periodic wrap in i for O-grids, in i and j for fully periodic grids, columns ascending)
struct Stencil {
    int Ni, Nj, per_i, per_j;
    int ja0, ja1;   // active element rows [ja0, ja1): rows outside are ghost rows of a slab (no matrix rows)
    __host__ __device__ bool active(int j) const { return j >= ja0 && j < ja1; }
    __host__ __device__ int bj(int j) const { return per_j ? 0 : ((j == 0) + (j == Nj - 1)); }
    __host__ __device__ int count(int i, int j) const {
        return active(j) ? 5 - bj(j) - (per_i ? 0 : ((i == 0) + (i == Ni - 1))) : 0;
    }
    __host__ __device__ long long row_start(int i, int j) const {
        if (j < ja0) return 0;
        const int jj = j < ja1 ? j : ja1;
        // blocks in the active element rows [ja0, jj), then in (i' < i, jj)
        const int nbrows = per_j ? 0 : ((ja0 == 0 && jj > 0) + (ja1 == Nj && jj >= Nj));
        long long s = (long long)(jj - ja0) * (5 * (long long)Ni - (per_i ? 0 : 2)) - (long long)Ni * nbrows;
        if (j < ja1) s += (long long)i * (5 - bj(j)) - ((per_i || i == 0) ? 0 : 1);
        return s;
    }
    // neighbour element index per slot {m, iL, iR, jL, jR} (-1 = Dirichlet boundary)
    __host__ __device__ void cols(int i, int j, int c[5]) const {
        const int m = j * Ni + i;
        c[0] = m;
        c[1] = i > 0 ? m - 1 : (per_i ? j * Ni + Ni - 1 : -1);
        c[2] = i < Ni - 1 ? m + 1 : (per_i ? j * Ni : -1);
        c[3] = j > 0 ? m - Ni : (per_j ? (Nj - 1) * Ni + i : -1);
        c[4] = j < Nj - 1 ? m + Ni : (per_j ? i : -1);
    }
};
// flags: DGB_FLAG_PERIODIC_I|J, DGB_FLAG_GHOST_LO|HI (include/dgb200.h)
__host__ __device__ inline Stencil make_stencil(int Ni, int Nj, int flags) {
    Stencil S;
    S.Ni = Ni; S.Nj = Nj;
    S.per_i = (flags & 1) ? 1 : 0;
    S.per_j = (flags & 2) ? 1 : 0;
    S.ja0 = (flags & 8) ? 1 : 0;
    S.ja1 = Nj - ((flags & 16) ? 1 : 0);
    return S;
}

// rank of each present slot in ascending column order; ties keep slot order (python's stable
// sorted(), dgfem/discrete_system.py:137-138)
__host__ __device__ __forceinline__ void slot_ranks(const int c[5], int rank[5]) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        int rk = 0;
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            if (c[t] < 0 || t == s) continue;
            if (c[t] < c[s] || (c[t] == c[s] && t < s)) ++rk;
        }
        rank[s] = c[s] < 0 ? -1 : rk;
    }
}

}  // namespace dgb
