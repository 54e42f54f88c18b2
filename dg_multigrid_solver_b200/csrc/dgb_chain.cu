// dgb_chain.cu -- lexicographic block Gauss-Seidel, split into a parallel part and a dependency chain.
//
// pyamg's block_gauss_seidel (dgfem/pyamg_relaxation.py:252-255) updates the block rows one after the
// other:  x_e <- Dinv_e (rhs_e - sum_{n != e} A_en x_n), every x_n being the newest value.  On the DG
// 5-point stencil two of the neighbours of e = (i, j) come EARLIER in the sweep (the previous element of
// the row and the element of the previous row), the others come LATER and still hold their old values:
//
//     x_e  =  c_e  -  M_row(e) x_(i-dir, j)  -  M_up(e) x_(i, j-dir),
//     c_e  =  Dinv_e (rhs_e - sum_{later n} A_en x_n^old),        M_* = Dinv_e A_e*   (pre-multiplied once)
//
//   k_gs_helper  computes every c_e -- no dependencies, a plain streaming kernel over 3 of the 5 blocks.
//                In a symmetric sweep (forward, backward, forward, ...) only the FIRST pass needs it: the chain
//                of one direction hands the next pass its c for free, because what it subtracted is exactly
//                what the opposite direction's c must leave out:
//                    c_next(e) = Dinv_e rhs_e - (c_e - x_e)        (d_e = Dinv_e rhs_e is kept in the record);
//   k_gs_chain   walks the dependency chain.  All it reads is one sequential record stream per element row
//                ({-M_row, -M_up, c}, brought in by TMA bulk copies), all it computes per element are two
//                b x b mat-vecs whose inputs are in registers: a warp owns R = 32/b consecutive rows,
//                lane (g, r) owns scalar row r of element row g, row g runs one element behind row g-1, so
//                both predecessor values are one warp shuffle away.  Bands of R rows hand their last row
//                over through a shared-memory ring (same CTA) or the level's global mailbox (next CTA).
//
// The sweep order, and therefore the result up to rounding of the re-associated products, is exactly the
// reference's.  O-grids (periodic in i): the LAST element of a row has a third earlier neighbour, the row's first
// element across the wrap; its pre-multiplied block lives in a small side array (one per row and direction) and
// is staged in shared memory and applied by the fill/drain path of the chain kernel.  Grids periodic in j keep the
// row-pipelined kernel.
#include <stdlib.h>

#include "dgb_async.cuh"
#include "dgb_common.cuh"
#include "dgb_mma.cuh"

namespace dgb {

int ensure_work(int n_rows);
int *work_ptr();
int *err_ptr();
extern int g_kernel_path;
extern int g_gs_variant;

// -DDGB_CHAIN_PRODUCER=0/1: force the refill scheme of every block size (default: ChainCfg<B>::PW)
#ifdef DGB_CHAIN_PRODUCER
#define DGB_CHAIN_PW(B) (DGB_CHAIN_PRODUCER)
#else
#define DGB_CHAIN_PW(B) ((B) <= 4 ? 1 : 0)
#endif

template <int B>
struct ChainCfg {
    static constexpr int B2 = B * B;
    // lanes per scalar row (each takes B/P columns).  b = 9: P = 3 / one element row per warp; P = 1 with three
    // element rows per warp was measured in round 2 (profiles/r02_probe_chain_variants.md): its step is 1.8x longer
    // (the shared-memory load rate of a lone warp, 8.4 cycles per LDS.128, bounds a step) and the pass 2.2 vs 1.76 ms
    static constexpr int P = B == 9 ? 3 : B == 16 ? 2 : 1;
    static constexpr int CW = B / P;                                  // matrix columns per lane
    static constexpr int LPR = B * P;                                 // lanes per element row
    static constexpr int R = LPR > 32 ? 1 : 32 / LPR;                 // element rows per warp (b=36: k_gs_chain_big)
    static constexpr int VN = P == 1 ? ((B + 1) & ~1) : CW;           // vector entries a lane loads
    // the two blocks of a record are stored lane-major, MV doubles at a time, so that the lanes of an element row
    // read consecutive shared-memory words (no bank conflicts): entry (r, c) of a block sits at mat_offset(r, c)
    static constexpr int MV = (CW % 2 == 0) ? 2 : 1;
    __host__ __device__ static constexpr int mat_offset(int r, int c) {
        return ((c % CW) / MV) * (LPR * MV) + (r * P + c / CW) * MV + (c % CW) % MV;
    }
    static constexpr int REC = (2 * B2 + 2 * B + 1) & ~1;             // doubles per record {M_row, M_up, (c, d) pairs}, 16-byte multiple
    // c_r and d_r of a record sit next to each other (one 16-byte load brings both)
    __host__ __device__ static constexpr int c_off(int r) { return 2 * B2 + 2 * r; }
    __host__ __device__ static constexpr int d_off(int r) { return 2 * B2 + 2 * r + 1; }
    static constexpr int CH = B <= 9 ? 8 : B <= 16 ? 4 : 1;           // steps per chunk (one bulk copy)
    static constexpr int NS = (B == 9 || B == 16 || B == 36) ? 2 : 3; // bulk-copy stages
    static constexpr int RING = B <= 9 ? 32 : 16;                     // columns per band hand-over ring
    static constexpr int RINGR = CH < 2 ? 2 : CH;                     // steps per row ring (rows of one warp)
    static constexpr int BP = (B + 1) & ~1;                           // doubles per ring slot
    static constexpr int PCH = 8;                                     // mailbox columns per poll (multiple of CH)
    static constexpr int PSL = (PCH * B + 31) / 32;                   // mailbox doubles per lane and poll
    static constexpr int WDEF = B == 9 ? 8 : B == 16 ? 5 : B <= 4 ? 3 : 4;   // warps (bands) per CTA (shared memory bound)
    // producer warp: one more warp per CTA refills the record stages of the W band warps (lane w serves warp w); a band
    // warp then only arrives on its stage's "empty" mbarrier -- the proxy fence, expect_tx and bulk copy leave its
    // critical path.  b = 4: pass -5.8 %; b = 9 (8 band warps per SM, two stages): +3 %, so it keeps refilling itself
    // (profiles/r02_probe_chain_variants.md)
    static constexpr int PW = DGB_CHAIN_PW(B);
    // doubles per row ring, padded so that the rows of a warp fall into different shared-memory banks
    static constexpr int RRS = RINGR * BP + (((RINGR * BP * 8) % 128) == 0 ? 4 : ((RINGR * BP * 8) % 128) == 64 ? 2 : 0);
    // O-grids: per warp, the wrap blocks of its R rows (staged once) and the rows' first new values
    static constexpr int WRAPD = LPR > 32 ? 0 : R * (B2 + BP);
    static constexpr int WR = RING * BP + R * RRS;                    // ring doubles per warp: incoming + rows
    static constexpr int SCR = RING * BP + 64;                        // scratch doubles per CTA: dummy store targets (benign races)
    __host__ __device__ static constexpr int stage_d(int W) { return W * NS * R * CH * REC; }
    __host__ __device__ static constexpr size_t o_ring(int W) { return sizeof(double) * stage_d(W); }
    // warp w: incoming ring at w * WR, then its R row rings; the CTA's outgoing ring is "warp W"'s incoming ring
    __host__ __device__ static constexpr size_t o_scr(int W) { return o_ring(W) + sizeof(double) * (W + 1) * WR; }
    __host__ __device__ static constexpr size_t o_bar(int W) { return o_scr(W) + sizeof(double) * SCR; }
    __host__ __device__ static constexpr size_t o_prog(int W) { return o_bar(W) + sizeof(uint64_t) * 2 * W * NS; }   // full + empty
    __host__ __device__ static constexpr size_t o_wrap(int W) { return (o_prog(W) + sizeof(int) * (W + 1) + 15) & ~(size_t)15; }
    __host__ __device__ static constexpr size_t smem(int W) { return o_wrap(W) + sizeof(double) * W * WRAPD; }
};

__device__ __forceinline__ bool chain_sentinel(double v) { return __double2hiint(v) == -1; }

// shared-memory accesses by 32-bit address (ptxas folds the constant offsets into the instruction)
__device__ __forceinline__ double lds1(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
// volatile flavours: ptxas keeps volatile accesses in program order (the delivery test must precede the data loads)
__device__ __forceinline__ double lds1v(uint32_t a) {
    double v;
    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double2 lds2v(uint32_t a) {
    double2 v;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts1(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// my part of one ring slot into registers (the whole slot when P == 1, else columns [part*CW, part*CW+CW): `a`
// already points at my part); *any_sentinel: some entry is the all-ones "not delivered" mark
template <int B>
__device__ __forceinline__ void chain_load_vec(uint32_t a, double (&v)[ChainCfg<B>::VN], bool *any_sentinel) {
    constexpr int VN = ChainCfg<B>::VN, P = ChainCfg<B>::P;
    if (VN % 2 == 0) {
#pragma unroll
        for (int c = 0; c < VN; c += 2) {
            const double2 t = lds2(a + c * 8);
            v[c] = t.x;
            v[c + 1 < VN ? c + 1 : c] = t.y;
        }
    } else {
#pragma unroll
        for (int c = 0; c < VN; ++c) v[c] = lds1(a + c * 8);
    }
    if (any_sentinel != nullptr) {
        unsigned mx = 0u;
#pragma unroll
        for (int c = 0; c < (P == 1 ? B : VN); ++c) mx = max(mx, (unsigned)__double2hiint(v[c]));
        *any_sentinel = mx == 0xffffffffu;
    }
}
// my part of one ring slot, volatile loads (ordered after the delivery test)
template <int B>
__device__ __forceinline__ void chain_load_vec_v(uint32_t a, double (&v)[ChainCfg<B>::VN]) {
    constexpr int VN = ChainCfg<B>::VN;
    if (VN % 2 == 0) {
#pragma unroll
        for (int c = 0; c < VN; c += 2) {
            const double2 t = lds2v(a + c * 8);
            v[c] = t.x;
            v[c + 1 < VN ? c + 1 : c] = t.y;
        }
    } else {
#pragma unroll
        for (int c = 0; c < VN; ++c) v[c] = lds1v(a + c * 8);
    }
}

// my part of my rows of the two (negated) pre-multiplied blocks and my entries of c and d, from the staged record
template <int B>
struct ChainRow {
    static constexpr int VN = ChainCfg<B>::VN, P = ChainCfg<B>::P, CW = ChainCfg<B>::CW;
    double ml[CW], mu[CW], c, d;
    // the same through volatile loads: ptxas keeps them behind the (volatile) loads of the ring slots
    __device__ __forceinline__ void load_v(uint32_t rm, uint32_t rc) {
        constexpr int B2 = B * B;
        const double2 cdv = lds2v(rc);
        c = cdv.x;
        d = cdv.y;
        constexpr int MV = ChainCfg<B>::MV, LS = ChainCfg<B>::LPR * MV;
        if (MV == 2) {
#pragma unroll
            for (int k = 0; k < CW; k += 2) {
                const double2 m0 = lds2v(rm + (k / 2) * LS * 8), m1 = lds2v(rm + (B2 + (k / 2) * LS) * 8);
                ml[k] = m0.x; ml[k + 1 < CW ? k + 1 : k] = m0.y;
                mu[k] = m1.x; mu[k + 1 < CW ? k + 1 : k] = m1.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < CW; ++k) {
                ml[k] = lds1v(rm + k * LS * 8);
                mu[k] = lds1v(rm + (B2 + k * LS) * 8);
            }
        }
    }
    __device__ __forceinline__ void load(uint32_t rm, uint32_t rc) {
        constexpr int B2 = B * B;
        const double2 cdv = lds2(rc);
        c = cdv.x;
        d = cdv.y;
        // lane-major layout: my k-th entry is LPR * MV doubles after my (k - MV)-th
        constexpr int MV = ChainCfg<B>::MV, LS = ChainCfg<B>::LPR * MV;
        if (MV == 2) {
#pragma unroll
            for (int k = 0; k < CW; k += 2) {
                const double2 m0 = lds2(rm + (k / 2) * LS * 8), m1 = lds2(rm + (B2 + (k / 2) * LS) * 8);
                ml[k] = m0.x; ml[k + 1 < CW ? k + 1 : k] = m0.y;
                mu[k] = m1.x; mu[k + 1 < CW ? k + 1 : k] = m1.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < CW; ++k) {
                ml[k] = lds1(rm + k * LS * 8);
                mu[k] = lds1(rm + (B2 + k * LS) * 8);
            }
        }
    }
    // x = c - M_row x_prev - M_up x_up   (records hold the negated products); with P > 1 the partial sums of the
    // P lanes of a scalar row are added by shuffles and the result is valid in the lane with part == 0
    // extra: this lane's share of one more (negated) product, added before the lanes of a row are summed
    __device__ __forceinline__ double eval(const double (&p)[VN], const double (&u)[VN]) const {
        if (P == 1) {
            // four independent accumulators, c folded into the first product (no c + 0.0)
            double a0 = fma(ml[0], p[0], c), a1 = CW > 1 ? ml[CW > 1 ? 1 : 0] * p[CW > 1 ? 1 : 0] : 0.0;
            double a2 = mu[0] * u[0], a3 = CW > 1 ? mu[CW > 1 ? 1 : 0] * u[CW > 1 ? 1 : 0] : 0.0;
#pragma unroll
            for (int k = 2; k < CW; ++k) {
                if (k & 1) {
                    a1 = fma(ml[k], p[k], a1);
                    a3 = fma(mu[k], u[k], a3);
                } else {
                    a0 = fma(ml[k], p[k], a0);
                    a2 = fma(mu[k], u[k], a2);
                }
            }
            return (a0 + a1) + (a2 + a3);
        }
        return eval(p, u, 0.0);
    }
    __device__ __forceinline__ double eval(const double (&p)[VN], const double (&u)[VN], double extra) const {
        if (P == 1) {
            double a0 = c + extra, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int k = 0; k < CW; ++k) {
                if (k & 1) {
                    a1 = fma(ml[k], p[k], a1);
                    a3 = fma(mu[k], u[k], a3);
                } else {
                    a0 = fma(ml[k], p[k], a0);
                    a2 = fma(mu[k], u[k], a2);
                }
            }
            return (a0 + a1) + (a2 + a3);
        }
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int k = 0; k < CW; ++k) {
            s1 = fma(ml[k], p[k], s1);
            s2 = fma(mu[k], u[k], s2);
        }
        const double mine = (s1 + s2) + extra;
        double sum = mine;
#pragma unroll
        for (int o = 1; o < P; ++o) sum += __shfl_down_sync(0xffffffffu, mine, o);
        return c + sum;
    }
};

// ---- thread-block cluster primitives (distributed shared memory) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t v;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(v));
    return v;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t v;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(v));
    return v;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t cluster_map(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(v) : "r"(local_smem_addr), "r"(rank));
    return v;
}
__device__ __forceinline__ void sts1_cluster(uint32_t a, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ int ldv_cluster_s32(uint32_t a) {
    int v;
    asm volatile("ld.volatile.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

__device__ __forceinline__ void stg1(double *p, double v) {
    asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// the next pass's c goes out together with d (same value as the one already there): a whole 16-byte (c, d) pair per
// store, so that the lanes of an element cover full 32-byte sectors -- c alone left every sector half written and
// the L2 fetched it from DRAM first (ncu: +0.3 GB read and +0.3 GB written per b = 9 pass at 2048^2)
__device__ __forceinline__ void stg2(double *p, double a, double b) {
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

// Records are stored in the order a band consumes them: rec[dir][band][t][g][REC], band = (row in sweep
// order) / R, g = row % R, t = (column in sweep order) + g, t < T = Ni + R - 1.  Unused (t, g) are zero.
template <int B>
__host__ __device__ __forceinline__ long long chain_loc(const Stencil &S_, int dir, int i, int j) {
    constexpr int R = ChainCfg<B>::R;
    const int sr = dir > 0 ? j - S_.ja0 : S_.ja1 - 1 - j;
    const int idx = dir > 0 ? i : S_.Ni - 1 - i;
    const int band = sr / R, g = sr - band * R;
    return ((long long)band * (S_.Ni + R - 1) + idx + g) * R + g;
}
template <int B>
__host__ __device__ __forceinline__ long long chain_dir_records(const Stencil &S_) {
    constexpr int R = ChainCfg<B>::R;
    const long long nbands = (S_.ja1 - S_.ja0 + R - 1) / R;
    return nbands * (S_.Ni + R - 1) * R;
}

// slow path of a step: spin until the row above has delivered this column (whole warp, uniform)
template <int B>
__device__ __noinline__ bool chain_wait_up(uint32_t uk, int *err) {
    constexpr int BP = ChainCfg<B>::BP;
    int spin = 0;
    bool bad;
    do {
        unsigned mx = 0u;
#pragma unroll
        for (int c = 0; c < BP; c += 2) {
            const double2 t = lds2(uk + c * 8);
            mx = max(mx, (unsigned)__double2hiint(t.x));
            if (c + 1 < B) mx = max(mx, (unsigned)__double2hiint(t.y));
        }
        bad = mx == 0xffffffffu;
        if (bad) __nanosleep(20);       // do not steal issue slots from the producing warp on the same sub-partition
        if (++spin > kSpinLimit || ((spin & 1023) == 1023 && *(volatile int *)err != 0)) {
            if ((threadIdx.x & 31) == 0) atomicExch(err, 2);
            return false;
        }
    } while (__any_sync(0xffffffffu, bad));
    return true;
}

#ifndef DGB_CHAIN_NOFENCE
#define DGB_CHAIN_NOFENCE 0
#endif


// -DDGB_CHAIN_TRACE: diagnostic build (tools/gpu/r02_chain_trace.sh) -- every band records when it started, got its
// first records, finished, and how long it sat in each kind of wait (16 x int64 per band, read by dgb_debug_chain_trace)
#ifdef DGB_CHAIN_TRACE
__device__ long long g_chain_trace[16 * 8192];
__device__ __forceinline__ long long trace_now() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define DGB_TRACE(x) x
#else
#define DGB_TRACE(x)
#endif
template <int B, int W, int DIR>
__global__ void __launch_bounds__((W + ChainCfg<B>::PW) * 32)
k_gs_chain(const double *__restrict__ rec, double *__restrict__ rec_other, const double *__restrict__ wrapm,
           double *__restrict__ x, double *mbox, Stencil S_, int *work, int *err, const int32_t *__restrict__ skip) {
    using C = ChainCfg<B>;
    constexpr int B2 = C::B2, R = C::R, REC = C::REC, CH = C::CH, NS = C::NS, RING = C::RING, BP = C::BP;
    constexpr int RINGR = C::RINGR, WR = C::WR, RRS = C::RRS, PCH = C::PCH, PSL = C::PSL;
    constexpr int P = C::P, CW = C::CW, LPR = C::LPR, VN = C::VN;
    constexpr uint32_t S = BP * 8;                 // bytes per ring slot
    constexpr uint32_t KS = R * REC * 8;           // bytes per step within a stage
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NOFENCE = DGB_CHAIN_NOFENCE;
    constexpr int PW = C::PW, NTH = (W + PW) * 32;
    if (skip != nullptr && *skip != 0) return;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_ticket;
    double *stages = reinterpret_cast<double *>(smem);
    double *rings = reinterpret_cast<double *>(smem + C::o_ring(W));
    double *scratch = reinterpret_cast<double *>(smem + C::o_scr(W));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::o_bar(W));
    volatile int *s_prog = reinterpret_cast<volatile int *>(smem + C::o_prog(W));
    const double sentinel = __longlong_as_double(-1LL);
    // one ticket per cluster: the CTAs of a cluster take consecutive groups of W bands and hand their last row
    // over through distributed shared memory; only the last CTA of a cluster uses the global mailbox
    const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
    if (crank == 0 && threadIdx.x == 0) s_ticket = atomicAdd(&work[0], 1);
    if (threadIdx.x <= W) s_prog[threadIdx.x] = 0;
    // incoming rings start empty (all sentinel), row rings start at zero (the value "before" column 0)
    for (int q = threadIdx.x; q < (W + 1) * WR; q += NTH) rings[q] = (q % WR) < RING * BP ? sentinel : 0.0;
    if ((threadIdx.x & 31) == 0 && threadIdx.x < W * 32) {
        uint64_t *bw = bars + (threadIdx.x >> 5) * NS;
        for (int s = 0; s < NS; ++s) {
            mbar_init(&bw[s], 1);
            mbar_init(&bw[W * NS + s], 1);         // "stage consumed" (producer warp build)
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    cluster_sync_all();            // rings initialised in every CTA of the cluster, rank 0 holds the ticket
    const int ticket = ldv_cluster_s32(cluster_map(smem_u32(&s_ticket), 0));
    cluster_sync_all();            // rank 0 may not leave before everybody has read its shared memory
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    asm volatile("" : "+r"(w));
    asm volatile("" : "+r"(lane));
    const int Ni = S_.Ni, nrows = S_.ja1 - S_.ja0;
    if (PW && w == W) {
        // ---- producer warp: lane cw refills the stages of band warp cw as they are handed back ----
        const int cband = (ticket * (int)csize + (int)crank) * W + lane;
        const int csr0 = cband * R;
        int cn = NS, cchunks = 0;                   // next chunk to bring in (the first NS are issued by the band warp)
        int cT = 0;
        if (lane < W && csr0 < nrows) {
            cT = Ni + min(R, nrows - csr0) - 1;
            cchunks = (cT + CH - 1) / CH;
        }
        const double *csrc = rec + (size_t)cband * (Ni + R - 1) * R * REC;
        double *cstage = stages + (size_t)lane * (NS * R * CH * REC);
        uint64_t *cfull = bars + (lane < W ? lane : 0) * NS, *cempty = cfull + W * NS;
        int spin = 0;
        while (__any_sync(FULL, cn < cchunks)) {
            const int cs = cn % NS;
            const bool ok = cn < cchunks && mbar_test_wait(&cempty[cs], (uint32_t)((cn / NS - 1) & 1));
            if (ok) {
                const uint32_t bytes = (uint32_t)min(CH, cT - cn * CH) * KS;
                mbar_expect_tx(&cfull[cs], bytes);
                bulk_g2s(cstage + (size_t)cs * (CH * R * REC), csrc + (size_t)cn * (CH * R * REC), bytes, &cfull[cs]);
                ++cn;
                spin = 0;
            }
            if (!__any_sync(FULL, ok)) {
                __nanosleep(40);
                if (++spin > kSpinLimit || ((spin & 1023) == 1023 && *(volatile int *)err != 0)) {
                    if (lane == 0) atomicExch(err, 1);
                    return;
                }
            }
        }
        return;
    }
    const int band = (ticket * (int)csize + (int)crank) * W + w;
    const int sr0 = band * R;                      // first row of the band, in sweep order
    if (sr0 >= nrows) return;
    DGB_TRACE(long long tr_wait_up = 0; long long tr_n_wait = 0; long long tr_mbar = 0; long long tr_flow = 0; long long tr_steps = 0; long long tr_epi = 0;)
    DGB_TRACE(if (lane == 0 && band < 8192) g_chain_trace[16 * band] = trace_now();)
    const int Rv = min(R, nrows - sr0);            // rows of this band
    uint64_t *full = bars + w * NS;
    // lane -> (element row g of the band, scalar row r, column part): lanes beyond the band's rows shadow row 0
    // (they load and compute, every store of theirs goes to a scratch area)
    const int g = lane / LPR, rem = lane - g * LPR;
    const int r = rem / P, part = rem - r * P;
    const bool live = g < Rv;
    const int gq = live ? g : 0;
    const bool fin = live && part == 0;            // this lane holds the finished x_r of its row
    const bool g0 = live && g == 0, lastg = fin && g == R - 1;
    const int j = DIR > 0 ? S_.ja0 + sr0 + gq : S_.ja1 - 1 - sr0 - gq;
    const int j0 = DIR > 0 ? S_.ja0 + sr0 : S_.ja1 - 1 - sr0;
    const int T = Ni + Rv - 1;                     // steps: row g handles sweep index t - g at step t
    const int nchunks = (T + CH - 1) / CH;
    // predecessor: 0 none, 1 shared-memory ring (previous warp, or previous CTA of the cluster), 2 global mailbox
    // successor:   0 none, 1 ring of the next warp, 3 ring of the next CTA of the cluster, 2 global mailbox
    int pred = sr0 == 0 ? 0 : ((w > 0 || crank > 0) ? 1 : 2);
    int succ = (band + 1) * R >= nrows ? 0 : (w < W - 1 ? 1 : (crank + 1 < csize ? 3 : 2));
    asm volatile("" : "+r"(pred), "+r"(succ));
    double *wstage = stages + (size_t)w * (NS * R * CH * REC);
    double *inring = rings + (size_t)w * WR;                                  // RING slots, indexed by column
    double *rowring = inring + RING * BP;                                     // R rings of RINGR slots, by step
    double *outr = rings + (size_t)(w + 1) * WR;                              // the next band's incoming ring

    // lane 0: one bulk copy per chunk (the records of CH steps of all R rows are contiguous)
    const double *bsrc = rec + (size_t)band * (Ni + R - 1) * R * REC;
    auto issue = [&](int n) {
        const int s = n % NS;
        const int steps = min(CH, T - n * CH);
        const uint32_t bytes = (uint32_t)steps * KS;
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(wstage + (size_t)s * (CH * R * REC), bsrc + (size_t)n * (CH * R * REC), bytes, &full[s]);
    };
    if (lane == 0)
        for (int n = 0; n < NS && n < nchunks; ++n) issue(n);

    // ---- mailbox polling (first band of a CTA): PCH columns of the predecessor row per poll ----
    const size_t mrow = (size_t)(j0 - DIR) * Ni;   // predecessor row of the band
    double PV[PSL];
    auto mb_load = [&](int m) {
#pragma unroll
        for (int sl = 0; sl < PSL; ++sl) {
            const int q = lane + 32 * sl;
            const int col = m * PCH + q / B;
            PV[sl] = 0.0;
            if (q < PCH * B && col < Ni) PV[sl] = __ldcg(mbox + (mrow + (DIR > 0 ? col : Ni - 1 - col)) * B + (q % B));
        }
    };
    auto mb_take = [&](int m) -> bool {            // wait for chunk m, move it into the incoming ring
        for (int spin = 0;; ++spin) {
            bool mine = true;
#pragma unroll
            for (int sl = 0; sl < PSL; ++sl) {
                const int q = lane + 32 * sl;
                if (q < PCH * B && m * PCH + q / B < Ni && chain_sentinel(PV[sl])) mine = false;
            }
            if (__all_sync(FULL, mine)) break;
            if (spin > kSpinLimit || ((spin & 63) == 63 && *(volatile int *)err != 0)) {
                if (lane == 0) atomicExch(err, 2);
                return false;
            }
            mb_load(m);
        }
#pragma unroll
        for (int sl = 0; sl < PSL; ++sl) {
            const int q = lane + 32 * sl;
            const int col = m * PCH + q / B;
            if (q < PCH * B && col < Ni) {
                inring[(col % RING) * BP + (q % B)] = PV[sl];
                __stcg(mbox + (mrow + (DIR > 0 ? col : Ni - 1 - col)) * B + (q % B), sentinel);
            }
        }
        if ((m + 1) * PCH < Ni) mb_load(m + 1);    // checked one poll later
        return true;
    };
    if (pred == 2) mb_load(0);

    // ---- per-lane shared-memory addresses (bytes), opaque so that they stay in registers ----
    // Row rings are indexed by the STEP that wrote the slot: row g > 0 at step t finds the value of the row
    // above (column t - g) and of its own previous column in the slots of step t - 1.  The incoming ring of
    // the band is indexed by column = the step at which row 0 reads it.
    const bool first_row = gq == 0;
    uint32_t in_b = smem_u32(inring);
    uint32_t up_b = smem_u32(rowring + (gq > 0 ? gq - 1 : 0) * RRS);   // ring of the row above (g > 0)
    uint32_t own_b = smem_u32(rowring + gq * RRS);
    const uint32_t scr = smem_u32(scratch);
    // last row -> next band's incoming ring (shared::cluster address: this CTA's, or warp 0 of the next CTA)
    uint32_t out_w = lastg ? (succ == 3 ? cluster_map(smem_u32(rings), crank + 1) : cluster_map(smem_u32(outr), crank)) + 8 * r
                           : cluster_map(scr, crank) + 8 * lane;
    const uint32_t po = (uint32_t)(part * CW * 8);                     // my columns inside a ring slot
    const uint32_t prog_next = succ == 3 ? cluster_map(smem_u32((const void *)&s_prog[0]), crank + 1)
                                         : cluster_map(smem_u32((const void *)&s_prog[w + 1]), crank);
    uint32_t rec_m = smem_u32(wstage) + (uint32_t)((gq * REC + C::mat_offset(r, part * CW)) * 8);   // my first matrix entry, stage 0 step 0
    uint32_t rec_c = smem_u32(wstage) + (uint32_t)((gq * REC + C::c_off(r)) * 8); // my (c, d) pair
    uint32_t own_w = fin ? own_b + 8 * r : scr + 8 * lane;
    asm volatile("" : "+r"(in_b), "+r"(up_b), "+r"(own_b), "+r"(out_w));
    asm volatile("" : "+r"(rec_m), "+r"(rec_c), "+r"(own_w));
    double *xrow = x + (size_t)j * Ni * B + r;
    // O-grid: the block that couples a row's last element to its first one (wrapm[row][B2], lane-major like the
    // record blocks) is staged in shared memory once per band; the first element's new value is parked beside it
    // one step after it is made and picked up again at the row's last element (fill/drain path only)
    const int per = S_.per_i;
    double *wrapw = reinterpret_cast<double *>(smem + C::o_wrap(W)) + (size_t)w * C::WRAPD;
    double *xfw = wrapw + R * B2;
    if (per) {
        for (int q = lane; q < Rv * B2; q += 32) {
            const int gg = q / B2;
            wrapw[q] = wrapm[(size_t)(DIR > 0 ? j0 + gg : j0 - gg) * B2 + (q - gg * B2)];
        }
        __syncwarp();
    }
    // my entry of c in the OTHER direction's record of sweep index idx: a constant stride of -R records per step
    // (the opposite sweep visits the rows and the columns in reverse order)
    double *corow;
    {
        const int sro = nrows - 1 - (sr0 + gq);                // my row in the opposite sweep order
        const int bo = sro / R, go = sro - bo * R;
        corow = rec_other + (((size_t)bo * (Ni + R - 1) + (Ni - 1) + go) * R + go) * REC + C::c_off(r);   // idx = 0
    }
    constexpr long long COS = -(long long)R * REC;          // doubles per step

    auto spin_fail = [=](int &spin) -> bool {
        if (++spin > kSpinLimit || ((spin & 1023) == 1023 && *(volatile int *)err != 0)) {
            if (lane == 0) atomicExch(err, 2);
            return true;
        }
        return false;
    };

    int prog_seen = 0;
    // this chunk's records are known to have landed: the mbarrier is probed (non-blocking) in the middle of the previous
    // chunk's steps, where the ~150 cycles of the probe overlap the dependent arithmetic (single band of b = 4: -5 %)
    bool have = false;
    for (int n = 0; n < nchunks; ++n) {
        const int s = n % NS;
        const int t0 = n * CH;
        DGB_TRACE(long long tr0 = clock64();)
        if (!have && !mbar_wait(&full[s], (uint32_t)((n / NS) & 1), err)) return;
        have = false;
        DGB_TRACE(if (n > 0) tr_mbar += clock64() - tr0; else if (lane == 0 && band < 8192) g_chain_trace[16 * band + 1] = trace_now();)
        DGB_TRACE(tr0 = clock64();)
        // ---- once per chunk: flow control and the global hand-overs ----
        if (pred == 1 && lane == 0) s_prog[w] = t0;                           // columns < t0 are consumed
        if (succ == 1 || succ == 3) {
            // the consumer's progress as sampled one chunk ago is usually enough (the sample of a remote CTA takes
            // about a microsecond to arrive: it must not sit on this warp's critical path every chunk)
            if (prog_seen < t0 + CH - RING) {
                int spin = 0;
                while ((prog_seen = ldv_cluster_s32(prog_next)) < t0 + CH - RING)
                    if (spin_fail(spin)) return;
            }
            prog_seen = ldv_cluster_s32(prog_next);      // consumed at the next chunk
        }
        if (pred != 1 && (t0 % PCH) == 0 && t0 < Ni) {
            if (pred == 2) {
                if (!mb_take(t0 / PCH)) return;
            } else {
                for (int q = lane; q < PCH * BP; q += 32) inring[((t0 + q / BP) % RING) * BP + (q % BP)] = 0.0;
            }
            __syncwarp();
        }
        DGB_TRACE(tr_flow += clock64() - tr0; tr0 = clock64();)
        const uint32_t si0 = (uint32_t)(t0 % RING) * S;                       // incoming-ring slot of column t0
        const uint32_t sr0b = (uint32_t)(t0 % RINGR) * S;                     // row-ring slot of step t0 (0 if RINGR == CH)
        uint32_t sm = rec_m + (uint32_t)s * (CH * KS), sc = rec_c + (uint32_t)s * (CH * KS);
        if (t0 >= R - 1 + 2 * per && t0 + CH <= Ni - per) {
            // ---- every row of the band is inside the grid for all CH steps (O-grid: and none of them is at the
            // row's second or last element): no predicates, immediates ----
            // vector of the row above: row 0 reads column t0 + k of the incoming ring, row g > 0 the slot of step t0 + k - 1
            uint32_t ua0, ua, pa0, pa, ow;
            if (RINGR == CH) {
                ua0 = first_row ? in_b + si0 : up_b + (RINGR - 1) * S;
                ua = first_row ? in_b + si0 : up_b - S;
                pa0 = own_b + (RINGR - 1) * S + po;
                pa = own_b - S + po;
                ow = own_w;
            } else {        // CH == 1: two-slot row rings
                ua0 = first_row ? in_b + si0 : up_b + (S - sr0b);
                ua = ua0;
                pa0 = own_b + (S - sr0b) + po;
                pa = pa0;
                ow = own_w + sr0b;
            }
            uint32_t oa = out_w + (uint32_t)((t0 - (R - 1)) % RING) * S;      // column-indexed (consumer's view)
            const uint32_t oend = out_w + RING * S;
            double *xp = xrow + (size_t)(DIR > 0 ? t0 - gq : Ni - 1 - (t0 - gq)) * B;
            double *cop = corow + (long long)(t0 - gq) * COS;
            asm volatile("" : "+r"(ua0), "+r"(ua), "+r"(pa0), "+r"(pa));
            asm volatile("" : "+r"(ow), "+r"(sm), "+r"(sc));
            asm volatile("" : "+l"(xp), "+l"(cop));
            // software pipeline: the record of step k + 1 is read (plain shared-memory loads, the stage is complete)
            // while the dependent arithmetic of step k waits for its operands
            ChainRow<B> rows[2];
            rows[0].load(sm, sc);
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const uint32_t uk = k == 0 ? ua0 : ua + k * S;     // slot of the row above
                double u[VN], p[VN];
                chain_load_vec_v<B>(uk + po, u);
                chain_load_vec_v<B>(k == 0 ? pa0 : pa + k * S, p);
                if (k + 1 < CH) rows[(k + 1) & 1].load_v(sm + (k + 1) * KS, sc + (k + 1) * KS);
                if (CH > 1 && k == CH / 2 && n + 1 < nchunks)
                    have = mbar_test_wait(&full[(n + 1) % NS], (uint32_t)(((n + 1) / NS) & 1));
                const ChainRow<B> &row = rows[k & 1];
                double xnew = row.eval(p, u);
                // the neighbour band has not delivered this column yet?  (rows above inside the warp always have)
                // delivery test on the loaded registers themselves (no ordering question): some entry of my part of the
                // slot still carries the all-ones mark?  (P == 1: every lane holds the whole slot)
                unsigned mxh = 0u;
#pragma unroll
                for (int c = 0; c < (P == 1 ? B : VN); ++c) mxh = max(mxh, (unsigned)__double2hiint(u[c]));
                const bool wait_up = __any_sync(FULL, mxh == 0xffffffffu);
                if (__builtin_expect(wait_up, 0)) {
                    DGB_TRACE(const long long tw = clock64();)
                    if (!chain_wait_up<B>(uk, err)) return;
                    DGB_TRACE(tr_wait_up += clock64() - tw; tr_n_wait += 1;)
                    chain_load_vec<B>(uk + po, u, nullptr);
                    xnew = row.eval(p, u);
                }
                sts1(ow + k * S, xnew);
                sts1_cluster(oa, xnew);
                oa += S;
                if (oa == oend) oa = out_w;
                if (fin) {
                    stg1(xp + k * DIR * B, xnew);
                    stg2(cop + k * COS, (row.d - row.c) + xnew, row.d);       // the next (opposite) pass's (c, d)
                }
                __syncwarp();
            }
            // hand the CH consumed slots of the incoming ring back in one go (they are contiguous: RING % CH == 0)
            for (int q = lane; q < CH * BP; q += 32) sts1(in_b + si0 + 8 * q, sentinel);
        } else {
            // ---- pipeline fill / drain: some rows are outside [0, Ni) ----
#pragma unroll 1
            for (int k = 0; k < CH && t0 + k < T; ++k) {
                const int t = t0 + k;
                const int idx = t - gq;
                const bool act = fin && idx >= 0 && idx < Ni;
                const bool poll = g0 && t < Ni;
                const uint32_t so = (uint32_t)(t % RINGR) * S, sop = (uint32_t)((t + RINGR - 1) % RINGR) * S;
                const uint32_t uk = first_row ? in_b + (uint32_t)(t % RING) * S : up_b + sop;
                double u[VN], p[VN];
                bool bad;
                chain_load_vec<B>(uk + po, u, &bad);
                chain_load_vec<B>(own_b + sop + po, p, nullptr);
                int spin = 0;
                while (__any_sync(FULL, poll && bad)) {
                    chain_load_vec<B>(uk + po, u, &bad);
                    if (spin_fail(spin)) return;
                }
                ChainRow<B> row;
                row.load(sm + k * KS, sc + k * KS);
                double extra = 0.0;
                if (per) {
                    if (idx == 1 && r == 0) {           // p holds the new value of the row's first element
#pragma unroll
                        for (int q = 0; q < CW; ++q) xfw[gq * BP + part * CW + q] = p[q];
                    }
                    if (idx == Ni - 1) {                // the wrap neighbour is an earlier element of the sweep
#pragma unroll
                        for (int q = 0; q < CW; ++q)
                            extra = fma(wrapw[gq * B2 + C::mat_offset(r, part * CW + q)], xfw[gq * BP + part * CW + q], extra);
                    }
                }
                const double xnew = row.eval(p, u, extra);
                if (poll && fin) sts1(uk + 8 * r, sentinel);
                if (act) {
                    sts1(own_b + so + 8 * r, xnew);
                    if (lastg) sts1_cluster(out_w + (uint32_t)(idx % RING) * S, xnew);
                    xrow[(size_t)(DIR > 0 ? idx : Ni - 1 - idx) * B] = xnew;
                    stg2(corow + (long long)idx * COS, (row.d - row.c) + xnew, row.d);
                }
                __syncwarp();
            }
        }
        DGB_TRACE(tr_steps += clock64() - tr0; tr0 = clock64();)
        // ---- the CTA's last row: this chunk's columns go to the global mailbox ----
        if (succ == 2) {
            const size_t lrow = (size_t)(j0 + DIR * (R - 1)) * Ni;
            const int tend = min(t0 + CH, T);
            for (int q = lane; q < CH * B; q += 32) {
                const int col = t0 - (R - 1) + q / B;
                if (col >= 0 && col < Ni && col <= tend - 1 - (R - 1))
                    __stcg(mbox + (lrow + (DIR > 0 ? col : Ni - 1 - col)) * B + (q % B), outr[(col % RING) * BP + (q % B)]);
            }
        }
        __syncwarp();
        if (lane == 0 && n + NS < nchunks) {
            if (PW) {
                mbar_arrive(&full[W * NS + s]);        // the stage is free: the producer warp refills it
            } else {
                if (NOFENCE == 0) fence_proxy_async();
                issue(n + NS);
            }
        }
        DGB_TRACE(__syncwarp(); tr_epi += clock64() - tr0;)
        // end of chunks 0, 1, 7, 63: how the distance to the band above develops along the row
        DGB_TRACE(if (lane == 0 && band < 8192 && (n == 0 || n == 1 || n == 7 || n == 63))
                      g_chain_trace[16 * band + 11 + (n == 0 ? 0 : n == 1 ? 1 : n == 7 ? 2 : 3)] = trace_now();)
    }
#ifdef DGB_CHAIN_TRACE
    if (lane == 0 && band < 8192) {
        long long *tr = g_chain_trace + 16 * band;
        tr[8] = tr_steps;
        tr[9] = tr_epi;
        tr[10] = nchunks;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tr[2] = trace_now();
        tr[3] = tr_n_wait;
        tr[4] = tr_wait_up;
        tr[5] = tr_mbar;
        tr[6] = tr_flow;
        tr[7] = (long long)smid | ((long long)pred << 16) | ((long long)succ << 20) | ((long long)w << 24) | ((long long)crank << 32);
    }
#endif
}


// =========================================================================================
// k_gs_chain_big: the same chain for blocks with more scalar rows than a warp has lanes (b = 36, p = 5).
// One element row per warp, lane q < b/2 owns the scalar rows 2q and 2q+1; the two b x b products are streamed
// through the lanes 16 bytes at a time (nothing but eight accumulators stays in registers); one record, i.e. one
// step, per bulk copy.  Rings, mailbox, clusters, tickets and the record stream are those of k_gs_chain.
// =========================================================================================
__device__ __forceinline__ void sts2(uint32_t a, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void sts2_cluster(uint32_t a, double2 v) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}

template <int B>
struct BigCfg {
    using C = ChainCfg<B>;
    static constexpr int LPL = B / 2;                                  // working lanes
    __host__ __device__ static constexpr size_t o_xf(int W) { return C::smem(W); }
    __host__ __device__ static constexpr size_t smem(int W) { return ((o_xf(W) + 15) & ~(size_t)15) + sizeof(double) * W * B; }
};

// rows 2q, 2q+1 of  c + M_row p + M_up u  for one staged record; *bad: some entry of u is the "not delivered" mark
template <int B>
__device__ __forceinline__ double2 big_eval(uint32_t rec_a /* record */, uint32_t pa, uint32_t ua, int q, bool *bad) {
    constexpr int B2 = B * B;
    const uint32_t m = rec_a + (uint32_t)(q * 32);                     // my two rows of a 16-byte column pair
    const double c0 = lds1(rec_a + (uint32_t)(ChainCfg<B>::c_off(2 * q) * 8)), c1 = lds1(rec_a + (uint32_t)(ChainCfg<B>::c_off(2 * q + 1) * 8));
    double a0 = c0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = c1, b1 = 0.0, b2 = 0.0, b3 = 0.0;
    unsigned mx = 0u;
#pragma unroll
    for (int k = 0; k < B / 2; ++k) {
        const double2 pp = lds2(pa + k * 16), uu = lds2(ua + k * 16);
        mx = max(mx, max((unsigned)__double2hiint(uu.x), (unsigned)__double2hiint(uu.y)));
        const double2 ra = lds2(m + k * (2 * B * 8)), rb = lds2(m + k * (2 * B * 8) + 16);
        const double2 sa = lds2(m + (B2 + k * 2 * B) * 8), sb = lds2(m + (B2 + k * 2 * B) * 8 + 16);
        a0 = fma(ra.x, pp.x, a0); a1 = fma(ra.y, pp.y, a1);
        b0 = fma(rb.x, pp.x, b0); b1 = fma(rb.y, pp.y, b1);
        a2 = fma(sa.x, uu.x, a2); a3 = fma(sa.y, uu.y, a3);
        b2 = fma(sb.x, uu.x, b2); b3 = fma(sb.y, uu.y, b3);
    }
    *bad = mx == 0xffffffffu;
    return make_double2((a0 + a1) + (a2 + a3), (b0 + b1) + (b2 + b3));
}

template <int B, int W, int DIR>
__global__ void __launch_bounds__(W * 32)
k_gs_chain_big(const double *__restrict__ rec, double *__restrict__ rec_other, const double *__restrict__ wrapm,
               double *__restrict__ x, double *mbox, Stencil S_, int *work, int *err,
               const int32_t *__restrict__ skip) {
    using C = ChainCfg<B>;
    constexpr int B2 = C::B2, REC = C::REC, NS = C::NS, RING = C::RING, BP = C::BP, WR = C::WR;
    constexpr int PCH = C::PCH, PSL = C::PSL, LPL = BigCfg<B>::LPL;
    constexpr uint32_t S = BP * 8;
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(C::R == 1 && C::CH == 1 && C::RINGR == 2 && B % 2 == 0, "one row and one step per chunk");
    if (skip != nullptr && *skip != 0) return;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_ticket;
    double *stages = reinterpret_cast<double *>(smem);
    double *rings = reinterpret_cast<double *>(smem + C::o_ring(W));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::o_bar(W));
    volatile int *s_prog = reinterpret_cast<volatile int *>(smem + C::o_prog(W));
    double *xfirst = reinterpret_cast<double *>(smem + ((BigCfg<B>::o_xf(W) + 15) & ~(size_t)15));
    const double sentinel = __longlong_as_double(-1LL);
    const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
    if (crank == 0 && threadIdx.x == 0) s_ticket = atomicAdd(&work[0], 1);
    if (threadIdx.x <= W) s_prog[threadIdx.x] = 0;
    for (int q = threadIdx.x; q < (W + 1) * WR; q += W * 32) rings[q] = (q % WR) < RING * BP ? sentinel : 0.0;
    if ((threadIdx.x & 31) == 0) {
        uint64_t *bw = bars + (threadIdx.x >> 5) * NS;
        for (int s = 0; s < NS; ++s) mbar_init(&bw[s], 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    cluster_sync_all();
    const int ticket = ldv_cluster_s32(cluster_map(smem_u32(&s_ticket), 0));
    cluster_sync_all();
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    asm volatile("" : "+r"(w));
    asm volatile("" : "+r"(lane));
    const int Ni = S_.Ni, nrows = S_.ja1 - S_.ja0;
    const int sr = (ticket * (int)csize + (int)crank) * W + w;       // my row, in sweep order
    if (sr >= nrows) return;
    uint64_t *full = bars + w * NS;
    const int q = lane % LPL;                        // lanes >= LPL shadow the first ones (they never store)
    const bool live = lane < LPL;
    const int j = DIR > 0 ? S_.ja0 + sr : S_.ja1 - 1 - sr;
    int pred = sr == 0 ? 0 : ((w > 0 || crank > 0) ? 1 : 2);
    int succ = sr + 1 >= nrows ? 0 : (w < W - 1 ? 1 : (crank + 1 < csize ? 3 : 2));
    asm volatile("" : "+r"(pred), "+r"(succ));
    double *wstage = stages + (size_t)w * (NS * REC);
    double *inring = rings + (size_t)w * WR;
    double *rowring = inring + RING * BP;                              // two slots: previous / current column
    double *outr = rings + (size_t)(w + 1) * WR;
    double *xf = xfirst + (size_t)w * B;
    const double *bsrc = rec + (size_t)sr * Ni * REC;                  // R = 1: band = row, T = Ni
    auto issue = [&](int n) {
        const int s = n % NS;
        mbar_expect_tx(&full[s], (uint32_t)(REC * 8));
        bulk_g2s(wstage + (size_t)s * REC, bsrc + (size_t)n * REC, (uint32_t)(REC * 8), &full[s]);
    };
    if (lane == 0)
        for (int n = 0; n < NS && n < Ni; ++n) issue(n);

    const size_t mrow = (size_t)(j - DIR) * Ni;
    double PV[PSL];
    auto mb_load = [&](int m) {
#pragma unroll
        for (int sl = 0; sl < PSL; ++sl) {
            const int e = lane + 32 * sl;
            const int col = m * PCH + e / B;
            PV[sl] = 0.0;
            if (e < PCH * B && col < Ni) PV[sl] = __ldcg(mbox + (mrow + (DIR > 0 ? col : Ni - 1 - col)) * B + (e % B));
        }
    };
    auto mb_take = [&](int m) -> bool {
        for (int spin = 0;; ++spin) {
            bool mine = true;
#pragma unroll
            for (int sl = 0; sl < PSL; ++sl) {
                const int e = lane + 32 * sl;
                if (e < PCH * B && m * PCH + e / B < Ni && chain_sentinel(PV[sl])) mine = false;
            }
            if (__all_sync(FULL, mine)) break;
            if (spin > kSpinLimit || ((spin & 63) == 63 && *(volatile int *)err != 0)) {
                if (lane == 0) atomicExch(err, 2);
                return false;
            }
            mb_load(m);
        }
#pragma unroll
        for (int sl = 0; sl < PSL; ++sl) {
            const int e = lane + 32 * sl;
            const int col = m * PCH + e / B;
            if (e < PCH * B && col < Ni) {
                inring[(col % RING) * BP + (e % B)] = PV[sl];
                __stcg(mbox + (mrow + (DIR > 0 ? col : Ni - 1 - col)) * B + (e % B), sentinel);
            }
        }
        if ((m + 1) * PCH < Ni) mb_load(m + 1);
        return true;
    };
    if (pred == 2) mb_load(0);

    const uint32_t in_b = smem_u32(inring), row_b = smem_u32(rowring);
    const uint32_t out_b = (succ == 3 ? cluster_map(smem_u32(rings), crank + 1) : cluster_map(smem_u32(outr), crank));
    const uint32_t prog_next = succ == 3 ? cluster_map(smem_u32((const void *)&s_prog[0]), crank + 1)
                                         : cluster_map(smem_u32((const void *)&s_prog[w + 1]), crank);
    const uint32_t stage0 = smem_u32(wstage);
    double *xrow = x + (size_t)j * Ni * B + 2 * q;
    const int sro = nrows - 1 - sr;                                    // my row in the opposite sweep order
    double *corow = rec_other + ((size_t)sro * Ni + (Ni - 1)) * REC + C::c_off(2 * q);     // idx = 0, stride -REC
    const double *wrow = wrapm + (size_t)j * B2;                       // O-grid: wrap block of my row
    const int per = S_.per_i;
    int prog_seen = 0;
    for (int t = 0; t < Ni; ++t) {
        const int s = t % NS;
        if (!mbar_wait(&full[s], (uint32_t)((t / NS) & 1), err)) return;
        if (pred == 1 && lane == 0) s_prog[w] = t;
        if (succ == 1 || succ == 3) {
            if (prog_seen < t + 1 - RING) {
                int spin = 0;
                while ((prog_seen = ldv_cluster_s32(prog_next)) < t + 1 - RING) {
                    if (++spin > kSpinLimit || ((spin & 1023) == 1023 && *(volatile int *)err != 0)) {
                        if (lane == 0) atomicExch(err, 2);
                        return;
                    }
                }
            }
            if ((t & 3) == 0) prog_seen = ldv_cluster_s32(prog_next);
        }
        if (pred != 1 && (t % PCH) == 0) {
            if (pred == 2) {
                if (!mb_take(t / PCH)) return;
            } else {
                for (int e = lane; e < PCH * BP; e += 32) inring[((t + e / BP) % RING) * BP + (e % BP)] = 0.0;
            }
            __syncwarp();
        }
        const uint32_t ua = in_b + (uint32_t)(t % RING) * S;           // the row above, this column
        const uint32_t pa = row_b + (uint32_t)((t + 1) & 1) * S;       // my row, previous column
        const uint32_t ra = stage0 + (uint32_t)s * (REC * 8);
        bool bad;
        double2 xn = big_eval<B>(ra, pa, ua, q, &bad);
        if (__any_sync(FULL, bad)) {
            if (!chain_wait_up<B>(ua, err)) return;
            xn = big_eval<B>(ra, pa, ua, q, &bad);
        }
        if (per) {
            if (t == 1 && live) {                                      // the first element's new value
                const double2 f = lds2(pa + q * 16);
                xf[2 * q] = f.x;
                xf[2 * q + 1] = f.y;
            }
            if (t == Ni - 1) {                                         // wrap neighbour: an earlier element
                __syncwarp();
                double ea = 0.0, eb = 0.0;
                for (int k = 0; k < B / 2; ++k) {
                    const double2 f = *reinterpret_cast<const double2 *>(xf + 2 * k);
                    const double2 wa = *reinterpret_cast<const double2 *>(wrow + C::mat_offset(2 * q, 2 * k));
                    const double2 wb = *reinterpret_cast<const double2 *>(wrow + C::mat_offset(2 * q + 1, 2 * k));
                    ea = fma(wa.x, f.x, ea); ea = fma(wa.y, f.y, ea);
                    eb = fma(wb.x, f.x, eb); eb = fma(wb.y, f.y, eb);
                }
                xn.x += ea;
                xn.y += eb;
            }
        }
        if (live) {
            const double2 v0 = lds2(ra + (uint32_t)(C::c_off(2 * q) * 8)), v1 = lds2(ra + (uint32_t)(C::c_off(2 * q + 1) * 8));
            const double2 cd = make_double2(v0.x, v1.x), dd = make_double2(v0.y, v1.y);
            sts2(ua + q * 16, make_double2(sentinel, sentinel));                  // hand the incoming slot back
            sts2(row_b + (uint32_t)(t & 1) * S + q * 16, xn);
            sts2_cluster(out_b + (uint32_t)(t % RING) * S + q * 16, xn);
            const int i = DIR > 0 ? t : Ni - 1 - t;
            *reinterpret_cast<double2 *>(xrow + (size_t)i * B) = xn;
            stg2(corow - (long long)t * REC, (dd.x - cd.x) + xn.x, dd.x);    // (c_next, d) of row 2q; row 2q+1 two doubles on
            stg2(corow - (long long)t * REC + 2, (dd.y - cd.y) + xn.y, dd.y);
            if (succ == 2) {
                __stcg(mbox + ((size_t)j * Ni + i) * B + 2 * q, xn.x);
                __stcg(mbox + ((size_t)j * Ni + i) * B + 2 * q + 1, xn.y);
            }
        }
        __syncwarp();
        if (lane == 0 && t + NS < Ni) {
            fence_proxy_async();
            issue(t + NS);
        }
    }
}

// ---- the parallel part: c_e = Dinv_e (rhs_e - sum over the neighbours the chain does not handle) ----
template <int B>
struct HelperCfg {
    static constexpr int EPB = (256 / B) > 0 ? (256 / B) : 1;
    static constexpr int NT = ((EPB * B + 31) / 32) * 32;
};

// RES: also evaluate the full residual r = rhs - A x (optionally stored) and its sum of squares (one partial per
// CTA) -- the smoother's entry residual test (dgfem/relaxation.py:202) shares the block reads of the first pass.
template <int B, bool RES>
__global__ void __launch_bounds__(HelperCfg<B>::NT, RES ? ((B > 4 && B <= 16) ? 5 : 6) : 8)
k_gs_helper(const double *__restrict__ data, const int32_t *__restrict__ indices, const int32_t *__restrict__ indptr,
            const double *__restrict__ dinv, const double *__restrict__ rhs, const double *__restrict__ x,
            double *rec, Stencil S_, int dir, const int32_t *__restrict__ skip, double *r_out,
            double *partials, int x_zero /* x == 0 (coarse-level initial guess, solver.py:171): no block is read */) {
    constexpr int EPB = HelperCfg<B>::EPB, REC = ChainCfg<B>::REC;
    if (skip != nullptr && *skip != 0) return;
    __shared__ double s_rsum[EPB * B];
    __shared__ double s_rhs[EPB * B];
    __shared__ double s_red[32];
    double sumsq = 0.0;
    const int el = threadIdx.x / B, r = threadIdx.x - el * B;
    const int Ni = S_.Ni;
    const int first = S_.ja0 * Ni, count = (S_.ja1 - S_.ja0) * Ni;
    const int ntiles = (count + EPB - 1) / EPB;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int idx = tile * EPB + el;
        const int e = (el < EPB && idx < count) ? first + idx : -1;
        // my row of Dinv_e: requested now (b <= 16: it fits in registers), used after the barrier
        // (RES variant, 4 < b <= 16; measured at 2048^2: b = 9 entry residual 3.72 -> 3.20 ms; b = 4 and the plain
        // helper at 8 CTAs per SM are faster without it)
        constexpr bool PRE = RES && B > 4 && B <= 16;
        double dv[PRE ? B : 1];
        if (e >= 0) {
            if (PRE) load_row<PRE ? B : 1>(dinv + ((size_t)e * B + r) * B, dv);
            const int j = e / Ni, i = e - j * Ni;
            const int e_row = (i - dir >= 0 && i - dir < Ni) ? e - dir : -1;       // handled by the chain
            const int e_up = S_.active(j - dir) ? e - dir * Ni : -1;               // handled by the chain
            // O-grid: the last element of the row meets the (already updated) first one across the wrap
            const int e_wrap = (S_.per_i && i == (dir > 0 ? Ni - 1 : 0)) ? e - dir * (Ni - 1) : -1;
            double acc = 0.0, acc_chain = 0.0;
            for (int jj = x_zero ? indptr[e + 1] : indptr[e]; jj < indptr[e + 1]; ++jj) {
                const int col = indices[jj];
                // RES (entry of a smoother call): neighbours in ghost rows are left to the edge kernel of the pass
                // that follows (dgb_block_gs_pass_seq), which sees the halo values of that moment
                const bool chain_part = col == e || col == e_row || col == e_up || col == e_wrap ||
                                        (RES && (col < first || col >= first + count));
                if (!RES && chain_part) continue;
                const double tt = row_dot<B>(data + ((size_t)jj * B + r) * B, x + (size_t)col * B);
                if (chain_part) acc_chain += tt;
                else acc += tt;
            }
            const double f = rhs[(size_t)e * B + r];
            if (RES) {
                const double res = f - (acc + acc_chain);
                if (r_out != nullptr) r_out[(size_t)e * B + r] = res;
                sumsq = fma(res, res, sumsq);
            }
            s_rhs[el * B + r] = f;
            s_rsum[el * B + r] = f - acc;
        }
        __syncthreads();
        if (e >= 0) {
            const double *d = dinv + ((size_t)e * B + r) * B;
            double tt = 0.0, td = 0.0;
#pragma unroll
            for (int c = 0; c < B; ++c) {
                const double dc = PRE ? dv[PRE ? c : 0] : d[c];
                tt = fma(dc, s_rsum[el * B + c], tt);
                td = fma(dc, s_rhs[el * B + c], td);
            }
            const int j = e / Ni, i = e - j * Ni;
            double *mine = rec + (size_t)chain_loc<B>(S_, dir, i, j) * REC + ChainCfg<B>::c_off(r);
            // c_e and d_e = Dinv_e rhs_e.  The opposite direction's record gets its (c, d) pair from the chain pass that
            // follows (its stores carry d along): an 8-byte store of d here would only half-fill sectors of that stream
            *reinterpret_cast<double2 *>(mine) = make_double2(tt, td);
        }
        __syncthreads();
    }
    if (RES) {
        const double t = block_sum<HelperCfg<B>::NT>(sumsq, s_red);
        if (threadIdx.x == 0) partials[blockIdx.x] = t;
    }
}

// ---- residual after a pass, from the record stream -------------------------------------------------------
// After a pass in direction D the records of direction -D hold, for every element e,
//     c_e = Dinv_e (rhs_e - sum_{n before e in D order} A_en x_n)        (what the next, opposite pass starts from)
// and the two pre-multiplied blocks of e's predecessors in -D order, i.e. of the neighbours that come AFTER e in D
// order.  Hence   Dinv_e r_e = c_e + M_row x_(row-pred) + M_up x_(up-pred) [+ M_wrap x_first] - x_e   and
//     r_e = A_ee (that)                                                   (r = rhs - A x, dgfem/relaxation.py:208)
// from 2 b^2 + 2 b record doubles plus the diagonal block, instead of all five blocks of the row: the residual
// test after every smoother iteration streams 0.64 of the bytes (b = 9).  Same mathematics as rhs - A x; the
// rounding differs (products with Dinv A instead of A), at the level the chained kernel's own already does.
// One thread per scalar row; a CTA stages TE consecutive records in shared memory with coalesced 16-byte loads.
template <int B>
struct ResRecCfg {
    static constexpr int TE = (256 / B) > 0 ? (256 / B) : 1;            // records per tile
    static constexpr int NT = ((TE * B + 31) / 32) * 32;
    static constexpr int NBUF = 2;                                       // tiles in flight per CTA (bulk copies)
    // staged records (NBUF tiles) | rho | x of the element, of its row predecessor and of its up predecessor | mbarriers
    static constexpr size_t o_bar = sizeof(double) * ((size_t)NBUF * TE * ChainCfg<B>::REC + 4 * TE * B);
    static constexpr size_t smem = o_bar + sizeof(uint64_t) * NBUF;
};

// The CTA's tiles arrive by TMA bulk copies (a tile = TE consecutive records = one contiguous piece of the stream),
// two tiles ahead of the arithmetic: the stream never waits for the three phases of a tile (stage x | rho | A_ee rho).
template <int B>
__global__ void __launch_bounds__(ResRecCfg<B>::NT)
k_residual_rec(const double *__restrict__ rec, const double *__restrict__ wrapm, const double *__restrict__ data,
               const double *__restrict__ x, Stencil S_, int dirp /* direction of the records = -D */,
               const int32_t *__restrict__ skip, double *r_out, double *partials, int *err) {
    using C = ChainCfg<B>;
    using RC = ResRecCfg<B>;
    constexpr int TE = RC::TE, NT = RC::NT, REC = C::REC, R = C::R, B2 = B * B, NBUF = RC::NBUF;
    if (skip != nullptr && *skip != 0) return;
    extern __shared__ __align__(128) unsigned char s_raw[];
    double *s_dyn = reinterpret_cast<double *>(s_raw);
    double *s_rho = s_dyn + (size_t)NBUF * TE * REC;    // [TE][B]
    double *s_x = s_rho + TE * B;                       // [3][TE][B]: x_e, x_row-pred, x_up-pred (zero when absent)
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_raw + RC::o_bar);
    __shared__ double s_red[32];
    const int Ni = S_.Ni, nrows = S_.ja1 - S_.ja0;
    // work item = (band, tile of TE consecutive records of that band): 32-bit index arithmetic
    const unsigned per_band = (unsigned)(Ni + R - 1) * R;                 // records of a band
    const unsigned tiles_band = (per_band + TE - 1) / TE;
    const unsigned nbands = (unsigned)((nrows + R - 1) / R);
    const unsigned nitems = nbands * tiles_band;
    const int el = threadIdx.x / B, r = threadIdx.x - el * B;
    // thread 0: bring the tile of work item `item` into buffer `buf`
    auto fetch = [&](unsigned item, int buf) {
        const unsigned band = item / tiles_band;
        const unsigned q0 = (item - band * tiles_band) * TE;
        const unsigned cnt = per_band - q0 < (unsigned)TE ? per_band - q0 : (unsigned)TE;
        const uint32_t bytes = cnt * (uint32_t)(REC * sizeof(double));
        mbar_expect_tx(&bars[buf], bytes);
        bulk_g2s(s_dyn + (size_t)buf * TE * REC, rec + ((size_t)band * per_band + q0) * REC, bytes, &bars[buf]);
    };
    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; ++b) mbar_init(&bars[b], 1);
        fence_barrier_init();
        fence_proxy_async();
        for (int b = 0; b < NBUF; ++b)
            if (blockIdx.x + (unsigned)b * gridDim.x < nitems) fetch(blockIdx.x + (unsigned)b * gridDim.x, b);
    }
    __syncthreads();
    double sumsq = 0.0;
    unsigned k = 0;
    for (unsigned item = blockIdx.x; item < nitems; item += gridDim.x, ++k) {
        const int buf = (int)(k % NBUF);
        const double *s_rec = s_dyn + (size_t)buf * TE * REC;
        const unsigned band = item / tiles_band;
        const unsigned q0 = (item - band * tiles_band) * TE;              // first record of the tile inside the band
        const int cnt = (int)(per_band - q0 < (unsigned)TE ? per_band - q0 : (unsigned)TE);
        long long e = -1;
        int i = 0, j = 0;
        if (el < cnt) {
            const unsigned rem = q0 + (unsigned)el;
            const int t = (int)(rem / R), g = (int)(rem - (unsigned)t * R);
            const int sr = (int)band * R + g, idx = t - g;
            if (sr < nrows && idx >= 0 && idx < Ni) {
                j = dirp > 0 ? S_.ja0 + sr : S_.ja1 - 1 - sr;
                i = dirp > 0 ? idx : Ni - 1 - idx;
                e = (long long)j * Ni + i;
            }
        }
        // my row of the diagonal block: requested now, used after the second barrier
        double av[B];
        if (e >= 0) {
            int c5[5], rk[5];
            S_.cols(i, j, c5);
            slot_ranks(c5, rk);
            load_row<B>(data + ((size_t)(S_.row_start(i, j) + rk[0]) * B + r) * B, av);
            const bool has_row = (i - dirp >= 0 && i - dirp < Ni), has_up = S_.active(j - dirp);
            s_x[el * B + r] = x[e * B + r];
            s_x[(TE + el) * B + r] = has_row ? x[(e - dirp) * B + r] : 0.0;
            s_x[(2 * TE + el) * B + r] = has_up ? x[(e - (long long)dirp * Ni) * B + r] : 0.0;
        }
        if (!mbar_wait(&bars[buf], (uint32_t)((k / NBUF) & 1), err)) return;
        __syncthreads();          // s_x complete; everybody is done with the previous tile's rho
        if (e >= 0) {
            const double *m = s_rec + (size_t)el * REC;
            const double *xr = s_x + (TE + el) * B, *xu = s_x + (2 * TE + el) * B;
            double a0 = m[C::c_off(r)] - s_x[el * B + r], a1 = 0.0;
#pragma unroll
            for (int c = 0; c < B; ++c) {
                a0 = fma(m[C::mat_offset(r, c)], xr[c], a0);
                a1 = fma(m[B2 + C::mat_offset(r, c)], xu[c], a1);
            }
            if (S_.per_i && i == (dirp > 0 ? Ni - 1 : 0)) {       // the row's last element in record order: wrap block
                const double *w = wrapm + (size_t)j * B2;
                const double *xf = x + (e - (long long)dirp * (Ni - 1)) * B;
                for (int c = 0; c < B; ++c) a1 = fma(w[C::mat_offset(r, c)], xf[c], a1);
            }
            s_rho[el * B + r] = a0 + a1;
        }
        __syncthreads();          // rho complete; the tile's buffer and s_x are free
        if (threadIdx.x == 0 && item + (unsigned)NBUF * gridDim.x < nitems) fetch(item + (unsigned)NBUF * gridDim.x, buf);
        if (e >= 0) {
            // the diagonal block of row e sits at row_start + (number of smaller columns)
            double res = 0.0;
#pragma unroll
            for (int c = 0; c < B; ++c) res = fma(av[c], s_rho[el * B + c], res);
            if (r_out != nullptr) r_out[e * B + r] = res;
            sumsq = fma(res, res, sumsq);
        }
    }
    const double tsum = block_sum<NT>(sumsq, s_red);
    if (threadIdx.x == 0) partials[blockIdx.x] = tsum;
}

// ---- slabs: the c a chain pass leaves for the opposite direction lacks the terms of neighbours in ghost rows
// (they are nobody's chain terms); after the halo exchange they are subtracted for the two edge rows:
//     c_e -= Dinv_e A_e,ghost x_ghost        (one thread per scalar row of the ghost-adjacent element rows)
template <int B>
__global__ void __launch_bounds__(HelperCfg<B>::NT)
k_gs_edge_helper(const double *__restrict__ data, const int32_t *__restrict__ indices,
                 const int32_t *__restrict__ indptr, const double *__restrict__ dinv, const double *__restrict__ x,
                 double *rec, Stencil S_, int dir, const int32_t *__restrict__ skip) {
    constexpr int EPB = HelperCfg<B>::EPB, REC = ChainCfg<B>::REC;
    if (skip != nullptr && *skip != 0) return;
    __shared__ double s_t[EPB * B];
    const int el = threadIdx.x / B, r = threadIdx.x - el * B;
    const int Ni = S_.Ni;
    const int nlo = S_.ja0 > 0 ? Ni : 0, nhi = S_.ja1 < S_.Nj ? Ni : 0;      // elements next to a ghost row
    const int count = nlo + nhi;
    const int ntiles = (count + EPB - 1) / EPB;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int idx = tile * EPB + el;
        int e = -1, eg = -1;
        if (el < EPB && idx < count) {
            const bool lo = idx < nlo;
            const int i = lo ? idx : idx - nlo;
            const int j = lo ? S_.ja0 : S_.ja1 - 1;
            e = j * Ni + i;
            eg = lo ? e - Ni : e + Ni;
        }
        // a slab of one active row has both ghost neighbours on the same element: handle lo then hi entries
        if (e >= 0) {
            double acc = 0.0;
            for (int jj = indptr[e]; jj < indptr[e + 1]; ++jj)
                if (indices[jj] == eg) acc += row_dot<B>(data + ((size_t)jj * B + r) * B, x + (size_t)eg * B);
            s_t[el * B + r] = acc;
        }
        __syncthreads();
        if (e >= 0) {
            const double *d = dinv + ((size_t)e * B + r) * B;
            double tt = 0.0;
#pragma unroll
            for (int c = 0; c < B; ++c) tt = fma(d[c], s_t[el * B + c], tt);
            const int j = e / Ni, i = e - j * Ni;
            double *mine = rec + (size_t)chain_loc<B>(S_, dir, i, j) * REC + ChainCfg<B>::c_off(r);
            if (S_.ja1 - S_.ja0 == 1 && nlo && nhi) atomicAdd(mine, -tt);   // both ghost terms land on one entry
            else *mine -= tt;
        }
        __syncthreads();
    }
}

// ---- records: { -Dinv_e A_e,row-predecessor | -Dinv_e A_e,previous-row | c | pad } at chain_loc(e) ----
template <int B>
__global__ void __launch_bounds__(256)
k_build_gs_chain(const double *__restrict__ data, const int32_t *__restrict__ indices,
                 const int32_t *__restrict__ indptr, const double *__restrict__ dinv, Stencil S_, double *rec) {
    constexpr int B2 = B * B, REC = ChainCfg<B>::REC;
    const int Ni = S_.Ni;
    const long long N = (long long)Ni * S_.Nj;
    const long long total = N * 4 * B2;
    for (long long tt = (long long)blockIdx.x * 256 + threadIdx.x; tt < total; tt += (long long)gridDim.x * 256) {
        const int e = (int)(tt / (4 * B2));
        const int q = (int)(tt - (long long)e * (4 * B2));
        const int slot = q / B2, rc = q - slot * B2;
        const int r = rc / B, c = rc - r * B;
        const int j = e / Ni, i = e - j * Ni;
        const int dir = slot < 2 ? 1 : -1;
        int col = -1;
        if (S_.active(j)) {
            if ((slot & 1) == 0) col = (i - dir >= 0 && i - dir < Ni) ? e - dir : -1;
            else col = S_.active(j - dir) ? e - dir * Ni : -1;
        }
        double v = 0.0;
        if (col >= 0) {
            for (int jj = indptr[e]; jj < indptr[e + 1]; ++jj) {
                if (indices[jj] != col) continue;
                const double *d = dinv + ((size_t)e * B + r) * B;
                const double *a = data + (size_t)jj * B2 + c;
                double s = 0.0;
                for (int k = 0; k < B; ++k) s = fma(d[k], a[k * B], s);
                v -= s;
            }
        }
        if (S_.active(j))
            rec[((size_t)(slot >> 1) * chain_dir_records<B>(S_) + (size_t)chain_loc<B>(S_, dir, i, j)) * REC + (slot & 1) * B2 + ChainCfg<B>::mat_offset(r, c)] = v;
    }
    if (!S_.per_i) return;
    // O-grid: wrap[dir][row j] = -Dinv_e A_e,first for the row's last element e in that sweep direction
    double *wrap = rec + 2 * (size_t)chain_dir_records<B>(S_) * REC;
    const long long wtotal = 2LL * S_.Nj * B2;
    for (long long tt = (long long)blockIdx.x * 256 + threadIdx.x; tt < wtotal; tt += (long long)gridDim.x * 256) {
        const int d01 = (int)(tt / ((long long)S_.Nj * B2));
        const int rem = (int)(tt - (long long)d01 * S_.Nj * B2);
        const int j = rem / B2, rc = rem - j * B2;
        const int r = rc / B, c = rc - r * B;
        const int dir = d01 == 0 ? 1 : -1;
        const int e = j * Ni + (dir > 0 ? Ni - 1 : 0), col = j * Ni + (dir > 0 ? 0 : Ni - 1);
        double v = 0.0;
        if (S_.active(j)) {
            for (int jj = indptr[e]; jj < indptr[e + 1]; ++jj) {
                if (indices[jj] != col) continue;
                const double *d = dinv + ((size_t)e * B + r) * B;
                const double *a = data + (size_t)jj * B2 + c;
                double sacc = 0.0;
                for (int k = 0; k < B; ++k) sacc = fma(d[k], a[k * B], sacc);
                v -= sacc;
            }
        }
        wrap[((size_t)d01 * S_.Nj + j) * B2 + ChainCfg<B>::mat_offset(r, c)] = v;
    }
}

// The same records for b >= 16 with the products -Dinv_e A_e,n on the FP64 tensor cores: one CTA (8 warps) per
// element, Dinv^T and the neighbour block staged in shared memory (C = Dinv A: L[t][k] = Dinv[k][t], R[t][l] = A[t][l]).
template <int B>
__global__ void __launch_bounds__(256)
k_build_gs_chain_mma(const double *__restrict__ data, const int32_t *__restrict__ indices,
                     const int32_t *__restrict__ indptr, const double *__restrict__ dinv, Stencil S_, double *rec) {
    using M = MmaCfg<B>;
    constexpr int B2 = B * B, REC = ChainCfg<B>::REC, BS = M::BS, B4 = M::B4;
    __shared__ __align__(16) double Lt[B4 * BS + 8], Rt[B4 * BS + 8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Ni = S_.Ni, N = Ni * S_.Nj;
    for (int t = tid; t < B4 * BS + 8; t += 256) Lt[t] = Rt[t] = 0.0;
    __syncthreads();
    const int nslots = S_.per_i ? 6 : 4;          // 4 record blocks + the two wrap blocks of the row ends
    for (int e = blockIdx.x; e < N; e += gridDim.x) {
        const int j = e / Ni, i = e - j * Ni;
        if (!S_.active(j)) continue;
        for (int t = tid; t < B2; t += 256) Lt[(t % B) * BS + t / B] = dinv[(size_t)e * B2 + t];     // Dinv^T
        for (int slot = 0; slot < nslots; ++slot) {
            const int dir = (slot & 2) ? -1 : 1;
            int col = -1;
            double *dst = nullptr;
            if (slot < 4) {
                if ((slot & 1) == 0) col = (i - dir >= 0 && i - dir < Ni) ? e - dir : -1;
                else col = S_.active(j - dir) ? e - dir * Ni : -1;
                dst = rec + ((size_t)(slot >> 1) * chain_dir_records<B>(S_) + (size_t)chain_loc<B>(S_, dir, i, j)) * REC + (slot & 1) * B2;
            } else {                               // wrap[d01][row j]: the row's last element meets its first one
                const int d01 = slot - 4, wd = d01 == 0 ? 1 : -1;
                if (i == (wd > 0 ? Ni - 1 : 0)) {
                    col = j * Ni + (wd > 0 ? 0 : Ni - 1);
                    dst = rec + 2 * (size_t)chain_dir_records<B>(S_) * REC + ((size_t)d01 * S_.Nj + j) * B2;
                }
            }
            int jj = -1;
            if (col >= 0)
                for (int q = indptr[e]; q < indptr[e + 1]; ++q)
                    if (indices[q] == col) jj = q;
            __syncthreads();                       // Lt complete / the previous slot's products are done with Rt
            if (jj < 0) continue;                  // no such neighbour: the record block stays zero (uniform over the CTA)
            for (int t = tid; t < B2; t += 256) Rt[(t / B) * BS + t % B] = data[(size_t)jj * B2 + t];
            __syncthreads();
            double acc[M::MAXT][2];
#pragma unroll
            for (int s = 0; s < M::MAXT; ++s) acc[s][0] = acc[s][1] = 0.0;
            mma_lr<B>(acc, Lt, Rt, B4, warp, lane);
            const int m = lane >> 2, kk = lane & 3;
#pragma unroll
            for (int s = 0; s < M::MAXT; ++s) {
                const int tile = warp + 8 * s;
                if (tile >= M::NT2) continue;
                const int row = 8 * (tile / M::NTL) + m, c0 = 8 * (tile % M::NTL) + 2 * kk;
                if (row < B && c0 < B) dst[ChainCfg<B>::mat_offset(row, c0)] = -acc[s][0];
                if (row < B && c0 + 1 < B) dst[ChainCfg<B>::mat_offset(row, c0 + 1)] = -acc[s][1];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// host side
// block sizes the chained kernel is used for: bit 0 b=4, bit 1 b=9, bit 2 b=16, bit 3 b=25, bit 4 b=36
// (dgb_set_kernel_path(300 + mask); the default follows the measurements in profiles/)
int g_chain_mask = 31;
bool chain_supported(int b, int flags) {
    if (g_gs_variant == 9) return false;            // tuning: force the row-pipelined kernel
    const int bit = b == 4 ? 1 : b == 9 ? 2 : b == 16 ? 4 : b == 25 ? 8 : b == 36 ? 16 : 0;
    // periodic in j (fully periodic grids) is not handled
    return flags >= 0 && (flags & DGB_FLAG_PERIODIC_J) == 0 && (g_chain_mask & bit) != 0;
}
// O-grids need three distinct elements per row (else the wrap neighbour coincides with the row neighbour)
static bool chain_shape_ok(int Ni, int flags) { return !(flags & DGB_FLAG_PERIODIC_I) || Ni >= 3; }
// doubles of one direction's record stream
static long long chain_dir_len(int b, const Stencil &S_) {
    switch (b) {
    case 4: return chain_dir_records<4>(S_) * ChainCfg<4>::REC;
    case 9: return chain_dir_records<9>(S_) * ChainCfg<9>::REC;
    case 16: return chain_dir_records<16>(S_) * ChainCfg<16>::REC;
    case 25: return chain_dir_records<25>(S_) * ChainCfg<25>::REC;
    case 36: return chain_dir_records<36>(S_) * ChainCfg<36>::REC;
    }
    return 0;
}

int g_chain_cluster = 0;        // CTAs per cluster of the chain kernel, 0 = automatic (tuning: dgb_set_kernel_path(400 + n))

template <int B, int W, int DIR>
static int chain_launch_d(const double *rec, double *rec_other, const double *wrapm, double *x, double *mbox,
                          Stencil S_, const int32_t *skip, cudaStream_t st) {
    using C = ChainCfg<B>;
    static bool configured = false;
    static int active[17] = {0};     // clusters of each size the device keeps resident at once
    auto kern = k_gs_chain<B, W, DIR>;
    if (!configured) {
        DGB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem(W)));
        (void)cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);    // sizes 9..16
        (void)cudaGetLastError();
        for (int cs = 16; cs >= 1; --cs) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs, 1, 1);
            cfg.blockDim = dim3((W + C::PW) * 32, 1, 1);
            cfg.dynamicSmemBytes = C::smem(W);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) active[cs] = n;
            (void)cudaGetLastError();
        }
        active[1] = active[1] > 0 ? active[1] : 1;
        configured = true;
        if (getenv("DGB_CHAIN_VERBOSE"))
            fprintf(stderr, "k_gs_chain<%d,%d>: resident clusters by size 1..16: %d %d %d %d %d %d %d %d | %d %d %d %d %d %d %d %d\n",
                    B, W, active[1], active[2], active[3], active[4], active[5], active[6], active[7], active[8], active[9],
                    active[10], active[11], active[12], active[13], active[14], active[15], active[16]);
    }
    DGB_CUDA_OK(cudaMemsetAsync(work_ptr(), 0, sizeof(int), st));
    const int nbands = (S_.ja1 - S_.ja0 + C::R - 1) / C::R;
    const int nctas = (nbands + W - 1) / W;
    // Cluster size.  A level whose CTAs are all resident at once takes the portable maximum (8: fewest hand-overs
    // through the global mailbox).  A larger level runs in waves -- a band starts when an earlier one has finished
    // its whole row -- so what counts is how many CTAs the device keeps resident: size 8 fits 15 clusters (120 of 148
    // SMs) on B200, size 9 also 15 (135 SMs), size 6 22 (132); measured on 2048^2 b = 9: 1.75 / 1.58 / 1.61 ms per pass
    // (profiles/r02_probe_chain_cluster.md).  g_chain_cluster > 0 forces a size (tuning).
    int cs = 8;
    if (g_chain_cluster > 0) {
        cs = g_chain_cluster < 16 ? g_chain_cluster : 16;
    } else if (nctas > active[8] * 8) {
        int best = active[8] * 8;
        for (int c = 6; c <= 16; ++c)
            if (active[c] * c > best || (active[c] * c == best && c > cs)) {
                best = active[c] * c;
                cs = c;
            }
    }
    if (cs > nctas) cs = nctas;
    while (cs > 1 && active[cs] == 0) --cs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(((nctas + cs - 1) / cs) * cs), 1, 1);
    cfg.blockDim = dim3((W + C::PW) * 32, 1, 1);
    cfg.dynamicSmemBytes = C::smem(W);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    DGB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, rec, rec_other, wrapm, x, mbox, S_, work_ptr(), err_ptr(), skip));
    DGB_LAUNCH_OK();
    return 0;
}

template <int B, int W>
static int chain_launch_w(const double *rec, double *rec_other, const double *wrapm, double *x, double *mbox,
                          Stencil S_, int dir, const int32_t *skip, cudaStream_t st) {
    return dir > 0 ? chain_launch_d<B, W, 1>(rec, rec_other, wrapm, x, mbox, S_, skip, st)
                   : chain_launch_d<B, W, -1>(rec, rec_other, wrapm, x, mbox, S_, skip, st);
}

bool chain_c_recurrence(int flags);

// have_c: the previous pass of this smoother call ran in the opposite direction on the same rhs and x has
// not changed since -- its chain left this pass's c in the record stream, the helper is not needed
template <int B>
static int chain_pass_t(const dgb_operator *op, const double *rhs, double *x, int dir, bool have_c,
                        const int32_t *skip, cudaStream_t st) {
    using C = ChainCfg<B>;
    using H = HelperCfg<B>;
    const Stencil S_ = make_stencil(op->Ni, op->Nj, op->stencil);
    double *rec = op->gs_chain + (dir > 0 ? 0 : chain_dir_len(B, S_));
    double *rec_other = op->gs_chain + (dir > 0 ? chain_dir_len(B, S_) : 0);
    const int count = (S_.ja1 - S_.ja0) * S_.Ni;
    int grid = (count + H::EPB - 1) / H::EPB;
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    if (have_c && !chain_c_recurrence(op->stencil)) {
        const int edge = ((S_.ja0 > 0) + (S_.ja1 < S_.Nj)) * S_.Ni;
        int ge = (edge + H::EPB - 1) / H::EPB;
        if (ge > sm_count() * 8) ge = sm_count() * 8;
        k_gs_edge_helper<B><<<ge, H::NT, 0, st>>>(op->data, op->indices, op->indptr, op->dinv, x, rec, S_, dir, skip);
        DGB_LAUNCH_OK();
    }
    if (!have_c && g_gs_variant != 22) {        // (21 / 22: time the two launches separately, results are then meaningless)
        k_gs_helper<B, false><<<grid, H::NT, 0, st>>>(op->data, op->indices, op->indptr, op->dinv, rhs, x, rec,
                                                      S_, dir, skip, nullptr, nullptr, 0);
        DGB_LAUNCH_OK();
    }
    if (g_gs_variant == 21) return 0;
    // O-grid: wrap blocks behind the two record streams, [direction][row][B2]
    const double *wrapm = op->gs_chain + 2 * chain_dir_len(B, S_) + (dir > 0 ? 0 : (long long)S_.Nj * C::B2);
    if (B <= 9 && g_gs_variant == 11) return chain_launch_w<B, 2>(rec, rec_other, wrapm, x, op->gs_mailbox, S_, dir, skip, st);
    return chain_launch_w<B, C::WDEF>(rec, rec_other, wrapm, x, op->gs_mailbox, S_, dir, skip, st);
}

// helper of the pass in direction `dir` fused with the residual r = rhs - A x (r may be NULL) and its per-CTA
// sums of squares; *grid_out = number of partials written
template <int B>
static int helper_residual_t(const dgb_operator *op, const double *rhs, const double *x, int dir, double *r,
                             double *partials, int *grid_out, cudaStream_t st, bool x_zero) {
    using H = HelperCfg<B>;
    const Stencil S_ = make_stencil(op->Ni, op->Nj, op->stencil);
    double *rec = op->gs_chain + (dir > 0 ? 0 : chain_dir_len(B, S_));
    const int count = (S_.ja1 - S_.ja0) * S_.Ni;
    int grid = (count + H::EPB - 1) / H::EPB;
    constexpr int occ = (B > 4 && B <= 16) ? 5 : 6;
    if (grid > sm_count() * occ) grid = sm_count() * occ;    // one wave at the occupancy __launch_bounds__ asks for
    if (grid > kMaxPartials) grid = kMaxPartials;
    k_gs_helper<B, true><<<grid, H::NT, 0, st>>>(op->data, op->indices, op->indptr, op->dinv, rhs, x, rec, S_,
                                                 dir, nullptr, r, partials, x_zero ? 1 : 0);
    DGB_LAUNCH_OK();
    *grid_out = grid;
    return 0;
}
// residual after a pass in direction `last_dir` from the records of the opposite direction; *grid_out partials
template <int B>
static int residual_rec_t(const dgb_operator *op, const double *x, int last_dir, double *r, double *partials,
                          int *grid_out, const int32_t *skip, cudaStream_t st) {
    using C = ChainCfg<B>;
    using RC = ResRecCfg<B>;
    static bool configured = false;
    static int occ = 1;
    if (!configured) {
        DGB_CUDA_OK(cudaFuncSetAttribute(k_residual_rec<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RC::smem));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_residual_rec<B>, RC::NT, RC::smem) != cudaSuccess || occ < 1)
            occ = 1;
        if (occ > 8) occ = 8;
        configured = true;
    }
    const Stencil S_ = make_stencil(op->Ni, op->Nj, op->stencil);
    const int dirp = -last_dir;
    const double *rec = op->gs_chain + (dirp > 0 ? 0 : chain_dir_len(B, S_));
    const double *wrapm = op->gs_chain + 2 * chain_dir_len(B, S_) + (dirp > 0 ? 0 : (long long)S_.Nj * C::B2);
    const long long per_band = (long long)(S_.Ni + C::R - 1) * C::R;
    const long long nbands = (S_.ja1 - S_.ja0 + C::R - 1) / C::R;
    long long grid = nbands * ((per_band + RC::TE - 1) / RC::TE);
    if (grid > (long long)sm_count() * occ) grid = (long long)sm_count() * occ;
    if (grid > kMaxPartials) grid = kMaxPartials;
    k_residual_rec<B><<<(int)grid, RC::NT, RC::smem, st>>>(rec, wrapm, op->data, x, S_, dirp, skip, r, partials, err_ptr());
    DGB_LAUNCH_OK();
    *grid_out = (int)grid;
    return 0;
}
// block sizes the record residual is compiled for (shared-memory tile of 256/b records)
bool chain_residual_supported(int b) { return b == 4 || b == 9 || b == 16; }
int gs_chain_residual(const dgb_operator *op, const double *x, int last_dir, double *r, double *partials,
                      int *grid_out, const int32_t *skip, cudaStream_t st) {
    switch (op->b) {
    case 4: return residual_rec_t<4>(op, x, last_dir, r, partials, grid_out, skip, st);
    case 9: return residual_rec_t<9>(op, x, last_dir, r, partials, grid_out, skip, st);
    case 16: return residual_rec_t<16>(op, x, last_dir, r, partials, grid_out, skip, st);
    }
    set_error("gs_chain_residual: unsupported block size b=%d", op->b);
    return 2;
}

int gs_chain_helper_residual(const dgb_operator *op, const double *rhs, const double *x, int dir, double *r,
                             double *partials, int *grid_out, cudaStream_t st, bool x_zero) {
    switch (op->b) {
    case 4: return helper_residual_t<4>(op, rhs, x, dir, r, partials, grid_out, st, x_zero);
    case 9: return helper_residual_t<9>(op, rhs, x, dir, r, partials, grid_out, st, x_zero);
    case 16: return helper_residual_t<16>(op, rhs, x, dir, r, partials, grid_out, st, x_zero);
    case 25: return helper_residual_t<25>(op, rhs, x, dir, r, partials, grid_out, st, x_zero);
    case 36: return helper_residual_t<36>(op, rhs, x, dir, r, partials, grid_out, st, x_zero);
    }
    set_error("gs_chain_helper_residual: unsupported block size b=%d", op->b);
    return 2;
}

template <int B, int W, int DIR>
static int big_launch_d(const double *rec, double *rec_other, const double *wrapm, double *x, double *mbox,
                        Stencil S_, const int32_t *skip, cudaStream_t st) {
    static bool configured = false;
    static int max_cluster = 1;
    auto kern = k_gs_chain_big<B, W, DIR>;
    const size_t smem = BigCfg<B>::smem(W);
    if (!configured) {
        DGB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int cs = 8; cs >= 1; cs >>= 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs, 1, 1);
            cfg.blockDim = dim3(W * 32, 1, 1);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) {
                max_cluster = cs;
                break;
            }
            (void)cudaGetLastError();
        }
        configured = true;
    }
    DGB_CUDA_OK(cudaMemsetAsync(work_ptr(), 0, sizeof(int), st));
    const int nctas = (S_.ja1 - S_.ja0 + W - 1) / W;
    int cs = 1;
    const int cs_max = g_chain_cluster > 0 ? g_chain_cluster : 8;
    while (cs * 2 <= max_cluster && cs * 2 <= cs_max && cs < nctas) cs *= 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(((nctas + cs - 1) / cs) * cs), 1, 1);
    cfg.blockDim = dim3(W * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    DGB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, rec, rec_other, wrapm, x, mbox, S_, work_ptr(), err_ptr(), skip));
    DGB_LAUNCH_OK();
    return 0;
}

template <int B>
static int chain_pass_big(const dgb_operator *op, const double *rhs, double *x, int dir, bool have_c,
                          const int32_t *skip, cudaStream_t st) {
    using C = ChainCfg<B>;
    using H = HelperCfg<B>;
    const Stencil S_ = make_stencil(op->Ni, op->Nj, op->stencil);
    double *rec = op->gs_chain + (dir > 0 ? 0 : chain_dir_len(B, S_));
    double *rec_other = op->gs_chain + (dir > 0 ? chain_dir_len(B, S_) : 0);
    const int count = (S_.ja1 - S_.ja0) * S_.Ni;
    int grid = (count + H::EPB - 1) / H::EPB;
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    if (have_c && !chain_c_recurrence(op->stencil)) {
        const int edge = ((S_.ja0 > 0) + (S_.ja1 < S_.Nj)) * S_.Ni;
        int ge = (edge + H::EPB - 1) / H::EPB;
        if (ge > sm_count() * 8) ge = sm_count() * 8;
        k_gs_edge_helper<B><<<ge, H::NT, 0, st>>>(op->data, op->indices, op->indptr, op->dinv, x, rec, S_, dir, skip);
        DGB_LAUNCH_OK();
    }
    if (!have_c && g_gs_variant != 22) {
        k_gs_helper<B, false><<<grid, H::NT, 0, st>>>(op->data, op->indices, op->indptr, op->dinv, rhs, x, rec,
                                                      S_, dir, skip, nullptr, nullptr, 0);
        DGB_LAUNCH_OK();
    }
    if (g_gs_variant == 21) return 0;
    const double *wrapm = op->gs_chain + 2 * chain_dir_len(B, S_) + (dir > 0 ? 0 : (long long)S_.Nj * C::B2);
    return dir > 0 ? big_launch_d<B, C::WDEF, 1>(rec, rec_other, wrapm, x, op->gs_mailbox, S_, skip, st)
                   : big_launch_d<B, C::WDEF, -1>(rec, rec_other, wrapm, x, op->gs_mailbox, S_, skip, st);
}

// the opposite-direction c left behind by a chain pass is complete only when no neighbour lives in a ghost row
bool chain_c_recurrence(int flags) { return (flags & (DGB_FLAG_GHOST_LO | DGB_FLAG_GHOST_HI)) == 0; }

int gs_chain_pass(const dgb_operator *op, const double *rhs, double *x, int dir, bool have_c, const int32_t *skip,
                  cudaStream_t st) {
    int rc = ensure_work(0);
    if (rc) return rc;
    switch (op->b) {
    case 4: return chain_pass_t<4>(op, rhs, x, dir, have_c, skip, st);
    case 9: return chain_pass_t<9>(op, rhs, x, dir, have_c, skip, st);
    case 16: return chain_pass_t<16>(op, rhs, x, dir, have_c, skip, st);
    case 25: return chain_pass_t<25>(op, rhs, x, dir, have_c, skip, st);
    case 36: return chain_pass_big<36>(op, rhs, x, dir, have_c, skip, st);
    }
    set_error("gs_chain_pass: unsupported block size b=%d", op->b);
    return 2;
}

}  // namespace dgb

using namespace dgb;

extern "C" {

int64_t dgb_gs_chain_len(int32_t b, int32_t Ni, int32_t Nj, int32_t stencil) {
    if (!chain_supported(b, stencil) || Ni <= 0 || Nj <= 0 || !chain_shape_ok(Ni, stencil)) return 0;
    return 2 * (int64_t)chain_dir_len(b, make_stencil(Ni, Nj, stencil)) +
           ((stencil & DGB_FLAG_PERIODIC_I) ? 2 * (int64_t)Nj * b * b : 0) + 2;
}

int dgb_build_gs_chain(const dgb_operator *op, void *stream) {
    DGB_ARG(op != nullptr && op->data && op->indices && op->indptr && op->dinv && op->gs_chain);
    DGB_ARG(chain_supported(op->b, op->stencil) && chain_shape_ok(op->Ni, op->stencil));
    cudaStream_t st = (cudaStream_t)stream;
    const Stencil S_ = make_stencil(op->Ni, op->Nj, op->stencil);
    const long long N = (long long)op->Ni * op->Nj;
    DGB_CUDA_OK(cudaMemsetAsync(op->gs_chain, 0, sizeof(double) * (size_t)dgb_gs_chain_len(op->b, op->Ni, op->Nj, op->stencil), st));
    long long g = (N * 4 * op->b * op->b + 255) / 256;
    if (g > sm_count() * 16) g = sm_count() * 16;
    switch (op->b) {
    case 4: k_build_gs_chain<4><<<(int)g, 256, 0, st>>>(op->data, op->indices, op->indptr, op->dinv, S_, op->gs_chain); break;
    case 9: k_build_gs_chain<9><<<(int)g, 256, 0, st>>>(op->data, op->indices, op->indptr, op->dinv, S_, op->gs_chain); break;
    // b >= 16: the products run on the FP64 tensor cores (dgb_set_kernel_path(100 + 51) keeps the per-entry kernel)
    case 16:
        if (g_gs_variant != 51) k_build_gs_chain_mma<16><<<(int)(N < sm_count() * 8 ? N : sm_count() * 8), 256, 0, st>>>(op->data, op->indices, op->indptr, op->dinv, S_, op->gs_chain);
        else k_build_gs_chain<16><<<(int)g, 256, 0, st>>>(op->data, op->indices, op->indptr, op->dinv, S_, op->gs_chain);
        break;
    case 25:
        if (g_gs_variant != 51) k_build_gs_chain_mma<25><<<(int)(N < sm_count() * 8 ? N : sm_count() * 8), 256, 0, st>>>(op->data, op->indices, op->indptr, op->dinv, S_, op->gs_chain);
        else k_build_gs_chain<25><<<(int)g, 256, 0, st>>>(op->data, op->indices, op->indptr, op->dinv, S_, op->gs_chain);
        break;
    case 36:
        if (g_gs_variant != 51) k_build_gs_chain_mma<36><<<(int)(N < sm_count() * 8 ? N : sm_count() * 8), 256, 0, st>>>(op->data, op->indices, op->indptr, op->dinv, S_, op->gs_chain);
        else k_build_gs_chain<36><<<(int)g, 256, 0, st>>>(op->data, op->indices, op->indptr, op->dinv, S_, op->gs_chain);
        break;
    }
    DGB_LAUNCH_OK();
    return 0;
}

}  // extern "C"

#ifdef DGB_CHAIN_TRACE
extern "C" int dgb_debug_chain_trace(long long *host_out, int n_bands) {
    if (n_bands > 8192) n_bands = 8192;
    return (int)cudaMemcpyFromSymbol(host_out, dgb::g_chain_trace, sizeof(long long) * 16 * n_bands);
}
#endif
