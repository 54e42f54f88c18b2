// dgb_stream.cu -- the row-pipelined lexicographic block Gauss-Seidel kernel, sm_100a, and the device-side
// bookkeeping (ticket / error words) shared by the single-launch smoother kernels.
//
//  k_gs_rows<B>      lexicographic block Gauss-Seidel, exact sweep order: one warp per element
//                    row j, rows pipelined against each other (row j may process element i once row j-1 has
//                    finished element i); each warp streams its row's blocks through its own
//                    TMA-fed ring.  CTAs take tickets so that a CTA only ever waits on CTAs
//                    that were started before it.  Used where the chained kernel (dgb_chain.cu) does not
//                    apply: grids periodic in j, b = 25 on O-grids, omega != 1.
//
// Reference semantics: pyamg amg_core.block_gauss_seidel (dgfem/pyamg_relaxation.py:252-255),
// dgfem/relaxation.py:170-195.  The kernel reads the "GS stream": the BSR data with each diagonal block
// replaced by its inverse (dgb_build_gs_stream), so one pass reads exactly nnzb blocks.
#include "dgb_async.cuh"
#include "dgb_common.cuh"

namespace dgb {

constexpr int gcd_c(int a, int b) { return b == 0 ? a : gcd_c(b, a % b); }

// Skewed column order: item t (consecutive items <-> consecutive lanes) reads row t of a B-wide
// row-major block from shared memory; rows are B doubles apart, so without a skew lanes t and
// t + PD hit the same bank pair.  Rotating the column order by q = (lane%16)/PD makes the 16 lanes
// of a half-warp hit 16 distinct 8-byte banks for every B used here.
template <int B>
__device__ __forceinline__ double skew_dot(const double *__restrict__ a, const double *v, int q) {
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < B; ++c) {
        int cc = c + q;
        cc = cc >= B ? cc - B : cc;
        acc = fma(a[cc], v[cc], acc);
    }
    return acc;
}

// =========================================================================================
// k_gs_rows
//
// Row j may process element i once row j-1 (in sweep order) has finished element i.  The value
// it needs from that row, x(i, j-1), is handed over
//   - through a shared-memory ring + progress counter when both rows live in the same CTA,
//   - through a global "mailbox" otherwise: the producer stores the 8-byte values themselves
//     (ld/st.cg, L2), the consumer spins until none of them is the sentinel (all-ones NaN) and
//     writes the sentinel back.  Data doubles as its own flag, so no release/acquire fence
//     (MEMBAR / CCTL.IVALL) sits on the critical path.  The mailbox is an [N*b] vector owned by
//     the level, all-sentinel outside a pass.
// Everything else a row reads is a sequential stream along the row and is prefetched deep:
//   - its block row (GS stream) through a TMA bulk-copy ring (S stages, mbarrier complete_tx),
//   - rhs(e), the old x of the next element of the row and of the next row through a cp.async
//     ring D elements ahead (HBM latency under load is ~2 us, an element takes ~0.3 us).
// =========================================================================================
template <int B>
struct GsDefault {
    static constexpr int W = B <= 9 ? 16 : B <= 16 ? 4 : B <= 25 ? 2 : 1;   // rows (warps) per CTA
    static constexpr int S = 3;                                             // TMA ring stages per warp
};
template <int B, int W_ = GsDefault<B>::W, int S_ = GsDefault<B>::S>
struct GsCfg {
    static constexpr int B2 = B * B;
    static constexpr int P = (B == 9) ? 3 : (B == 4) ? 4 : 1;       // lanes per scalar row
    static constexpr int CW = (B + P - 1) / P;                      // columns per lane
    static constexpr int RS = (B + 31) / 32;                        // row slots per lane (P == 1)
    static constexpr int W = W_;                                    // rows (warps) per CTA
    static constexpr int S = S_;                                    // TMA ring stages per warp
    static constexpr int D = B <= 9 ? 8 : 4;                        // cp.async prefetch distance (elements)
    static constexpr int RING = 8;                                  // x hand-over ring slots per warp
    static constexpr int STAGE_D = (5 * B2 + 2 + 1) & ~1;           // doubles per stage, even
    static constexpr int BP = (B + 1) & ~1;
    static constexpr int PD = 16 / gcd_c(B, 16);
    // per warp: stages | vs[5] | xprev | rs | pvs | ring[RING] | vring[D+1][4]
    static constexpr int NV = 4;                                    // rhs | x side | x next row | x previous row (ghost)
    static constexpr int WARP_D = S * STAGE_D + 8 * BP + RING * BP + (D + 1) * NV * BP;
    static constexpr size_t oBar = sizeof(double) * W * WARP_D;          // full[W][S], hand[W][RING]
    static constexpr size_t oProg = oBar + sizeof(uint64_t) * W * (S + RING);
    static constexpr size_t SMEM = oProg + sizeof(int) * W;
};

struct GsElem {
    int e, n, tdiag;
    int col[5];
};

template <int B>
__device__ __forceinline__ GsElem gs_elem(const Stencil &S_, int i, int j) {
    GsElem E;
    int c[5], rk[5];
    S_.cols(i, j, c);
    slot_ranks(c, rk);
    E.e = c[0];
    E.n = 0;
#pragma unroll
    for (int t = 0; t < 5; ++t) {       // sorted position t holds the slot whose rank is t
        int v = -1;
#pragma unroll
        for (int s = 0; s < 5; ++s) v = (rk[s] == t) ? c[s] : v;
        E.col[t] = v;
        E.n += (v >= 0);
    }
    E.tdiag = rk[0];
    return E;
}

__device__ __forceinline__ bool is_sentinel(double v) { return __double_as_longlong(v) == -1LL; }
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// work[0] = ticket counter
template <int B, int WW = GsDefault<B>::W, int SS = GsDefault<B>::S>
__global__ void __launch_bounds__(WW * 32)
k_gs_rows(const double *__restrict__ gs, const double *__restrict__ rhs, double *x, double *mbox, Stencil S_,
          int dir, double omega, int *work, int *err, const int32_t *__restrict__ skip) {
    using C = GsCfg<B, WW, SS>;
    constexpr int B2 = C::B2, S = C::S, P = C::P, CW = C::CW, RS = C::RS, BP = C::BP, RING = C::RING, W = C::W;
    constexpr int D = C::D;
    if (skip != nullptr && *skip != 0) return;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_ticket;
    volatile int *s_prog = reinterpret_cast<volatile int *>(smem + C::oProg);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::oBar);
    if (threadIdx.x == 0) s_ticket = atomicAdd(&work[0], 1);
    if (threadIdx.x < W) s_prog[threadIdx.x] = 0;
    // one thread per warp initialises that warp's barriers (before any warp may touch a neighbour's)
    if ((threadIdx.x & 31) == 0) {
        uint64_t *bw = bars + (threadIdx.x >> 5) * (S + RING);
        for (int s = 0; s < S + RING; ++s) mbar_init(&bw[s], 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // opaque to the compiler from here on: keep them in registers (ptxas otherwise re-derives every
    // shared-memory address from SR_TID in the latency-critical element loop)
    asm volatile("" : "+r"(w));
    asm volatile("" : "+r"(lane));
    const int Ni = S_.Ni, Nj = S_.Nj;
    const int nrows = S_.ja1 - S_.ja0;         // active element rows (ghost rows of a slab are only read)
    const int sr = s_ticket * W + w;           // row index in sweep order
    if (sr >= nrows) return;
    const int j = dir > 0 ? S_.ja0 + sr : S_.ja1 - 1 - sr;
    uint32_t woff = (uint32_t)w * (uint32_t)(C::WARP_D * sizeof(double));
    asm volatile("" : "+r"(woff));
    double *wbase = reinterpret_cast<double *>(smem + woff);
    double *vs = wbase + S * C::STAGE_D;       // [5][BP]  directly loaded neighbour vectors (wrap cases)
    double *xprev = vs + 5 * BP;               // previous element of this row (new value)
    double *rsv = xprev + BP;                  // rhs - sum offdiag
    double *pvs = rsv + BP;                    // predecessor-row value when it came through the mailbox
    double *ring = pvs + BP;                   // [RING][BP], written by this warp, read by warp w+1
    double *vring = ring + RING * BP;          // [D+1][NV][BP]: rhs | x side | x next row | x previous row (old)
    const double *ring_pred = ring - C::WARP_D;
    uint64_t *full = bars + w * (S + RING);    // TMA stage barriers of this warp
    uint64_t *hand = full + S;                 // hand[slot]: this warp has written ring slot `slot`
    uint64_t *hand_pred = hand - (S + RING);

    // ---- row constants: block counts and the interior column pattern ----
    const int n_first = S_.count(0, j), n_last = S_.count(Ni - 1, j), n_int = Ni > 2 ? S_.count(1, j) : 0;
    const long long k_row = S_.row_start(0, j);
    auto k0_of = [&](int i) -> long long { return k_row + (i > 0 ? n_first + (long long)(i - 1) * n_int : 0); };
    auto cnt_of = [&](int i) -> int { return i == 0 ? n_first : (i == Ni - 1 ? n_last : n_int); };
    GsElem tmpl;                               // columns of an interior element, relative to e
    tmpl.e = 0; tmpl.n = 0; tmpl.tdiag = 0;
#pragma unroll
    for (int t = 0; t < 5; ++t) tmpl.col[t] = -1;
    if (Ni > 2) {
        tmpl = gs_elem<B>(S_, 1, j);
#pragma unroll
        for (int t = 0; t < 5; ++t) tmpl.col[t] = tmpl.col[t] >= 0 ? tmpl.col[t] - tmpl.e : (1 << 30);
    }
    auto elem_at = [&](int i) -> GsElem {
        if (i > 0 && i < Ni - 1) {
            GsElem E;
            E.e = j * Ni + i;
            E.n = tmpl.n;
            E.tdiag = tmpl.tdiag;
#pragma unroll
            for (int t = 0; t < 5; ++t) E.col[t] = t < tmpl.n ? E.e + tmpl.col[t] : -1;
            return E;
        }
        return gs_elem<B>(S_, i, j);
    };
    const char *gbytes = reinterpret_cast<const char *>(gs);
    auto issue = [&](int idx) {     // lane 0: bulk copy of element idx's block row into its stage
        const int i = dir > 0 ? idx : Ni - 1 - idx;
        const size_t byte0 = (size_t)k0_of(i) * B2 * 8, byte1 = byte0 + (size_t)cnt_of(i) * B2 * 8;
        const size_t a0 = byte0 & ~(size_t)15, a1 = (byte1 + 15) & ~(size_t)15;
        const int s = idx % S;
        mbar_expect_tx(&full[s], (uint32_t)(a1 - a0));
        bulk_g2s(wbase + (size_t)s * C::STAGE_D, gbytes + a0, (uint32_t)(a1 - a0), &full[s]);
    };
    if (lane == 0)
        for (int idx = 0; idx < S && idx < Ni; ++idx) issue(idx);
    const int jn = j + dir;                                  // next row in sweep order (old values)
    const bool next_row_ok = jn >= 0 && jn < Nj;
    // previous row holds OLD values when this row starts a sweep inside a slab (ghost row of the neighbour slab)
    const bool prev_row_old = (sr == 0) && (j - dir >= 0) && (j - dir < Nj);
    constexpr int NV = C::NV;
    // cp.async prefetch of the sequential vector streams for sweep index n (one commit per call)
    auto prefetch_vectors = [&](int n) {
        if (n < Ni) {
            const int i = dir > 0 ? n : Ni - 1 - n;
            const int e = j * Ni + i;
            double *slot = vring + (size_t)(n % (D + 1)) * NV * BP;
            const int iside = i + dir;
#pragma unroll
            for (int sl = 0; sl < RS; ++sl) {
                const int c = lane + 32 * sl;
                if (c < B) {
                    cp_async8(slot + c, rhs + (size_t)e * B + c);
                    if (iside >= 0 && iside < Ni) cp_async8(slot + BP + c, x + (size_t)(e + dir) * B + c);
                    if (next_row_ok) cp_async8(slot + 2 * BP + c, x + (size_t)(e + dir * Ni) * B + c);
                    if (prev_row_old) cp_async8(slot + 3 * BP + c, x + (size_t)(e - dir * Ni) * B + c);
                }
            }
        }
        cp_async_commit();
    };
    for (int n = 0; n < D; ++n) prefetch_vectors(n);

    // lane -> (scalar row, column range)
    const int r0 = P > 1 ? lane / P : lane;
    const int part = P > 1 ? lane % P : 0;
    const int c0 = part * CW, c1 = min(B, c0 + CW);
    const int q = (lane & 15) / C::PD;
    // hand-over topology
    const int pred = sr == 0 ? 0 : (w > 0 ? 1 : 2);                       // 0 none, 1 smem ring, 2 global mailbox
    const int succ = sr == nrows - 1 ? 0 : (w < W - 1 ? 1 : 2);
    const int pred_off = -dir * Ni;                                        // element offset to the predecessor row
    const double sentinel = __longlong_as_double(-1LL);

    double PV[RS];
    auto pred_fetch_global = [&](int e_pred, double (&pv)[RS]) {
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) {
            const int c = lane + 32 * sl;
            pv[sl] = (c < B) ? __ldcg(mbox + (size_t)e_pred * B + c) : 0.0;
        }
    };
    auto pred_ready_global = [&](int e_pred, double (&pv)[RS]) -> bool {
        bool mine = true;
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) {
            const int c = lane + 32 * sl;
            if (c < B && is_sentinel(pv[sl])) mine = false;
        }
        const bool ok = __all_sync(0xffffffffu, mine);
        if (ok) {
#pragma unroll
            for (int sl = 0; sl < RS; ++sl) {
                const int c = lane + 32 * sl;
                if (c < B) __stcg(mbox + (size_t)e_pred * B + c, sentinel);      // hand the slot back
            }
        }
        return ok;
    };

    // ---- interior pattern of this row: which sorted block position holds which neighbour ----
    int tP = -1, tV = -1, tS = -1, tN = -1, tD = 0;
    bool fast_ok = Ni > 2;
    if (Ni > 2) {
        tD = tmpl.tdiag;
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            if (t < tmpl.n && t != tD) {
                const int off = tmpl.col[t];
                if (off == -dir) tV = t;                                   // previous element (new value)
                else if ((pred != 0 || prev_row_old) && off == pred_off) tP = t;   // previous row (new value, or ghost)
                else if (off == dir) tS = t;                               // next element of the row (old)
                else if (next_row_ok && off == dir * Ni) tN = t;           // next row (old)
                else fast_ok = false;                                      // periodic wrap: generic path
            }
        }
    }
    const uint32_t full_a = smem_u32(full), hand_a = smem_u32(hand), handp_a = smem_u32(hand_pred);
    const uint32_t stage_a = smem_u32(wbase);
    auto wait_a = [&](uint32_t bar, uint32_t parity) -> bool {
        for (int spin = 0; spin < kSpinLimit; ++spin) {
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            if (ok) return true;
            if ((spin & 1023) == 1023 && *(volatile int *)err != 0) return false;
        }
        atomicExch(err, 1);
        return false;
    };
    // running TMA source offset of the element to be issued next (sweep index idx + S)
    auto issue_a = [&](int idx) {
        const int i = dir > 0 ? idx : Ni - 1 - idx;
        const size_t byte0 = (size_t)k0_of(i) * B2 * 8, byte1 = byte0 + (size_t)cnt_of(i) * B2 * 8;
        const size_t a0 = byte0 & ~(size_t)15, a1 = (byte1 + 15) & ~(size_t)15;
        const uint32_t bar = full_a + 8u * (uint32_t)(idx % S);
        const uint32_t dst = stage_a + (uint32_t)((idx % S) * C::STAGE_D * 8);
        const uint32_t bytes = (uint32_t)(a1 - a0);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(gbytes + a0), "r"(bytes), "r"(bar)
                     : "memory");
    };

    GsElem cur = elem_at(dir > 0 ? 0 : Ni - 1);
    int e_prev = -1;
    bool have_pv = false;
    auto generic_step = [&](int idx) -> bool {
        const int i = dir > 0 ? idx : Ni - 1 - idx;
        constexpr bool fast = false;
        const int e = j * Ni + i;
        // ---- predecessor-row value ----
        if (pred == 1) {
            if (!wait_a(handp_a + 8u * (uint32_t)(idx % RING), (idx / RING) & 1)) return false;
        } else if (pred == 2 && !have_pv) {
            int spin = 0;
            for (;;) {
                pred_fetch_global(e + pred_off, PV);
                if (pred_ready_global(e + pred_off, PV)) break;
                if (++spin > kSpinLimit || ((spin & 63) == 63 && *(volatile int *)err != 0)) {
                    if (lane == 0) atomicExch(err, 2);
                    return false;
                }
            }
        }
        // ---- vectors of this element have landed (cp.async groups complete in order) ----
        cp_async_wait<D - 1>();
        const double *vslot = vring + (size_t)(idx % (D + 1)) * NV * BP;
        if (pred == 2) {
#pragma unroll
            for (int sl = 0; sl < RS; ++sl) {
                const int c = lane + 32 * sl;
                if (c < B) pvs[c] = PV[sl];
            }
        }
        const double *pvec = pred == 1 ? ring_pred + (idx % RING) * BP : (pred == 2 ? pvs : vslot + 3 * BP);
        if (!fast) {
            // neighbour vectors that are neither streamed nor handed over (periodic wraps): load now
            const int e_pred = (pred || prev_row_old) ? e + pred_off : -2;
            const int e_side = (i + dir >= 0 && i + dir < Ni) ? e + dir : -3;
            const int e_next = next_row_ok ? e + dir * Ni : -4;
#pragma unroll
            for (int t = 0; t < 5; ++t) {
                const int col = cur.col[t];
                if (t < cur.n && col != e && col != e_prev && col != e_pred && col != e_side && col != e_next) {
#pragma unroll
                    for (int sl = 0; sl < RS; ++sl) {
                        const int c = lane + 32 * sl;
                        if (c < B) vs[t * BP + c] = __ldcg(x + (size_t)col * B + c);
                    }
                }
            }
        }
        const int s = idx % S;
        if (!wait_a(full_a + 8u * (uint32_t)s, (idx / S) & 1)) return false;
        const double *st = wbase + (size_t)s * C::STAGE_D + (int)((k0_of(i) * B2) & 1);
        __syncwarp();
        // ---- early, non-blocking poll of the mailbox for the next element ----
        double PVn[RS];
        bool polled = false;
        if (pred == 2 && idx + 1 < Ni) {
            pred_fetch_global(e + dir + pred_off, PVn);      // evaluated after the arithmetic below
            polled = true;
        }
        // ---- phase A: acc_r = sum over off-diagonal blocks of A[t][r][:] . x_col(t) ----
        double acc[RS], acc2 = 0.0;
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) acc[sl] = 0.0;
        int tdiag = tD;
        if (fast) {
            if (P > 1) {
                if (r0 < B) {
                    const double *Ar = st + r0 * B;
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                    if (tV >= 0)
                        for (int c = c0; c < c1; ++c) a0 = fma(Ar[tV * B2 + c], xprev[c], a0);
                    if (tP >= 0)
                        for (int c = c0; c < c1; ++c) a1 = fma(Ar[tP * B2 + c], pvec[c], a1);
                    if (tS >= 0)
                        for (int c = c0; c < c1; ++c) a2 = fma(Ar[tS * B2 + c], vslot[BP + c], a2);
                    if (tN >= 0)
                        for (int c = c0; c < c1; ++c) a3 = fma(Ar[tN * B2 + c], vslot[2 * BP + c], a3);
                    acc[0] = (a1 + a2) + a3;
                    acc2 = a0;
                }
            } else {
#pragma unroll
                for (int sl = 0; sl < RS; ++sl) {
                    const int r = lane + 32 * sl;
                    if (r < B) {
                        const double *Ar = st + r * B;
                        double a = 0.0;
                        if (tP >= 0) a += skew_dot<B>(Ar + tP * B2, pvec, q);
                        if (tS >= 0) a += skew_dot<B>(Ar + tS * B2, vslot + BP, q);
                        if (tN >= 0) a += skew_dot<B>(Ar + tN * B2, vslot + 2 * BP, q);
                        if (tV >= 0) a += skew_dot<B>(Ar + tV * B2, xprev, q);
                        acc[sl] = a;
                    }
                }
            }
        } else {
            tdiag = cur.tdiag;
            const int e_pred = (pred || prev_row_old) ? e + pred_off : -2;
            const int e_side = (i + dir >= 0 && i + dir < Ni) ? e + dir : -3;
            const int e_next = next_row_ok ? e + dir * Ni : -4;
#pragma unroll
            for (int t = 0; t < 5; ++t) {
                if (t >= cur.n || t == cur.tdiag) continue;
                const int col = cur.col[t];
                const double *v = (col == e_prev) ? xprev
                                : (col == e_pred) ? pvec
                                : (col == e_side) ? vslot + BP
                                : (col == e_next) ? vslot + 2 * BP
                                                  : vs + t * BP;
                const double *A = st + t * B2;
                if (P > 1) {
                    if (r0 < B)
                        for (int c = c0; c < c1; ++c) acc[0] = fma(A[r0 * B + c], v[c], acc[0]);
                } else {
#pragma unroll
                    for (int sl = 0; sl < RS; ++sl) {
                        const int r = lane + 32 * sl;
                        if (r < B) acc[sl] += skew_dot<B>(A + r * B, v, q);
                    }
                }
            }
        }
        if (P > 1) {
            const double mine = acc[0] + acc2;
            acc[0] = mine;
#pragma unroll
            for (int o = 1; o < P; ++o) acc[0] += __shfl_down_sync(0xffffffffu, mine, o);
        }
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) {
            const int r = P > 1 ? r0 : lane + 32 * sl;
            const bool own = (P > 1) ? (part == 0 && r < B) : (r < B);
            if (own) rsv[r] = vslot[r] - acc[sl];
        }
        __syncwarp();
        // ---- phase B: x_i = Dinv_i * rsum ----
        const double *Dm = st + tdiag * B2;
        double xn[RS];
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) xn[sl] = 0.0;
        if (P > 1) {
            if (r0 < B)
                for (int c = c0; c < c1; ++c) xn[0] = fma(Dm[r0 * B + c], rsv[c], xn[0]);
            const double mine = xn[0];
#pragma unroll
            for (int o = 1; o < P; ++o) xn[0] += __shfl_down_sync(0xffffffffu, mine, o);
        } else {
#pragma unroll
            for (int sl = 0; sl < RS; ++sl) {
                const int r = lane + 32 * sl;
                if (r < B) xn[sl] = skew_dot<B>(Dm + r * B, rsv, q);
            }
        }
        // ---- hand the result over: x, own scratch, successor row ----
        if (succ == 1 && s_prog[w + 1] < idx - RING + 1) {   // ring slot not yet consumed by warp w+1
            int spin = 0;
            while (s_prog[w + 1] < idx - RING + 1) {
                __nanosleep(40);
                if (++spin > kSpinLimit) { if (lane == 0) atomicExch(err, 2); return false; }
            }
        }
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) {
            const int r = P > 1 ? r0 : lane + 32 * sl;
            const bool own = (P > 1) ? (part == 0 && r < B) : (r < B);
            if (own) {
                double v = xn[sl];
                if (omega != 1.0) v = omega * v + (1.0 - omega) * __ldcg(x + (size_t)e * B + r);
                __stcg(x + (size_t)e * B + r, v);
                xprev[r] = v;
                if (succ == 1) ring[(idx % RING) * BP + r] = v;
                if (succ == 2) __stcg(mbox + (size_t)e * B + r, v);
            }
        }
        __syncwarp();
        if (lane == 0) {
            if (succ == 1)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hand_a + 8u * (uint32_t)(idx % RING)) : "memory");
            s_prog[w] = idx + 1;
            if (idx + S < Ni) {
                fence_proxy_async();
                issue_a(idx + S);
            }
        }
        prefetch_vectors(idx + D);       // reuses the ring slot of element idx - 1
        have_pv = polled ? pred_ready_global(e + dir + pred_off, PVn) : false;
        // ---- rotate ----
        e_prev = e;
        if (idx + 1 < Ni && !(fast_ok && idx + 1 >= 1 && idx + 1 <= Ni - 2)) cur = elem_at(dir > 0 ? idx + 1 : Ni - 2 - idx);
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) PV[sl] = PVn[sl];
            return true;
    };

    // ---- sweep: first element (generic), interior elements (lean loop), last element (generic) ----
    int idx = 0;
    if (!generic_step(0)) return;
    idx = 1;
    if (fast_ok) {
        // loop-carried counters / running pointers instead of idx % , idx / and address arithmetic
        int s = 1 % S, ph = (1 / S) & 1;                   // TMA stage and its parity
        int rg = 1 % RING, rph = (1 / RING) & 1;           // hand-over ring slot and parity
        int vi = 1 % (D + 1);                              // vector ring slot of the current element
        int vp = (1 + D) % (D + 1);                        // vector ring slot being prefetched (element idx + D)
        int e = j * Ni + (dir > 0 ? 1 : Ni - 2);
        int shift = (int)((k0_of(dir > 0 ? 1 : Ni - 2) * B2) & 1);
        const int flip = (n_int & B2) & 1;
        const int oV = tV * B2, oP = tP * B2, oS = tS * B2, oN = tN * B2, oD = tD * B2;
        const int rowoff = (P > 1 ? r0 : 0) * B;
        const bool rowok = P > 1 ? (r0 < B) : true;
        const bool own1 = P > 1 && part == 0 && r0 < B;
        // vector streams of element idx + D (lane c < B moves entry c; RS == 1 in the lean path for B <= 32)
        const int step = dir * B;
        const double *p_rhs = rhs + (size_t)(j * Ni + (dir > 0 ? 1 + D : Ni - 2 - D)) * B + lane;
        const double *p_side = x + (size_t)(j * Ni + (dir > 0 ? 1 + D : Ni - 2 - D) + dir) * B + lane;
        const double *p_next = x + (size_t)(j * Ni + (dir > 0 ? 1 + D : Ni - 2 - D) + dir * Ni) * B + lane;
        const double *p_pold = x + ((long long)(j * Ni + (dir > 0 ? 1 + D : Ni - 2 - D)) - (long long)dir * Ni) * B + lane;
        // TMA source of element idx + S
        long long k_issue = k0_of(dir > 0 ? 1 + S : Ni - 2 - S);      // only meaningful while idx + S < Ni
        for (; idx <= Ni - 2; ++idx) {
            // predecessor-row value
            if (pred == 1) {
                if (!wait_a(handp_a + 8u * (uint32_t)rg, (uint32_t)rph)) return;
            } else if (pred == 2 && !have_pv) {
                int spin = 0;
                for (;;) {
                    pred_fetch_global(e + pred_off, PV);
                    if (pred_ready_global(e + pred_off, PV)) break;
                    if (++spin > kSpinLimit || ((spin & 63) == 63 && *(volatile int *)err != 0)) {
                        if (lane == 0) atomicExch(err, 2);
                        return;
                    }
                }
            }
            cp_async_wait<D - 1>();
            const double *vslot = vring + vi * (NV * BP);
            if (pred == 2) {
#pragma unroll
                for (int sl = 0; sl < RS; ++sl) {
                    const int c = lane + 32 * sl;
                    if (c < B) pvs[c] = PV[sl];
                }
            }
            const double *pvec = pred == 1 ? ring_pred + rg * BP : (pred == 2 ? pvs : vslot + 3 * BP);
            if (!wait_a(full_a + 8u * (uint32_t)s, (uint32_t)ph)) return;
            const double *st = wbase + s * C::STAGE_D + shift;
            __syncwarp();
            double PVn[RS];
            if (pred == 2) pred_fetch_global(e + dir + pred_off, PVn);     // idx + 1 <= Ni - 1 always exists
            // phase A
            double acc[RS];
            if (P > 1) {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                if (rowok) {
                    const double *Ar = st + rowoff;
                    if (tV >= 0) {
#pragma unroll
                        for (int c = 0; c < CW; ++c) if (c0 + c < B) a0 = fma(Ar[oV + c0 + c], xprev[c0 + c], a0);
                    }
                    if (tP >= 0) {
#pragma unroll
                        for (int c = 0; c < CW; ++c) if (c0 + c < B) a1 = fma(Ar[oP + c0 + c], pvec[c0 + c], a1);
                    }
#pragma unroll
                    for (int c = 0; c < CW; ++c) if (c0 + c < B) a2 = fma(Ar[oS + c0 + c], vslot[BP + c0 + c], a2);
                    if (tN >= 0) {
#pragma unroll
                        for (int c = 0; c < CW; ++c) if (c0 + c < B) a3 = fma(Ar[oN + c0 + c], vslot[2 * BP + c0 + c], a3);
                    }
                }
                const double mine = (a1 + a2) + (a3 + a0);
                acc[0] = mine;
#pragma unroll
                for (int o = 1; o < P; ++o) acc[0] += __shfl_down_sync(0xffffffffu, mine, o);
                if (own1) rsv[r0] = vslot[r0] - acc[0];
            } else {
#pragma unroll
                for (int sl = 0; sl < RS; ++sl) {
                    const int r = lane + 32 * sl;
                    if (r < B) {
                        const double *Ar = st + r * B;
                        double a = skew_dot<B>(Ar + oS, vslot + BP, q);
                        if (tP >= 0) a += skew_dot<B>(Ar + oP, pvec, q);
                        if (tN >= 0) a += skew_dot<B>(Ar + oN, vslot + 2 * BP, q);
                        if (tV >= 0) a += skew_dot<B>(Ar + oV, xprev, q);
                        rsv[r] = vslot[r] - a;
                    }
                }
            }
            __syncwarp();
            // phase B
            double xn[RS];
            if (P > 1) {
                double v = 0.0;
                if (rowok) {
                    const double *Dr = st + oD + rowoff;
#pragma unroll
                    for (int c = 0; c < CW; ++c) if (c0 + c < B) v = fma(Dr[c0 + c], rsv[c0 + c], v);
                }
                xn[0] = v;
#pragma unroll
                for (int o = 1; o < P; ++o) xn[0] += __shfl_down_sync(0xffffffffu, v, o);
            } else {
#pragma unroll
                for (int sl = 0; sl < RS; ++sl) {
                    const int r = lane + 32 * sl;
                    xn[sl] = (r < B) ? skew_dot<B>(st + oD + r * B, rsv, q) : 0.0;
                }
            }
            if (succ == 1 && s_prog[w + 1] < idx - RING + 1) {
                int spin = 0;
                while (s_prog[w + 1] < idx - RING + 1) {
                    __nanosleep(40);
                    if (++spin > kSpinLimit) { if (lane == 0) atomicExch(err, 2); return; }
                }
            }
#pragma unroll
            for (int sl = 0; sl < RS; ++sl) {
                const int r = P > 1 ? r0 : lane + 32 * sl;
                const bool own = (P > 1) ? own1 : (r < B);
                if (own) {
                    double v = xn[sl];
                    if (omega != 1.0) v = omega * v + (1.0 - omega) * __ldcg(x + (size_t)e * B + r);
                    __stcg(x + (size_t)e * B + r, v);
                    xprev[r] = v;
                    if (succ == 1) ring[rg * BP + r] = v;
                    if (succ == 2) __stcg(mbox + (size_t)e * B + r, v);
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (succ == 1)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hand_a + 8u * (uint32_t)rg) : "memory");
                s_prog[w] = idx + 1;
                if (idx + S < Ni) {
                    // element idx + S: interior unless it is the last of the sweep
                    const int cnt = (idx + S == Ni - 1) ? (dir > 0 ? n_last : n_first) : n_int;
                    const size_t byte0 = (size_t)k_issue * B2 * 8, byte1 = byte0 + (size_t)cnt * B2 * 8;
                    const size_t a0 = byte0 & ~(size_t)15, a1 = (byte1 + 15) & ~(size_t)15;
                    const uint32_t bar = full_a + 8u * (uint32_t)s;          // same stage as the one just consumed
                    const uint32_t dst = stage_a + (uint32_t)(s * C::STAGE_D * 8);
                    const uint32_t bytes = (uint32_t)(a1 - a0);
                    fence_proxy_async();
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                        "l"(gbytes + a0), "r"(bytes), "r"(bar)
                        : "memory");
                }
            }
            // next TMA source: fwd: += count of the element just issued; bwd: -= count of the one before it
            if (dir > 0) k_issue += (idx + S == Ni - 1) ? n_last : n_int;
            else k_issue -= (idx + S + 1 == Ni - 1) ? n_first : n_int;
            // prefetch the vector streams of element idx + D into the slot of element idx - 1
            if (idx + D < Ni) {
                double *slot = vring + vp * (NV * BP);
                if (RS == 1) {
                    if (lane < B) {
                        cp_async8(slot + lane, p_rhs);
                        if (idx + D < Ni - 1) cp_async8(slot + BP + lane, p_side);     // not for the last element of the row
                        if (next_row_ok) cp_async8(slot + 2 * BP + lane, p_next);
                        if (prev_row_old) cp_async8(slot + 3 * BP + lane, p_pold);
                    }
                } else {
#pragma unroll
                    for (int sl = 0; sl < RS; ++sl) {
                        const int c = lane + 32 * sl;
                        if (c < B) {
                            cp_async8(slot + c, p_rhs + 32 * sl);
                            if (idx + D < Ni - 1) cp_async8(slot + BP + c, p_side + 32 * sl);
                            if (next_row_ok) cp_async8(slot + 2 * BP + c, p_next + 32 * sl);
                            if (prev_row_old) cp_async8(slot + 3 * BP + c, p_pold + 32 * sl);
                        }
                    }
                }
            }
            cp_async_commit();
            p_rhs += step; p_side += step; p_next += step; p_pold += step;
            if (pred == 2) {
                have_pv = pred_ready_global(e + dir + pred_off, PVn);
#pragma unroll
                for (int sl = 0; sl < RS; ++sl) PV[sl] = PVn[sl];
            }
            // advance the counters
            e_prev = e;
            e += dir;
            shift ^= flip;
            if (++s == S) { s = 0; ph ^= 1; }
            if (++rg == RING) { rg = 0; rph ^= 1; }
            if (++vi == D + 1) vi = 0;
            if (++vp == D + 1) vp = 0;
        }
        if (idx < Ni) cur = elem_at(dir > 0 ? idx : Ni - 1 - idx);
    }
    for (; idx < Ni; ++idx)
        if (!generic_step(idx)) return;
    cp_async_wait<0>();
}

// structure check: BSR (indices, indptr) == closed-form 5-point stencil?
__global__ void __launch_bounds__(256)
k_check_stencil(const int32_t *__restrict__ indices, const int32_t *__restrict__ indptr, Stencil S_,
                int32_t *mismatch) {
    const int N = S_.Ni * S_.Nj;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < N; e += gridDim.x * blockDim.x) {
        const int i = e % S_.Ni, j = e / S_.Ni;
        int c[5], rk[5];
        S_.cols(i, j, c);
        slot_ranks(c, rk);
        const long long k0 = S_.row_start(i, j);
        bool bad = indptr[e] != (int32_t)k0 || indptr[e + 1] - indptr[e] != S_.count(i, j);
        if (!bad && S_.active(j))           // ghost rows of a slab have no blocks
            for (int s = 0; s < 5; ++s)
                if (rk[s] >= 0 && indices[k0 + rk[s]] != c[s]) bad = true;
        if (bad) atomicAdd(mismatch, 1);
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ticket word and error flag of the single-launch smoother kernels: one pair per device, allocated on first
// use on that device.  The ticket is reset (stream-ordered) before every launch, so smoother launches of one
// device must be issued on one stream at a time (include/dgb200.h, "Conventions").
constexpr int kMaxDevices = 64;
struct DevState {
    int *work = nullptr;   // [0] ticket
    int *err = nullptr;    // device-side error flag of the asynchronous kernels
};
static DevState g_dev[kMaxDevices];
int g_kernel_path = 0;             // 0 auto (single-launch smoother kernels where available), 1 generic only

static DevState *dev_state() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    return &g_dev[dev];
}
int *work_ptr() { DevState *d = dev_state(); return d ? d->work : nullptr; }
int *err_ptr() { DevState *d = dev_state(); return d ? d->err : nullptr; }
int ensure_work(int) {
    DevState *d = dev_state();
    if (d == nullptr) {
        set_error("ensure_work: no current CUDA device");
        return -1;
    }
    if (d->err == nullptr) {
        DGB_CUDA_OK(cudaMalloc(&d->err, sizeof(int)));
        DGB_CUDA_OK(cudaMemset(d->err, 0, sizeof(int)));
    }
    if (d->work == nullptr) {
        DGB_CUDA_OK(cudaMalloc(&d->work, sizeof(int) * 64));
        DGB_CUDA_OK(cudaMemset(d->work, 0, sizeof(int) * 64));
    }
    return 0;
}

bool stream_supported(int b) { return b == 1 || b == 4 || b == 9 || b == 16 || b == 22 || b == 25 || b == 36; }

template <int B, int WW, int SS>
static int gs_rows_launch_c(const double *gs, const double *rhs, double *x, double *mbox, Stencil S_, int dir,
                            double omega, const int32_t *skip, cudaStream_t st) {
    using C = GsCfg<B, WW, SS>;
    static bool configured = false;
    if (!configured) {
        DGB_CUDA_OK(cudaFuncSetAttribute(k_gs_rows<B, WW, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::SMEM));
        configured = true;
    }
    DGB_CUDA_OK(cudaMemsetAsync(work_ptr(), 0, sizeof(int), st));
    const int grid = (S_.ja1 - S_.ja0 + C::W - 1) / C::W;
    k_gs_rows<B, WW, SS><<<grid, C::W * 32, C::SMEM, st>>>(gs, rhs, x, mbox, S_, dir, omega, work_ptr(), err_ptr(), skip);
    DGB_LAUNCH_OK();
    return 0;
}
int g_gs_variant = 0;   // experiment switch (DGB_GS_VARIANT)
extern int g_chain_mask;
extern int g_chain_cluster;
template <int B>
static int gs_rows_launch_t(const double *gs, const double *rhs, double *x, double *mbox, Stencil S_, int dir,
                            double omega, const int32_t *skip, cudaStream_t st) {
    if (B == 9 && g_gs_variant == 1) return gs_rows_launch_c<9, 8, 6>(gs, rhs, x, mbox, S_, dir, omega, skip, st);
    if (B == 9 && g_gs_variant == 2) return gs_rows_launch_c<9, 4, 12>(gs, rhs, x, mbox, S_, dir, omega, skip, st);
    return gs_rows_launch_c<B, GsDefault<B>::W, GsDefault<B>::S>(gs, rhs, x, mbox, S_, dir, omega, skip, st);
}

int gs_rows_launch(int b, const double *gs, const double *rhs, double *x, double *mbox, int Ni, int Nj, int flags,
                   int dir, double omega, const int32_t *skip, cudaStream_t st) {
    int rc = ensure_work(0);
    if (rc) return rc;
    Stencil S_ = make_stencil(Ni, Nj, flags);
    DGB_DISPATCH_B(b, return (gs_rows_launch_t<B>(gs, rhs, x, mbox, S_, dir, omega, skip, st)));
    return 0;
}

}  // namespace dgb

using namespace dgb;

extern "C" {

int dgb_set_kernel_path(int32_t path) {
    const int old = g_kernel_path;
    if (path == 0 || path == 1) g_kernel_path = path;
    if (path >= 100 && path < 200) g_gs_variant = path - 100;      // tuning experiments only
    if (path >= 300 && path < 332) g_chain_mask = path - 300;      // block sizes of the chained GS kernel
    if (path >= 400 && path <= 416) g_chain_cluster = path - 400;  // CTAs per cluster of the chained GS kernel
    return old;
}

int dgb_device_error(int32_t reset) {
    int *err = err_ptr();
    if (err == nullptr) return 0;
    int v = 0;
    if (cudaMemcpy(&v, err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (reset && v != 0) cudaMemset(err, 0, sizeof(int));
    return v;
}

int dgb_check_stencil(const int32_t *indices, const int32_t *indptr, int32_t Ni, int32_t Nj, int32_t flags,
                      int32_t *mismatch, void *stream) {
    DGB_ARG(indices && indptr && mismatch && Ni > 0 && Nj > 0 && flags >= 0);
    cudaStream_t st = (cudaStream_t)stream;
    Stencil S_ = make_stencil(Ni, Nj, flags);
    DGB_CUDA_OK(cudaMemsetAsync(mismatch, 0, sizeof(int32_t), st));
    int g = (Ni * Nj + 255) / 256;
    if (g > sm_count() * 8) g = sm_count() * 8;
    k_check_stencil<<<g, 256, 0, st>>>(indices, indptr, S_, mismatch);
    DGB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
