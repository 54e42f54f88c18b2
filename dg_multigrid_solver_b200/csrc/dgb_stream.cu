// dgb_stream.cu -- the HBM-streaming kernels of the V-cycle, sm_100a.
//
//  k_stream<B,MODE>  persistent CTAs; one producer warp feeds a ring of shared-memory stages with
//                    TMA bulk copies (cp.async.bulk + mbarrier complete_tx) of whole block rows,
//                    eight consumer warps do block-row x vector products out of shared memory.
//                    MODE: apply y=Ax | residual r=b-Ax with fused sum(r^2) | relaxation
//                    (block-Jacobi, or one colour of the red-black block-GS, in place).
//  k_gs_rows<B>      lexicographic block Gauss-Seidel, exact sweep order: one warp per element
//                    row j, rows pipelined against each other through release/acquire progress
//                    counters in global memory (row j may process element i once row j-1 has
//                    finished element i); each warp streams its row's blocks through its own
//                    TMA-fed ring.  CTAs take tickets so that a CTA only ever waits on CTAs
//                    that were started before it.
//
// Reference semantics: scipy bsr_matvec (dgfem/solver.py:117,119,150), pyamg
// amg_core.block_gauss_seidel (dgfem/pyamg_relaxation.py:252-255), dgfem/relaxation.py:123-195.
// The smoother kernels read the "GS stream": the BSR data with each diagonal block replaced by
// its inverse (dgb_build_gs_stream), so one pass reads exactly nnzb blocks.
#include "dgb_async.cuh"
#include "dgb_common.cuh"

namespace dgb {

constexpr int gcd_c(int a, int b) { return b == 0 ? a : gcd_c(b, a % b); }

enum { S_APPLY = 0, S_RESIDUAL = 1, S_RELAX = 2 };

// =========================================================================================
// k_stream
// =========================================================================================
template <int B>
struct StreamCfg {
    static constexpr int B2 = B * B;
    static constexpr int T = B <= 4 ? 32 : B <= 9 ? 8 : B <= 16 ? 4 : B <= 25 ? 2 : 1;   // block rows per stage
    static constexpr int S = B <= 4 ? 4 : 3;                                            // stages
    static constexpr int MAXBR = 5;                                                     // blocks per row
    static constexpr int ROWSLOT = (MAXBR * B2 + 2 + 1) & ~1;                            // doubles, even
    static constexpr int STAGE_D = T * ROWSLOT;
    static constexpr int NCW = 8, NC = NCW * 32, NT = NC + 32;
    static constexpr int PD = 16 / gcd_c(B, 16);          // period of (B*t mod 16): bank skew, see below
    // dynamic shared memory layout (bytes)
    static constexpr size_t oStage = 0;
    static constexpr size_t oPartial = oStage + sizeof(double) * S * STAGE_D;
    static constexpr size_t oRsum = oPartial + sizeof(double) * 2 * T * MAXBR * B;
    static constexpr size_t oBar = oRsum + sizeof(double) * 2 * T * B;                   // full[S], empty[S]
    static constexpr size_t oInts = oBar + sizeof(uint64_t) * 2 * S;
    // ints: row_off[S][T], row_n[S][T], row_cix[S][T], cols[S][T*MAXBR], diag_t[2][T]
    static constexpr size_t nInts = 3 * S * T + S * T * MAXBR + 2 * T;
    static constexpr size_t SMEM = oInts + sizeof(int) * nInts;
};

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

template <int B, int MODE>
__global__ void __launch_bounds__(StreamCfg<B>::NT)
k_stream(const double *__restrict__ data, const int32_t *__restrict__ indices,
         const int32_t *__restrict__ indptr, int N, int Ni, const double *__restrict__ rhs,
         const double *x_in, double *x_out, double *partials, double omega, int colour, int *err,
         const int32_t *__restrict__ skip) {
    using C = StreamCfg<B>;
    constexpr int T = C::T, S = C::S, B2 = C::B2, NC = C::NC, MAXBR = C::MAXBR;
    if (skip != nullptr && *skip != 0) return;
    extern __shared__ __align__(128) unsigned char smem[];
    double *stage = reinterpret_cast<double *>(smem + C::oStage);
    double *partial = reinterpret_cast<double *>(smem + C::oPartial);
    double *rsum = reinterpret_cast<double *>(smem + C::oRsum);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::oBar);
    uint64_t *empty = full + S;
    int *row_off = reinterpret_cast<int *>(smem + C::oInts);
    int *row_n = row_off + S * T;
    int *row_cix = row_n + S * T;
    int *cols = row_cix + S * T;
    int *diag_t = cols + S * T * MAXBR;
    __shared__ double s_red[C::NCW];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], C::NCW);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();
    const int ntiles = (N + T - 1) / T;

    if (warp == 0) {
        // ------------------------------- producer warp ------------------------------------
        const char *gbytes = reinterpret_cast<const char *>(data);
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            bool ok = true;
            if (lane == 0) ok = mbar_wait(&empty[s], ph ^ 1, err);
            ok = __shfl_sync(0xffffffffu, ok, 0);
            if (!ok) return;
            const int e0 = tile * T;
            const int nrow = min(T, N - e0);
            int ip_lo = 0, ip_hi = 0;
            if (lane < nrow) {
                ip_lo = indptr[e0 + lane];
                ip_hi = indptr[e0 + lane + 1];
            }
            double *sbase = stage + (size_t)s * C::STAGE_D;
            if (colour < 0) {
                // one bulk copy for the whole tile (its blocks are contiguous in the BSR data)
                const int k0 = __shfl_sync(0xffffffffu, ip_lo, 0);
                const int k1 = __shfl_sync(0xffffffffu, ip_hi, nrow - 1);
                const size_t byte0 = (size_t)k0 * B2 * 8, byte1 = (size_t)k1 * B2 * 8;
                const size_t a0 = byte0 & ~(size_t)15, a1 = (byte1 + 15) & ~(size_t)15;
                const int shift = (int)((byte0 - a0) >> 3);
                if (lane < T) {
                    row_off[s * T + lane] = lane < nrow ? (ip_lo - k0) * B2 + shift : 0;
                    row_n[s * T + lane] = lane < nrow ? ip_hi - ip_lo : 0;
                    row_cix[s * T + lane] = ip_lo - k0;
                }
                for (int k = lane; k < k1 - k0; k += 32) cols[s * T * MAXBR + k] = indices[k0 + k];
                __syncwarp();
                if (lane == 0) {
                    mbar_expect_tx(&full[s], (uint32_t)(a1 - a0));
                    bulk_g2s(sbase, gbytes + a0, (uint32_t)(a1 - a0), &full[s]);
                }
            } else {
                // one colour class: one bulk copy per active block row into its own slot
                const int e = e0 + lane;
                const bool active = lane < nrow && ((((e % Ni) + (e / Ni)) & 1) == colour);
                const int nb = active ? ip_hi - ip_lo : 0;
                const size_t byte0 = (size_t)ip_lo * B2 * 8, byte1 = (size_t)ip_hi * B2 * 8;
                const size_t a0 = byte0 & ~(size_t)15, a1 = (byte1 + 15) & ~(size_t)15;
                const int shift = (int)((byte0 - a0) >> 3);
                uint32_t bytes = active ? (uint32_t)(a1 - a0) : 0u;
                if (lane < T) {
                    row_off[s * T + lane] = lane * C::ROWSLOT + shift;
                    row_n[s * T + lane] = nb;
                    row_cix[s * T + lane] = lane * MAXBR;
                }
                for (int t = 0; t < nb; ++t) cols[s * T * MAXBR + lane * MAXBR + t] = indices[ip_lo + t];
                uint32_t total = bytes;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                __syncwarp();
                if (lane == 0) mbar_expect_tx(&full[s], total);
                __syncwarp();
                if (active) bulk_g2s(sbase + (size_t)lane * C::ROWSLOT, gbytes + a0, bytes, &full[s]);
            }
        }
        return;
    }

    // ----------------------------------- consumer warps -------------------------------------
    const int ctid = tid - 32;
    const int q = (ctid & 15) / C::PD;
    double sumsq = 0.0;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        if (!mbar_wait(&full[s], ph, err)) return;
        const int e0 = tile * T;
        const int nrow = min(T, N - e0);
        const double *st = stage + (size_t)s * C::STAGE_D;
        const int pb = it & 1;
        double *part = partial + (size_t)pb * T * MAXBR * B;
        double *rs = rsum + (size_t)pb * T * B;
        // phase 1: one item = (row l, block t, scalar row r): dot of one block row with x[col]
        for (int item = ctid; item < nrow * MAXBR * B; item += NC) {
            const int l = item / (MAXBR * B);
            const int rem = item - l * (MAXBR * B);
            const int t = rem / B, r = rem - t * B;
            if (t < row_n[s * T + l]) {
                const int e = e0 + l;
                const int col = cols[s * T * MAXBR + row_cix[s * T + l] + t];
                double acc = 0.0;
                if (MODE == S_RELAX && col == e) {
                    if (r == 0) diag_t[pb * T + l] = t;
                } else {
                    acc = skew_dot<B>(st + row_off[s * T + l] + t * B2 + r * B, x_in + (size_t)col * B, q);
                }
                part[item] = acc;
            }
        }
        named_bar_sync<1, NC>();
        // phase 2: sum the row's blocks in stored (ascending column) order
        for (int item = ctid; item < nrow * B; item += NC) {
            const int l = item / B, r = item - l * B;
            const int n = row_n[s * T + l];
            if (n > 0) {
                double acc = 0.0;
                for (int t = 0; t < n; ++t) acc += part[(l * MAXBR + t) * B + r];
                const size_t o = (size_t)(e0 + l) * B + r;
                if (MODE == S_APPLY) {
                    x_out[o] = acc;
                } else if (MODE == S_RESIDUAL) {
                    const double res = rhs[o] - acc;
                    if (x_out != nullptr) x_out[o] = res;
                    sumsq = fma(res, res, sumsq);
                } else {
                    rs[item] = rhs[o] - acc;
                }
            }
        }
        if (MODE == S_RELAX) {
            named_bar_sync<1, NC>();
            // phase 3: x_i = omega * Dinv_i * rsum_i + (1 - omega) * x_i
            for (int item = ctid; item < nrow * B; item += NC) {
                const int l = item / B, r = item - l * B;
                if (row_n[s * T + l] > 0) {
                    const double *d = st + row_off[s * T + l] + diag_t[pb * T + l] * B2 + r * B;
                    const double xn = skew_dot<B>(d, rs + l * B, q);
                    const size_t o = (size_t)(e0 + l) * B + r;
                    x_out[o] = (omega == 1.0) ? xn : omega * xn + (1.0 - omega) * x_in[o];
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (MODE == S_RESIDUAL) {
        sumsq = warp_sum(sumsq);
        if (lane == 0) s_red[warp - 1] = sumsq;
        named_bar_sync<1, NC>();
        if (ctid == 0) {
            double t = 0.0;
            for (int w = 0; w < C::NCW; ++w) t += s_red[w];
            partials[blockIdx.x] = t;
        }
    }
}

// =========================================================================================
// k_gs_rows  (v2: no fences)
//
// Row j may process element i once row j-1 (in sweep order) has finished element i.  The value
// it needs from that row, x(i, j-1), is handed over
//   - through a shared-memory ring + progress counter when both rows live in the same CTA,
//   - through a global "mailbox" otherwise: the producer stores the 8-byte values themselves
//     (ld/st.cg, L2), the consumer spins until none of them is the sentinel (all-ones NaN) and
//     writes the sentinel back.  Data doubles as its own flag, so no release/acquire fence
//     (MEMBAR / CCTL.IVALL) sits on the critical path.  The mailbox is an [N*b] vector owned by
//     the level, all-sentinel outside a pass.
// =========================================================================================
template <int B>
struct GsCfg {
    static constexpr int B2 = B * B;
    static constexpr int P = (B == 9) ? 3 : (B == 4) ? 4 : 1;       // lanes per scalar row
    static constexpr int CW = (B + P - 1) / P;                      // columns per lane
    static constexpr int RS = (B + 31) / 32;                        // row slots per lane (P == 1)
    static constexpr int W = B <= 9 ? 16 : B <= 16 ? 4 : B <= 25 ? 2 : 1;   // rows (warps) per CTA
    static constexpr int S = 3;                                     // TMA ring stages per warp
    static constexpr int RING = 8;                                  // x hand-over ring slots per warp
    static constexpr int STAGE_D = (5 * B2 + 2 + 1) & ~1;           // doubles per stage, even
    static constexpr int BP = (B + 1) & ~1;
    static constexpr int PD = 16 / gcd_c(B, 16);
    static constexpr int WARP_D = S * STAGE_D + 7 * BP + RING * BP; // stages | vs[5] | xprev | rs | ring
    static constexpr size_t oBar = sizeof(double) * W * WARP_D;
    static constexpr size_t oProg = oBar + sizeof(uint64_t) * W * S;
    static constexpr size_t SMEM = oProg + sizeof(int) * W;
};

struct GsElem {
    int e, n, tdiag, shift;
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
    E.shift = (int)((S_.row_start(i, j) * GsCfg<B>::B2) & 1);
    return E;
}

__device__ __forceinline__ bool is_sentinel(double v) { return __double_as_longlong(v) == -1LL; }

// work[0] = ticket counter
template <int B>
__global__ void __launch_bounds__(GsCfg<B>::W * 32)
k_gs_rows(const double *__restrict__ gs, const double *__restrict__ rhs, double *x, double *mbox, Stencil S_,
          int dir, double omega, int *work, int *err, const int32_t *__restrict__ skip) {
    using C = GsCfg<B>;
    constexpr int B2 = C::B2, S = C::S, P = C::P, CW = C::CW, RS = C::RS, BP = C::BP, RING = C::RING, W = C::W;
    if (skip != nullptr && *skip != 0) return;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_ticket;
    volatile int *s_prog = reinterpret_cast<volatile int *>(smem + C::oProg);
    if (threadIdx.x == 0) s_ticket = atomicAdd(&work[0], 1);
    if (threadIdx.x < W) s_prog[threadIdx.x] = 0;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Ni = S_.Ni, Nj = S_.Nj;
    const int sr = s_ticket * W + w;           // row index in sweep order
    if (sr >= Nj) return;
    const int j = dir > 0 ? sr : Nj - 1 - sr;
    double *wbase = reinterpret_cast<double *>(smem) + (size_t)w * C::WARP_D;
    double *vs = wbase + S * C::STAGE_D;       // [5][BP]
    double *xprev = vs + 5 * BP;
    double *rsv = xprev + BP;
    double *ring = rsv + BP;                   // [RING][BP], written by this warp, read by warp w+1
    const double *ring_pred = ring - C::WARP_D;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::oBar) + w * S;
    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncwarp();
    const char *gbytes = reinterpret_cast<const char *>(gs);
    auto issue = [&](int idx) {     // lane 0: bulk copy of element idx's block row into its stage
        const int i = dir > 0 ? idx : Ni - 1 - idx;
        const long long k0 = S_.row_start(i, j);
        const int cnt = S_.count(i, j);
        const size_t byte0 = (size_t)k0 * B2 * 8, byte1 = byte0 + (size_t)cnt * B2 * 8;
        const size_t a0 = byte0 & ~(size_t)15, a1 = (byte1 + 15) & ~(size_t)15;
        const int s = idx % S;
        mbar_expect_tx(&full[s], (uint32_t)(a1 - a0));
        bulk_g2s(wbase + (size_t)s * C::STAGE_D, gbytes + a0, (uint32_t)(a1 - a0), &full[s]);
    };
    if (lane == 0)
        for (int idx = 0; idx < S && idx < Ni; ++idx) issue(idx);

    // lane -> (scalar row, column range)
    const int r0 = P > 1 ? lane / P : lane;
    const int part = P > 1 ? lane % P : 0;
    const int c0 = part * CW, c1 = min(B, c0 + CW);
    const int q = (lane & 15) / C::PD;
    // hand-over topology
    const int pred = sr == 0 ? 0 : (w > 0 ? 1 : 2);                       // 0 none, 1 smem ring, 2 global mailbox
    const int succ = sr == Nj - 1 ? 0 : (w < W - 1 ? 1 : 2);
    const int pred_off = -dir * Ni;                                        // element offset to the predecessor row
    const double sentinel = __longlong_as_double(-1LL);

    double V[5][RS], rhsv[RS], xold[RS], PV[RS];
    // neighbour vectors that do not come from the predecessor row (old values, or this row's own)
    auto load_vectors = [&](const GsElem &E, int e_prev, double (&Vv)[5][RS], double (&rh)[RS], double (&xo)[RS]) {
        const int e_pred = pred ? E.e + pred_off : -2;
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const int col = E.col[t];
#pragma unroll
            for (int sl = 0; sl < RS; ++sl) {
                const int c = lane + 32 * sl;
                Vv[t][sl] = 0.0;
                if (t < E.n && col != E.e && col != e_prev && col != e_pred && c < B)
                    Vv[t][sl] = __ldcg(x + (size_t)col * B + c);
            }
        }
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) {
            const int r = P > 1 ? r0 : lane + 32 * sl;
            const bool own = (P > 1) ? (part == 0 && r < B) : (r < B);
            rh[sl] = own ? rhs[(size_t)E.e * B + r] : 0.0;
            xo[sl] = (own && omega != 1.0) ? __ldcg(x + (size_t)E.e * B + r) : 0.0;
        }
    };
    // predecessor-row value for sweep index idx: raw fetch (may hold sentinels for the mailbox)
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
    auto pred_read_ring = [&](int idx, double (&pv)[RS]) {
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) {
            const int c = lane + 32 * sl;
            pv[sl] = (c < B) ? ring_pred[(idx % RING) * BP + c] : 0.0;
        }
    };

    GsElem cur = gs_elem<B>(S_, dir > 0 ? 0 : Ni - 1, j);
    int e_prev = -1;
    bool have_pv = false;
    load_vectors(cur, e_prev, V, rhsv, xold);
    for (int idx = 0; idx < Ni; ++idx) {
        // ---- predecessor-row value (blocking only if the prefetch below did not get it) ----
        if (pred != 0 && !have_pv) {
            int spin = 0;
            if (pred == 1) {
                while (s_prog[w - 1] < idx + 1) {
                    if (++spin > kSpinLimit) { if (lane == 0) atomicExch(err, 2); return; }
                }
                __threadfence_block();
                pred_read_ring(idx, PV);
            } else {
                for (;;) {
                    pred_fetch_global(cur.e + pred_off, PV);
                    if (pred_ready_global(cur.e + pred_off, PV)) break;
                    if (++spin > kSpinLimit || ((spin & 63) == 63 && *(volatile int *)err != 0)) {
                        if (lane == 0) atomicExch(err, 2);
                        return;
                    }
                }
            }
        }
        const int s = idx % S;
        if (!mbar_wait(&full[s], (idx / S) & 1, err)) return;
        const double *st = wbase + (size_t)s * C::STAGE_D + cur.shift;
        // ---- publish the neighbour vectors to this warp's scratch ----
        const int e_pred = pred ? cur.e + pred_off : -2;
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const int col = cur.col[t];
            if (t < cur.n && col != cur.e && col != e_prev) {
#pragma unroll
                for (int sl = 0; sl < RS; ++sl) {
                    const int c = lane + 32 * sl;
                    if (c < B) vs[t * BP + c] = (col == e_pred) ? PV[sl] : V[t][sl];
                }
            }
        }
        __syncwarp();
        // ---- prefetch for the next element ----
        GsElem nxt = cur;
        double Vn[5][RS], rhsn[RS], xoldn[RS], PVn[RS];
        bool have_next = false, polled = false;
        if (idx + 1 < Ni) {
            nxt = gs_elem<B>(S_, dir > 0 ? idx + 1 : Ni - 2 - idx, j);
            load_vectors(nxt, cur.e, Vn, rhsn, xoldn);
            if (pred == 1) {
                if (s_prog[w - 1] >= idx + 2) {
                    __threadfence_block();
                    pred_read_ring(idx + 1, PVn);
                    have_next = true;
                }
            } else if (pred == 2) {
                pred_fetch_global(nxt.e + pred_off, PVn);      // evaluated after the arithmetic below
                polled = true;
            }
        }
        // ---- phase A: acc_r = sum over off-diagonal blocks of A[t][r][:] . x_col(t) ----
        double acc[RS];
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) acc[sl] = 0.0;
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            if (t >= cur.n || t == cur.tdiag) continue;
            const double *v = (cur.col[t] == e_prev) ? xprev : vs + t * BP;
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
        if (P > 1) {
            const double mine = acc[0];
#pragma unroll
            for (int o = 1; o < P; ++o) acc[0] += __shfl_down_sync(0xffffffffu, mine, o);
        }
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) {
            const int r = P > 1 ? r0 : lane + 32 * sl;
            const bool own = (P > 1) ? (part == 0 && r < B) : (r < B);
            if (own) rsv[r] = rhsv[sl] - acc[sl];
        }
        __syncwarp();
        // ---- phase B: x_i = Dinv_i * rsum ----
        const double *D = st + cur.tdiag * B2;
        double xn[RS];
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) xn[sl] = 0.0;
        if (P > 1) {
            if (r0 < B)
                for (int c = c0; c < c1; ++c) xn[0] = fma(D[r0 * B + c], rsv[c], xn[0]);
            const double mine = xn[0];
#pragma unroll
            for (int o = 1; o < P; ++o) xn[0] += __shfl_down_sync(0xffffffffu, mine, o);
        } else {
#pragma unroll
            for (int sl = 0; sl < RS; ++sl) {
                const int r = lane + 32 * sl;
                if (r < B) xn[sl] = skew_dot<B>(D + r * B, rsv, q);
            }
        }
        // ---- hand the result over: x, own scratch, successor row ----
        if (succ == 1) {     // ring slot must have been consumed: warp w+1 finished element idx - RING
            int spin = 0;
            while (s_prog[w + 1] < idx - RING + 1) {
                if (++spin > kSpinLimit) { if (lane == 0) atomicExch(err, 2); return; }
            }
        }
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) {
            const int r = P > 1 ? r0 : lane + 32 * sl;
            const bool own = (P > 1) ? (part == 0 && r < B) : (r < B);
            if (own) {
                const double v = (omega == 1.0) ? xn[sl] : omega * xn[sl] + (1.0 - omega) * xold[sl];
                __stcg(x + (size_t)cur.e * B + r, v);
                xprev[r] = v;
                if (succ == 1) ring[(idx % RING) * BP + r] = v;
                if (succ == 2) __stcg(mbox + (size_t)cur.e * B + r, v);
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            s_prog[w] = idx + 1;
            if (idx + S < Ni) {
                fence_proxy_async();
                issue(idx + S);
            }
        }
        if (polled) have_next = pred_ready_global(nxt.e + pred_off, PVn);
        // ---- rotate ----
        e_prev = cur.e;
        cur = nxt;
        have_pv = have_next;
#pragma unroll
        for (int t = 0; t < 5; ++t)
#pragma unroll
            for (int sl = 0; sl < RS; ++sl) V[t][sl] = Vn[t][sl];
#pragma unroll
        for (int sl = 0; sl < RS; ++sl) {
            rhsv[sl] = rhsn[sl];
            xold[sl] = xoldn[sl];
            PV[sl] = PVn[sl];
        }
    }
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
        if (!bad)
            for (int s = 0; s < 5; ++s)
                if (rk[s] >= 0 && indices[k0 + rk[s]] != c[s]) bad = true;
        if (bad) atomicAdd(mismatch, 1);
    }
}

// ---------------------------------------------------------------------------------------
// host side
static int *g_work = nullptr;      // [0] ticket, [1..] progress
static int g_work_cap = 0;
static int *g_err = nullptr;       // device-side error flag of the async kernels
int g_kernel_path = 0;             // 0 auto (streaming kernels where available), 1 generic only

static int ensure_work(int n_rows) {
    if (g_err == nullptr) {
        DGB_CUDA_OK(cudaMalloc(&g_err, sizeof(int)));
        DGB_CUDA_OK(cudaMemset(g_err, 0, sizeof(int)));
    }
    if (n_rows + 2 > g_work_cap) {
        if (g_work) cudaFree(g_work);
        g_work_cap = n_rows + 2 + 4096;
        DGB_CUDA_OK(cudaMalloc(&g_work, sizeof(int) * g_work_cap));
    }
    return 0;
}

bool stream_supported(int b) { return b == 1 || b == 4 || b == 9 || b == 16 || b == 22 || b == 25 || b == 36; }

template <int B, int MODE>
static int stream_launch_t(const double *data, const int32_t *indices, const int32_t *indptr, int N, int Ni,
                           const double *rhs, const double *x_in, double *x_out, double *partials, double omega,
                           int colour, const int32_t *skip, cudaStream_t st, int *grid_out) {
    using C = StreamCfg<B>;
    static bool configured = false;
    static int occ = 1;
    if (!configured) {
        DGB_CUDA_OK(cudaFuncSetAttribute(k_stream<B, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        DGB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stream<B, MODE>, C::NT, C::SMEM));
        if (occ < 1) occ = 1;
        if (occ > 2) occ = 2;
        configured = true;
    }
    const int ntiles = (N + C::T - 1) / C::T;
    int grid = sm_count() * occ;
    if (grid > ntiles) grid = ntiles;
    if (grid > kMaxPartials) grid = kMaxPartials;
    k_stream<B, MODE><<<grid, C::NT, C::SMEM, st>>>(data, indices, indptr, N, Ni, rhs, x_in, x_out, partials, omega,
                                                   colour, g_err, skip);
    DGB_LAUNCH_OK();
    if (grid_out) *grid_out = grid;
    return 0;
}

// mode: S_APPLY / S_RESIDUAL / S_RELAX
int stream_launch(int mode, int b, const double *data, const int32_t *indices, const int32_t *indptr, int N,
                  int Ni, const double *rhs, const double *x_in, double *x_out, double *partials, double omega,
                  int colour, const int32_t *skip, cudaStream_t st, int *grid_out) {
    int rc = ensure_work(0);
    if (rc) return rc;
    if (mode == S_APPLY) {
        DGB_DISPATCH_B(b, return (stream_launch_t<B, S_APPLY>(data, indices, indptr, N, Ni, rhs, x_in, x_out, partials,
                                                             omega, colour, skip, st, grid_out)));
    } else if (mode == S_RESIDUAL) {
        DGB_DISPATCH_B(b, return (stream_launch_t<B, S_RESIDUAL>(data, indices, indptr, N, Ni, rhs, x_in, x_out,
                                                                partials, omega, colour, skip, st, grid_out)));
    } else {
        DGB_DISPATCH_B(b, return (stream_launch_t<B, S_RELAX>(data, indices, indptr, N, Ni, rhs, x_in, x_out, partials,
                                                             omega, colour, skip, st, grid_out)));
    }
    return 0;
}

template <int B>
static int gs_rows_launch_t(const double *gs, const double *rhs, double *x, double *mbox, Stencil S_, int dir,
                            double omega, const int32_t *skip, cudaStream_t st) {
    using C = GsCfg<B>;
    static bool configured = false;
    if (!configured) {
        DGB_CUDA_OK(cudaFuncSetAttribute(k_gs_rows<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured = true;
    }
    DGB_CUDA_OK(cudaMemsetAsync(g_work, 0, sizeof(int), st));
    const int grid = (S_.Nj + C::W - 1) / C::W;
    k_gs_rows<B><<<grid, C::W * 32, C::SMEM, st>>>(gs, rhs, x, mbox, S_, dir, omega, g_work, g_err, skip);
    DGB_LAUNCH_OK();
    return 0;
}

int gs_rows_launch(int b, const double *gs, const double *rhs, double *x, double *mbox, int Ni, int Nj, int flags,
                   int dir, double omega, const int32_t *skip, cudaStream_t st) {
    int rc = ensure_work(0);
    if (rc) return rc;
    Stencil S_{Ni, Nj, (flags & DGB_FLAG_PERIODIC_I) ? 1 : 0, (flags & DGB_FLAG_PERIODIC_J) ? 1 : 0};
    DGB_DISPATCH_B(b, return (gs_rows_launch_t<B>(gs, rhs, x, mbox, S_, dir, omega, skip, st)));
    return 0;
}

}  // namespace dgb

using namespace dgb;

extern "C" {

int dgb_set_kernel_path(int32_t path) {
    const int old = g_kernel_path;
    if (path == 0 || path == 1) g_kernel_path = path;
    return old;
}

int dgb_device_error(int32_t reset) {
    if (g_err == nullptr) return 0;
    int v = 0;
    if (cudaMemcpy(&v, g_err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (reset && v != 0) cudaMemset(g_err, 0, sizeof(int));
    return v;
}

int dgb_check_stencil(const int32_t *indices, const int32_t *indptr, int32_t Ni, int32_t Nj, int32_t flags,
                      int32_t *mismatch, void *stream) {
    DGB_ARG(indices && indptr && mismatch && Ni > 0 && Nj > 0 && flags >= 0);
    cudaStream_t st = (cudaStream_t)stream;
    Stencil S_{Ni, Nj, (flags & DGB_FLAG_PERIODIC_I) ? 1 : 0, (flags & DGB_FLAG_PERIODIC_J) ? 1 : 0};
    DGB_CUDA_OK(cudaMemsetAsync(mismatch, 0, sizeof(int32_t), st));
    int g = (Ni * Nj + 255) / 256;
    if (g > sm_count() * 8) g = sm_count() * 8;
    k_check_stencil<<<g, 256, 0, st>>>(indices, indptr, S_, mismatch);
    DGB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
