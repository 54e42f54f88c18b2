// dgb_comm.cu -- element-slab multi-GPU path over peer memory (NVLink / NVSwitch), sm_100a.
//
// One process per GPU.  Every rank allocates one "arena" (cudaMalloc), exports it with CUDA IPC and maps the arenas
// of all other ranks: a device pointer into a peer's arena is  peer_base + (local pointer - my_base)  because all
// ranks lay their arenas out identically (symmetric allocation).  Inside the arena live
//   * a control block: flags written by the neighbours, sequence counters, all-reduce slots, the error word;
//   * the iterate u of every distributed level, as a block of (rows + 2) element rows: [ghost below | owned rows |
//     ghost above] -- a rank without a lower / upper neighbour simply never reads that ghost row;
//   * the gathered right-hand side of the first replicated (coarse) level.
// Three collectives, each ONE kernel launch whose blocks talk to the peers with plain stores / system-scope flags:
//   halo exchange   my edge rows -> the neighbours' ghost rows, after both sides said "arrived" (their ghost rows are
//                   free, my edge rows are final); "pushed" flags tell the neighbour the data has landed
//   all-reduce      every rank stores its partial sum into a slot on every rank; each rank adds the slots in rank
//                   order (bitwise the same result everywhere) and applies the smoother's residual test with it
//   all-gather      barrier, every rank stores its chunk into every rank's buffer, flags, wait
// The reference has no parallel path: this replaces nothing in dgfem, it carries
// Solver.multigrid_V_cycle (dgfem/solver.py:141-207) across slabs of whole element rows (m = j*Ni + i,
// utils/helpers.py:14).  The V-cycle itself (dgb_vcycle_slab) is sequenced here, in C++, with no host
// synchronisation and no Python between the launches.
#include <vector>

#include "dgb_common.cuh"

namespace dgb {

constexpr int kMaxWorld = 64;
constexpr unsigned long long kWaitNs = 20ull * 1000ull * 1000ull * 1000ull;    // bounded waits: 20 s

struct CommCtl {                       // at the start of every arena; peer-visible
    // written by the neighbours (lo = rank - 1, hi = rank + 1)
    unsigned arrived_from_lo, arrived_from_hi, pushed_from_lo, pushed_from_hi;
    // sequence numbers of the collectives, kept on the device (only this rank's kernels touch them) so that a
    // captured CUDA graph of the cycle can be replayed: every block advances its own counter
    unsigned seq_halo[2], seq_red, pad0;
    unsigned seq_gather[kMaxWorld];
    int err;                           // 3 = a peer did not show up within kWaitNs
    int pad1[7];
    unsigned bar_arrived[kMaxWorld];   // all-gather: written by rank p into slot p
    unsigned gat_pushed[kMaxWorld];
    double red[2][kMaxWorld];          // all-reduce slots, all-ones NaN = empty
};
constexpr size_t kCtlBytes = (sizeof(CommCtl) + 4095) & ~(size_t)4095;

}  // namespace dgb

struct dgb_comm {
    int rank, world, device;
    size_t bytes;
    char *base;                        // my arena
    char *peer[dgb::kMaxWorld];        // peer[p] = base of rank p's arena in my address space (peer[rank] = base)
    char **d_peer;                     // the same table in device memory
    bool connected;
};

namespace dgb {

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned ld_acq_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_rel_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_vol_f64_sys(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// one thread spins until *flag >= want (sequence numbers only grow); false + err on timeout
__device__ bool wait_flag(const unsigned *flag, unsigned want, int *err) {
    const unsigned long long t0 = gtimer();
    while ((int)(ld_acq_sys(flag) - want) < 0) {
        if (gtimer() - t0 > kWaitNs || *(volatile int *)err != 0) {
            atomicExch(err, 3);
            return false;
        }
        __nanosleep(100);
    }
    return true;
}

__device__ __forceinline__ CommCtl *ctl_of(char *base) { return reinterpret_cast<CommCtl *>(base); }

// ---- halo exchange: block 0 talks to rank - 1, block 1 to rank + 1 ---------------------------------------------
// blk: byte offset of the level's vector block in every arena; layout [ghost below][rows owned][ghost above]
__global__ void __launch_bounds__(1024)
k_halo_exchange(char *const *__restrict__ peers, int rank, int world, size_t blk, long long row_doubles, int rows) {
    const bool hi = blockIdx.x == 1;
    const int nb = hi ? rank + 1 : rank - 1;
    char *me = peers[rank];
    CommCtl *mine = ctl_of(me);
    __shared__ int s_ok;
    __shared__ unsigned s_seq;
    if (threadIdx.x == 0) {
        s_ok = 1;
        s_seq = ++mine->seq_halo[hi ? 1 : 0];       // all ranks issue the same sequence of exchanges
    }
    __syncthreads();
    const unsigned seq = s_seq;
    if (nb >= 0 && nb < world) {
        CommCtl *theirs = ctl_of(peers[nb]);
        if (threadIdx.x == 0) {
            // everything before this kernel on my stream is complete: my ghost rows are free, my edge rows final
            st_rel_sys(hi ? &theirs->arrived_from_lo : &theirs->arrived_from_hi, seq);
            s_ok = wait_flag(hi ? &mine->arrived_from_hi : &mine->arrived_from_lo, seq, &mine->err) ? 1 : 0;
        }
        __syncthreads();
        if (s_ok) {
            // my last (first) owned row -> the upper (lower) neighbour's ghost row below (above) its rows
            const double *src = reinterpret_cast<const double *>(me + blk) + (hi ? (long long)rows : 1LL) * row_doubles;
            double *dst = reinterpret_cast<double *>(peers[nb] + blk) + (hi ? 0LL : (long long)(rows + 1)) * row_doubles;
            if ((row_doubles & 1) == 0 && (blk & 15) == 0) {
                const double2 *s2 = reinterpret_cast<const double2 *>(src);
                double2 *d2 = reinterpret_cast<double2 *>(dst);
                for (long long t = threadIdx.x; t < row_doubles / 2; t += 1024) d2[t] = s2[t];
            } else {
                for (long long t = threadIdx.x; t < row_doubles; t += 1024) dst[t] = src[t];
            }
            __threadfence_system();
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_ok) {
            st_rel_sys(hi ? &theirs->pushed_from_lo : &theirs->pushed_from_hi, seq);
            wait_flag(hi ? &mine->pushed_from_hi : &mine->pushed_from_lo, seq, &mine->err);
        }
    }
}

// ---- all-reduce of one double + the smoother's residual test ----------------------------------------------------
// mode 0: *value <- global sum;  1: and dgb_smoother_begin;  2: and dgb_smoother_check (dgfem/relaxation.py:202-216)
__global__ void __launch_bounds__(64)
k_allreduce_ctl(char *const *__restrict__ peers, int rank, int world, double *value, int mode, dgb_smoother_ctl *sctl,
                double n_global) {
    CommCtl *mine = ctl_of(peers[rank]);
    __shared__ double s_v[kMaxWorld];
    const unsigned seq = mine->seq_red;             // advanced by thread 0 after the barrier below
    const int buf = (int)(seq & 1u);
    double mine_v = *value;
    if (__double_as_longlong(mine_v) == -1LL) mine_v = __longlong_as_double(0x7ff8000000000000LL);   // keep the empty mark free
    if ((int)threadIdx.x < world) {
        const int p = threadIdx.x;
        double *slot = &ctl_of(peers[p])->red[buf][rank];
        asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(slot), "d"(mine_v) : "memory");
        // my copy of rank p's contribution
        const double *in = &mine->red[buf][p];
        const unsigned long long t0 = gtimer();
        double v = ld_vol_f64_sys(in);
        while (__double_as_longlong(v) == -1LL) {
            if (gtimer() - t0 > kWaitNs || *(volatile int *)&mine->err != 0) {
                atomicExch(&mine->err, 3);
                v = 0.0;
                break;
            }
            __nanosleep(50);
            v = ld_vol_f64_sys(in);
        }
        s_v[p] = v;
        mine->red[buf][p] = __longlong_as_double(-1LL);      // empty again for the reduction after the next one
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int p = 0; p < world; ++p) s += s_v[p];          // rank order on every rank: identical bits
        *value = s;
        mine->seq_red = seq + 1;
        if (mode == 1) {
            sctl->res0 = sqrt(s / n_global);
            sctl->ratio = 1.0;
            sctl->skip = sctl->diverged;
            sctl->iters = 0;
            sctl->calls += 1;
        } else if (mode == 2 && !sctl->skip) {
            const double ratio = sqrt(s / n_global) / sctl->res0;
            sctl->ratio = ratio;
            sctl->iters += 1;
            if (ratio < 1e-6) {
                sctl->skip = 1;
            } else if (ratio > 1e10) {
                sctl->diverged = 1;
                sctl->skip = 1;
            }
        }
        __threadfence_system();
    }
}

// ---- all-gather: dst block (same offset in every arena) <- [chunk of rank 0 | rank 1 | ...] ----------------------
__global__ void __launch_bounds__(1024)
k_allgather(char *const *__restrict__ peers, int rank, int world, const double *__restrict__ src, size_t dst_off,
            long long chunk) {
    const int p = blockIdx.x;                       // one block per destination rank
    CommCtl *mine = ctl_of(peers[rank]);
    CommCtl *theirs = ctl_of(peers[p]);
    __shared__ int s_ok;
    __shared__ unsigned s_seq;
    if (threadIdx.x == 0) s_seq = ++mine->seq_gather[p];
    __syncthreads();
    const unsigned seq = s_seq;
    if (threadIdx.x == 0) {
        s_ok = 1;
        if (p != rank) {
            st_rel_sys(&theirs->bar_arrived[rank], seq);                       // my previous gather buffer is consumed
            s_ok = wait_flag(&mine->bar_arrived[p], seq, &mine->err) ? 1 : 0;  // ... and so is rank p's
        }
    }
    __syncthreads();
    if (s_ok) {
        double *dst = reinterpret_cast<double *>(peers[p] + dst_off) + (long long)rank * chunk;
        for (long long t = threadIdx.x; t < chunk; t += 1024) dst[t] = src[t];
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_ok && p != rank) {
        st_rel_sys(&theirs->gat_pushed[rank], seq);
        wait_flag(&mine->gat_pushed[p], seq, &mine->err);
    }
}

__global__ void k_ctl_init(CommCtl *c) {
    const int t = threadIdx.x;
    if (t == 0) {
        c->arrived_from_lo = c->arrived_from_hi = c->pushed_from_lo = c->pushed_from_hi = 0;
        c->seq_halo[0] = c->seq_halo[1] = c->seq_red = 0;
        c->err = 0;
    }
    if (t < kMaxWorld) {
        c->bar_arrived[t] = c->gat_pushed[t] = c->seq_gather[t] = 0;
        c->red[0][t] = c->red[1][t] = __longlong_as_double(-1LL);
    }
}

static int comm_ok(const dgb_comm *c) {
    DGB_ARG(c != nullptr && c->connected);
    return 0;
}

}  // namespace dgb

using namespace dgb;

extern "C" {

int dgb_comm_create(int32_t rank, int32_t world, int64_t arena_bytes, dgb_comm **out) {
    DGB_ARG(out && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world && arena_bytes >= 0);
    dgb_comm *c = new dgb_comm();
    c->rank = rank;
    c->world = world;
    c->connected = false;
    c->d_peer = nullptr;
    DGB_CUDA_OK(cudaGetDevice(&c->device));
    c->bytes = kCtlBytes + (((size_t)arena_bytes + 255) & ~(size_t)255);
    DGB_CUDA_OK(cudaMalloc(&c->base, c->bytes));
    DGB_CUDA_OK(cudaMemset(c->base, 0, c->bytes));
    k_ctl_init<<<1, kMaxWorld>>>(reinterpret_cast<CommCtl *>(c->base));
    DGB_LAUNCH_OK();
    DGB_CUDA_OK(cudaDeviceSynchronize());
    for (int p = 0; p < kMaxWorld; ++p) c->peer[p] = nullptr;
    c->peer[rank] = c->base;
    *out = c;
    return 0;
}

int dgb_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int dgb_comm_export(dgb_comm *c, void *h_handle) {
    DGB_ARG(c && h_handle);
    cudaIpcMemHandle_t h;
    DGB_CUDA_OK(cudaIpcGetMemHandle(&h, c->base));
    memcpy(h_handle, &h, sizeof(h));
    return 0;
}

int dgb_comm_connect(dgb_comm *c, const void *h_handles) {
    DGB_ARG(c && (h_handles || c->world == 1));
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)h_handles + (size_t)p * sizeof(h), sizeof(h));
        void *ptr = nullptr;
        DGB_CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[p] = (char *)ptr;
    }
    DGB_CUDA_OK(cudaMalloc(&c->d_peer, sizeof(char *) * kMaxWorld));
    DGB_CUDA_OK(cudaMemcpy(c->d_peer, c->peer, sizeof(char *) * kMaxWorld, cudaMemcpyHostToDevice));
    c->connected = true;
    return 0;
}

void *dgb_comm_arena(dgb_comm *c, int64_t *bytes) {
    if (c == nullptr) return nullptr;
    if (bytes) *bytes = (int64_t)(c->bytes - kCtlBytes);
    return c->base + kCtlBytes;
}

int dgb_comm_error(dgb_comm *c, int32_t reset) {
    if (c == nullptr) return 0;
    int v = 0;
    if (cudaMemcpy(&v, &reinterpret_cast<CommCtl *>(c->base)->err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (reset && v) cudaMemset(&reinterpret_cast<CommCtl *>(c->base)->err, 0, sizeof(int));
    return v;
}

void dgb_comm_destroy(dgb_comm *c) {
    if (c == nullptr) return;
    cudaDeviceSynchronize();
    for (int p = 0; p < c->world; ++p)
        if (p != c->rank && c->peer[p]) cudaIpcCloseMemHandle(c->peer[p]);
    if (c->d_peer) cudaFree(c->d_peer);
    cudaFree(c->base);
    delete c;
}

int dgb_halo_exchange(dgb_comm *c, double *block, int64_t row_doubles, int32_t rows, void *stream) {
    int rc = comm_ok(c);
    if (rc) return rc;
    DGB_ARG(block && row_doubles > 0 && rows > 0);
    const size_t off = (size_t)((char *)block - c->base);
    DGB_ARG((char *)block >= c->base + kCtlBytes && off + (size_t)(rows + 2) * row_doubles * 8 <= c->bytes);
    if (c->world == 1) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    k_halo_exchange<<<2, 1024, 0, st>>>(c->d_peer, c->rank, c->world, off, row_doubles, rows);
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_allreduce_sum(dgb_comm *c, double *value, int32_t mode, dgb_smoother_ctl *ctl, int64_t n_global, void *stream) {
    int rc = comm_ok(c);
    if (rc) return rc;
    DGB_ARG(value && mode >= 0 && mode <= 2 && (mode == 0 || (ctl && n_global > 0)));
    k_allreduce_ctl<<<1, 64, 0, (cudaStream_t)stream>>>(c->d_peer, c->rank, c->world, value, mode, ctl, (double)n_global);
    DGB_LAUNCH_OK();
    return 0;
}

int dgb_allgather(dgb_comm *c, const double *src, double *dst_block, int64_t chunk, void *stream) {
    int rc = comm_ok(c);
    if (rc) return rc;
    DGB_ARG(src && dst_block && chunk > 0);
    const size_t off = (size_t)((char *)dst_block - c->base);
    DGB_ARG((char *)dst_block >= c->base + kCtlBytes && off + (size_t)chunk * c->world * 8 <= c->bytes);
    cudaStream_t st = (cudaStream_t)stream;
    k_allgather<<<c->world, 1024, 0, st>>>(c->d_peer, c->rank, c->world, src, off, chunk);
    DGB_LAUNCH_OK();
    return 0;
}

}  // extern "C"

// =========================================================================================================
// the V-cycle across slabs
// =========================================================================================================
namespace dgb {

int vcycle_entry(const dgb_level *lv, int k, const dgb_vcycle_opts &o, dgb_smoother_ctl *ctl, double *partials,
                 double *sumsq, void *stream, bool u_zero);

struct SlabCtx {
    dgb_comm *comm;
    const dgb_slab_level *lv;
    int n;
    const dgb_slab_opts *o;
    dgb_smoother_ctl *ctl;
    double *partials, *sumsq;
    void *stream;
    // level whose ghost rows are current (no owned row of it changed since its last exchange), -1 none: every rank
    // takes the same decisions, so an exchange that would move the same bits again is dropped on all of them
    mutable int halo_fresh = -1;
};

static int halo(const SlabCtx &s, int k) {
    if (s.halo_fresh == k) return 0;
    const dgb_slab_level &L = s.lv[k];
    const int rows = L.lev.op.Nj - L.ghost_lo - L.ghost_hi;
    s.halo_fresh = k;
    return dgb_halo_exchange(s.comm, L.u_block, (int64_t)L.lev.op.Ni * L.lev.op.b, rows, s.stream);
}
static void touched(const SlabCtx &s) { s.halo_fresh = -1; }

// global sum of squares of rhs - A u over the owned rows -> *sumsq, then the residual test `mode` (0 none)
// relaxed_colour >= 0: that colour was relaxed last and nothing changed since -- its rows are skipped
static int residual_test(const SlabCtx &s, int k, double *r, const int32_t *skip, int mode, int first_direction,
                         bool *fused, int relaxed_colour = -1) {
    const dgb_slab_level &L = s.lv[k];
    int rc = halo(s, k);
    if (rc) return rc;
    bool f = false;
    if (relaxed_colour >= 0) {
        if ((rc = dgb_bsr_residual_colour(&L.lev.op, L.lev.rhs, L.lev.u, r, relaxed_colour, L.colour_shift, s.partials,
                                          s.sumsq, skip, s.stream)))
            return rc;
        if (fused) *fused = false;
        return dgb_allreduce_sum(s.comm, s.sumsq, mode, s.ctl + k, L.n_global, s.stream);
    }
    if (first_direction != 0 && skip == nullptr) {
        rc = dgb_block_gs_entry_residual(&L.lev.op, L.lev.rhs, L.lev.u, first_direction, r, s.partials, s.sumsq, s.stream);
        if (rc == 0) f = true;
        else if (rc != DGB_UNSUPPORTED) return rc;
    }
    if (!f && (rc = dgb_bsr_residual(&L.lev.op, L.lev.rhs, L.lev.u, r, s.partials, s.sumsq, skip, s.stream))) return rc;
    if (fused) *fused = f;
    return dgb_allreduce_sum(s.comm, s.sumsq, mode, s.ctl + k, L.n_global, s.stream);
}

// *last_colour: the colour relaxed last with nothing changed since (-1 none): relaxing it again would recompute the
// same bits, so that pass (and its halo exchange) is dropped
static int gs_pass(const SlabCtx &s, int k, int direction, int prev, const int32_t *skip, int *last_colour) {
    const dgb_slab_level &L = s.lv[k];
    int rc;
    if (s.o->gs_mode == DGB_GS_REDBLACK) {
        for (int c = 0; c < 2; ++c) {
            const int colour = direction > 0 ? c : 1 - c;
            if (colour == *last_colour) continue;
            if ((rc = halo(s, k))) return rc;
            if ((rc = dgb_block_gs_colour(&L.lev.op, L.lev.rhs, L.lev.u, colour, L.colour_shift, skip, s.stream))) return rc;
            touched(s);
            *last_colour = colour;
        }
        return 0;
    }
    // slab_lexicographic: lexicographic inside the slab, the neighbours' rows as they were before the pass
    if ((rc = halo(s, k))) return rc;
    touched(s);
    return dgb_block_gs_pass_seq(&L.lev.op, L.lev.rhs, L.lev.u, direction, prev, skip, s.stream);
}

// one smoother call (Relaxation.block_gauss_seidel_pyamg, dgfem/relaxation.py:198-218, across slabs);
// *have_r: r holds rhs - A u of the returned u
static int smooth(const SlabCtx &s, int k, bool post, double *r_keep, bool *have_r) {
    const dgb_slab_level &L = s.lv[k];
    const int smoother = post ? L.lev.post_smoother : L.lev.smoother;
    const int direction = post ? L.lev.post_direction : L.lev.direction;
    const int iterations = post ? L.lev.post_iterations : L.lev.pre_iterations;
    if (have_r) *have_r = false;
    if (iterations <= 0) return 0;
    if (smoother != DGB_SMOOTHER_BLOCK_GS_PYAMG) {
        set_error("dgb_vcycle_slab: only block_gauss_seidel_pyamg runs across slabs (smoother id %d)", smoother);
        return 3;
    }
    const bool check = s.o->check_residual != 0;
    const int32_t *skip = nullptr;
    int prev = 0, rc;
    int last_colour = -1;
    if (check) {
        const int first = direction >= 0 ? +1 : -1;
        bool fused = false;
        const bool lexi = s.o->gs_mode != DGB_GS_REDBLACK;
        if (!lexi && L.lev.op.stencil >= 0) {
            // 2-colour mode: the entry residual kernel also relaxes the first colour of the first pass
            const int c0 = first > 0 ? 0 : 1;
            if ((rc = halo(s, k))) return rc;
            if ((rc = dgb_block_gs_colour_entry(&L.lev.op, L.lev.rhs, L.lev.u, r_keep, c0, L.colour_shift, s.partials, s.sumsq,
                                                &s.ctl[k].diverged, s.stream)))
                return rc;
            touched(s);
            if ((rc = dgb_allreduce_sum(s.comm, s.sumsq, 1, s.ctl + k, L.n_global, s.stream))) return rc;
            last_colour = c0;
        } else {
            if ((rc = residual_test(s, k, r_keep, nullptr, 1, lexi ? first : 0, &fused))) return rc;
            if (fused) prev = -first;
        }
        skip = &s.ctl[k].skip;
    }
    for (int it = 0; it < iterations; ++it) {
        if (direction >= 0) {
            if ((rc = gs_pass(s, k, +1, prev, skip, &last_colour))) return rc;
            prev = +1;
        }
        if (direction <= 0) {
            if ((rc = gs_pass(s, k, -1, prev, skip, &last_colour))) return rc;
            prev = -1;
        }
        if (check && (rc = residual_test(s, k, r_keep, skip, 2, 0, nullptr, last_colour))) return rc;
    }
    if (have_r) *have_r = check && r_keep != nullptr;
    return 0;
}

static int vcycle_slab(const SlabCtx &s, int k) {
    const dgb_slab_level &L = s.lv[k];
    cudaStream_t st = (cudaStream_t)s.stream;
    int rc;
    bool have_r = false;
    if ((rc = smooth(s, k, false, L.lev.r, &have_r))) return rc;
    if (!have_r && (rc = residual_test(s, k, L.lev.r, nullptr, 0, 0, nullptr))) return rc;
    if (k > 0) {
        const dgb_slab_level &C = s.lv[k - 1];
        const dgb_level &c = C.lev;
        if ((rc = dgb_restrict_slab(c.transfer_kind, c.R, c.nc, c.nf, c.op.Ni, c.op.Nj, C.ghost_lo, C.ghost_hi, L.ghost_lo,
                                    L.lev.r, c.rhs, s.stream)))
            return rc;
        // solver.py:171 -- the whole block (ghost rows included) starts from zero
        DGB_CUDA_OK(cudaMemsetAsync(C.u_block, 0, sizeof(double) * (size_t)c.op.Ni * c.op.b * (c.op.Nj - C.ghost_lo - C.ghost_hi + 2), st));
        s.halo_fresh = k - 1;       // zero everywhere, the neighbours' rows included
        if ((rc = vcycle_slab(s, k - 1))) return rc;
        if ((rc = dgb_prolong_add_slab(c.transfer_kind, c.P, c.nc, c.nf, c.op.Ni, c.op.Nj, C.ghost_lo, C.ghost_hi, L.ghost_lo,
                                       c.u, L.lev.u, s.stream)))
            return rc;
        touched(s);
    } else {
        // the link to the replicated hierarchy: restrict into an un-ghosted chunk, gather the chunks of all ranks on
        // every rank, run the rest of the cycle redundantly (identical bits everywhere), prolong my chunk back
        const dgb_slab_opts &o = *s.o;
        const dgb_level &top = o.coarse_levels[o.n_coarse - 1];
        if ((rc = dgb_restrict_slab(o.link_kind, o.link_R, o.link_nc, o.link_nf, o.link_Ni_c, o.link_rows_c, 0, 0, L.ghost_lo,
                                    L.lev.r, o.link_rhs_local, s.stream)))
            return rc;
        const int64_t chunk = (int64_t)o.link_Ni_c * o.link_rows_c * o.link_nc;
        if ((rc = dgb_allgather(s.comm, o.link_rhs_local, top.rhs, chunk, s.stream))) return rc;
        DGB_CUDA_OK(cudaMemsetAsync(top.u, 0, sizeof(double) * (size_t)top.op.Ni * top.op.Nj * top.op.b, st));
        if ((rc = vcycle_entry(o.coarse_levels, o.n_coarse - 1, o.coarse_opts, o.coarse_ctl, s.partials, s.sumsq, s.stream, true)))
            return rc;
        if ((rc = dgb_prolong_add_slab(o.link_kind, o.link_P, o.link_nc, o.link_nf, o.link_Ni_c, o.link_rows_c, 0, 0, L.ghost_lo,
                                       top.u + (size_t)s.comm->rank * chunk, L.lev.u, s.stream)))
            return rc;
        touched(s);
    }
    return smooth(s, k, true, nullptr, nullptr);
}

}  // namespace dgb

extern "C" int dgb_vcycle_slab(dgb_comm *comm, const dgb_slab_level *h_levels, int32_t nlevels, const dgb_slab_opts *h_opts,
                               dgb_smoother_ctl *ctl, double *partials, double *sumsq, void *stream) {
    int rc = comm_ok(comm);
    if (rc) return rc;
    DGB_ARG(h_levels && h_opts && ctl && partials && sumsq && nlevels >= 1);
    DGB_ARG(h_opts->gs_mode == DGB_GS_REDBLACK || h_opts->gs_mode == DGB_GS_SLAB_LEXICOGRAPHIC);
    DGB_ARG(h_opts->coarse_levels && h_opts->n_coarse >= 1 && h_opts->coarse_ctl && h_opts->link_R && h_opts->link_P &&
            h_opts->link_rhs_local);
    for (int k = 0; k < nlevels; ++k) {
        const dgb_slab_level &L = h_levels[k];
        DGB_ARG(L.u_block && L.lev.u == L.u_block + (size_t)(1 - L.ghost_lo) * L.lev.op.Ni * L.lev.op.b);
        DGB_ARG(L.lev.rhs && L.lev.r && L.n_global > 0);
    }
    SlabCtx s{comm, h_levels, nlevels, h_opts, ctl, partials, sumsq, stream};
    return vcycle_slab(s, nlevels - 1);
}
