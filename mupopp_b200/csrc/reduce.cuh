// Deterministic grid-wide reductions shared by the Krylov kernels (krylov.cu) and the functional kernels (functional.cu):
// device scalar slots, last-block-done partial sums without floating-point atomics, and -- on several GPUs -- the in-kernel
// allreduce through peer-memory mailboxes (NVLink stores + sequence-number flags).
#pragma once
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 4096;  // partial sums per reduction slot
constexpr int kMaxRestart = 4;    // residual-replacement restarts of the Krylov recurrences
inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// device scalar slots
// layout: (S_A0,S_RR) and (S_RR,S_A1) are contiguous pairs so that the ping-ponged recurrence scalar and
// the residual norm travel in ONE allreduce per reduction point on multi-GPU runs.
enum {
    S_BB = 0, S_TOL2 = 1, S_ITER = 2,
    S_A0 = 3, S_RR = 4, S_A1 = 5,       // A0/A1: r.z (CG) or rhat.r (BiCGStab), ping-pong
    S_PQ = 6,                           // p.Ap (CG)  /  rhat.v (BiCGStab)
    S_TS = 7, S_TT = 8,                 // BiCGStab t.s, t.t
    S_TMP0 = 9,
    S_NCOL = 10,                        // GMRES: Givens columns completed in the current cycle
    S_DELTA = 11,                       // single-reduction CG: w.u
    S_G0 = 12, S_G1 = 13,               // single-reduction CG: gamma = r.u, ping-pong (current / previous)
    S_AL0 = 14, S_AL1 = 15,             // single-reduction CG: alpha, ping-pong
    S_COUNT = 16,
    S_GLOC = 40, S_RRLOC = 41           // single-reduction CG: rank-local r.u and r.r of the vector kernel, folded into the SpMV's exchange
};
constexpr int S_RZ0 = S_A0, S_RZ1 = S_A1, S_RHO0 = S_A0, S_RHO1 = S_A1, S_RV = S_PQ;

__device__ __forceinline__ double safe_div(double a, double b) { return b == 0.0 ? 0.0 : a / b; }

// ---- system-scope flag primitives of the peer-memory collectives (NVLink stores + sequence-number flags)
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;   // a peer that stays silent this long is dead

// spin until *flag >= seq; false on time-out (never hang the GPU: a hung box is a lost run)
__device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long seq) {
    if (ld_acquire_sys(flag) >= seq) return true;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flag) < seq) {
        if (global_timer_ns() - t0 > kSpinTimeoutNs) return false;
        __nanosleep(64);
    }
    return true;
}

// Block-level sum of NR accumulators, partials to global, the last block to arrive adds the partials in a
// fixed order and stores the results: no floating-point atomics => bitwise reproducible for a given grid.
// Multi-GPU, two transports:
//  * peer mode (ctl->comm != null): the last block ALSO allreduces the NR sums over NVLink inside the kernel.  Thread t < nranks
//    stores the sums into rank t's mailbox window (mapped with CUDA IPC) and releases a sequence-number flag, then waits for
//    rank t's contribution in the local window; thread 0 adds the contributions in RANK ORDER, so every rank computes bitwise the
//    same scalars (the ranks therefore agree on "converged" and keep issuing the same sequence of exchanges).  Mailboxes are
//    double-buffered by the parity of the sequence number: a rank can be at most one exchange ahead of its slowest peer.
//  * NCCL mode: rank-local sums go to the LOCAL copy of the scalar block (offset ctl->local_off = kLocalOff) and an out-of-place
//    ncclAllReduce(local -> global) follows on the host side (reduce_slots); kernels only ever READ the global copy.  This keeps
//    the allreduce idempotent when the producing kernel has become a no-op after convergence.
template <int NR, int BT = kThreads>   // BT = blockDim.x of the calling kernel (a multiple of 32)
__device__ __forceinline__ void grid_reduce(double (&acc)[NR], double *__restrict__ partials, mp::ReduceCtl *ctl,
                                            double *__restrict__ scal, const int (&slots)[NR], bool bump_iter, bool use_off = true,
                                            bool exchange = true) {
    static_assert(NR <= mp::kMboxVals, "mailbox too small");
    __shared__ double sm[NR][BT / 32];
    __shared__ double res[NR];
    __shared__ double xin[mp::kMaxRanks][NR];
    __shared__ bool last;
    __shared__ int failed;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double v = acc[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sm[r][w] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double v = 0.0;
            for (int k = 0; k < BT / 32; ++k) v += sm[r][k];
            partials[r * kMaxBlocks + blockIdx.x] = v;
        }
        __threadfence();
        unsigned int t = atomicAdd(&ctl->count, 1u);
        last = (t == gridDim.x - 1);
        failed = 0;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    mp::CommDev *comm = exchange ? ctl->comm : nullptr;      // exchange == false: keep the rank-local sums (folded into a later exchange)
    if (!exchange) use_off = false;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double v = 0.0;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += BT) v += __ldcg(&partials[r * kMaxBlocks + k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) sm[r][w] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int k = 0; k < BT / 32; ++k) s += sm[r][k];
            if (comm) res[r] = s;
            else scal[slots[r] + (use_off ? (int)ctl->local_off : 0)] = s;
        }
    }
    if (comm) {
        __syncthreads();
        const unsigned long long seq = comm->seq;     // only thread 0 writes it, after the barrier below
        const int par = (int)(seq & 1ull), nr = comm->nranks, me = comm->rank;
        if ((int)threadIdx.x < nr) {
            mp::MboxWindow *dst = comm->win[threadIdx.x];
#pragma unroll
            for (int r = 0; r < NR; ++r) *(volatile double *)&dst->val[par][me][r] = res[r];
            st_release_sys(&dst->flag[par][me], seq);                      // orders the value stores before the flag
            const mp::MboxWindow *mine = comm->win[me];
            if (!wait_flag(&mine->flag[par][threadIdx.x], seq)) failed = 1;
#pragma unroll
            for (int r = 0; r < NR; ++r) xin[threadIdx.x][r] = *(const volatile double *)&mine->val[par][threadIdx.x][r];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                double s = 0.0;
                for (int t = 0; t < nr; ++t) s += xin[t][r];               // rank order: identical bits on every rank
                scal[slots[r]] = s;
            }
            comm->seq = seq + 1ull;
            if (failed) {                                                  // a dead peer: poison the residual so that the host sees it
                comm->error = 1;
                scal[S_RR] = __longlong_as_double(0x7ff8000000000000ll);
            }
        }
    }
    if (threadIdx.x == 0) {
        ctl->count = 0u;
        if (bump_iter) scal[S_ITER] += 1.0;
    }
}

__device__ __forceinline__ bool solver_active(const double *scal) { return scal[S_RR] > scal[S_TOL2]; }

}  // namespace
