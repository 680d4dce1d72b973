// Bandwidth-bound linear algebra: CSR SpMV (sub-warp per row), fused BLAS-1 with deterministic
// two-stage reductions, Jacobi-preconditioned CG and BiCGStab whose scalars (alpha, beta, omega,
// convergence test) live on the device, so the host only polls every `check_every` iterations.
//
// Replaces PETSc MatMult/VecDot/VecAXPY/KSP reached through dolfin.KrylovSolver
// (/root/reference/src/modules/mupopp.py:289-303, 410-425) and the sparse LU inside
// solve(a==L, ...) (/root/reference/src/run/CCS2D/CCS_3comp.py:261): LU returns the exact solution of
// the assembled system, which a tightly converged Krylov iteration reproduces.
#include <string.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 4096;  // partial sums per reduction slot
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
    S_COUNT = 16
};
constexpr int S_RZ0 = S_A0, S_RZ1 = S_A1, S_RHO0 = S_A0, S_RHO1 = S_A1, S_RV = S_PQ;

__device__ __forceinline__ double safe_div(double a, double b) { return b == 0.0 ? 0.0 : a / b; }

// Block-level sum of NR accumulators, partials to global, the last block to arrive adds the partials in a
// fixed order and stores the results: no floating-point atomics => bitwise reproducible for a given grid.
template <int NR>
__device__ __forceinline__ void grid_reduce(double (&acc)[NR], double *__restrict__ partials, unsigned int *counter,
                                            double *__restrict__ scal, const int (&slots)[NR], bool bump_iter) {
    __shared__ double sm[NR][kThreads / 32];
    __shared__ bool last;
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
            for (int k = 0; k < kThreads / 32; ++k) v += sm[r][k];
            partials[r * kMaxBlocks + blockIdx.x] = v;
        }
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double v = 0.0;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += kThreads) v += __ldcg(&partials[r * kMaxBlocks + k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) sm[r][w] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int k = 0; k < kThreads / 32; ++k) s += sm[r][k];
            scal[slots[r]] = s;
        }
    }
    if (threadIdx.x == 0) {
        *counter = 0u;
        if (bump_iter) scal[S_ITER] += 1.0;
    }
}

__device__ __forceinline__ bool solver_active(const double *scal) { return scal[S_RR] > scal[S_TOL2]; }

// ------------------------------------------------------------------ SpMV: S lanes per row
// DOT: 0 none | 1: slot0 = sum w[row]*y[row] | 2: slot0 = sum y*w, slot1 = sum y*y
template <int S, int DOT, bool GUARD>
__global__ void __launch_bounds__(kThreads)
k_spmv(int64_t nrow, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, const double *__restrict__ vals,
       const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ w, double *partials,
       unsigned int *counter, double *scal, int slot0, int slot1) {
    if (GUARD && !solver_active(scal)) return;
    constexpr int RPW = 32 / S;  // rows per warp per pass
    const int lane = threadIdx.x & 31, sub = lane / S, sl = lane % S;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    double acc[DOT == 0 ? 1 : DOT];
#pragma unroll
    for (int r = 0; r < (DOT == 0 ? 1 : DOT); ++r) acc[r] = 0.0;
    (void)acc;
    for (int64_t base = warp * RPW; base < nrow; base += nwarp * RPW) {
        const int64_t row = base + sub;
        double sum = 0.0;
        if (row < nrow) {
            const int32_t b = rowptr[row], e = rowptr[row + 1];
            for (int32_t k = b + sl; k < e; k += S) sum += vals[k] * x[colidx[k]];
        }
#pragma unroll
        for (int o = S / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (row < nrow && sl == 0) {
            y[row] = sum;
            if constexpr (DOT == 1) acc[0] += w[row] * sum;
            if constexpr (DOT == 2) { acc[0] += sum * w[row]; acc[1] += sum * sum; }
        }
    }
    if constexpr (DOT == 1) { const int sl_[1] = {slot0}; grid_reduce<1>(acc, partials, counter, scal, sl_, false); }
    if constexpr (DOT == 2) { const int sl_[2] = {slot0, slot1}; grid_reduce<2>(acc, partials, counter, scal, sl_, false); }
}

// ------------------------------------------------------------------ BLAS-1
__global__ void __launch_bounds__(kThreads) k_axpby(int64_t n, double a, const double *__restrict__ x, double b, double *__restrict__ y) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x)
        y[i] = (b == 0.0) ? a * x[i] : a * x[i] + b * y[i];
}

__global__ void __launch_bounds__(kThreads) k_dot(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                                                  double *partials, unsigned int *counter, double *scal, int slot) {
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) acc[0] += x[i] * y[i];
    const int sl_[1] = {slot};
    grid_reduce<1>(acc, partials, counter, scal, sl_, false);
}

__global__ void __launch_bounds__(kThreads) k_jacobi_setup(int64_t nrow, const int32_t *__restrict__ rowptr,
                                                           const int32_t *__restrict__ colidx, const double *__restrict__ vals,
                                                           double *__restrict__ dinv) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nrow) return;
    double d = 0.0;
    for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
        if (colidx[k] == r) d = vals[k];
    dinv[r] = d != 0.0 ? 1.0 / d : 1.0;
}

// r = b - q ; (CG) p = dinv*r ; dots: bb, rr, rz
__global__ void __launch_bounds__(kThreads) k_init_residual(int64_t n, const double *__restrict__ b, const double *__restrict__ q,
                                                            const double *__restrict__ dinv, double *__restrict__ r,
                                                            double *__restrict__ p, double *__restrict__ rhat, double *partials,
                                                            unsigned int *counter, double *scal, int slot_rz) {
    double acc[3] = {0.0, 0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        const double bi = b[i], ri = bi - q[i];
        r[i] = ri;
        const double zi = dinv[i] * ri;
        if (p) p[i] = zi;
        if (rhat) rhat[i] = ri;
        acc[0] += bi * bi;
        acc[1] += ri * ri;
        acc[2] += rhat ? ri * ri : ri * zi;
    }
    const int sl_[3] = {S_BB, S_RR, slot_rz};
    grid_reduce<3>(acc, partials, counter, scal, sl_, false);
}

// CG: alpha = rz/pq ; x += alpha p ; r -= alpha q ; rz' = r.(dinv r) ; rr = r.r ; iter++
__global__ void __launch_bounds__(kThreads) k_cg_update(int64_t n, const double *__restrict__ p, const double *__restrict__ q,
                                                        const double *__restrict__ dinv, double *__restrict__ x,
                                                        double *__restrict__ r, double *partials, unsigned int *counter,
                                                        double *scal, int rz_cur, int rz_nxt) {
    if (!solver_active(scal)) return;
    const double alpha = safe_div(scal[rz_cur], scal[S_PQ]);
    double acc[2] = {0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        acc[0] += ri * ri * dinv[i];
        acc[1] += ri * ri;
    }
    const int sl_[2] = {rz_nxt, S_RR};
    grid_reduce<2>(acc, partials, counter, scal, sl_, true);
}

// CG: beta = rz'/rz ; p = dinv r + beta p
__global__ void __launch_bounds__(kThreads) k_cg_pupdate(int64_t n, const double *__restrict__ r, const double *__restrict__ dinv,
                                                         double *__restrict__ p, const double *scal, int rz_cur, int rz_nxt) {
    if (!solver_active(scal)) return;
    const double beta = safe_div(scal[rz_nxt], scal[rz_cur]);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x)
        p[i] = dinv[i] * r[i] + beta * p[i];
}

// BiCGStab A: beta = rho'/rv * tt/ts ; omega = ts/tt ; p = r + beta (p - omega v) ; y = dinv p
__global__ void __launch_bounds__(kThreads) k_bicg_p(int64_t n, int first, const double *__restrict__ r, const double *__restrict__ v,
                                                     const double *__restrict__ dinv, double *__restrict__ p, double *__restrict__ y,
                                                     const double *scal, int rho_cur) {
    if (!solver_active(scal)) return;
    const double omega = first ? 0.0 : safe_div(scal[S_TS], scal[S_TT]);
    const double beta = first ? 0.0 : safe_div(scal[rho_cur], scal[S_RV]) * safe_div(scal[S_TT], scal[S_TS]);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        const double pi = first ? r[i] : r[i] + beta * (p[i] - omega * v[i]);
        p[i] = pi;
        y[i] = dinv[i] * pi;
    }
}

// BiCGStab C: alpha = rho/rv ; s = r - alpha v ; z = dinv s
__global__ void __launch_bounds__(kThreads) k_bicg_s(int64_t n, const double *__restrict__ r, const double *__restrict__ v,
                                                     const double *__restrict__ dinv, double *__restrict__ s, double *__restrict__ z,
                                                     const double *scal, int rho_cur) {
    if (!solver_active(scal)) return;
    const double alpha = safe_div(scal[rho_cur], scal[S_RV]);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        const double si = r[i] - alpha * v[i];
        s[i] = si;
        z[i] = dinv[i] * si;
    }
}

// BiCGStab E: omega = ts/tt ; x += alpha y + omega z ; r = s - omega t ; rho' = rhat.r ; rr = r.r ; iter++
__global__ void __launch_bounds__(kThreads) k_bicg_x(int64_t n, const double *__restrict__ y, const double *__restrict__ z,
                                                     const double *__restrict__ s, const double *__restrict__ t,
                                                     const double *__restrict__ rhat, double *__restrict__ x, double *__restrict__ r,
                                                     double *partials, unsigned int *counter, double *scal, int rho_cur, int rho_nxt) {
    if (!solver_active(scal)) return;
    const double alpha = safe_div(scal[rho_cur], scal[S_RV]);
    const double omega = safe_div(scal[S_TS], scal[S_TT]);
    double acc[2] = {0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        x[i] += alpha * y[i] + omega * z[i];
        const double ri = s[i] - omega * t[i];
        r[i] = ri;
        acc[0] += rhat[i] * ri;
        acc[1] += ri * ri;
    }
    const int sl_[2] = {rho_nxt, S_RR};
    grid_reduce<2>(acc, partials, counter, scal, sl_, true);
}

int grid_for(const mp_ctx *ctx, int64_t n_threads_wanted) {
    int64_t g = cdiv(n_threads_wanted, kThreads);
    int64_t cap = (int64_t)ctx->num_sm * 8;  // 8 x 256 threads = 2048 resident threads per SM
    if (cap > kMaxBlocks) cap = kMaxBlocks;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

int pick_lanes(const mp_csr *A, int64_t nnz_hint) {
    double avg = A->nrow > 0 ? (double)nnz_hint / (double)A->nrow : 1.0;
    if (avg <= 3.0) return 2;
    if (avg <= 6.0) return 4;
    if (avg <= 24.0) return 8;
    if (avg <= 48.0) return 16;
    return 32;
}

template <int DOT, bool GUARD>
int launch_spmv(mp_ctx *ctx, const mp_csr *A, int lanes, const double *x, double *y, const double *w, int slot0, int slot1) {
    if (A->nrow == 0) return 0;
    int grid = grid_for(ctx, A->nrow * lanes);
#define L(S_)                                                                                                           \
    k_spmv<S_, DOT, GUARD><<<grid, kThreads, 0, ctx->stream>>>(A->nrow, A->d_rowptr, A->d_colidx, A->d_vals, x, y, w,  \
                                                                ctx->d_partials, ctx->d_counter, ctx->d_scalars, slot0, slot1)
    switch (lanes) {
        case 2: L(2); break;
        case 4: L(4); break;
        case 8: L(8); break;
        case 16: L(16); break;
        default: L(32); break;
    }
#undef L
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return 0;
}

int nnz_of(mp_ctx *ctx, const mp_csr *A, int64_t *nnz) {
    (void)ctx;
    *nnz = A->nnz;
    return 0;
}

int poll(mp_ctx *ctx) {
    MP_CUDA(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, ctx->stream));
    MP_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int check_csr(const mp_csr *A) {
    MP_CHECK(A && A->d_rowptr && A->d_colidx && A->d_vals, "csr: null pointer");
    MP_CHECK(A->nrow >= 0 && A->ncol >= A->nrow, "csr: need ncol >= nrow >= 0 (owned rows first, then ghosts)");
    return 0;
}

// set tolerance and initial scalars from the host after the one synchronising poll at solve start
int start_solve(mp_ctx *ctx, mp_solve_info *info, bool *done) {
    MP_TRY(mp_allreduce_sum(ctx, ctx->d_scalars + S_BB, 1));
    MP_TRY(mp_allreduce_sum(ctx, ctx->d_scalars + S_RR, 1));
    MP_TRY(poll(ctx));
    double bb = ctx->h_scalars[S_BB], rr = ctx->h_scalars[S_RR];
    double tol = info->rtol * sqrt(bb);
    if (info->atol > tol) tol = info->atol;
    double hs[3] = {bb, tol * tol, 0.0};  // S_BB, S_TOL2, S_ITER are contiguous; S_RR is already on the device
    memcpy(ctx->h_scalars, hs, sizeof(hs));
    MP_CUDA(cudaMemcpyAsync(ctx->d_scalars, ctx->h_scalars, sizeof(hs), cudaMemcpyHostToDevice, ctx->stream));
    MP_CUDA(cudaStreamSynchronize(ctx->stream));
    info->niter = 0;
    info->relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    *done = !(rr > tol * tol);
    info->converged = *done ? 1 : 0;
    return 0;
}

int finish_solve(mp_ctx *ctx, mp_solve_info *info) {
    MP_TRY(poll(ctx));
    const double bb = ctx->h_scalars[S_BB], rr = ctx->h_scalars[S_RR];
    info->niter = (int)ctx->h_scalars[S_ITER];
    info->relres = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr);
    info->converged = (rr <= ctx->h_scalars[S_TOL2]) ? 1 : 0;
    return 0;
}

}  // namespace

extern "C" int mp_spmv(mp_ctx *ctx, const mp_csr *A, const double *d_x, double *d_y) {
    MP_CHECK(ctx && d_x && d_y, "mp_spmv: null argument");
    MP_TRY(check_csr(A));
    int64_t nnz = 0;
    MP_TRY(nnz_of(ctx, A, &nnz));
    return launch_spmv<0, false>(ctx, A, pick_lanes(A, nnz), d_x, d_y, nullptr, 0, 0);
}

extern "C" int mp_axpby(mp_ctx *ctx, int64_t n, double a, const double *d_x, double b, double *d_y) {
    MP_CHECK(ctx && d_x && d_y && n >= 0, "mp_axpby: bad argument");
    if (n == 0) return 0;
    k_axpby<<<grid_for(ctx, n), kThreads, 0, ctx->stream>>>(n, a, d_x, b, d_y);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mp_dot(mp_ctx *ctx, int64_t n, const double *d_x, const double *d_y, double *result) {
    MP_CHECK(ctx && d_x && d_y && result && n >= 0, "mp_dot: bad argument");
    k_dot<<<grid_for(ctx, n > 0 ? n : 1), kThreads, 0, ctx->stream>>>(n, d_x, d_y, ctx->d_partials, ctx->d_counter, ctx->d_scalars, S_TMP0);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    MP_TRY(mp_allreduce_sum(ctx, ctx->d_scalars + S_TMP0, 1));
    MP_CUDA(cudaMemcpyAsync(ctx->h_scalars + S_TMP0, ctx->d_scalars + S_TMP0, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MP_CUDA(cudaStreamSynchronize(ctx->stream));
    *result = ctx->h_scalars[S_TMP0];
    return 0;
}

extern "C" int mp_jacobi_setup(mp_ctx *ctx, const mp_csr *A, double *d_dinv) {
    MP_CHECK(ctx && d_dinv, "mp_jacobi_setup: null argument");
    MP_TRY(check_csr(A));
    if (A->nrow == 0) return 0;
    k_jacobi_setup<<<(unsigned)cdiv(A->nrow, kThreads), kThreads, 0, ctx->stream>>>(A->nrow, A->d_rowptr, A->d_colidx, A->d_vals, d_dinv);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mp_solve_pcg(mp_ctx *ctx, const mp_csr *A, const double *d_dinv, const double *d_b, double *d_x, mp_halo *halo,
                            mp_solve_info *info) {
    MP_CHECK(ctx && d_dinv && d_b && d_x && info, "mp_solve_pcg: null argument");
    MP_TRY(check_csr(A));
    const int64_t n = A->nrow, nc = A->ncol;
    int64_t nnz = 0;
    MP_TRY(nnz_of(ctx, A, &nnz));
    const int lanes = pick_lanes(A, nnz);
    double *w = nullptr;
    MP_TRY(mp::ws_get(ctx, 0, sizeof(double) * (size_t)(2 * n + nc + 8), (void **)&w));
    double *r = w, *q = w + n, *p = w + 2 * n;  // p has ghost entries
    const int g1 = grid_for(ctx, n);
    // r = b - A x
    MP_TRY(mp_halo_exchange(ctx, halo, d_x));
    MP_TRY((launch_spmv<0, false>(ctx, A, lanes, d_x, q, nullptr, 0, 0)));
    k_init_residual<<<g1, kThreads, 0, ctx->stream>>>(n, d_b, q, d_dinv, r, p, nullptr, ctx->d_partials, ctx->d_counter, ctx->d_scalars, S_RZ0);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    MP_TRY(mp_allreduce_sum(ctx, ctx->d_scalars + S_RZ0, 1));
    bool done = false;
    MP_TRY(start_solve(ctx, info, &done));
    if (done) return 0;
    int every = info->check_every > 0 ? info->check_every : (n >= (1 << 20) ? 10 : 50);
    int cur = S_RZ0, nxt = S_RZ1;
    for (int it = 0; it < info->maxit; ++it) {
        MP_TRY(mp_halo_exchange(ctx, halo, p));
        MP_TRY((launch_spmv<1, true>(ctx, A, lanes, p, q, p, S_PQ, 0)));
        MP_TRY(mp_allreduce_sum(ctx, ctx->d_scalars + S_PQ, 1));
        k_cg_update<<<g1, kThreads, 0, ctx->stream>>>(n, p, q, d_dinv, d_x, r, ctx->d_partials, ctx->d_counter, ctx->d_scalars, cur, nxt);
        MP_TRY(mp_allreduce_sum(ctx, ctx->d_scalars + (nxt < S_RR ? nxt : S_RR), 2));
        k_cg_pupdate<<<g1, kThreads, 0, ctx->stream>>>(n, r, d_dinv, p, ctx->d_scalars, cur, nxt);
        mp::count_launch(ctx, 2);
        int tmp = cur; cur = nxt; nxt = tmp;
        if ((it + 1) % every == 0 || it + 1 == info->maxit) {
            MP_CUDA(cudaGetLastError());
            MP_TRY(poll(ctx));
            if (!(ctx->h_scalars[S_RR] > ctx->h_scalars[S_TOL2])) break;
            if (!(ctx->h_scalars[S_RR] == ctx->h_scalars[S_RR])) { mp::set_error("mp_solve_pcg: NaN residual at iteration %d", it + 1); return 4; }
        }
    }
    return finish_solve(ctx, info);
}

extern "C" int mp_solve_bicgstab(mp_ctx *ctx, const mp_csr *A, const double *d_dinv, const double *d_b, double *d_x, mp_halo *halo,
                                 mp_solve_info *info) {
    MP_CHECK(ctx && d_dinv && d_b && d_x && info, "mp_solve_bicgstab: null argument");
    MP_TRY(check_csr(A));
    const int64_t n = A->nrow, nc = A->ncol;
    int64_t nnz = 0;
    MP_TRY(nnz_of(ctx, A, &nnz));
    const int lanes = pick_lanes(A, nnz);
    double *w = nullptr;
    MP_TRY(mp::ws_get(ctx, 0, sizeof(double) * (size_t)(6 * n + 2 * nc + 8), (void **)&w));
    double *r = w, *rhat = w + n, *p = w + 2 * n, *v = w + 3 * n, *s = w + 4 * n, *t = w + 5 * n, *y = w + 6 * n, *z = w + 6 * n + nc;
    const int g1 = grid_for(ctx, n);
    MP_TRY(mp_halo_exchange(ctx, halo, d_x));
    MP_TRY((launch_spmv<0, false>(ctx, A, lanes, d_x, v, nullptr, 0, 0)));
    k_init_residual<<<g1, kThreads, 0, ctx->stream>>>(n, d_b, v, d_dinv, r, nullptr, rhat, ctx->d_partials, ctx->d_counter, ctx->d_scalars, S_RHO0);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    MP_TRY(mp_allreduce_sum(ctx, ctx->d_scalars + S_RHO0, 1));
    bool done = false;
    MP_TRY(start_solve(ctx, info, &done));
    if (done) return 0;
    int every = info->check_every > 0 ? info->check_every : (n >= (1 << 20) ? 5 : 25);
    int cur = S_RHO0, nxt = S_RHO1;
    for (int it = 0; it < info->maxit; ++it) {
        k_bicg_p<<<g1, kThreads, 0, ctx->stream>>>(n, it == 0 ? 1 : 0, r, v, d_dinv, p, y, ctx->d_scalars, cur);
        MP_TRY(mp_halo_exchange(ctx, halo, y));
        MP_TRY((launch_spmv<1, true>(ctx, A, lanes, y, v, rhat, S_RV, 0)));
        MP_TRY(mp_allreduce_sum(ctx, ctx->d_scalars + S_RV, 1));
        k_bicg_s<<<g1, kThreads, 0, ctx->stream>>>(n, r, v, d_dinv, s, z, ctx->d_scalars, cur);
        MP_TRY(mp_halo_exchange(ctx, halo, z));
        MP_TRY((launch_spmv<2, true>(ctx, A, lanes, z, t, s, S_TS, S_TT)));
        MP_TRY(mp_allreduce_sum(ctx, ctx->d_scalars + S_TS, 2));
        k_bicg_x<<<g1, kThreads, 0, ctx->stream>>>(n, y, z, s, t, rhat, d_x, r, ctx->d_partials, ctx->d_counter, ctx->d_scalars, cur, nxt);
        MP_TRY(mp_allreduce_sum(ctx, ctx->d_scalars + (nxt < S_RR ? nxt : S_RR), 2));
        mp::count_launch(ctx, 3);
        int tmp = cur; cur = nxt; nxt = tmp;
        if ((it + 1) % every == 0 || it + 1 == info->maxit) {
            MP_CUDA(cudaGetLastError());
            MP_TRY(poll(ctx));
            if (!(ctx->h_scalars[S_RR] > ctx->h_scalars[S_TOL2])) break;
            if (!(ctx->h_scalars[S_RR] == ctx->h_scalars[S_RR])) { mp::set_error("mp_solve_bicgstab: NaN residual at iteration %d", it + 1); return 4; }
        }
    }
    return finish_solve(ctx, info);
}
