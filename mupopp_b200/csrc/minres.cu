// Device-resident preconditioned MINRES over a BLOCK operator, with a block-diagonal Chebyshev-Jacobi preconditioner.
//
// Replaces dolfin.KrylovSolver("minres", "amg") on the systems produced by assemble_system(a, L, bcs) / assemble_system(b, L, bcs) in
// compaction.momentum_solver (/root/reference/src/modules/mupopp.py:383-429), the MUMPS solve of the Stokes block of
// stokes_ADR_precipitation (/root/reference/src/run/stokes_ADR_3D/master_advective_stokes.py:213) and the mixed Darcy saddle system of
// the multi-component forms (mupopp.py:804-807): symmetric indefinite systems whose operator is a sum of CSR blocks
//     y[rb] += scale * A_t x[cb]      (+ the symmetric Dirichlet elimination  y = M A M x + (I - M) x  through a 0/1 mask).
// Round 1 sequenced this recurrence from Python with two blocking host dot products per iteration.  Here every scalar of the Lanczos
// / Givens recurrence lives in a ping-pong pair of device structs (iteration k reads set k&1 and writes set (k+1)&1, so a late
// thread block can never see a half-updated state); the host only enqueues kernels and polls every `check_every` iterations.
// The preconditioner applies, per diagonal block, z = p_k(D^-1 P) D^-1 v with the Chebyshev polynomial of degree k on
// [lmin, lmax] (k = 1: Jacobi) -- a fixed SPD operator, as MINRES requires -- using the reference's own preconditioner form `b`
// (mupopp.py:379-380) for P.  The reference hands that form to hypre AMG; AMG is outside BASELINE.json's list ("Jacobi or Chebyshev").
#include <math.h>
#include <string.h>

#include "reduce.cuh"

namespace {

struct MrState {          // one set of recurrence scalars (device)
    double gamma, gamma_old, eta, s, s_old, c, c_old, rr;     // rr = eta^2 (the convergence test), iteration count in `iter`
    double tol2, iter, delta, gnew2, bnorm, pad[3];
};

__device__ __forceinline__ bool mr_active(const MrState *st) { return st->rr > st->tol2; }

// y = beta y + alpha A x, CSR with S lanes per row; guarded by the solver state (st == nullptr: always runs)
template <int S>
__global__ void __launch_bounds__(kThreads)
k_csr_axpby(int64_t nrow, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, const double *__restrict__ vals,
            const double *__restrict__ x, double *__restrict__ y, double alpha, double beta, const MrState *st) {
    if (st && !mr_active(st)) return;
    constexpr int RPW = 32 / S;
    const int lane = threadIdx.x & 31, sub = lane / S, sl = lane % S;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t base = warp * RPW; base < nrow; base += nwarp * RPW) {
        const int64_t row = base + sub;
        double sum = 0.0;
        if (row < nrow) {
            const int32_t b = rowptr[row], e = rowptr[row + 1];
            for (int32_t k = b + sl; k < e; k += S) sum += vals[k] * x[colidx[k]];
        }
#pragma unroll
        for (int o = S / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (row < nrow && sl == 0) y[row] = beta == 0.0 ? alpha * sum : beta * y[row] + alpha * sum;
    }
}

// z = a * x .* m  (m == nullptr: z = a x);  z may alias x
__global__ void __launch_bounds__(kThreads) k_mr_scale(int64_t n, double a, const double *x, const double *__restrict__ m, double *z,
                                                       const MrState *st, int use_inv_gamma) {
    if (st && !mr_active(st)) return;
    const double f = use_inv_gamma ? a / st->gamma : a;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x)
        z[i] = m ? f * x[i] * m[i] : f * x[i];
}

// Dirichlet pass-through of the masked operator: y = m .* y + (x - xm)
__global__ void __launch_bounds__(kThreads) k_mr_unmask(int64_t n, const double *__restrict__ m, const double *__restrict__ x,
                                                        const double *__restrict__ xm, double *__restrict__ y, const MrState *st) {
    if (st && !mr_active(st)) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x)
        y[i] = m[i] * y[i] + (x[i] - xm[i]);
}

// scal[S_TMP0] <- x . y   (deterministic two-stage reduction; consumed by the kernels that follow on the stream)
__global__ void __launch_bounds__(kThreads) k_mr_dot(int64_t n, const double *__restrict__ x, const double *__restrict__ y, double *partials,
                                                     mp::ReduceCtl *ctl, double *scal, const MrState *st) {
    if (!mr_active(st)) return;
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) acc[0] += x[i] * y[i];
    const int sl_[1] = {S_TMP0};
    grid_reduce<1>(acc, partials, ctl, scal, sl_, false);
}

// v_new = Az - (delta/gamma) v - (gamma/gamma_old) v_old      (delta = scal[S_TMP0] of the preceding dot)
__global__ void __launch_bounds__(kThreads) k_mr_lanczos(int64_t n, const double *__restrict__ Az, const double *__restrict__ v,
                                                         const double *__restrict__ v_old, double *__restrict__ v_new, const MrState *st,
                                                         const double *scal) {
    if (!mr_active(st)) return;
    const double a = scal[S_TMP0] / st->gamma, b = st->gamma / st->gamma_old;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x)
        v_new[i] = Az[i] - a * v[i] - b * v_old[i];
}

// Chebyshev-Jacobi block preconditioner: first step z = d = (1/theta) dinv v (.* mask)
__global__ void __launch_bounds__(kThreads) k_mr_cheb_first(int64_t n, double inv_theta, const double *__restrict__ v, const double *__restrict__ dinv,
                                                            const double *__restrict__ m, double *__restrict__ z, double *__restrict__ d,
                                                            const MrState *st) {
    if (st && !mr_active(st)) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        double val = inv_theta * dinv[i] * v[i];
        if (m) val *= m[i];
        z[i] = val;
        if (d) d[i] = val;
    }
}

// d = c1 d + c2 dinv (v - t) (.* mask) ; z += d        (t = P z by the caller)
__global__ void __launch_bounds__(kThreads) k_mr_cheb_step(int64_t n, double c1, double c2, const double *__restrict__ v, const double *__restrict__ t,
                                                           const double *__restrict__ dinv, const double *__restrict__ m, double *__restrict__ z,
                                                           double *__restrict__ d, const MrState *st) {
    if (st && !mr_active(st)) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        double dn = c1 * d[i] + c2 * dinv[i] * (v[i] - t[i]);
        if (m) dn *= m[i];
        d[i] = dn;
        z[i] += dn;
    }
}

// Givens update and the solution update; thread 0 of block 0 writes the NEXT state set.
//   gamma_new = sqrt(z_new . v_new) ; a0..a3 ; w_new = (z - a3 w_old - a2 w)/a1 ; x += c_new eta w_new ; eta' = -s_new eta
__global__ void __launch_bounds__(kThreads) k_mr_update(int64_t n, const double *__restrict__ z, const double *__restrict__ w_old,
                                                        const double *__restrict__ w, double *__restrict__ w_new, double *__restrict__ x,
                                                        const MrState *cur, MrState *nxt, const double *scal) {
    if (!mr_active(cur)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *nxt = *cur;     // keep the converged state alive in both sets
        return;
    }
    const double gamma = cur->gamma, delta = cur->delta;
    const double gnew = sqrt(fmax(scal[S_TMP0], 0.0));             // z_new . v_new of the dot issued just before this kernel
    const double a0 = cur->c * delta - cur->c_old * cur->s * gamma;
    const double a1 = sqrt(a0 * a0 + gnew * gnew);
    const double a2 = cur->s * delta + cur->c_old * cur->c * gamma;
    const double a3 = cur->s_old * gamma;
    const double inv1 = a1 > 0.0 ? 1.0 / a1 : 0.0;
    const double cn = a0 * inv1, sn = gnew * inv1;
    const double step = cn * cur->eta;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        const double wn = (z[i] - a3 * w_old[i] - a2 * w[i]) * inv1;
        w_new[i] = wn;
        x[i] += step * wn;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        MrState o = *cur;
        o.gamma_old = gamma;
        o.gamma = gnew;
        o.s_old = cur->s; o.s = sn;
        o.c_old = cur->c; o.c = cn;
        o.eta = -sn * cur->eta;
        o.rr = gnew > 0.0 ? o.eta * o.eta : 0.0;                   // gamma_new == 0: the Krylov space is exhausted, the iterate is exact
        o.iter = cur->iter + 1.0;
        *nxt = o;
    }
}

// copy the freshly reduced delta (scal[S_TMP0]) into the current state (single thread; runs between the dot and its consumers)
__global__ void k_mr_store_delta(MrState *cur, const double *scal) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && mr_active(cur)) cur->delta = scal[S_TMP0];
}

int grid_n(const mp_ctx *ctx, int64_t n) {
    int64_t g = cdiv(n > 0 ? n : 1, kThreads);
    int64_t cap = (int64_t)ctx->num_sm * 8;
    if (cap > kMaxBlocks) cap = kMaxBlocks;
    return (int)(g < cap ? g : cap);
}

int csr_axpby(mp_ctx *ctx, const mp_csr *A, const double *x, double *y, double alpha, double beta, const MrState *st) {
    if (A->nrow == 0) return 0;
    const double avg = A->nrow > 0 ? (double)A->nnz / (double)A->nrow : 1.0;
    const int lanes = avg <= 3.0 ? 2 : (avg <= 6.0 ? 4 : (avg <= 24.0 ? 8 : (avg <= 48.0 ? 16 : 32)));
    const int g = grid_n(ctx, A->nrow * lanes);
#define L(S_) k_csr_axpby<S_><<<g, kThreads, 0, ctx->stream>>>(A->nrow, A->d_rowptr, A->d_colidx, A->d_vals, x, y, alpha, beta, st)
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

int check_op(const mp_block_op *op) {
    MP_CHECK(op && op->nblk >= 1 && op->nblk <= 8 && op->nterm >= 1 && op->terms, "block operator: bad descriptor");
    for (int b = 0; b < op->nblk; ++b) MP_CHECK(op->off[b + 1] >= op->off[b], "block operator: offsets must ascend");
    for (int t = 0; t < op->nterm; ++t) {
        const mp_block_term &T = op->terms[t];
        MP_CHECK(T.rb >= 0 && T.rb < op->nblk && T.cb >= 0 && T.cb < op->nblk && T.A, "block operator: bad term %d", t);
        MP_CHECK(T.A->nrow == op->off[T.rb + 1] - op->off[T.rb] && T.A->ncol == op->off[T.cb + 1] - op->off[T.cb],
                 "block operator: term %d is %lld x %lld but its blocks are %lld x %lld", t, (long long)T.A->nrow, (long long)T.A->ncol,
                 (long long)(op->off[T.rb + 1] - op->off[T.rb]), (long long)(op->off[T.cb + 1] - op->off[T.cb]));
        MP_CHECK(T.A->d_rowptr && T.A->d_colidx && T.A->d_vals, "block operator: term %d has null CSR arrays", t);
    }
    return 0;
}

// y = (M A M + I - M) x ; xm: scratch of size n (only used with a mask)
int block_apply(mp_ctx *ctx, const mp_block_op *op, const double *x, double *y, double *xm, const MrState *st) {
    const int64_t n = op->off[op->nblk];
    const double *src = x;
    if (op->d_mask) {
        k_mr_scale<<<grid_n(ctx, n), kThreads, 0, ctx->stream>>>(n, 1.0, x, op->d_mask, xm, st, 0);
        mp::count_launch(ctx);
        src = xm;
    }
    bool touched[8] = {false, false, false, false, false, false, false, false};
    for (int t = 0; t < op->nterm; ++t) {
        const mp_block_term &T = op->terms[t];
        MP_TRY(csr_axpby(ctx, T.A, src + op->off[T.cb], y + op->off[T.rb], T.scale, touched[T.rb] ? 1.0 : 0.0, st));
        touched[T.rb] = true;
    }
    for (int b = 0; b < op->nblk; ++b)
        if (!touched[b] && op->off[b + 1] > op->off[b])
            MP_CUDA(cudaMemsetAsync(y + op->off[b], 0, sizeof(double) * (size_t)(op->off[b + 1] - op->off[b]), ctx->stream));
    if (op->d_mask) {
        k_mr_unmask<<<grid_n(ctx, n), kThreads, 0, ctx->stream>>>(n, op->d_mask, x, xm, y, st);
        mp::count_launch(ctx);
    }
    MP_CUDA(cudaGetLastError());
    return 0;
}

// z = P^-1 v, block by block: Jacobi, or the degree-k Chebyshev polynomial of D^-1 P_b on [lmin, lmax]
int block_prec(mp_ctx *ctx, const mp_block_op *op, const mp_block_prec *P, const double *v, double *z, double *d, double *t, const MrState *st) {
    for (int b = 0; b < op->nblk; ++b) {
        const int64_t o = op->off[b], nb = op->off[b + 1] - o;
        if (nb == 0) continue;
        const double *m = op->d_mask ? op->d_mask + o : nullptr;
        const int g = grid_n(ctx, nb);
        const int deg = P->P[b] ? P->degree[b] : 1;
        if (deg <= 1) {
            // Dirichlet dofs carry dinv = 1 and a zero residual, so no mask is needed for plain Jacobi
            k_mr_cheb_first<<<g, kThreads, 0, ctx->stream>>>(nb, 1.0, v + o, P->d_dinv + o, nullptr, z + o, nullptr, st);
            mp::count_launch(ctx);
            continue;
        }
        const double lmin = P->lmin[b], lmax = P->lmax[b];
        const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
        k_mr_cheb_first<<<g, kThreads, 0, ctx->stream>>>(nb, 1.0 / theta, v + o, P->d_dinv + o, m, z + o, d + o, st);
        mp::count_launch(ctx);
        double rho0 = 1.0 / sigma;
        for (int i = 1; i < deg; ++i) {
            const double rho1 = 1.0 / (2.0 * sigma - rho0);
            MP_TRY(csr_axpby(ctx, P->P[b], z + o, t + o, 1.0, 0.0, st));
            k_mr_cheb_step<<<g, kThreads, 0, ctx->stream>>>(nb, rho1 * rho0, 2.0 * rho1 / delta, v + o, t + o, P->d_dinv + o, m, z + o, d + o, st);
            mp::count_launch(ctx);
            rho0 = rho1;
        }
    }
    MP_CUDA(cudaGetLastError());
    return 0;
}

int host_dot(mp_ctx *ctx, int64_t n, const double *x, const double *y, double *out) { return mp_dot(ctx, n, x, y, out); }

}  // namespace

extern "C" int mp_block_apply(mp_ctx *ctx, const mp_block_op *op, const double *d_x, double *d_y) {
    MP_CHECK(ctx && d_x && d_y, "mp_block_apply: null argument");
    MP_TRY(check_op(op));
    const int64_t n = op->off[op->nblk];
    double *xm = nullptr;
    MP_TRY(mp::ws_get(ctx, 3, sizeof(double) * (size_t)(n + 8), (void **)&xm));
    return block_apply(ctx, op, d_x, d_y, xm, nullptr);
}

extern "C" int mp_solve_minres(mp_ctx *ctx, const mp_block_op *op, const mp_block_prec *prec, const double *d_b, double *d_x, mp_solve_info *info) {
    MP_CHECK(ctx && prec && d_b && d_x && info && prec->d_dinv, "mp_solve_minres: null argument");
    MP_CHECK(ctx->nranks == 1, "mp_solve_minres: the block operators are single-GPU (the partitioned path is mode F)");
    MP_TRY(check_op(op));
    for (int b = 0; b < op->nblk; ++b)
        if (prec->P[b] && prec->degree[b] > 1) {
            MP_CHECK(prec->degree[b] <= 32 && prec->lmin[b] > 0.0 && prec->lmax[b] > prec->lmin[b], "mp_solve_minres: block %d needs 0 < lmin < lmax", b);
            MP_CHECK(prec->P[b]->nrow == op->off[b + 1] - op->off[b] && prec->P[b]->ncol == prec->P[b]->nrow, "mp_solve_minres: preconditioner block %d has the wrong size", b);
        }
    const int64_t n = op->off[op->nblk];
    double *wk = nullptr;
    MP_TRY(mp::ws_get(ctx, 2, sizeof(double) * (size_t)(12 * n + 64), (void **)&wk));
    double *v_old = wk, *v = wk + n, *v_new = wk + 2 * n, *z = wk + 3 * n, *z_new = wk + 4 * n, *w_old = wk + 5 * n, *w = wk + 6 * n,
           *w_new = wk + 7 * n, *Az = wk + 8 * n, *xm = wk + 9 * n, *t = wk + 10 * n, *d = wk + 11 * n;
    MrState *st = reinterpret_cast<MrState *>(wk + 12 * n);          // two sets (2 x 16 doubles)
    static_assert(sizeof(MrState) == 16 * sizeof(double), "MrState layout");
    MrState h[2];
    const int g = grid_n(ctx, n);
    constexpr int kMrGraphIters = 6;
    bool use_graph = ctx->minres_graph != 0;
    const int every = info->check_every > 0 ? info->check_every : (use_graph ? 4 * kMrGraphIters : 25);
    struct GraphHolder {                 // released on every exit path of the solve (the MP_TRY / MP_CUDA macros return early)
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaStream_t stream = nullptr;
        void release() {
            if (exec) cudaGraphExecDestroy(exec);
            if (graph) cudaGraphDestroy(graph);
            if (stream) cudaStreamDestroy(stream);
            exec = nullptr; graph = nullptr; stream = nullptr;
        }
        ~GraphHolder() { release(); }
    } gh;
    cudaGraph_t &graph = gh.graph;
    cudaGraphExec_t &gexec = gh.exec;
    cudaStream_t &cstream = gh.stream;
    int64_t per_graph = 0;
    auto release_graph = [&]() { gh.release(); };
    int total = 0;
    double bnorm = 0.0, relres = 0.0;
    bool converged = false;
    for (int restart = 0; restart < 4 && !converged && total < info->maxit; ++restart) {
        // (re)start from the TRUE residual v = b - A x ; z = P^-1 v ; gamma = sqrt(z.v)
        MP_TRY(block_apply(ctx, op, d_x, Az, xm, nullptr));
        MP_CUDA(cudaMemcpyAsync(v, d_b, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
        MP_TRY(mp_axpby(ctx, n, -1.0, Az, 1.0, v));
        MP_TRY(block_prec(ctx, op, prec, v, z, d, t, nullptr));
        double zv = 0.0;
        MP_TRY(host_dot(ctx, n, z, v, &zv));
        if (restart == 0) {
            MP_TRY(block_prec(ctx, op, prec, d_b, Az, d, t, nullptr));
            double bb = 0.0;
            MP_TRY(host_dot(ctx, n, Az, d_b, &bb));
            bnorm = sqrt(fmax(bb, 1e-300));
        }
        const double gamma = sqrt(fmax(zv, 0.0));
        double tol = info->rtol * bnorm;
        if (info->atol > tol) tol = info->atol;
        relres = gamma / bnorm;
        if (!(gamma > tol)) { converged = true; break; }
        memset(h, 0, sizeof(h));
        h[0].gamma = gamma; h[0].gamma_old = 1.0; h[0].eta = gamma; h[0].s = h[0].s_old = 0.0; h[0].c = h[0].c_old = 1.0;
        h[0].rr = gamma * gamma; h[0].tol2 = tol * tol; h[0].iter = 0.0; h[0].bnorm = bnorm;
        h[1] = h[0];
        MP_CUDA(cudaMemcpyAsync(st, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
        MP_CUDA(cudaMemsetAsync(v_old, 0, sizeof(double) * (size_t)n, ctx->stream));
        MP_CUDA(cudaMemsetAsync(w_old, 0, sizeof(double) * (size_t)n, ctx->stream));
        MP_CUDA(cudaMemsetAsync(w, 0, sizeof(double) * (size_t)n, ctx->stream));
        int cur = 0, inner = 0, last_poll = total;
        bool inner_done = false;
        // one iteration of the recurrence on the current pointer configuration (v's rotate with period 3, z and the state set with period 2)
        auto one_iteration = [&]() -> int {
            MrState *sc = st + cur, *sn = st + (cur ^ 1);
            k_mr_scale<<<g, kThreads, 0, ctx->stream>>>(n, 1.0, z, nullptr, z, sc, 1);                  // z <- z / gamma
            MP_TRY(block_apply(ctx, op, z, Az, xm, sc));
            k_mr_dot<<<g, kThreads, 0, ctx->stream>>>(n, Az, z, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, sc);   // delta -> scal[S_TMP0]
            k_mr_store_delta<<<1, 32, 0, ctx->stream>>>(sc, ctx->d_scalars);
            k_mr_lanczos<<<g, kThreads, 0, ctx->stream>>>(n, Az, v, v_old, v_new, sc, ctx->d_scalars);
            mp::count_launch(ctx, 4);
            MP_TRY(block_prec(ctx, op, prec, v_new, z_new, d, t, sc));
            k_mr_dot<<<g, kThreads, 0, ctx->stream>>>(n, z_new, v_new, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, sc);   // z_new.v_new
            k_mr_update<<<g, kThreads, 0, ctx->stream>>>(n, z, w_old, w, w_new, d_x, sc, sn, ctx->d_scalars);
            mp::count_launch(ctx, 2);
            double *tp = v_old; v_old = v; v = v_new; v_new = tp;
            tp = w_old; w_old = w; w = w_new; w_new = tp;
            tp = z; z = z_new; z_new = tp;
            cur ^= 1;
            return 0;
        };
        while (total < info->maxit && !inner_done) {
            // An iteration is ~100 launches of kernels that run a few microseconds each on these block systems (3 x 10^4 .. 10^6 rows): the
            // recurrence is launch bound.  Six iterations bring every rotating pointer back to where it started, so they are captured ONCE
            // into a CUDA graph (on a side stream: the context's stream may be the legacy default stream, which cannot be captured) and
            // replayed; every kernel reads its scalars from device memory and is a no-op once converged, so the replay needs no host input.
            if (use_graph && !gexec && info->maxit - total >= kMrGraphIters) {
                cudaStream_t orig = ctx->stream;
                const int64_t l0 = ctx->launches;
                int rc = 0;
                if (!cstream && cudaStreamCreateWithFlags(&cstream, cudaStreamNonBlocking) != cudaSuccess) rc = 1;
                if (!rc && cudaStreamBeginCapture(cstream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) rc = 1;
                if (!rc) {
                    ctx->stream = cstream;
                    for (int k = 0; k < kMrGraphIters && !rc; ++k) rc = one_iteration();
                    ctx->stream = orig;
                    if (cudaStreamEndCapture(cstream, &graph) != cudaSuccess || !graph) rc = 1;
                }
                per_graph = ctx->launches - l0;
                ctx->launches = l0;                                      // captured, not run
                if (!rc && cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) rc = 1;
                if (rc) {                                                // no graph: plain launches (the pointers are back in place either way)
                    cudaGetLastError();
                    use_graph = false;
                    gexec = nullptr;
                }
            }
            if (use_graph && gexec && info->maxit - total >= kMrGraphIters) {
                MP_CUDA(cudaGraphLaunch(gexec, ctx->stream));
                ctx->launches += per_graph;
                total += kMrGraphIters;
                inner += kMrGraphIters;
            } else {
                MP_TRY(one_iteration());
                ++total;
                ++inner;
            }
            if (total - last_poll >= every || total >= info->maxit) {
                last_poll = total;
                MP_CUDA(cudaGetLastError());
                MP_CUDA(cudaMemcpyAsync(h, st, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
                MP_CUDA(cudaStreamSynchronize(ctx->stream));
                const MrState &sv = h[cur];
                if (!(sv.rr == sv.rr)) { release_graph(); mp::set_error("mp_solve_minres: NaN residual at iteration %d", total); return 4; }
                if (!(sv.rr > sv.tol2)) {
                    inner_done = true;
                    total = total - inner + (int)sv.iter;        // the recurrence stopped counting when it converged
                }
            }
        }
        if (inner % kMrGraphIters != 0) release_graph();         // the pointers did not end where the graph expects them: capture again
    }
    release_graph();
    if (!converged) {                                                // report the TRUE preconditioned residual of the final iterate
        MP_TRY(block_apply(ctx, op, d_x, Az, xm, nullptr));
        MP_CUDA(cudaMemcpyAsync(v, d_b, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
        MP_TRY(mp_axpby(ctx, n, -1.0, Az, 1.0, v));
        MP_TRY(block_prec(ctx, op, prec, v, z, d, t, nullptr));
        double zv = 0.0;
        MP_TRY(host_dot(ctx, n, z, v, &zv));
        relres = sqrt(fmax(zv, 0.0)) / bnorm;
        double tol = info->rtol * bnorm;
        if (info->atol > tol) tol = info->atol;
        converged = !(relres * bnorm > tol);
    }
    info->niter = total;
    info->relres = relres;
    info->converged = converged ? 1 : 0;
    return 0;
}

extern "C" int mp_schur_darcy_solve(mp_ctx *ctx, int32_t dim, const mp_csr *M, const mp_csr *const *BT, const mp_csr *const *B, const mp_csr *Sp,
                                    double K, const double *d_mask, const double *d_dinv, int32_t deg_u, double lmin_u, double lmax_u,
                                    int32_t deg_p, double lmin_p, double lmax_p, const double *d_b, double *d_x, mp_solve_info *info) {
    MP_CHECK(ctx && M && BT && B && Sp && d_dinv && d_b && d_x && info && (dim == 2 || dim == 3), "mp_schur_darcy_solve: bad argument");
    const int64_t n2 = M->nrow, n1 = Sp->nrow;
    mp_block_term terms[9];
    mp_block_op op;
    memset(&op, 0, sizeof(op));
    op.nblk = dim + 1;
    for (int a = 0; a <= dim; ++a) op.off[a] = a * n2;
    op.off[dim + 1] = dim * n2 + n1;
    int nt = 0;
    for (int a = 0; a < dim; ++a) {
        MP_CHECK(BT[a] && B[a], "mp_schur_darcy_solve: null divergence block");
        terms[nt++] = mp_block_term{a, a, M, 1.0};
        terms[nt++] = mp_block_term{a, dim, BT[a], -K};
        terms[nt++] = mp_block_term{dim, a, B[a], -K};
    }
    op.nterm = nt;
    op.terms = terms;
    op.d_mask = d_mask;
    mp_block_prec pr;
    memset(&pr, 0, sizeof(pr));
    for (int a = 0; a < dim; ++a) { pr.P[a] = M; pr.degree[a] = deg_u; pr.lmin[a] = lmin_u; pr.lmax[a] = lmax_u; }
    pr.P[dim] = Sp; pr.degree[dim] = deg_p; pr.lmin[dim] = lmin_p; pr.lmax[dim] = lmax_p;
    pr.d_dinv = d_dinv;
    return mp_solve_minres(ctx, &op, &pr, d_b, d_x, info);
}
