// Bandwidth-bound linear algebra: CSR SpMV (sub-warp per row), fused BLAS-1 with deterministic
// two-stage reductions, Jacobi-preconditioned CG and BiCGStab whose scalars (alpha, beta, omega,
// convergence test) live on the device, so the host only polls every `check_every` iterations.
//
// Replaces PETSc MatMult/VecDot/VecAXPY/KSP reached through dolfin.KrylovSolver
// (/root/reference/src/modules/mupopp.py:289-303, 410-425) and the sparse LU inside
// solve(a==L, ...) (/root/reference/src/run/CCS2D/CCS_3comp.py:261): LU returns the exact solution of
// the assembled system, which a tightly converged Krylov iteration reproduces.
#include <string.h>

#include "reduce.cuh"

namespace {

// ------------------------------------------------------------------ SpMV: S lanes per row
// DOT: 0 none | 1: slot0 = sum w[row]*y[row] | 2: slot0 = sum y*w, slot1 = sum y*y
template <int S, int DOT, bool GUARD>
__global__ void __launch_bounds__(kThreads)
k_spmv(int64_t nrow, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, const double *__restrict__ vals,
       const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ w, double *partials,
       mp::ReduceCtl *ctl, double *scal, int slot0, int slot1) {
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
    if constexpr (DOT == 1) { const int sl_[1] = {slot0}; grid_reduce<1>(acc, partials, ctl, scal, sl_, false); }
    if constexpr (DOT == 2) { const int sl_[2] = {slot0, slot1}; grid_reduce<2>(acc, partials, ctl, scal, sl_, false); }
}


// ------------------------------------------------------------------ SpMV, CSR-stream variant (short, regular rows: P1 operators)
// A CTA owns `rpb` consecutive rows.  Phase 1 streams the block's nonzeros fully coalesced (every load independent => deep
// memory-level parallelism) and parks the products val*x[col] in shared memory; phase 2 sums each row's segment in a fixed
// order with one thread per row.  Same DOT/GUARD modes as k_spmv.
template <int DOT, bool GUARD>
__global__ void __launch_bounds__(kThreads)
k_spmv_stream(int64_t nrow, int rpb, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
              const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ w,
              double *partials, mp::ReduceCtl *ctl, double *scal, int slot0, int slot1) {
    if (GUARD && !solver_active(scal)) return;
    extern __shared__ double prod[];
    double acc[DOT == 0 ? 1 : DOT];
#pragma unroll
    for (int r = 0; r < (DOT == 0 ? 1 : DOT); ++r) acc[r] = 0.0;
    (void)acc;
    const int64_t nblk = (nrow + rpb - 1) / rpb;
    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int64_t r0 = blk * rpb;
        const int64_t r1 = min(r0 + (int64_t)rpb, nrow);
        const int32_t k0 = rowptr[r0], k1 = rowptr[r1];
#pragma unroll 4
        for (int32_t k = k0 + threadIdx.x; k < k1; k += kThreads) prod[k - k0] = vals[k] * x[colidx[k]];
        __syncthreads();
        for (int64_t r = r0 + threadIdx.x; r < r1; r += kThreads) {
            const int32_t b = rowptr[r] - k0, e = rowptr[r + 1] - k0;
            double sum = 0.0;
            for (int32_t k = b; k < e; ++k) sum += prod[k];
            y[r] = sum;
            if constexpr (DOT == 1) acc[0] += w[r] * sum;
            if constexpr (DOT == 2) { acc[0] += sum * w[r]; acc[1] += sum * sum; }
        }
        __syncthreads();
    }
    if constexpr (DOT == 1) { const int sl_[1] = {slot0}; grid_reduce<1>(acc, partials, ctl, scal, sl_, false); }
    if constexpr (DOT == 2) { const int sl_[2] = {slot0, slot1}; grid_reduce<2>(acc, partials, ctl, scal, sl_, false); }
}

// ------------------------------------------------------------------ SpMV, warp-stream variant (short regular rows: P1 operators)
// Each warp owns tiles of 32 consecutive rows.  Phase 1: the lanes pull the tile's (col,val) pairs fully coalesced and park
// them in the warp's private shared-memory strip (a transposition; no block barrier).  Phase 2: lane l walks row l of the
// tile, so that at step j the 32 lanes gather x at 32 NEIGHBOURING columns (FEM rows of consecutive vertices are shifted
// copies of one another): ~2-3 L1 wavefronts per gather instruction instead of ~18 when lanes hold consecutive nonzeros.
// The L1 wavefront rate of the gather, not DRAM, was what bounded the nonzero-parallel variants (profiles/r1_spmv.md).
template <int MAXJ, int DOT, bool GUARD>
__global__ void __launch_bounds__(kThreads)
k_spmv_warp(int64_t nrow, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
            const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ w,
            double *partials, mp::ReduceCtl *ctl, double *scal, int slot0, int slot1) {
    if (GUARD && !solver_active(scal)) return;
    extern __shared__ double smem_d[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    double *sval = smem_d + wib * (32 * MAXJ);
    int32_t *scol = reinterpret_cast<int32_t *>(smem_d + (kThreads / 32) * (32 * MAXJ)) + wib * (32 * MAXJ);
    double acc[DOT == 0 ? 1 : DOT];
#pragma unroll
    for (int r = 0; r < (DOT == 0 ? 1 : DOT); ++r) acc[r] = 0.0;
    (void)acc;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    const int64_t ntile = (nrow + 31) >> 5;
    for (int64_t tile = warp; tile < ntile; tile += nwarp) {
        const int64_t row = (tile << 5) + lane;
        const int32_t b = rowptr[row < nrow ? row : nrow];
        const int32_t e = rowptr[row < nrow ? row + 1 : nrow];
        const int32_t k0 = __shfl_sync(0xffffffffu, b, 0);
        const int32_t k1 = __shfl_sync(0xffffffffu, e, 31);
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            const int32_t k = k0 + lane + 32 * j;
            if (k < k1) {
                scol[k - k0] = colidx[k];
                sval[k - k0] = vals[k];
            }
        }
        __syncwarp();
        const int len = e - b, off = b - k0;
        double sum = 0.0;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j)
            if (j < len) sum += sval[off + j] * x[scol[off + j]];
        if (row < nrow) {
            y[row] = sum;
            if constexpr (DOT == 1) acc[0] += w[row] * sum;
            if constexpr (DOT == 2) { acc[0] += sum * w[row]; acc[1] += sum * sum; }
        }
        __syncwarp();
    }
    if constexpr (DOT == 1) { const int sl_[1] = {slot0}; grid_reduce<1>(acc, partials, ctl, scal, sl_, false); }
    if constexpr (DOT == 2) { const int sl_[2] = {slot0, slot1}; grid_reduce<2>(acc, partials, ctl, scal, sl_, false); }
}

// ------------------------------------------------------------------ SpMV, SELL-32 (production kernels for P1 operators)
// Slice entries are stored [j][lane], so the (col,val) loads are perfectly coalesced and at step j the lanes of a warp gather x at
// neighbouring columns.  Measured 0.25 ms at 200^3 with int32 columns (6.4 TB/s algorithmic, 0.98 of the measured copy peak) against
// 0.34 ms for the best CSR variant (tools/spmv_lab.cu, profiles/r1_spmv_lab.txt).  __launch_bounds__(256, 8) keeps 2048 resident
// threads per SM: the kernel is latency-bound below that.  Two refinements, both bit-for-bit neutral for a fixed (C8, LPR) choice:
//  * C8: compressed column indices ("slice-uniform offsets").  Rows numbered along a mesh line share their stencil offsets, so for most
//    (slice, j) all 32 rows have col = row + d: ONE int32 per (slice, j) replaces 32 column indices, the rest keep explicit columns
//    (slices that straddle the end of a mesh line, unstructured parts).  ~8.2-9 instead of 12 bytes per nonzero, and fewer per-lane
//    instructions than the plain kernel (no column load for uniform positions).  A per-entry dictionary code (1 byte + shared-memory
//    lookup) was tried first and was SLOWER than int32 columns (0.264 vs 0.254 ms at 200^3): the decode sat on the gather's critical path.
//  * LPR lanes per row (1, 2, 4): a warp then covers 32/LPR rows of a slice and the lanes of a row take every LPR-th entry; the
//    work unit shrinks to a half / quarter slice.  Meant for the 1 M-row slabs of an 8-GPU run (3.3 slices per resident warp); it
//    measured slower than one lane per row at every size, so it is an experiment knob ("spmv_lpr"), not the default.
template <bool GHOST>
__device__ __forceinline__ double sell_x(const double *__restrict__ x, int32_t c, int64_t nown, const double *__restrict__ gw) {
    if constexpr (GHOST) return c < nown ? x[c] : __ldcg(gw + c);     // ghost data arrived during this kernel: read it from L2
    else return x[c];
}

// predicated loads (result 0 when the predicate is off; the address is then never dereferenced): they keep the batched inner loop of
// the compressed kernel free of branches, so that the four gathers of a batch are in flight together
__device__ __forceinline__ int32_t ld_nc_i32_if(const int32_t *p, bool pred) {
    int32_t v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.b32 %0, 0;\n\t@q ld.global.nc.b32 %0, [%1];\n\t}" : "=r"(v) : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ double ld_nc_f64_if(const double *p, bool pred) {
    double v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@q ld.global.nc.f64 %0, [%1];\n\t}" : "=d"(v) : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ double ld_cg_f64_if(const double *p, bool pred) {
    double v;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@q ld.global.cg.f64 %0, [%1];\n\t}" : "=d"(v) : "l"(p), "r"((int)pred) : "memory");
    return v;
}
template <bool GHOST>
__device__ __forceinline__ double sell_x_if(const double *__restrict__ x, int32_t c, bool pred, int64_t nown, const double *__restrict__ gw) {
    if constexpr (GHOST) {
        const bool own = c < nown;
        return ld_nc_f64_if(x + c, pred && own) + ld_cg_f64_if(gw + c, pred && !own);     // exactly one of the two is loaded
    } else {
        return ld_nc_f64_if(x + c, pred);
    }
}

// Row sums of one work unit.  The loads are issued in explicit batches of four (columns, values, then the four gathers): what bounds
// this kernel is memory-level parallelism, and a plain `#pragma unroll 4` of a loop with a per-lane trip count compiles to four
// SERIAL load->gather->FMA chains (seen in the SASS: 0.31 instead of 0.25 ms at 200^3).  The additions keep the sequential order.
template <bool C8, int LPR, bool GHOST>
__device__ __forceinline__ double sell_unit(int64_t unit, int lane, int64_t nrow, const int32_t *__restrict__ sliceptr,
                                            const int32_t *__restrict__ scol, const double *__restrict__ sval,
                                            const int32_t *__restrict__ meta, const int32_t *__restrict__ xcol,
                                            const double *__restrict__ x, int64_t nown, const double *__restrict__ gw, int64_t &row_out) {
    constexpr int RPU = 32 / LPR;                       // rows per unit
    const int64_t s = unit / LPR;
    const int rl = (int)(unit % LPR) * RPU + (lane % RPU);   // row inside the slice
    const int part = lane / RPU;                        // which share of the row's entries this lane sums
    const int32_t b = sliceptr[s];
    const int64_t row = (s << 5) + rl;
    const int wfull = (sliceptr[s + 1] - b) >> 5;
    const bool live = row < nrow;                       // rows past the end of the last slice do nothing
    const int32_t kb = b + rl;
    double sum = 0.0;
    if constexpr (C8) {
        // slice-uniform offsets: one int32 per (slice, j); the lanes fetch up to 32 of them with one coalesced load and hand them
        // around with shuffles, so a uniform position costs NO per-lane column load at all
        for (int j0 = 0; j0 < wfull; j0 += 32) {
            const int32_t mine = (j0 + lane < wfull) ? meta[(b >> 5) + j0 + lane] : 0;
            const int jend = min(wfull - j0, 32);
            for (int base = 0; base < jend; base += 4 * LPR) {          // warp-uniform trip count: every lane takes part in the shuffles
                int32_t m[4], c[4];
                double v[4], xv[4];
                bool ok[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int jj = base + part + t * LPR;
                    m[t] = __shfl_sync(0xffffffffu, mine, jj & 31);
                    ok[t] = live && jj < jend;
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) c[t] = ld_nc_i32_if(xcol + (((int64_t)(m[t] >> 1) << 5) + rl), ok[t] && !(m[t] & 1));
#pragma unroll
                for (int t = 0; t < 4; ++t) v[t] = ld_nc_f64_if(sval + (kb + 32 * (j0 + base + part + t * LPR)), ok[t]);
#pragma unroll
                for (int t = 0; t < 4; ++t) c[t] = (m[t] & 1) ? (int32_t)row + (m[t] >> 1) : c[t];
#pragma unroll
                for (int t = 0; t < 4; ++t) xv[t] = sell_x_if<GHOST>(x, c[t], ok[t], nown, gw);
#pragma unroll
                for (int t = 0; t < 4; ++t) sum += v[t] * xv[t];
            }
        }
    } else {
        const int w = live ? wfull : 0;
        int j = part;
        for (; j + 3 * LPR < w; j += 4 * LPR) {
            const int32_t k = kb + 32 * j;
            int32_t c[4];
            double v[4], xv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) c[t] = scol[k + 32 * LPR * t];
#pragma unroll
            for (int t = 0; t < 4; ++t) v[t] = sval[k + 32 * LPR * t];
#pragma unroll
            for (int t = 0; t < 4; ++t) xv[t] = sell_x<GHOST>(x, c[t], nown, gw);
#pragma unroll
            for (int t = 0; t < 4; ++t) sum += v[t] * xv[t];
        }
        if (j < w) {                                    // up to three left: one more (predicated) batch
            const int32_t k = kb + 32 * j;
            const bool o1 = j + LPR < w, o2 = j + 2 * LPR < w;
            const int32_t c0 = scol[k], c1 = o1 ? scol[k + 32 * LPR] : c0, c2 = o2 ? scol[k + 64 * LPR] : c0;
            const double v0 = sval[k], v1 = o1 ? sval[k + 32 * LPR] : 0.0, v2 = o2 ? sval[k + 64 * LPR] : 0.0;
            const double x0 = sell_x<GHOST>(x, c0, nown, gw), x1 = sell_x<GHOST>(x, c1, nown, gw), x2 = sell_x<GHOST>(x, c2, nown, gw);
            sum += v0 * x0;
            if (o1) sum += v1 * x1;
            if (o2) sum += v2 * x2;
        }
    }
#pragma unroll
    for (int o = RPU; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    row_out = (part == 0 && live) ? row : -1;           // the lane that owns the row's result (dead lanes: none)
    return sum;
}

struct SellArgs {
    int64_t nrow;
    const int32_t *sliceptr, *scol;
    const double *sval;
    const int32_t *meta, *xcol;
};

template <int DOT, bool GUARD, bool C8, int LPR>
__global__ void __launch_bounds__(kThreads, 8)
k_spmv_sell(SellArgs A, const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ w, double *partials,
            mp::ReduceCtl *ctl, double *scal, int slot0, int slot1) {
    if (GUARD && !solver_active(scal)) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    const int64_t nunit = ((A.nrow + 31) >> 5) * LPR;
    double acc[DOT == 0 ? 1 : DOT];
#pragma unroll
    for (int r = 0; r < (DOT == 0 ? 1 : DOT); ++r) acc[r] = 0.0;
    (void)acc;
    for (int64_t u = warp; u < nunit; u += nwarp) {
        int64_t row;
        const double sum = sell_unit<C8, LPR, false>(u, lane, A.nrow, A.sliceptr, A.scol, A.sval, A.meta, A.xcol, x, 0, nullptr, row);
        if (row >= 0) {
            y[row] = sum;
            if constexpr (DOT == 1 || DOT == 3) acc[0] += w[row] * sum;
            if constexpr (DOT == 2) { acc[0] += sum * w[row]; acc[1] += sum * sum; }
        }
    }
    if constexpr (DOT == 1) { const int sl_[1] = {slot0}; grid_reduce<1>(acc, partials, ctl, scal, sl_, false); }
    if constexpr (DOT == 2) { const int sl_[2] = {slot0, slot1}; grid_reduce<2>(acc, partials, ctl, scal, sl_, false); }
    if constexpr (DOT == 3) {   // single-reduction CG: w.u plus the vector kernel's rank-local r.u, r.r in ONE exchange; iter++
        if (blockIdx.x == 0 && threadIdx.x == 0) { acc[1] += scal[S_GLOC]; acc[2] += scal[S_RRLOC]; }
        const int sl_[3] = {slot0, slot1, S_RR};
        grid_reduce<3>(acc, partials, ctl, scal, sl_, true);
    }
}

// ------------------------------------------------------------------ SELL-32 SpMV fused with the halo exchange over peer memory
// Multi-GPU production kernel (peer mode).  One launch does what k_pack + ncclSend/ncclRecv + k_spmv_sell did:
//   1. the first blocks PUSH this rank's interface values of x straight into the neighbours' ghost windows (plain stores over
//      NVLink into CUDA-IPC mapped memory), the last pusher releases a sequence-number flag in every neighbour's window;
//   2. every warp multiplies its INTERIOR slices (no ghost column) while those stores are in flight;
//   3. warps that own boundary slices wait for the neighbours' flags and multiply them, reading ghost columns (col >= nown)
//      from the local ghost window.  Ghost windows are double-buffered by the parity of the push number: a neighbour starts
//      push s+1 only after its own SpMV s, for which it needed MY push s, which I issue after finishing SpMV s-1.
// Row sums and dot partials are accumulated in the same fixed order on every run (interior slices ascending, then the boundary
// list ascending), so results stay bitwise reproducible.
template <int DOT, bool GUARD, bool C8, int LPR>
__global__ void __launch_bounds__(kThreads, 8)
k_spmv_sell_peer(SellArgs A, const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ w, double *partials,
                 mp::ReduceCtl *ctl, double *scal, int slot0, int slot1, mp::HaloDev *H) {
    if (GUARD && !solver_active(scal)) return;
    const unsigned long long seq = H->seq;            // bumped by the last block to finish, i.e. after every block has read it
    const int par = (int)(seq & 1ull);
    const int lane = threadIdx.x & 31;
    // ---- 1. push
    const int nsend = H->nsend, nneigh = H->nneigh;
    const int npb = min((int)gridDim.x, max(1, (nsend + kThreads - 1) / kThreads));
    if ((int)blockIdx.x < npb) {
        for (int t = blockIdx.x * kThreads + threadIdx.x; t < nsend; t += npb * kThreads) {
            int k = 0;
            while (t >= H->send_off[k + 1]) ++k;
            H->peer_gwin[k][(int64_t)par * H->peer_gcap[k] + H->peer_goff[k] + (t - H->send_off[k])] = x[H->send_idx[t]];
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int old = atomicAdd(&ctl->push_count, 1u);
            if (old == (unsigned int)npb - 1u) {
                ctl->push_count = 0u;
                __threadfence_system();               // the other pushers' fenced stores are ordered before the flags below
                for (int k = 0; k < nneigh; ++k) st_release_sys(H->peer_flag[k] + par, seq);
            }
        }
    }
    // ---- 2. interior slices
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    const int64_t nunit = ((A.nrow + 31) >> 5) * LPR;
    const uint8_t *__restrict__ kind = H->slice_kind;
    double acc[DOT == 0 ? 1 : DOT];
#pragma unroll
    for (int r = 0; r < (DOT == 0 ? 1 : DOT); ++r) acc[r] = 0.0;
    (void)acc;
    for (int64_t u = warp; u < nunit; u += nwarp) {
        if (kind[u / LPR]) continue;
        int64_t row;
        const double sum = sell_unit<C8, LPR, false>(u, lane, A.nrow, A.sliceptr, A.scol, A.sval, A.meta, A.xcol, x, 0, nullptr, row);
        if (row >= 0) {
            y[row] = sum;
            if constexpr (DOT == 1 || DOT == 3) acc[0] += w[row] * sum;
            if constexpr (DOT == 2) { acc[0] += sum * w[row]; acc[1] += sum * sum; }
        }
    }
    // ---- 3. boundary slices (handed to the warps from the top end: those got one interior unit fewer)
    const int64_t rwarp = nwarp - 1 - warp;
    const int64_t nbu = (int64_t)H->nbslice * LPR;
    if (rwarp < nbu) {
        int ok = 1;
        if (lane == 0)
            for (int k = 0; k < nneigh; ++k) ok &= wait_flag(H->flag + 2 * k + par, seq) ? 1 : 0;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok && lane == 0) scal[S_RR] = __longlong_as_double(0x7ff8000000000000ll);   // dead neighbour: the host sees a NaN residual
        const int64_t nown = H->nown;
        const double *__restrict__ gw = H->gwin + (int64_t)par * H->nghost - nown;
        for (int64_t i = rwarp; i < nbu; i += nwarp) {
            const int64_t u = (int64_t)H->bslices[i / LPR] * LPR + (i % LPR);
            int64_t row;
            const double sum = sell_unit<C8, LPR, true>(u, lane, A.nrow, A.sliceptr, A.scol, A.sval, A.meta, A.xcol, x, nown, gw, row);
            if (row >= 0) {
                y[row] = sum;
                if constexpr (DOT == 1 || DOT == 3) acc[0] += w[row] * sum;
                if constexpr (DOT == 2) { acc[0] += sum * w[row]; acc[1] += sum * sum; }
            }
        }
    }
    // ---- 4. completion: the last block to get here opens the next push number
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int old = atomicAdd(&ctl->done_count, 1u);
        if (old == gridDim.x - 1) {
            ctl->done_count = 0u;
            H->seq = seq + 1ull;
        }
    }
    if constexpr (DOT == 1) { const int sl_[1] = {slot0}; grid_reduce<1>(acc, partials, ctl, scal, sl_, false); }
    if constexpr (DOT == 2) { const int sl_[2] = {slot0, slot1}; grid_reduce<2>(acc, partials, ctl, scal, sl_, false); }
    if constexpr (DOT == 3) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { acc[1] += scal[S_GLOC]; acc[2] += scal[S_RRLOC]; }
        const int sl_[3] = {slot0, slot1, S_RR};
        grid_reduce<3>(acc, partials, ctl, scal, sl_, true);
    }
}

// ------------------------------------------------------------------ BLAS-1
// x, y (and z below) may alias: no __restrict__ here
__global__ void __launch_bounds__(kThreads) k_axpby(int64_t n, double a, const double *x, double b, double *y) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x)
        y[i] = (b == 0.0) ? a * x[i] : a * x[i] + b * y[i];
}

__global__ void __launch_bounds__(kThreads) k_vmul(int64_t n, double a, const double *x, const double *y, double *z) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) z[i] = a * x[i] * y[i];
}

__global__ void __launch_bounds__(kThreads) k_dot(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                                                  double *partials, mp::ReduceCtl *ctl, double *scal, int slot) {
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) acc[0] += x[i] * y[i];
    const int sl_[1] = {slot};
    grid_reduce<1>(acc, partials, ctl, scal, sl_, false);
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
                                                            mp::ReduceCtl *ctl, double *scal, int slot_rz) {
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
    grid_reduce<3>(acc, partials, ctl, scal, sl_, false);
}

// CG: alpha = rz/pq ; x += alpha p ; r -= alpha q ; rz' = r.(dinv r) ; rr = r.r ; iter++
// (4 independent element groups per thread and trip => 20 loads in flight per thread)
__global__ void __launch_bounds__(kThreads) k_cg_update(int64_t n, const double *__restrict__ p, const double *__restrict__ q,
                                                        const double *__restrict__ dinv, double *__restrict__ x,
                                                        double *__restrict__ r, double *partials, mp::ReduceCtl *ctl,
                                                        double *scal, int rz_cur, int rz_nxt) {
    if (!solver_active(scal)) return;
    const double alpha = safe_div(scal[rz_cur], scal[S_PQ]);
    double acc[2] = {0.0, 0.0};
    const int64_t stride = gridDim.x * (int64_t)blockDim.x;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        double pp[4], qq[4], dd[4], xx[4], rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n) { pp[u] = p[i]; qq[u] = q[i]; dd[u] = dinv[i]; xx[u] = x[i]; rr[u] = r[i]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n) {
                x[i] = xx[u] + alpha * pp[u];
                const double ri = rr[u] - alpha * qq[u];
                r[i] = ri;
                acc[0] += ri * ri * dd[u];
                acc[1] += ri * ri;
            }
        }
    }
    const int sl_[2] = {rz_nxt, S_RR};
    grid_reduce<2>(acc, partials, ctl, scal, sl_, true);
}

// CG: beta = rz'/rz ; p = dinv r + beta p
__global__ void __launch_bounds__(kThreads) k_cg_pupdate(int64_t n, const double *__restrict__ r, const double *__restrict__ dinv,
                                                         double *__restrict__ p, const double *scal, int rz_cur, int rz_nxt) {
    if (!solver_active(scal)) return;
    const double beta = safe_div(scal[rz_nxt], scal[rz_cur]);
    const int64_t stride = gridDim.x * (int64_t)blockDim.x;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        double rr[4], dd[4], pp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n) { rr[u] = r[i]; dd[u] = dinv[i]; pp[u] = p[i]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n) p[i] = dd[u] * rr[u] + beta * pp[u];
        }
    }
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
                                                     double *partials, mp::ReduceCtl *ctl, double *scal, int rho_cur, int rho_nxt) {
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
    grid_reduce<2>(acc, partials, ctl, scal, sl_, true);
}


// ------------------------------------------------------------------ single-reduction PCG (Chronopoulos & Gear), vector kernel
// State: x, r, u = D^-1 r, w = A u, p, s = A p (by recurrence).  Per iteration ONE vector kernel and ONE SpMV kernel:
//   beta = gamma/gamma_old ; alpha = gamma / (delta - beta gamma / alpha_old)           (scalars recomputed by every thread)
//   p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = D^-1 r ; local sums gamma' = r.u, rr = r.r
// then w = A u with delta' = w.u, and the three sums travel in ONE exchange inside the SpMV kernel (k_spmv_sell* DOT == 3).
// Standard PCG needs two reduction points and three kernels per iteration; on 8 GPUs with 1 M rows each an iteration is ~50 us of
// HBM traffic, so every exchange (~4 us over NVLink) and every launch gap counts.
__global__ void __launch_bounds__(kThreads) k_cgcg_update(int64_t n, int first, const double *__restrict__ w, const double *__restrict__ dinv,
                                                          double *__restrict__ u, double *__restrict__ p, double *__restrict__ s,
                                                          double *__restrict__ x, double *__restrict__ r, double *partials,
                                                          mp::ReduceCtl *ctl, double *scal, int cur, int nxt) {
    if (!solver_active(scal)) return;
    const double gamma = scal[S_G0 + cur], delta = scal[S_DELTA];
    const double beta = first ? 0.0 : safe_div(gamma, scal[S_G0 + nxt]);
    const double alpha = first ? safe_div(gamma, delta) : safe_div(gamma, delta - beta * safe_div(gamma, scal[S_AL0 + nxt]));
    if (blockIdx.x == 0 && threadIdx.x == 0) scal[S_AL0 + cur] = alpha;          // nobody reads AL[cur] in this kernel
    double acc[2] = {0.0, 0.0};
    const int64_t stride = gridDim.x * (int64_t)blockDim.x;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += 2 * stride) {
        double uu[2], ww[2], dd[2], pp[2], ss[2], xx[2], rr[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int64_t i = i0 + k * stride;
            if (i < n) {
                uu[k] = u[i]; ww[k] = w[i]; dd[k] = dinv[i]; xx[k] = x[i]; rr[k] = r[i];
                pp[k] = first ? 0.0 : p[i];
                ss[k] = first ? 0.0 : s[i];
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int64_t i = i0 + k * stride;
            if (i < n) {
                const double pi = uu[k] + beta * pp[k], si = ww[k] + beta * ss[k];
                p[i] = pi;
                s[i] = si;
                x[i] = xx[k] + alpha * pi;
                const double ri = rr[k] - alpha * si;
                r[i] = ri;
                const double un = dd[k] * ri;
                u[i] = un;
                acc[0] += ri * un;
                acc[1] += ri * ri;
            }
        }
    }
    const int sl_[2] = {S_GLOC, S_RRLOC};
    grid_reduce<2>(acc, partials, ctl, scal, sl_, false, false, /*exchange=*/false);
}

// r = b - q ; u = dinv r ; dots bb, rr, gamma = r.u  (start of a single-reduction CG cycle)
__global__ void __launch_bounds__(kThreads) k_cgcg_init(int64_t n, const double *__restrict__ b, const double *__restrict__ q,
                                                        const double *__restrict__ dinv, double *__restrict__ r, double *__restrict__ u,
                                                        double *partials, mp::ReduceCtl *ctl, double *scal, int slot_gamma) {
    double acc[3] = {0.0, 0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        const double bi = b[i], ri = bi - q[i];
        r[i] = ri;
        const double ui = dinv[i] * ri;
        u[i] = ui;
        acc[0] += bi * bi;
        acc[1] += ri * ri;
        acc[2] += ri * ui;
    }
    const int sl_[3] = {S_BB, S_RR, slot_gamma};
    grid_reduce<3>(acc, partials, ctl, scal, sl_, false);
}

// ------------------------------------------------------------------ Chebyshev polynomial preconditioner (Jacobi-scaled)
// z = p_k(D^-1 A) D^-1 r on [lmin, lmax]:  the "Chebyshev" option of BASELINE.json's north_star.  k-1 SpMVs per application and
// no inner products, so a PCG iteration does k SpMVs per pair of reductions (what matters once the allreduce latency dominates).
__global__ void __launch_bounds__(kThreads) k_cheb_first(int64_t n, double inv_theta, const double *__restrict__ r,
                                                         const double *__restrict__ dinv, double *__restrict__ z,
                                                         double *__restrict__ d, const double *scal) {
    if (!solver_active(scal)) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        const double v = inv_theta * dinv[i] * r[i];
        z[i] = v;
        d[i] = v;
    }
}

// t = A z (done by the caller): d = c1 d + c2 dinv (r - t) ; z += d
__global__ void __launch_bounds__(kThreads) k_cheb_step(int64_t n, double c1, double c2, const double *__restrict__ r,
                                                        const double *__restrict__ t, const double *__restrict__ dinv,
                                                        double *__restrict__ z, double *__restrict__ d, const double *scal) {
    if (!solver_active(scal)) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        const double dn = c1 * d[i] + c2 * dinv[i] * (r[i] - t[i]);
        d[i] = dn;
        z[i] += dn;
    }
}

// general-preconditioner CG: x += alpha p ; r -= alpha q ; rr ; iter++
__global__ void __launch_bounds__(kThreads) k_cgp_update(int64_t n, const double *__restrict__ p, const double *__restrict__ q,
                                                         double *__restrict__ x, double *__restrict__ r, double *partials,
                                                         mp::ReduceCtl *ctl, double *scal, int rz_cur) {
    if (!solver_active(scal)) return;
    const double alpha = safe_div(scal[rz_cur], scal[S_PQ]);
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        acc[0] += ri * ri;
    }
    const int sl_[1] = {S_RR};
    grid_reduce<1>(acc, partials, ctl, scal, sl_, true);
}

// rz' = r.z  (guarded: the slot keeps its value once converged)
__global__ void __launch_bounds__(kThreads) k_cgp_rz(int64_t n, const double *__restrict__ r, const double *__restrict__ z,
                                                     double *partials, mp::ReduceCtl *ctl, double *scal, int slot) {
    if (!solver_active(scal)) return;
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) acc[0] += r[i] * z[i];
    const int sl_[1] = {slot};
    grid_reduce<1>(acc, partials, ctl, scal, sl_, false);
}

// p = z + beta p   (first == 1: p = z)
__global__ void __launch_bounds__(kThreads) k_cgp_pupdate(int64_t n, int first, const double *__restrict__ z, double *__restrict__ p,
                                                          const double *scal, int rz_cur, int rz_nxt) {
    if (!solver_active(scal)) return;
    const double beta = first ? 0.0 : safe_div(scal[rz_nxt], scal[rz_cur]);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x)
        p[i] = first ? z[i] : z[i] + beta * p[i];
}

// Gershgorin bound of lambda_max(D^-1 A): max_i dinv_i * sum_j |a_ij|  (max is order independent; positive doubles compare as integers)
__global__ void __launch_bounds__(kThreads) k_gershgorin(int64_t nrow, const int32_t *__restrict__ rowptr, const double *__restrict__ vals,
                                                         const double *__restrict__ dinv, unsigned long long *out) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double v = 0.0;
    if (r < nrow) {
        double s = 0.0;
        for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k) s += fabs(vals[k]);
        v = s * fabs(dinv[r]);
    }
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(v));
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

// grid / variant choice of the SELL kernels: compressed columns when the operator has them, lanes per row from the unit count
inline int sell_lpr(const mp_ctx *ctx, int64_t nslice) {
    // measured (profiles/r2_spmv_variants.md): 2 or 4 lanes per row are SLOWER than one at every size tried, 8 M rows down to the 1 M-row
    // slab of an 8-GPU run (0.0329 / 0.0386 / 0.0508 ms), so one lane per row is the default and the others stay as an experiment knob
    (void)nslice;
    if (ctx->spmv_lpr == 2 || ctx->spmv_lpr == 4) return ctx->spmv_lpr;
    return 1;
}

template <int DOT, bool GUARD>
int launch_sell(mp_ctx *ctx, const mp_csr *A, mp_halo *halo, const double *x, double *y, const double *w, int slot0, int slot1) {
    const int64_t nslice = (A->nrow + 31) / 32;
    const int lpr = sell_lpr(ctx, nslice);
    const bool c8 = ctx->spmv_c8 && A->d_sell_meta && A->d_sell_xcol;
    const int64_t nunit = nslice * lpr;
    int64_t nblk = (nunit + (kThreads / 32) - 1) / (kThreads / 32);
    int64_t cap = (int64_t)ctx->num_sm * 8;
    if (cap > kMaxBlocks) cap = kMaxBlocks;
    const int g = (int)(nblk < cap ? nblk : cap);
    SellArgs S{A->nrow, A->d_sell_sliceptr, A->d_sell_col, A->d_sell_val, A->d_sell_meta, A->d_sell_xcol};
    const bool timed = ctx->timing && ctx->tev_used + 2 <= ctx->tev.size();
    if (timed) cudaEventRecord(ctx->tev[ctx->tev_used], ctx->stream);
#define MP_SELL(C8_, LPR_)                                                                                                              \
    do {                                                                                                                                \
        if (halo) k_spmv_sell_peer<DOT, GUARD, C8_, LPR_><<<g, kThreads, 0, ctx->stream>>>(S, x, y, w, ctx->d_partials, ctx->d_ctl,     \
                                                                                            ctx->d_scalars, slot0, slot1, halo->d_dev); \
        else k_spmv_sell<DOT, GUARD, C8_, LPR_><<<g, kThreads, 0, ctx->stream>>>(S, x, y, w, ctx->d_partials, ctx->d_ctl,               \
                                                                                  ctx->d_scalars, slot0, slot1);                        \
    } while (0)
    if (c8) { if (lpr == 1) MP_SELL(true, 1); else if (lpr == 2) MP_SELL(true, 2); else MP_SELL(true, 4); }
    else { if (lpr == 1) MP_SELL(false, 1); else if (lpr == 2) MP_SELL(false, 2); else MP_SELL(false, 4); }
#undef MP_SELL
    if (timed) {
        cudaEventRecord(ctx->tev[ctx->tev_used + 1], ctx->stream);
        ctx->tev_used += 2;
    }
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return 0;
}

constexpr int kStreamCap = 2048;  // products (doubles) parked in shared memory per CTA

template <int DOT, bool GUARD>
int launch_spmv(mp_ctx *ctx, const mp_csr *A, int lanes, const double *x, double *y, const double *w, int slot0, int slot1) {
    if (A->nrow == 0) return 0;
    int grid = grid_for(ctx, A->nrow * lanes);
    const double avg = (double)A->nnz / (double)A->nrow;
    if (A->d_sell_val && A->d_sell_col && A->d_sell_sliceptr && ctx->spmv_variant == 0)
        return launch_sell<DOT, GUARD>(ctx, A, nullptr, x, y, w, slot0, slot1);
    const bool timed = ctx->timing && ctx->tev_used + 2 <= ctx->tev.size();
    if (timed) cudaEventRecord(ctx->tev[ctx->tev_used], ctx->stream);
    if (ctx->spmv_variant == 3 && A->max_row_nnz > 0 && A->max_row_nnz <= 32 && A->max_row_nnz <= 3.0 * avg + 2.0) {
        // warp-stream kernel: MAXJ = row-length bound rounded up to 8/16/32
        const int maxj = A->max_row_nnz <= 8 ? 8 : (A->max_row_nnz <= 16 ? 16 : 32);
        const size_t smem = (sizeof(double) + sizeof(int32_t)) * 32 * (size_t)maxj * (kThreads / 32);
        const int64_t ntile = (A->nrow + 31) / 32;
        int64_t nblk = (ntile + (kThreads / 32) - 1) / (kThreads / 32);
        const int per_sm = maxj == 8 ? 8 : (maxj == 16 ? 6 : 3);   // CTAs resident per SM (registers / shared memory)
        int64_t cap = (int64_t)ctx->num_sm * per_sm;
        int g = (int)(nblk < cap ? nblk : cap);
#define LW(MJ)                                                                                                              \
        do {                                                                                                                \
            static bool attr_set = false;                                                                                   \
            if (!attr_set) {                                                                                                \
                MP_CUDA(cudaFuncSetAttribute(k_spmv_warp<MJ, DOT, GUARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                attr_set = true;                                                                                            \
            }                                                                                                               \
            k_spmv_warp<MJ, DOT, GUARD><<<g, kThreads, smem, ctx->stream>>>(A->nrow, A->d_rowptr, A->d_colidx, A->d_vals, x, y, w, \
                                                                             ctx->d_partials, ctx->d_ctl, ctx->d_scalars, slot0, slot1); \
        } while (0)
        if (maxj == 8) LW(8); else if (maxj == 16) LW(16); else LW(32);
#undef LW
        if (timed) {
            cudaEventRecord(ctx->tev[ctx->tev_used + 1], ctx->stream);
            ctx->tev_used += 2;
        }
        mp::count_launch(ctx);
        MP_CUDA(cudaGetLastError());
        return 0;
    }
    if (ctx->spmv_variant <= 1 && A->max_row_nnz > 0 && A->max_row_nnz <= 32 && A->max_row_nnz <= 3.0 * avg + 2.0) {
        int rpb = kStreamCap / A->max_row_nnz;
        if (rpb > kThreads) rpb = kThreads;
        rpb &= ~31;
        const size_t smem = sizeof(double) * (size_t)rpb * A->max_row_nnz;
        int64_t nblk = (A->nrow + rpb - 1) / rpb;
        int64_t cap = (int64_t)ctx->num_sm * 8;
        int g = (int)(nblk < cap ? nblk : cap);
        k_spmv_stream<DOT, GUARD><<<g, kThreads, smem, ctx->stream>>>(A->nrow, rpb, A->d_rowptr, A->d_colidx, A->d_vals, x, y, w,
                                                                       ctx->d_partials, ctx->d_ctl, ctx->d_scalars, slot0, slot1);
        if (timed) {
            cudaEventRecord(ctx->tev[ctx->tev_used + 1], ctx->stream);
            ctx->tev_used += 2;
        }
        mp::count_launch(ctx);
        MP_CUDA(cudaGetLastError());
        return 0;
    }
#define L(S_)                                                                                                           \
    k_spmv<S_, DOT, GUARD><<<grid, kThreads, 0, ctx->stream>>>(A->nrow, A->d_rowptr, A->d_colidx, A->d_vals, x, y, w,  \
                                                                ctx->d_partials, ctx->d_ctl, ctx->d_scalars, slot0, slot1)
    switch (lanes) {
        case 2: L(2); break;
        case 4: L(4); break;
        case 8: L(8); break;
        case 16: L(16); break;
        default: L(32); break;
    }
#undef L
    if (timed) {
        cudaEventRecord(ctx->tev[ctx->tev_used + 1], ctx->stream);
        ctx->tev_used += 2;
    }
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return 0;
}

// y = A x with the ghost update of x: peer mode = ONE kernel (push + interior + wait + boundary, k_spmv_sell_peer);
// otherwise k_pack + grouped ncclSend/ncclRecv followed by the single-GPU kernels.
template <int DOT, bool GUARD>
int spmv_halo(mp_ctx *ctx, const mp_csr *A, mp_halo *halo, int lanes, double *x, double *y, const double *w, int slot0, int slot1) {
    const bool sell = A->d_sell_val && A->d_sell_col && A->d_sell_sliceptr && ctx->spmv_variant == 0;
    if (halo && halo->peer && halo->nneigh > 0 && ctx->nranks > 1 && ctx->peer && sell && A->nrow > 0 && halo->nown == A->nrow) {
        MP_TRY(mp::halo_classify(ctx, halo, A->nrow, A->d_sell_sliceptr, A->d_sell_col));
        return launch_sell<DOT, GUARD>(ctx, A, halo, x, y, w, slot0, slot1);
    }
    MP_TRY(mp_halo_exchange(ctx, halo, x));
    return launch_spmv<DOT, GUARD>(ctx, A, lanes, x, y, w, slot0, slot1);
}

int nnz_of(mp_ctx *ctx, const mp_csr *A, int64_t *nnz) {
    (void)ctx;
    *nnz = A->nnz;
    return 0;
}

// rank-local slot values (written at +kLocalOff by grid_reduce) -> global slots on every rank; no-op on one GPU
int reduce_slots(mp_ctx *ctx, int slot, int count) {
    if (ctx->nranks <= 1 || ctx->peer) return 0;   // peer mode: grid_reduce has already allreduced inside the kernel
    return mp::allreduce_oop(ctx, ctx->d_scalars + mp::kLocalOff + slot, ctx->d_scalars + slot, count);
}

}  // namespace
int mp::reduce_slots_public(mp_ctx *ctx, int slot, int count) { return reduce_slots(ctx, slot, count); }
namespace {

int poll(mp_ctx *ctx) {
    MP_CUDA(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, ctx->stream));
    int *h_err = reinterpret_cast<int *>(ctx->h_scalars + 48);
    *h_err = 0;
    if (ctx->peer && ctx->d_comm)
        MP_CUDA(cudaMemcpyAsync(h_err, &ctx->d_comm->error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MP_CUDA(cudaStreamSynchronize(ctx->stream));
    MP_CHECK(*h_err == 0, "peer-memory exchange timed out: a neighbouring rank stopped answering");
    return 0;
}

int check_csr(const mp_csr *A, bool for_solver = true) {
    MP_CHECK(A && A->d_rowptr && A->d_colidx && A->d_vals, "csr: null pointer");
    MP_CHECK(A->nrow >= 0 && A->ncol >= 0, "csr: negative size");
    MP_CHECK(!for_solver || A->ncol >= A->nrow, "csr: a solver needs ncol >= nrow (owned rows first, then ghosts)");
    return 0;
}

// first == true: set the tolerance from ||b|| and zero the iteration counter (one synchronising poll per solve);
// first == false (restart): only test the freshly evaluated TRUE residual against the stored tolerance.
int start_solve(mp_ctx *ctx, mp_solve_info *info, bool first, bool *done, bool *zero_rhs = nullptr) {
    MP_TRY(reduce_slots(ctx, S_BB, 1));
    MP_TRY(reduce_slots(ctx, S_RR, 1));
    MP_TRY(poll(ctx));
    const double bb = ctx->h_scalars[S_BB], rr = ctx->h_scalars[S_RR];
    if (!first) {
        *done = !(rr > ctx->h_scalars[S_TOL2]);
        return 0;
    }
    double tol = info->rtol * sqrt(bb);
    if (info->atol > tol) tol = info->atol;
    double hs[3] = {bb, tol * tol, 0.0};  // S_BB, S_TOL2, S_ITER are contiguous; S_RR is already on the device
    memcpy(ctx->h_scalars, hs, sizeof(hs));
    MP_CUDA(cudaMemcpyAsync(ctx->d_scalars, ctx->h_scalars, sizeof(hs), cudaMemcpyHostToDevice, ctx->stream));
    MP_CUDA(cudaStreamSynchronize(ctx->stream));
    info->niter = 0;
    info->relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    if (zero_rhs) *zero_rhs = !(bb > 0.0);
    *done = !(rr > tol * tol);
    info->converged = *done ? 1 : 0;
    return 0;
}

// ||b|| == 0 (allreduced, so every rank agrees): the solution is x = 0.  Without this the relative test has tolerance 0 and the
// recurrence would spin to maxit from a non-zero initial guess (PETSc avoids it with its absolute floor atol = 1e-50).
int zero_solution(mp_ctx *ctx, double *d_x, int64_t n, mp_solve_info *info) {
    MP_CUDA(cudaMemsetAsync(d_x, 0, sizeof(double) * (size_t)n, ctx->stream));
    info->niter = 0; info->converged = 1; info->relres = 0.0;
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

namespace {

// ------------------------------------------------------------------ GMRES(m), right Jacobi preconditioning
// Small dense data of one cycle, all on the device: R (upper triangular, column major, ld = m+1), Givens (cs, sn),
// rotated rhs g, solution y, and the two Gram-Schmidt coefficient columns h1, h2 (classical GS applied twice).
struct GmresSmall {
    double *R, *cs, *sn, *g, *y, *h1, *h2, *h1loc, *h2loc;
    int m;
};

// V0 = r / ||r||, z = dinv * V0, g = (||r||, 0, ...), ncol = 0
__global__ void __launch_bounds__(kThreads) k_gmres_start(int64_t n, const double *__restrict__ r, const double *__restrict__ dinv,
                                                          double *__restrict__ V0, double *__restrict__ z, double *scal, GmresSmall S) {
    if (!solver_active(scal)) return;
    const double beta = sqrt(scal[S_RR]);
    const double inv = beta > 0.0 ? 1.0 / beta : 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        const double v = r[i] * inv;
        V0[i] = v;
        z[i] = dinv[i] * v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S.g[0] = beta;
        scal[S_NCOL] = 0.0;
    }
}

// out[k0+v] = V_{k0+v} . w   for v < cnt <= 8
__global__ void __launch_bounds__(kThreads) k_multidot8(int64_t n, const double *__restrict__ V, int64_t ld, int k0, int cnt,
                                                        const double *__restrict__ w, double *partials, mp::ReduceCtl *ctl,
                                                        double *out, const double *scal) {
    if (!solver_active(scal)) return;
    double acc[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) acc[v] = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        const double wi = w[i];
#pragma unroll
        for (int v = 0; v < 8; ++v)
            if (v < cnt) acc[v] += V[(int64_t)(k0 + v) * ld + i] * wi;
    }
    const int sl_[8] = {k0, k0 + 1, k0 + 2, k0 + 3, k0 + 4, k0 + 5, k0 + 6, k0 + 7};
    grid_reduce<8>(acc, partials, ctl, out, sl_, false, false);
}

// w -= sum_{k<=j} h[k] V_k ; optionally slot S_TMP0 = ||w||^2
template <bool NORM>
__global__ void __launch_bounds__(kThreads) k_gmres_orth(int64_t n, const double *__restrict__ V, int64_t ld, int j,
                                                         const double *__restrict__ h, double *__restrict__ w, double *partials,
                                                         mp::ReduceCtl *ctl, double *scal) {
    if (!solver_active(scal)) return;
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        double wi = w[i];
        for (int k = 0; k <= j; ++k) wi -= h[k] * V[(int64_t)k * ld + i];
        w[i] = wi;
        if (NORM) acc[0] += wi * wi;
    }
    if (NORM) {
        const int sl_[1] = {S_TMP0};
        grid_reduce<1>(acc, partials, ctl, scal, sl_, false);
    }
}

// V_{j+1} = w / ||w|| ; z = dinv * V_{j+1}
__global__ void __launch_bounds__(kThreads) k_gmres_normalize(int64_t n, const double *__restrict__ w, const double *__restrict__ dinv,
                                                              double *__restrict__ Vn, double *__restrict__ z, const double *scal) {
    if (!solver_active(scal)) return;
    const double nrm = sqrt(scal[S_TMP0]);
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        const double v = w[i] * inv;
        Vn[i] = v;
        z[i] = dinv[i] * v;
    }
}

// one thread: new Hessenberg column -> Givens -> residual estimate
__global__ void k_gmres_givens(int j, GmresSmall S, double *scal) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (!solver_active(scal)) return;
    const int ld = S.m + 1;
    double *col = S.R + (int64_t)j * ld;
    for (int k = 0; k <= j; ++k) col[k] = S.h1[k] + S.h2[k];
    double hn = sqrt(scal[S_TMP0]);
    for (int i = 0; i < j; ++i) {
        const double t = S.cs[i] * col[i] + S.sn[i] * col[i + 1];
        col[i + 1] = -S.sn[i] * col[i] + S.cs[i] * col[i + 1];
        col[i] = t;
    }
    const double a = col[j], b = hn;
    const double d = sqrt(a * a + b * b);
    const double c = d > 0.0 ? a / d : 1.0, s = d > 0.0 ? b / d : 0.0;
    S.cs[j] = c; S.sn[j] = s;
    col[j] = d;
    const double gj = S.g[j];
    S.g[j] = c * gj;
    S.g[j + 1] = -s * gj;
    scal[S_RR] = S.g[j + 1] * S.g[j + 1];
    scal[S_ITER] += 1.0;
    scal[S_NCOL] = (double)(j + 1);
}

// one thread: back substitution R y = g over the completed columns
__global__ void k_gmres_solve(GmresSmall S, const double *scal) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nc = (int)scal[S_NCOL];
    const int ld = S.m + 1;
    for (int k = nc - 1; k >= 0; --k) {
        double v = S.g[k];
        for (int l = k + 1; l < nc; ++l) v -= S.R[(int64_t)l * ld + k] * S.y[l];
        const double d = S.R[(int64_t)k * ld + k];
        S.y[k] = d != 0.0 ? v / d : 0.0;
    }
}

// x += dinv * sum_k y[k] V_k
__global__ void __launch_bounds__(kThreads) k_gmres_xupdate(int64_t n, const double *__restrict__ V, int64_t ld,
                                                            const double *__restrict__ y, const double *__restrict__ dinv,
                                                            double *__restrict__ x, const double *scal) {
    const int nc = (int)scal[S_NCOL];
    if (nc == 0) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (int64_t)blockDim.x) {
        double a = 0.0;
        for (int k = 0; k < nc; ++k) a += y[k] * V[(int64_t)k * ld + i];
        x[i] += dinv[i] * a;
    }
}

}  // namespace

extern "C" int mp_solve_gmres(mp_ctx *ctx, const mp_csr *A, const double *d_dinv, const double *d_b, double *d_x, mp_halo *halo,
                              int restart_m, mp_solve_info *info) {
    MP_CHECK(ctx && d_dinv && d_b && d_x && info, "mp_solve_gmres: null argument");
    MP_TRY(check_csr(A));
    const int m = restart_m > 0 ? (restart_m > 200 ? 200 : restart_m) : 50;
    const int64_t n = A->nrow, nc = A->ncol;
    int64_t nnz = 0;
    MP_TRY(nnz_of(ctx, A, &nnz));
    const int lanes = pick_lanes(A, nnz);
    double *wk = nullptr, *sm = nullptr;
    MP_TRY(mp::ws_get(ctx, 0, sizeof(double) * (size_t)((m + 3) * n + nc + 8), (void **)&wk));
    const int pad = m + 16;  // k_multidot8 always stores 8 slots
    const size_t nsmall = (size_t)(m + 1) * m + 9 * (size_t)pad;
    MP_TRY(mp::ws_get(ctx, 1, sizeof(double) * nsmall, (void **)&sm));
    double *V = wk, *w = wk + (int64_t)(m + 1) * n, *r = w + n, *z = r + n;  // z has ghost entries
    GmresSmall S;
    S.m = m;
    S.R = sm; S.cs = S.R + (size_t)(m + 1) * m; S.sn = S.cs + pad; S.g = S.sn + pad; S.y = S.g + pad;
    S.h1 = S.y + pad; S.h2 = S.h1 + pad; S.h1loc = S.h2 + pad; S.h2loc = S.h1loc + pad;
    const int g1 = grid_for(ctx, n);
    const int every = info->check_every > 0 ? info->check_every : 10;   // rank-independent (collective control flow)
    int total = 0;
    for (int cycle = 0; cycle < 100000; ++cycle) {
        // true residual r = b - A x at the start of every cycle
        MP_TRY((spmv_halo<0, false>(ctx, A, halo, lanes, d_x, w, nullptr, 0, 0)));
        k_init_residual<<<g1, kThreads, 0, ctx->stream>>>(n, d_b, w, d_dinv, r, nullptr, nullptr, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, S_TMP0);
        mp::count_launch(ctx);
        MP_CUDA(cudaGetLastError());
        bool done = false;
        MP_TRY(start_solve(ctx, info, cycle == 0, &done));
        if (done || total >= info->maxit) break;
        k_gmres_start<<<g1, kThreads, 0, ctx->stream>>>(n, r, d_dinv, V, z, ctx->d_scalars, S);
        mp::count_launch(ctx);
        bool conv = false;
        for (int j = 0; j < m && total < info->maxit; ++j) {
            MP_TRY((spmv_halo<0, true>(ctx, A, halo, lanes, z, w, nullptr, 0, 0)));
            for (int pass = 0; pass < 2; ++pass) {   // classical Gram-Schmidt, applied twice
                double *h = pass == 0 ? S.h1 : S.h2;
                double *hl = (ctx->nranks > 1 && !ctx->peer) ? (pass == 0 ? S.h1loc : S.h2loc) : h;   // rank-local dots, then out-of-place allreduce
                for (int k0 = 0; k0 <= j; k0 += 8) {
                    const int cnt = (j + 1 - k0) < 8 ? (j + 1 - k0) : 8;
                    k_multidot8<<<g1, kThreads, 0, ctx->stream>>>(n, V, n, k0, cnt, w, ctx->d_partials, ctx->d_ctl, hl, ctx->d_scalars);
                    mp::count_launch(ctx);
                }
                if (ctx->nranks > 1 && !ctx->peer) MP_TRY(mp::allreduce_oop(ctx, hl, h, j + 1));
                if (pass == 0) k_gmres_orth<false><<<g1, kThreads, 0, ctx->stream>>>(n, V, n, j, h, w, ctx->d_partials, ctx->d_ctl, ctx->d_scalars);
                else k_gmres_orth<true><<<g1, kThreads, 0, ctx->stream>>>(n, V, n, j, h, w, ctx->d_partials, ctx->d_ctl, ctx->d_scalars);
                mp::count_launch(ctx);
            }
            MP_TRY(reduce_slots(ctx, S_TMP0, 1));
            k_gmres_normalize<<<g1, kThreads, 0, ctx->stream>>>(n, w, d_dinv, V + (int64_t)(j + 1) * n, z, ctx->d_scalars);
            k_gmres_givens<<<1, 32, 0, ctx->stream>>>(j, S, ctx->d_scalars);
            mp::count_launch(ctx, 2);
            ++total;
            if (total % every == 0 || total == info->maxit || j + 1 == m) {
                MP_CUDA(cudaGetLastError());
                MP_TRY(poll(ctx));
                if (!(ctx->h_scalars[S_RR] == ctx->h_scalars[S_RR])) { mp::set_error("mp_solve_gmres: NaN residual at iteration %d", total); return 4; }
                if (!(ctx->h_scalars[S_RR] > ctx->h_scalars[S_TOL2])) { conv = true; break; }
            }
        }
        k_gmres_solve<<<1, 32, 0, ctx->stream>>>(S, ctx->d_scalars);
        k_gmres_xupdate<<<g1, kThreads, 0, ctx->stream>>>(n, V, n, S.y, d_dinv, d_x, ctx->d_scalars);
        mp::count_launch(ctx, 2);
        (void)conv;
    }
    return finish_solve(ctx, info);
}

namespace {
}  // namespace

extern "C" int mp_spmv(mp_ctx *ctx, const mp_csr *A, const double *d_x, double *d_y) {
    MP_CHECK(ctx && d_x && d_y, "mp_spmv: null argument");
    MP_TRY(check_csr(A, false));
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

extern "C" int mp_vmul(mp_ctx *ctx, int64_t n, double a, const double *d_x, const double *d_y, double *d_z) {
    MP_CHECK(ctx && d_x && d_y && d_z && n >= 0, "mp_vmul: bad argument");
    if (n == 0) return 0;
    k_vmul<<<grid_for(ctx, n), kThreads, 0, ctx->stream>>>(n, a, d_x, d_y, d_z);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mp_dot(mp_ctx *ctx, int64_t n, const double *d_x, const double *d_y, double *result) {
    MP_CHECK(ctx && d_x && d_y && result && n >= 0, "mp_dot: bad argument");
    k_dot<<<grid_for(ctx, n > 0 ? n : 1), kThreads, 0, ctx->stream>>>(n, d_x, d_y, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, S_TMP0);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    MP_TRY(reduce_slots(ctx, S_TMP0, 1));
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

namespace {
// Single-reduction PCG driver (see k_cgcg_update).  Needs the SELL mirror (the fused three-sum exchange lives in the SELL kernels) and,
// on several GPUs, the peer-memory transport; mp_solve_pcg falls back to the classic recurrence otherwise.
int solve_pcg_single_reduction(mp_ctx *ctx, const mp_csr *A, const double *d_dinv, const double *d_b, double *d_x, mp_halo *halo,
                               mp_solve_info *info) {
    const int64_t n = A->nrow, nc = A->ncol;
    const bool use_halo = halo && halo->peer && halo->nneigh > 0 && ctx->nranks > 1;
    if (use_halo) MP_TRY(mp::halo_classify(ctx, halo, A->nrow, A->d_sell_sliceptr, A->d_sell_col));
    mp_halo *H = use_halo ? halo : nullptr;
    double *wk = nullptr;
    MP_TRY(mp::ws_get(ctx, 0, sizeof(double) * (size_t)(5 * n + nc + 8), (void **)&wk));
    double *r = wk, *w = wk + n, *p = wk + 2 * n, *s = wk + 3 * n, *q = wk + 4 * n, *u = wk + 5 * n;   // u carries ghost entries
    const int g1 = grid_for(ctx, n);
    const int every = info->check_every > 0 ? info->check_every : 20;
    int total = 0;
    for (int restart = 0; restart <= kMaxRestart + 1; ++restart) {
        // true residual r = b - A x, u = D^-1 r, gamma = r.u
        MP_TRY((launch_sell<0, false>(ctx, A, H, d_x, q, nullptr, 0, 0)));
        k_cgcg_init<<<g1, kThreads, 0, ctx->stream>>>(n, d_b, q, d_dinv, r, u, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, S_G0);
        mp::count_launch(ctx);
        MP_CUDA(cudaGetLastError());
        bool done = false, zero_rhs = false;
        MP_TRY(start_solve(ctx, info, restart == 0, &done, &zero_rhs));
        if (zero_rhs) return zero_solution(ctx, d_x, n, info);
        if (done || total >= info->maxit || restart > kMaxRestart) break;
        // w = A u, delta = w.u
        MP_TRY((launch_sell<1, true>(ctx, A, H, u, w, u, S_DELTA, 0)));
        int cur = 0, nxt = 1;
        for (int it = 0; total < info->maxit; ++it) {
            k_cgcg_update<<<g1, kThreads, 0, ctx->stream>>>(n, it == 0 ? 1 : 0, w, d_dinv, u, p, s, d_x, r, ctx->d_partials, ctx->d_ctl,
                                                             ctx->d_scalars, cur, nxt);
            mp::count_launch(ctx);
            MP_TRY((launch_sell<3, true>(ctx, A, H, u, w, u, S_DELTA, S_G0 + nxt)));
            int tmp = cur; cur = nxt; nxt = tmp;
            ++total;
            if (total % every == 0 || total == info->maxit) {
                MP_CUDA(cudaGetLastError());
                MP_TRY(poll(ctx));
                if (!(ctx->h_scalars[S_RR] == ctx->h_scalars[S_RR])) { mp::set_error("mp_solve_pcg: NaN residual at iteration %d", total); return 4; }
                if (!(ctx->h_scalars[S_RR] > ctx->h_scalars[S_TOL2])) break;
            }
        }
    }
    return finish_solve(ctx, info);
}
}  // namespace

extern "C" int mp_solve_pcg(mp_ctx *ctx, const mp_csr *A, const double *d_dinv, const double *d_b, double *d_x, mp_halo *halo,
                            mp_solve_info *info) {
    MP_CHECK(ctx && d_dinv && d_b && d_x && info, "mp_solve_pcg: null argument");
    MP_TRY(check_csr(A));
    {
        const bool sell = A->d_sell_val && A->d_sell_col && A->d_sell_sliceptr && ctx->spmv_variant == 0 && A->nrow > 0;
        const bool multi = ctx->nranks > 1 && halo && halo->nneigh > 0;
        const bool comm_ok = !multi || (ctx->peer && halo->peer && halo->nown == A->nrow);
        const int variant = ctx->pcg_variant < 0 ? (multi ? 1 : 0) : ctx->pcg_variant;    // default: single reduction on several GPUs
        if (variant == 1 && sell && comm_ok && (ctx->nranks == 1 || ctx->peer)) return solve_pcg_single_reduction(ctx, A, d_dinv, d_b, d_x, halo, info);
    }
    const int64_t n = A->nrow, nc = A->ncol;
    int64_t nnz = 0;
    MP_TRY(nnz_of(ctx, A, &nnz));
    const int lanes = pick_lanes(A, nnz);
    double *w = nullptr;
    MP_TRY(mp::ws_get(ctx, 0, sizeof(double) * (size_t)(2 * n + nc + 8), (void **)&w));
    double *r = w, *q = w + n, *p = w + 2 * n;  // p has ghost entries
    const int g1 = grid_for(ctx, n);
    const int every = info->check_every > 0 ? info->check_every : 20;   // must NOT depend on the rank-local size: all ranks poll together
    int total = 0;
    // outer loop = residual replacement: after the recurrence says "converged" the TRUE residual b - A x is
    // evaluated and the iteration restarted from x if it is not below the tolerance yet.
    for (int restart = 0; restart <= kMaxRestart + 1; ++restart) {   // the last pass only re-evaluates the TRUE residual
        MP_TRY((spmv_halo<0, false>(ctx, A, halo, lanes, d_x, q, nullptr, 0, 0)));
        k_init_residual<<<g1, kThreads, 0, ctx->stream>>>(n, d_b, q, d_dinv, r, p, nullptr, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, S_RZ0);
        mp::count_launch(ctx);
        MP_CUDA(cudaGetLastError());
        MP_TRY(reduce_slots(ctx, S_RZ0, 1));
        bool done = false, zero_rhs = false;
        MP_TRY(start_solve(ctx, info, restart == 0, &done, &zero_rhs));
        if (zero_rhs) return zero_solution(ctx, d_x, n, info);
        if (done || total >= info->maxit || restart > kMaxRestart) break;
        int cur = S_RZ0, nxt = S_RZ1;
        for (; total < info->maxit;) {
            MP_TRY((spmv_halo<1, true>(ctx, A, halo, lanes, p, q, p, S_PQ, 0)));
            MP_TRY(reduce_slots(ctx, S_PQ, 1));
            k_cg_update<<<g1, kThreads, 0, ctx->stream>>>(n, p, q, d_dinv, d_x, r, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, cur, nxt);
            MP_TRY(reduce_slots(ctx, (nxt < S_RR ? nxt : S_RR), 2));
            k_cg_pupdate<<<g1, kThreads, 0, ctx->stream>>>(n, r, d_dinv, p, ctx->d_scalars, cur, nxt);
            mp::count_launch(ctx, 2);
            int tmp = cur; cur = nxt; nxt = tmp;
            ++total;
            if (total % every == 0 || total == info->maxit) {
                MP_CUDA(cudaGetLastError());
                MP_TRY(poll(ctx));
                if (!(ctx->h_scalars[S_RR] == ctx->h_scalars[S_RR])) { mp::set_error("mp_solve_pcg: NaN residual at iteration %d", total); return 4; }
                if (!(ctx->h_scalars[S_RR] > ctx->h_scalars[S_TOL2])) break;
            }
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
    const int every = info->check_every > 0 ? info->check_every : 10;   // rank-independent (collective control flow)
    int total = 0;
    for (int restart = 0; restart <= kMaxRestart + 1; ++restart) {   // the last pass only re-evaluates the TRUE residual   // residual replacement, see mp_solve_pcg
        MP_TRY((spmv_halo<0, false>(ctx, A, halo, lanes, d_x, v, nullptr, 0, 0)));
        k_init_residual<<<g1, kThreads, 0, ctx->stream>>>(n, d_b, v, d_dinv, r, nullptr, rhat, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, S_RHO0);
        mp::count_launch(ctx);
        MP_CUDA(cudaGetLastError());
        MP_TRY(reduce_slots(ctx, S_RHO0, 1));
        bool done = false, zero_rhs = false;
        MP_TRY(start_solve(ctx, info, restart == 0, &done, &zero_rhs));
        if (zero_rhs) return zero_solution(ctx, d_x, n, info);
        if (done || total >= info->maxit || restart > kMaxRestart) break;
        int cur = S_RHO0, nxt = S_RHO1;
        for (int it = 0; total < info->maxit; ++it) {
            k_bicg_p<<<g1, kThreads, 0, ctx->stream>>>(n, it == 0 ? 1 : 0, r, v, d_dinv, p, y, ctx->d_scalars, cur);
            MP_TRY((spmv_halo<1, true>(ctx, A, halo, lanes, y, v, rhat, S_RV, 0)));
            MP_TRY(reduce_slots(ctx, S_RV, 1));
            k_bicg_s<<<g1, kThreads, 0, ctx->stream>>>(n, r, v, d_dinv, s, z, ctx->d_scalars, cur);
            MP_TRY((spmv_halo<2, true>(ctx, A, halo, lanes, z, t, s, S_TS, S_TT)));
            MP_TRY(reduce_slots(ctx, S_TS, 2));
            k_bicg_x<<<g1, kThreads, 0, ctx->stream>>>(n, y, z, s, t, rhat, d_x, r, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, cur, nxt);
            MP_TRY(reduce_slots(ctx, (nxt < S_RR ? nxt : S_RR), 2));
            mp::count_launch(ctx, 3);
            int tmp = cur; cur = nxt; nxt = tmp;
            ++total;
            if (total % every == 0 || total == info->maxit) {
                MP_CUDA(cudaGetLastError());
                MP_TRY(poll(ctx));
                if (!(ctx->h_scalars[S_RR] == ctx->h_scalars[S_RR])) { mp::set_error("mp_solve_bicgstab: NaN residual at iteration %d", total); return 4; }
                if (!(ctx->h_scalars[S_RR] > ctx->h_scalars[S_TOL2])) break;
            }
        }
    }
    return finish_solve(ctx, info);
}

extern "C" int mp_spectral_bound(mp_ctx *ctx, const mp_csr *A, const double *d_dinv, double *lmax) {
    MP_CHECK(ctx && d_dinv && lmax, "mp_spectral_bound: null argument");
    MP_TRY(check_csr(A));
    unsigned long long *d_out = reinterpret_cast<unsigned long long *>(ctx->d_scalars + S_COUNT + mp::kLocalOff);   // scratch slot
    MP_CUDA(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), ctx->stream));
    if (A->nrow > 0) {
        k_gershgorin<<<(unsigned)cdiv(A->nrow, kThreads), kThreads, 0, ctx->stream>>>(A->nrow, A->d_rowptr, A->d_vals, d_dinv, d_out);
        mp::count_launch(ctx);
    }
    double h = 0.0;
    MP_CUDA(cudaMemcpyAsync(&h, d_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MP_CUDA(cudaStreamSynchronize(ctx->stream));
    MP_TRY(mp::allreduce_max_host(ctx, &h));
    *lmax = h;
    return 0;
}

// Chebyshev-Jacobi preconditioned CG.  degree >= 1 (1 == plain Jacobi scaling by 1/theta); [lmin, lmax] must contain the top of the
// spectrum of D^-1 A (mp_spectral_bound gives a safe lmax; lmin = lmax / 30 is the customary smoother window).
extern "C" int mp_solve_pcg_chebyshev(mp_ctx *ctx, const mp_csr *A, const double *d_dinv, const double *d_b, double *d_x, mp_halo *halo,
                                      int degree, double lmin, double lmax, mp_solve_info *info) {
    MP_CHECK(ctx && d_dinv && d_b && d_x && info, "mp_solve_pcg_chebyshev: null argument");
    MP_CHECK(degree >= 1 && degree <= 16 && lmin > 0.0 && lmax > lmin, "mp_solve_pcg_chebyshev: need 1 <= degree <= 16 and 0 < lmin < lmax");
    MP_TRY(check_csr(A));
    const int64_t n = A->nrow, nc = A->ncol;
    int64_t nnz = 0;
    MP_TRY(nnz_of(ctx, A, &nnz));
    const int lanes = pick_lanes(A, nnz);
    double *w = nullptr;
    MP_TRY(mp::ws_get(ctx, 0, sizeof(double) * (size_t)(4 * n + 2 * nc + 8), (void **)&w));
    double *r = w, *q = w + n, *d = w + 2 * n, *t = w + 3 * n, *p = w + 4 * n, *z = w + 4 * n + nc;   // p, z carry ghost entries
    const int g1 = grid_for(ctx, n);
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    auto precond = [&]() -> int {   // z = p_k(D^-1 A) D^-1 r
        k_cheb_first<<<g1, kThreads, 0, ctx->stream>>>(n, 1.0 / theta, r, d_dinv, z, d, ctx->d_scalars);
        mp::count_launch(ctx);
        double rho0 = 1.0 / sigma;
        for (int i = 1; i < degree; ++i) {
            const double rho1 = 1.0 / (2.0 * sigma - rho0);
            MP_TRY((spmv_halo<0, true>(ctx, A, halo, lanes, z, t, nullptr, 0, 0)));
            k_cheb_step<<<g1, kThreads, 0, ctx->stream>>>(n, rho1 * rho0, 2.0 * rho1 / delta, r, t, d_dinv, z, d, ctx->d_scalars);
            mp::count_launch(ctx);
            rho0 = rho1;
        }
        return 0;
    };
    const int every = info->check_every > 0 ? info->check_every : 10;
    int total = 0;
    for (int restart = 0; restart <= kMaxRestart + 1; ++restart) {   // the last pass only re-evaluates the TRUE residual
        MP_TRY((spmv_halo<0, false>(ctx, A, halo, lanes, d_x, q, nullptr, 0, 0)));
        k_init_residual<<<g1, kThreads, 0, ctx->stream>>>(n, d_b, q, d_dinv, r, nullptr, nullptr, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, S_TMP0);
        mp::count_launch(ctx);
        MP_CUDA(cudaGetLastError());
        bool done = false, zero_rhs = false;
        MP_TRY(start_solve(ctx, info, restart == 0, &done, &zero_rhs));
        if (zero_rhs) return zero_solution(ctx, d_x, n, info);
        if (done || total >= info->maxit || restart > kMaxRestart) break;
        int cur = S_RZ0, nxt = S_RZ1;
        MP_TRY(precond());
        k_cgp_rz<<<g1, kThreads, 0, ctx->stream>>>(n, r, z, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, cur);
        MP_TRY(reduce_slots(ctx, cur, 1));
        k_cgp_pupdate<<<g1, kThreads, 0, ctx->stream>>>(n, 1, z, p, ctx->d_scalars, cur, cur);
        mp::count_launch(ctx, 2);
        for (; total < info->maxit;) {
            MP_TRY((spmv_halo<1, true>(ctx, A, halo, lanes, p, q, p, S_PQ, 0)));
            MP_TRY(reduce_slots(ctx, S_PQ, 1));
            k_cgp_update<<<g1, kThreads, 0, ctx->stream>>>(n, p, q, d_x, r, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, cur);
            MP_TRY(reduce_slots(ctx, S_RR, 1));
            MP_TRY(precond());
            k_cgp_rz<<<g1, kThreads, 0, ctx->stream>>>(n, r, z, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, nxt);
            MP_TRY(reduce_slots(ctx, nxt, 1));
            k_cgp_pupdate<<<g1, kThreads, 0, ctx->stream>>>(n, 0, z, p, ctx->d_scalars, cur, nxt);
            mp::count_launch(ctx, 3);
            int tmp = cur; cur = nxt; nxt = tmp;
            ++total;
            if (total % every == 0 || total == info->maxit) {
                MP_CUDA(cudaGetLastError());
                MP_TRY(poll(ctx));
                if (!(ctx->h_scalars[S_RR] == ctx->h_scalars[S_RR])) { mp::set_error("mp_solve_pcg_chebyshev: NaN residual at iteration %d", total); return 4; }
                if (!(ctx->h_scalars[S_RR] > ctx->h_scalars[S_TOL2])) break;
            }
        }
    }
    return finish_solve(ctx, info);
}

// ------------------------------------------------------------------ line (block-tridiagonal) Jacobi preconditioner
// Rows are numbered plane by plane on structured lattices (BoxMesh: vertex id iz*npl + iy*(nx+1) + ix, mesh.py), so the dofs of the
// grid line through in-plane position l are the rows l, l+stride, l+2 stride, ...  M = the tridiagonal part of A along those lines
// (diagonal + the (i, i -+ stride) entries): "block Jacobi" with one block per line.  The reference's boxes are thin in z
// ([0,4]x[0,4]x[0,1] cut into n^3 hexes, run/CCS/CCS_3D_top_source.py), i.e. the z-coupling of div(k grad p) is 16x the in-plane
// coupling and point Jacobi leaves exactly that direction unpreconditioned; solving the z-lines exactly removes the 1/hz^2 part of
// the condition number (2.5-3x fewer PCG iterations for the same tolerance, measured in profiles/).
// M = L D L^T is factored once (the pressure operator is constant in time): invw = 1/D, and the two sweeps are fused with the CG
// vector updates so that a line-PCG iteration moves FEWER bytes than a Jacobi-PCG iteration (104 N against 116 N):
//   forward  (k_line_fwd):  x += alpha p ; r -= alpha q ; rr ; s = D^-1 L^-1 r ; rz = (L^-1 r)^T D^-1 (L^-1 r)
//   backward (k_line_bwd):  z = L^-T s   ; p = z + beta p          (z is never stored)
// One thread per line, lanes of a warp on neighbouring lines => every access is a coalesced 256-byte row of a plane; the recurrence
// along the line is two dependent DFMAs per plane, the loads of two 4-plane batches are kept in flight per thread.
namespace {
// One warp per CTA and at most 224 registers: 9 CTAs per SM = 1332 slots, so the 1250 warps of a 200^3 sweep (40 000 lines) run as ONE
// wave with 8 or 9 warps per SM.  (64-thread CTAs at 255 registers gave 4 per SM = 592 slots for 625 CTAs: a second wave of 33 CTAs
// doubled the kernel time, ncu launch__waves_per_multiprocessor 1.06, profiles/r2_ncu_line_kernels.md.)
constexpr int kLineThreads = 32;
constexpr int kLineCtasPerSm = 9;
constexpr int kLineU = 4;           // planes per batch

// one thread per 128-byte line asks L2 for a row of a plane that the sweep reaches `pd` planes later: the registers of a thread hold
// only two batches of loads, too few bytes in flight for HBM with 40 000 threads; the prefetches cost no registers
__device__ __forceinline__ void prefetch_l2(const double *a) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a)); }

__global__ void __launch_bounds__(kThreads) k_line_extract(int64_t nrow, int64_t stride, const int32_t *__restrict__ rowptr,
                                                           const int32_t *__restrict__ colidx, const double *__restrict__ vals,
                                                           double *__restrict__ lo, double *__restrict__ diag, double *__restrict__ up) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nrow) return;
    double l = 0.0, d = 0.0, u = 0.0;
    const int64_t below = i - stride, above = i + stride;
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int64_t c = colidx[k];
        if (c == below) l = vals[k];
        if (c == i) d = vals[k];
        if (c == above && above < nrow) u = vals[k];      // columns >= nrow are ghosts of a neighbouring slab: outside the block
    }
    lo[i] = l;
    diag[i] = d;
    up[i] = u;
}

// LU of every line block: invw (in: diagonal of A, out: 1 / pivot), mb (in: the (i, i + stride) entries, out: invw[i] * upper[i], the
// backward multiplier); *bad |= a pivot that is not positive (spd) / zero or NaN (general)
__global__ void __launch_bounds__(kLineThreads) k_line_factor(int64_t stride, int64_t nplanes, const double *__restrict__ lo,
                                                              double *__restrict__ invw, double *__restrict__ mb, int spd, int *bad) {
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < stride; l += gridDim.x * (int64_t)blockDim.x) {
        double ip = 0.0, up_prev = 0.0;
        for (int64_t k = 0; k < nplanes; ++k) {
            const int64_t i = k * stride + l;
            const double w = invw[i] - lo[i] * up_prev * ip;
            if (spd ? !(w > 0.0) : !(w != 0.0)) *bad = 1;
            ip = 1.0 / w;
            invw[i] = ip;
            up_prev = mb[i];
            mb[i] = ip * up_prev;
        }
    }
}

template <bool UPDATE>
__global__ void __launch_bounds__(kLineThreads, kLineCtasPerSm) k_line_fwd(int stride, int nplanes, const double *__restrict__ p,
                                                           const double *__restrict__ q, double *__restrict__ x, double *__restrict__ r,
                                                           const double *__restrict__ lo, const double *__restrict__ invw,
                                                           double *__restrict__ s, double *partials, mp::ReduceCtl *ctl, double *scal,
                                                           int rz_cur, int rz_nxt, int pd) {
    if (!solver_active(scal)) return;
    const double alpha = UPDATE ? safe_div(scal[rz_cur], scal[S_PQ]) : 0.0;
    double acc[2] = {0.0, 0.0};
    const bool pf_lane = pd > 0 && (threadIdx.x & 15) == 0;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < stride; l += gridDim.x * blockDim.x) {
        double sp = 0.0;     // s of the plane below
        auto prefetch = [&](int k0) {
            if (!pf_lane) return;
#pragma unroll
            for (int u = 0; u < kLineU; ++u) {
                const int k = k0 + pd + u;
                if (k < nplanes) {
                    const int i = k * stride + l;
                    if (UPDATE) { prefetch_l2(p + i); prefetch_l2(q + i); prefetch_l2(x + i); }
                    prefetch_l2(r + i); prefetch_l2(lo + i); prefetch_l2(invw + i);
                }
            }
        };
        double pa[kLineU], qa[kLineU], xa[kLineU], ra[kLineU], la[kLineU], wa[kLineU];
        double pb[kLineU], qb[kLineU], xb[kLineU], rb[kLineU], lb[kLineU], wb[kLineU];
        auto load = [&](int k0, double (&pp)[kLineU], double (&qq)[kLineU], double (&xx)[kLineU], double (&rr)[kLineU],
                        double (&ll)[kLineU], double (&ww)[kLineU]) {
#pragma unroll
            for (int u = 0; u < kLineU; ++u) {
                const int k = k0 + u;
                if (k < nplanes) {
                    const int i = k * stride + l;
                    if (UPDATE) { pp[u] = p[i]; qq[u] = q[i]; xx[u] = x[i]; }
                    rr[u] = r[i]; ll[u] = lo[i]; ww[u] = invw[i];
                }
            }
        };
        auto work = [&](int k0, double (&pp)[kLineU], double (&qq)[kLineU], double (&xx)[kLineU], double (&rr)[kLineU],
                        double (&ll)[kLineU], double (&ww)[kLineU]) {
#pragma unroll
            for (int u = 0; u < kLineU; ++u) {
                const int k = k0 + u;
                if (k < nplanes) {
                    const int i = k * stride + l;
                    double ri = rr[u];
                    if (UPDATE) {
                        x[i] = xx[u] + alpha * pp[u];
                        ri -= alpha * qq[u];
                        r[i] = ri;
                        acc[1] += ri * ri;
                    }
                    const double y = ri - ll[u] * sp;
                    sp = ww[u] * y;
                    s[i] = sp;
                    acc[0] += y * sp;
                }
            }
        };
        if (pf_lane)
            for (int k0 = -pd; k0 < 0; k0 += kLineU) prefetch(k0);     // the first pd planes
        load(0, pa, qa, xa, ra, la, wa);
        for (int k0 = 0; k0 < nplanes; k0 += 2 * kLineU) {
            prefetch(k0);
            load(k0 + kLineU, pb, qb, xb, rb, lb, wb);
            work(k0, pa, qa, xa, ra, la, wa);
            prefetch(k0 + kLineU);
            load(k0 + 2 * kLineU, pa, qa, xa, ra, la, wa);
            work(k0 + kLineU, pb, qb, xb, rb, lb, wb);
        }
    }
    if (UPDATE) {
        const int sl_[2] = {rz_nxt, S_RR};
        grid_reduce<2, kLineThreads>(acc, partials, ctl, scal, sl_, true);
    } else {
        double a1[1] = {acc[0]};
        const int sl_[1] = {rz_nxt};
        grid_reduce<1, kLineThreads>(a1, partials, ctl, scal, sl_, false);
    }
}

// OUT == 0: p = z + beta p (FIRST: p = z), the CG direction update;  OUT == 1: z stored to p (plain application z = M^-1 r)
template <bool FIRST>
__global__ void __launch_bounds__(kLineThreads, kLineCtasPerSm) k_line_bwd(int stride, int nplanes, const double *__restrict__ s,
                                                           const double *__restrict__ mb, double *__restrict__ p, const double *scal,
                                                           int rz_cur, int rz_nxt, int guarded, int pd) {
    if (guarded && !solver_active(scal)) return;
    const double beta = FIRST ? 0.0 : safe_div(scal[rz_nxt], scal[rz_cur]);
    constexpr int U = 2 * kLineU;
    const bool pf_lane = pd > 0 && (threadIdx.x & 15) == 0;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < stride; l += gridDim.x * blockDim.x) {
        double zn = 0.0;     // z of the plane above
        auto prefetch = [&](int k0) {        // planes k0 - pd, k0 - pd - 1, ...
            if (!pf_lane) return;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = k0 - pd - u;
                if (k >= 0 && k < nplanes) {
                    const int i = k * stride + l;
                    prefetch_l2(s + i); prefetch_l2(mb + i);
                    if (!FIRST) prefetch_l2(p + i);
                }
            }
        };
        double sa[U], ma[U], pa[U], sb[U], mbb[U], pb[U];
        auto load = [&](int k0, double (&ss)[U], double (&mm)[U], double (&pp)[U]) {   // planes k0, k0-1, ...
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = k0 - u;
                if (k >= 0) {
                    const int i = k * stride + l;
                    ss[u] = s[i]; mm[u] = mb[i];
                    if (!FIRST) pp[u] = p[i];
                }
            }
        };
        auto work = [&](int k0, double (&ss)[U], double (&mm)[U], double (&pp)[U]) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = k0 - u;
                if (k >= 0) {
                    const int i = k * stride + l;
                    zn = ss[u] - mm[u] * zn;
                    p[i] = FIRST ? zn : zn + beta * pp[u];
                }
            }
        };
        if (pf_lane)
            for (int k0 = nplanes - 1 + pd; k0 > nplanes - 1; k0 -= U) prefetch(k0);     // the last pd planes
        load(nplanes - 1, sa, ma, pa);
        for (int k0 = nplanes - 1; k0 >= 0; k0 -= 2 * U) {
            prefetch(k0);
            load(k0 - U, sb, mbb, pb);
            work(k0, sa, ma, pa);
            prefetch(k0 - U);
            load(k0 - 2 * U, sa, ma, pa);
            work(k0 - U, sb, mbb, pb);
        }
    }
}


// BiCGStab with the line preconditioner (right preconditioning: v = A M^-1 p, t = A M^-1 s).  The forward sweep is fused with the
// vector update that produces its input:  MODE 0: e = r + beta (p - omega v), stored to p (first: e = r);  MODE 1: e = r - alpha v,
// stored to s.  out <- D^-1 L^-1 e; the plain backward sweep (k_line_bwd<true>) then gives y = M^-1 p / z = M^-1 s.
template <int MODE>
__global__ void __launch_bounds__(kLineThreads, kLineCtasPerSm) k_line_fwd_bicg(int stride, int nplanes, int first, const double *__restrict__ r,
                                                                                const double *__restrict__ v, double *__restrict__ e,
                                                                                const double *__restrict__ lo, const double *__restrict__ invw,
                                                                                double *__restrict__ out, const double *scal, int rho_cur) {
    if (!solver_active(scal)) return;
    double cb = 0.0, cv = 0.0;     // e = r + cb e_old + cv v
    if (MODE == 0) {
        if (!first) {
            const double omega = safe_div(scal[S_TS], scal[S_TT]);
            cb = safe_div(scal[rho_cur], scal[S_RV]) * safe_div(scal[S_TT], scal[S_TS]);
            cv = -cb * omega;
        }
    } else {
        cv = -safe_div(scal[rho_cur], scal[S_RV]);
    }
    const bool use_e = MODE == 0 && !first, use_v = MODE == 1 || !first;
    constexpr int U = kLineU;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < stride; l += gridDim.x * blockDim.x) {
        double sp = 0.0;
        double ra[U], va[U], ea[U], la[U], wa[U], rb[U], vb[U], eb[U], lb[U], wb[U];
        auto load = [&](int k0, double (&rr)[U], double (&vv)[U], double (&ee)[U], double (&ll)[U], double (&ww)[U]) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = k0 + u;
                if (k < nplanes) {
                    const int i = k * stride + l;
                    rr[u] = r[i]; ll[u] = lo[i]; ww[u] = invw[i];
                    vv[u] = use_v ? v[i] : 0.0;
                    ee[u] = use_e ? e[i] : 0.0;
                }
            }
        };
        auto work = [&](int k0, double (&rr)[U], double (&vv)[U], double (&ee)[U], double (&ll)[U], double (&ww)[U]) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = k0 + u;
                if (k < nplanes) {
                    const int i = k * stride + l;
                    const double ei = rr[u] + cb * ee[u] + cv * vv[u];
                    e[i] = ei;
                    sp = ww[u] * (ei - ll[u] * sp);
                    out[i] = sp;
                }
            }
        };
        load(0, ra, va, ea, la, wa);
        for (int k0 = 0; k0 < nplanes; k0 += 2 * U) {
            load(k0 + U, rb, vb, eb, lb, wb);
            work(k0, ra, va, ea, la, wa);
            load(k0 + 2 * U, ra, va, ea, la, wa);
            work(k0 + U, rb, vb, eb, lb, wb);
        }
    }
}

// Single-reduction (Chronopoulos-Gear) PCG with the line preconditioner, for several GPUs: the vector update of k_cgcg_update fused with
// the upward sweep (p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; f = D^-1 L^-1 r) and the rank-LOCAL sums
// gamma' = r.M^-1 r (= y^T D^-1 y), rr = r.r left for the SpMV kernel, which folds them into its own exchange (DOT == 3): ONE exchange
// point per iteration instead of the classic recurrence's three (halo, p.Ap, {r.z, r.r}).
__global__ void __launch_bounds__(kLineThreads, kLineCtasPerSm) k_line_fwd_cgcg(int stride, int nplanes, int first, const double *__restrict__ w,
                                                                                const double *__restrict__ u, double *__restrict__ p,
                                                                                double *__restrict__ s, double *__restrict__ x, double *__restrict__ r,
                                                                                const double *__restrict__ lo, const double *__restrict__ invw,
                                                                                double *__restrict__ f, double *partials, mp::ReduceCtl *ctl,
                                                                                double *scal, int cur, int nxt) {
    if (!solver_active(scal)) return;
    const double gamma = scal[S_G0 + cur], delta = scal[S_DELTA];
    const double beta = first ? 0.0 : safe_div(gamma, scal[S_G0 + nxt]);
    const double alpha = first ? safe_div(gamma, delta) : safe_div(gamma, delta - beta * safe_div(gamma, scal[S_AL0 + nxt]));
    if (blockIdx.x == 0 && threadIdx.x == 0) scal[S_AL0 + cur] = alpha;          // nobody reads AL[cur] in this kernel
    double acc[2] = {0.0, 0.0};
    constexpr int U = kLineU;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < stride; l += gridDim.x * blockDim.x) {
        double sp = 0.0;
        double ua[U], wa[U], pa[U], sa[U], xa[U], ra[U], la[U], ia[U], ub[U], wb[U], pb[U], sb[U], xb[U], rb[U], lb[U], ib[U];
        auto load = [&](int k0, double (&uu)[U], double (&ww)[U], double (&pp)[U], double (&ss)[U], double (&xx)[U], double (&rr)[U],
                        double (&ll)[U], double (&ii)[U]) {
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int k = k0 + q;
                if (k < nplanes) {
                    const int i = k * stride + l;
                    uu[q] = u[i]; ww[q] = w[i]; xx[q] = x[i]; rr[q] = r[i]; ll[q] = lo[i]; ii[q] = invw[i];
                    pp[q] = first ? 0.0 : p[i];
                    ss[q] = first ? 0.0 : s[i];
                }
            }
        };
        auto work = [&](int k0, double (&uu)[U], double (&ww)[U], double (&pp)[U], double (&ss)[U], double (&xx)[U], double (&rr)[U],
                        double (&ll)[U], double (&ii)[U]) {
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int k = k0 + q;
                if (k < nplanes) {
                    const int i = k * stride + l;
                    const double pi = uu[q] + beta * pp[q], si = ww[q] + beta * ss[q];
                    p[i] = pi;
                    s[i] = si;
                    x[i] = xx[q] + alpha * pi;
                    const double ri = rr[q] - alpha * si;
                    r[i] = ri;
                    acc[1] += ri * ri;
                    const double y = ri - ll[q] * sp;
                    sp = ii[q] * y;
                    f[i] = sp;
                    acc[0] += y * sp;
                }
            }
        };
        load(0, ua, wa, pa, sa, xa, ra, la, ia);
        for (int k0 = 0; k0 < nplanes; k0 += 2 * U) {
            load(k0 + U, ub, wb, pb, sb, xb, rb, lb, ib);
            work(k0, ua, wa, pa, sa, xa, ra, la, ia);
            load(k0 + 2 * U, ua, wa, pa, sa, xa, ra, la, ia);
            work(k0 + U, ub, wb, pb, sb, xb, rb, lb, ib);
        }
    }
    const int sl_[2] = {S_GLOC, S_RRLOC};
    grid_reduce<2, kLineThreads>(acc, partials, ctl, scal, sl_, false, false, /*exchange=*/false);
}

int line_grid(const mp_ctx *ctx, int64_t stride) {
    int64_t g = cdiv(stride, kLineThreads);
    int64_t cap = (int64_t)ctx->num_sm * kLineCtasPerSm;     // beyond one wave the lines are spread evenly by the grid-stride loop
    if (cap > kMaxBlocks) cap = kMaxBlocks;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}
}  // namespace

extern "C" int mp_line_factor(mp_ctx *ctx, int64_t nrow, int64_t stride, int spd, const double *d_lo, double *d_invw, double *d_mb) {
    MP_CHECK(ctx && d_lo && d_invw && d_mb, "mp_line_factor: null argument");
    MP_CHECK(stride > 0 && nrow > 0 && nrow % stride == 0, "mp_line_factor: the row count must be a positive multiple of the line stride");
    int *d_bad = reinterpret_cast<int *>(ctx->d_scalars + S_COUNT + mp::kLocalOff);   // scratch slot (as mp_spectral_bound)
    MP_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    k_line_factor<<<line_grid(ctx, stride), kLineThreads, 0, ctx->stream>>>(stride, nrow / stride, d_lo, d_invw, d_mb, spd, d_bad);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    int bad = 0;
    MP_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MP_CUDA(cudaStreamSynchronize(ctx->stream));
    MP_CHECK(!bad, spd ? "mp_line_factor: a line block is not positive definite (the operator must be SPD)"
                       : "mp_line_factor: zero pivot in a line block");
    return 0;
}

extern "C" int mp_line_setup(mp_ctx *ctx, const mp_csr *A, int64_t stride, int spd, double *d_lo, double *d_invw, double *d_mb) {
    MP_CHECK(ctx && d_lo && d_invw && d_mb, "mp_line_setup: null argument");
    MP_TRY(check_csr(A));
    MP_CHECK(stride > 0 && A->nrow > 0 && A->nrow % stride == 0, "mp_line_setup: the row count must be a positive multiple of the line stride");
    const int64_t n = A->nrow;
    k_line_extract<<<(unsigned)cdiv(n, kThreads), kThreads, 0, ctx->stream>>>(n, stride, A->d_rowptr, A->d_colidx, A->d_vals, d_lo, d_invw, d_mb);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return mp_line_factor(ctx, n, stride, spd, d_lo, d_invw, d_mb);
}

/* z = M^-1 r for the factors of mp_line_setup (tests, other drivers); d_work: nrow doubles */
extern "C" int mp_line_apply(mp_ctx *ctx, int64_t nrow, int64_t stride, const double *d_lo, const double *d_invw, const double *d_mb,
                             const double *d_r, double *d_z, double *d_work) {
    MP_CHECK(ctx && d_lo && d_invw && d_mb && d_r && d_z && d_work, "mp_line_apply: null argument");
    MP_CHECK(stride > 0 && nrow > 0 && nrow % stride == 0, "mp_line_apply: the row count must be a positive multiple of the line stride");
    MP_CHECK(nrow < (int64_t)1 << 31, "mp_line_apply: row count beyond 32-bit indexing");
    // the sweeps are guarded by the solver-active flag: make it true for a stand-alone application
    const double hs[2] = {1.0, 0.0};
    MP_CUDA(cudaMemcpyAsync(ctx->d_scalars + S_RR, &hs[0], sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MP_CUDA(cudaMemcpyAsync(ctx->d_scalars + S_TOL2, &hs[1], sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MP_CUDA(cudaStreamSynchronize(ctx->stream));
    const int g = line_grid(ctx, stride);
    k_line_fwd<false><<<g, kLineThreads, 0, ctx->stream>>>((int)stride, (int)(nrow / stride), nullptr, nullptr, nullptr, const_cast<double *>(d_r), d_lo,
                                                            d_invw, d_work, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, S_TMP0, S_TMP0, ctx->line_prefetch);
    k_line_bwd<true><<<g, kLineThreads, 0, ctx->stream>>>((int)stride, (int)(nrow / stride), d_work, d_mb, d_z, ctx->d_scalars, S_TMP0, S_TMP0, 0, ctx->line_prefetch);
    mp::count_launch(ctx, 2);
    MP_CUDA(cudaGetLastError());
    return 0;
}

namespace {
// Single-reduction line-PCG driver (see k_line_fwd_cgcg); same preconditions as solve_pcg_single_reduction.
int solve_pcg_line_single_reduction(mp_ctx *ctx, const mp_csr *A, int64_t stride, const double *d_lo, const double *d_invw, const double *d_mb,
                                    const double *d_b, double *d_x, mp_halo *halo, mp_solve_info *info) {
    const int64_t n = A->nrow, nc = A->ncol;
    const int st = (int)stride, nplanes = (int)(n / stride);
    const bool use_halo = halo && halo->peer && halo->nneigh > 0 && ctx->nranks > 1;
    if (use_halo) MP_TRY(mp::halo_classify(ctx, halo, A->nrow, A->d_sell_sliceptr, A->d_sell_col));
    mp_halo *H = use_halo ? halo : nullptr;
    double *wk = nullptr;
    MP_TRY(mp::ws_get(ctx, 0, sizeof(double) * (size_t)(6 * n + nc + 8), (void **)&wk));
    double *r = wk, *w = wk + n, *p = wk + 2 * n, *s = wk + 3 * n, *q = wk + 4 * n, *f = wk + 5 * n, *u = wk + 6 * n;   // u carries ghost entries
    const int g1 = grid_for(ctx, n), gl = line_grid(ctx, stride);
    const int every = info->check_every > 0 ? info->check_every : 20;
    int total = 0;
    for (int restart = 0; restart <= kMaxRestart + 1; ++restart) {
        MP_TRY((launch_sell<0, false>(ctx, A, H, d_x, q, nullptr, 0, 0)));
        k_init_residual<<<g1, kThreads, 0, ctx->stream>>>(n, d_b, q, d_invw, r, nullptr, nullptr, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, S_TMP0);
        mp::count_launch(ctx);
        MP_CUDA(cudaGetLastError());
        bool done = false, zero_rhs = false;
        MP_TRY(start_solve(ctx, info, restart == 0, &done, &zero_rhs));
        if (zero_rhs) return zero_solution(ctx, d_x, n, info);
        if (done || total >= info->maxit || restart > kMaxRestart) break;
        // u = M^-1 r, gamma = r.u (allreduced by the sweep's own reduction), w = A u, delta = w.u
        k_line_fwd<false><<<gl, kLineThreads, 0, ctx->stream>>>(st, nplanes, nullptr, nullptr, nullptr, r, d_lo, d_invw, f, ctx->d_partials, ctx->d_ctl,
                                                                 ctx->d_scalars, S_G0, S_G0, 0);
        k_line_bwd<true><<<gl, kLineThreads, 0, ctx->stream>>>(st, nplanes, f, d_mb, u, ctx->d_scalars, S_G0, S_G0, 1, 0);
        mp::count_launch(ctx, 2);
        MP_TRY((launch_sell<1, true>(ctx, A, H, u, w, u, S_DELTA, 0)));
        int cur = 0, nxt = 1;
        for (int it = 0; total < info->maxit; ++it) {
            k_line_fwd_cgcg<<<gl, kLineThreads, 0, ctx->stream>>>(st, nplanes, it == 0 ? 1 : 0, w, u, p, s, d_x, r, d_lo, d_invw, f, ctx->d_partials,
                                                                   ctx->d_ctl, ctx->d_scalars, cur, nxt);
            k_line_bwd<true><<<gl, kLineThreads, 0, ctx->stream>>>(st, nplanes, f, d_mb, u, ctx->d_scalars, S_G0, S_G0, 1, 0);
            mp::count_launch(ctx, 2);
            MP_TRY((launch_sell<3, true>(ctx, A, H, u, w, u, S_DELTA, S_G0 + nxt)));
            int tmp = cur; cur = nxt; nxt = tmp;
            ++total;
            if (total % every == 0 || total == info->maxit) {
                MP_CUDA(cudaGetLastError());
                MP_TRY(poll(ctx));
                if (!(ctx->h_scalars[S_RR] == ctx->h_scalars[S_RR])) { mp::set_error("mp_solve_pcg_line: NaN residual at iteration %d", total); return 4; }
                if (!(ctx->h_scalars[S_RR] > ctx->h_scalars[S_TOL2])) break;
            }
        }
    }
    return finish_solve(ctx, info);
}
}  // namespace

extern "C" int mp_solve_pcg_line(mp_ctx *ctx, const mp_csr *A, int64_t stride, const double *d_lo, const double *d_invw,
                                 const double *d_mb, const double *d_b, double *d_x, mp_halo *halo, mp_solve_info *info) {
    MP_CHECK(ctx && d_lo && d_invw && d_mb && d_b && d_x && info, "mp_solve_pcg_line: null argument");
    MP_TRY(check_csr(A));
    MP_CHECK(stride > 0 && A->nrow > 0 && A->nrow % stride == 0, "mp_solve_pcg_line: the row count must be a positive multiple of the line stride");
    MP_CHECK(A->nrow < (int64_t)1 << 31, "mp_solve_pcg_line: row count beyond 32-bit indexing");
    {   // option pcg_variant = 1: the single-reduction recurrence, one exchange point per iteration.  NOT the multi-GPU default here:
        // measured at 200^3 it is slower on 2 GPUs (50.2 vs 47.5 ms/step: 128 N instead of 104 N bytes per iteration) and no faster on 8
        // (21.3 vs 21.0 ms/step) -- with a third of the iterations left, the exchanges are no longer what limits the line-PCG
        const bool sell = A->d_sell_val && A->d_sell_col && A->d_sell_sliceptr && ctx->spmv_variant == 0;
        const bool multi = ctx->nranks > 1 && halo && halo->nneigh > 0;
        const bool comm_ok = !multi || (ctx->peer && halo->peer && halo->nown == A->nrow);
        const int variant = ctx->pcg_variant == 1 ? 1 : 0;
        if (variant == 1 && sell && comm_ok && (ctx->nranks == 1 || ctx->peer))
            return solve_pcg_line_single_reduction(ctx, A, stride, d_lo, d_invw, d_mb, d_b, d_x, halo, info);
    }
    const int64_t n = A->nrow, nc = A->ncol, nplanes = n / stride;
    int64_t nnz = 0;
    MP_TRY(nnz_of(ctx, A, &nnz));
    const int lanes = pick_lanes(A, nnz);
    double *w = nullptr;
    MP_TRY(mp::ws_get(ctx, 0, sizeof(double) * (size_t)(3 * n + nc + 8), (void **)&w));
    double *r = w, *q = w + n, *s = w + 2 * n, *p = w + 3 * n;   // p carries ghost entries
    const int g1 = grid_for(ctx, n), gl = line_grid(ctx, stride), pd = ctx->line_prefetch;
    const int every = info->check_every > 0 ? info->check_every : 10;   // must not depend on the rank-local size
    int total = 0;
    for (int restart = 0; restart <= kMaxRestart + 1; ++restart) {   // residual replacement, see mp_solve_pcg
        MP_TRY((spmv_halo<0, false>(ctx, A, halo, lanes, d_x, q, nullptr, 0, 0)));
        k_init_residual<<<g1, kThreads, 0, ctx->stream>>>(n, d_b, q, d_invw, r, nullptr, nullptr, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, S_TMP0);
        mp::count_launch(ctx);
        MP_CUDA(cudaGetLastError());
        bool done = false, zero_rhs = false;
        MP_TRY(start_solve(ctx, info, restart == 0, &done, &zero_rhs));
        if (zero_rhs) return zero_solution(ctx, d_x, n, info);
        if (done || total >= info->maxit || restart > kMaxRestart) break;
        int cur = S_RZ0, nxt = S_RZ1;
        k_line_fwd<false><<<gl, kLineThreads, 0, ctx->stream>>>((int)stride, (int)nplanes, nullptr, nullptr, nullptr, r, d_lo, d_invw, s, ctx->d_partials,
                                                                 ctx->d_ctl, ctx->d_scalars, cur, cur, pd);
        MP_TRY(reduce_slots(ctx, cur, 1));
        k_line_bwd<true><<<gl, kLineThreads, 0, ctx->stream>>>((int)stride, (int)nplanes, s, d_mb, p, ctx->d_scalars, cur, cur, 1, pd);
        mp::count_launch(ctx, 2);
        for (; total < info->maxit;) {
            MP_TRY((spmv_halo<1, true>(ctx, A, halo, lanes, p, q, p, S_PQ, 0)));
            MP_TRY(reduce_slots(ctx, S_PQ, 1));
            k_line_fwd<true><<<gl, kLineThreads, 0, ctx->stream>>>((int)stride, (int)nplanes, p, q, d_x, r, d_lo, d_invw, s, ctx->d_partials, ctx->d_ctl,
                                                                    ctx->d_scalars, cur, nxt, pd);
            MP_TRY(reduce_slots(ctx, (nxt < S_RR ? nxt : S_RR), 2));
            k_line_bwd<false><<<gl, kLineThreads, 0, ctx->stream>>>((int)stride, (int)nplanes, s, d_mb, p, ctx->d_scalars, cur, nxt, 1, pd);
            mp::count_launch(ctx, 2);
            int tmp = cur; cur = nxt; nxt = tmp;
            ++total;
            if (total % every == 0 || total == info->maxit) {
                MP_CUDA(cudaGetLastError());
                MP_TRY(poll(ctx));
                if (!(ctx->h_scalars[S_RR] == ctx->h_scalars[S_RR])) { mp::set_error("mp_solve_pcg_line: NaN residual at iteration %d", total); return 4; }
                if (!(ctx->h_scalars[S_RR] > ctx->h_scalars[S_TOL2])) break;
            }
        }
    }
    return finish_solve(ctx, info);
}

/* BiCGStab with the line preconditioner for the non-symmetric SUPG transport row (factors from mp_line_setup(..., spd = 0, ...)). */
extern "C" int mp_solve_bicgstab_line(mp_ctx *ctx, const mp_csr *A, int64_t stride, const double *d_lo, const double *d_invw,
                                      const double *d_mb, const double *d_b, double *d_x, mp_halo *halo, mp_solve_info *info) {
    MP_CHECK(ctx && d_lo && d_invw && d_mb && d_b && d_x && info, "mp_solve_bicgstab_line: null argument");
    MP_TRY(check_csr(A));
    MP_CHECK(stride > 0 && A->nrow > 0 && A->nrow % stride == 0, "mp_solve_bicgstab_line: the row count must be a positive multiple of the line stride");
    MP_CHECK(A->nrow < (int64_t)1 << 31, "mp_solve_bicgstab_line: row count beyond 32-bit indexing");
    const int64_t n = A->nrow, nc = A->ncol;
    const int st = (int)stride, nplanes = (int)(n / stride);
    int64_t nnz = 0;
    MP_TRY(nnz_of(ctx, A, &nnz));
    const int lanes = pick_lanes(A, nnz);
    double *w = nullptr;
    MP_TRY(mp::ws_get(ctx, 0, sizeof(double) * (size_t)(7 * n + 2 * nc + 8), (void **)&w));
    double *r = w, *rhat = w + n, *p = w + 2 * n, *v = w + 3 * n, *s = w + 4 * n, *t = w + 5 * n, *f = w + 6 * n, *y = w + 7 * n, *z = w + 7 * n + nc;
    const int g1 = grid_for(ctx, n), gl = line_grid(ctx, stride);
    const int every = info->check_every > 0 ? info->check_every : 10;   // rank-independent (collective control flow)
    int total = 0;
    for (int restart = 0; restart <= kMaxRestart + 1; ++restart) {   // residual replacement, see mp_solve_pcg
        MP_TRY((spmv_halo<0, false>(ctx, A, halo, lanes, d_x, v, nullptr, 0, 0)));
        k_init_residual<<<g1, kThreads, 0, ctx->stream>>>(n, d_b, v, d_invw, r, nullptr, rhat, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, S_RHO0);
        mp::count_launch(ctx);
        MP_CUDA(cudaGetLastError());
        MP_TRY(reduce_slots(ctx, S_RHO0, 1));
        bool done = false, zero_rhs = false;
        MP_TRY(start_solve(ctx, info, restart == 0, &done, &zero_rhs));
        if (zero_rhs) return zero_solution(ctx, d_x, n, info);
        if (done || total >= info->maxit || restart > kMaxRestart) break;
        int cur = S_RHO0, nxt = S_RHO1;
        for (int it = 0; total < info->maxit; ++it) {
            k_line_fwd_bicg<0><<<gl, kLineThreads, 0, ctx->stream>>>(st, nplanes, it == 0 ? 1 : 0, r, v, p, d_lo, d_invw, f, ctx->d_scalars, cur);
            k_line_bwd<true><<<gl, kLineThreads, 0, ctx->stream>>>(st, nplanes, f, d_mb, y, ctx->d_scalars, cur, cur, 1, 0);
            MP_TRY((spmv_halo<1, true>(ctx, A, halo, lanes, y, v, rhat, S_RV, 0)));
            MP_TRY(reduce_slots(ctx, S_RV, 1));
            k_line_fwd_bicg<1><<<gl, kLineThreads, 0, ctx->stream>>>(st, nplanes, 0, r, v, s, d_lo, d_invw, f, ctx->d_scalars, cur);
            k_line_bwd<true><<<gl, kLineThreads, 0, ctx->stream>>>(st, nplanes, f, d_mb, z, ctx->d_scalars, cur, cur, 1, 0);
            MP_TRY((spmv_halo<2, true>(ctx, A, halo, lanes, z, t, s, S_TS, S_TT)));
            MP_TRY(reduce_slots(ctx, S_TS, 2));
            k_bicg_x<<<g1, kThreads, 0, ctx->stream>>>(n, y, z, s, t, rhat, d_x, r, ctx->d_partials, ctx->d_ctl, ctx->d_scalars, cur, nxt);
            MP_TRY(reduce_slots(ctx, (nxt < S_RR ? nxt : S_RR), 2));
            mp::count_launch(ctx, 5);
            int tmp = cur; cur = nxt; nxt = tmp;
            ++total;
            if (total % every == 0 || total == info->maxit) {
                MP_CUDA(cudaGetLastError());
                MP_TRY(poll(ctx));
                if (!(ctx->h_scalars[S_RR] == ctx->h_scalars[S_RR])) { mp::set_error("mp_solve_bicgstab_line: NaN residual at iteration %d", total); return 4; }
                if (!(ctx->h_scalars[S_RR] > ctx->h_scalars[S_TOL2])) break;
            }
        }
    }
    return finish_solve(ctx, info);
}
