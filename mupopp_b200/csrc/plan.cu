// Scatter plan: CSR pattern + dof->(cell,local) incidence lists + slot map, all built on the GPU,
// and the deterministic (atomic-free) row-gather scatter of element matrices / vectors.
//
// Replaces DOLFIN's SparsityPatternBuilder / GenericDofMap / MatSetValuesLocal (reached from
// assemble_system, e.g. /root/reference/src/modules/mupopp.py:295-297,415-417).
// CUB (part of the CUDA toolkit) is used for the one-time radix sorts / scans of plan construction;
// everything that runs per timestep is a hand-written kernel.
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"

struct mp_plan {
    int lr = 0, lc = 0;
    int64_t ncell = 0, nrow = 0, ncol = 0, nnz = 0;
    int32_t *d_rowptr = nullptr;
    int32_t *d_colidx = nullptr;
    int32_t *d_inc_ptr = nullptr;  // [nrow+1]
    int32_t *d_inc = nullptr;      // [ninc]  id = cell*lr + li, ascending per row
    uint16_t *d_slot = nullptr;    // [ninc*lc] (incidence order) offset of (row(cell,li), col(cell,j)) inside its CSR row
    int32_t *d_pos = nullptr;      // [ncell*lr] incidence index of (cell,li), -1 if the row is not assembled
    int64_t ninc = 0;
    int32_t *d_sell_ptr = nullptr; // [nslice+1] SELL-32 slice offsets (entries)
    int32_t *d_sell_col = nullptr; // [nsell]
    int64_t nslice = 0, nsell = 0;
    // compressed column indices of the SELL mirror ("slice-uniform offsets"): for column position j of slice s, meta[sell_ptr[s]/32 + j] is
    //   (d << 1) | 1   when every row of the slice has col = row + d there (rows numbered along a mesh line share their stencil offsets), or
    //   (k << 1)       otherwise, the 32 explicit columns being xcol[32 k .. 32 k + 31].
    // One int32 per (slice, j) instead of 32: ~8.2-9 bytes per nonzero streamed by SpMV instead of 12 on structured numberings.
    int32_t *d_sell_meta = nullptr;    // [nsell / 32]
    int32_t *d_sell_xcol = nullptr;    // [32 * nexplicit]
    int64_t nexplicit = 0;
    int rows_per_block = 128;      // scatter kernel tiling: nnz of a row block must fit in shared memory
    int smem_doubles = 0;
    int max_row_nnz = 0;
};

namespace {

constexpr int kThreads = 256;
inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

__global__ void k_inc_keys(const int32_t *__restrict__ row_dofs, int64_t n, int32_t nrow, int32_t *keys, int32_t *ids,
                           int32_t *cnt) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    int32_t r = row_dofs[t];
    bool ok = (r >= 0 && r < nrow);
    keys[t] = ok ? r : nrow;  // invalid rows sort to the end
    ids[t] = (int32_t)t;
    if (ok) atomicAdd(&cnt[r], 1);  // integer atomics: result is order independent
}

__global__ void k_pattern_keys(const int32_t *__restrict__ row_dofs, const int32_t *__restrict__ col_dofs, int lr, int lc,
                               int64_t ncell, int32_t nrow, int64_t ncol, uint64_t *keys) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // one thread per (cell, li)
    if (t >= ncell * lr) return;
    int64_t cell = t / lr;
    int32_t r = row_dofs[t];
    bool ok = (r >= 0 && r < nrow);
    for (int j = 0; j < lc; ++j) {
        int32_t c = col_dofs[cell * lc + j];
        keys[t * lc + j] = (ok && c >= 0 && c < ncol) ? (uint64_t)r * (uint64_t)ncol + (uint64_t)c : ~0ull;
    }
}

__global__ void k_unpack_unique(const uint64_t *__restrict__ ukeys, int64_t nu, int64_t ncol, int32_t *colidx, int32_t *rowcnt) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nu) return;
    uint64_t k = ukeys[t];
    if (k == ~0ull) return;
    colidx[t] = (int32_t)(k % (uint64_t)ncol);
    atomicAdd(&rowcnt[k / (uint64_t)ncol], 1);
}

// slot map in INCIDENCE order (thread per incidence entry): the scatter kernels then stream it
__global__ void k_slots(const int32_t *__restrict__ inc, int64_t ninc, const int32_t *__restrict__ row_dofs,
                        const int32_t *__restrict__ col_dofs, int lr, int lc, const int32_t *__restrict__ rowptr,
                        const int32_t *__restrict__ colidx, uint16_t *slot, int32_t *pos) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ninc) return;
    const int32_t id = inc[e];
    const int64_t cell = id / lr;
    const int32_t r = row_dofs[id];
    pos[id] = (int32_t)e;
    const int32_t b = rowptr[r], en = rowptr[r + 1];
    for (int j = 0; j < lc; ++j) {
        const int32_t c = col_dofs[cell * lc + j];
        int32_t lo = b, hi = en;
        while (lo < hi) {  // lower_bound
            const int32_t mid = (lo + hi) >> 1;
            if (colidx[mid] < c) lo = mid + 1; else hi = mid;
        }
        slot[e * lc + j] = (lo < en && colidx[lo] == c) ? (uint16_t)(lo - b) : (uint16_t)0xffff;
    }
}

__global__ void k_inc_pos(const int32_t *__restrict__ inc, int64_t ninc, int32_t *pos) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < ninc) pos[inc[e]] = (int32_t)e;
}

__global__ void k_max_rowlen(const int32_t *__restrict__ rowptr, int32_t nrow, int *out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int v = 0;
    if (t < nrow) v = rowptr[t + 1] - rowptr[t];
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}

// SELL-32: per-slice width = longest row of the slice
__global__ void k_sell_width(const int32_t *__restrict__ rowptr, int32_t nrow, int64_t nslice, int32_t *width32) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (s >= nslice) return;
    const int64_t r = s * 32 + lane;
    int v = r < nrow ? rowptr[r + 1] - rowptr[r] : 0;
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) width32[s] = 32 * v;
}

// fill SELL columns or values.  Padding entries carry value 0 and a column that is VALID for the operator: the row's first
// column (0 for an empty / out-of-range row), never the row index itself -- rectangular blocks (P2xP1: nrow > ncol) would
// otherwise gather x[] past its end, and 0 * NaN/Inf from a neighbouring allocation poisons the row.
template <bool VALS>
__global__ void k_sell_fill(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, const double *__restrict__ vals,
                            int32_t nrow, int64_t nslice, const int32_t *__restrict__ sell_ptr, int32_t *__restrict__ scol,
                            double *__restrict__ sval) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (s >= nslice) return;
    const int64_t r = s * 32 + lane;
    const int32_t b = sell_ptr[s], w = (sell_ptr[s + 1] - b) >> 5;
    const int32_t rb = r < nrow ? rowptr[r] : 0, len = r < nrow ? rowptr[r + 1] - rb : 0;
    for (int j = 0; j < w; ++j) {
        if (VALS) sval[b + 32 * j + lane] = j < len ? vals[rb + j] : 0.0;
        else scol[b + 32 * j + lane] = j < len ? colidx[rb + j] : (len > 0 ? colidx[rb] : 0);
    }
}

// ---- slice-uniform column offsets.  One warp per slice; lane = row.  For column position j the slice is UNIFORM when all real
// entries share d = col - row and row + d is a valid column for the padding lanes too (they multiply by 0 but still gather).
// pass 0: flag[s, j] = 1 if NOT uniform (then exclusive-scanned into explicit-block indices); pass 1: write meta and the explicit columns.
template <int PASS>
__global__ void k_sell_uniform(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ sell_ptr, const int32_t *__restrict__ scol,
                               int32_t nrow, int64_t ncol, int64_t nslice, int32_t *__restrict__ flag_or_meta,
                               const int32_t *__restrict__ xindex, int32_t *__restrict__ xcol) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (s >= nslice) return;
    const int64_t r = s * 32 + lane;
    const int32_t b = sell_ptr[s], w = (sell_ptr[s + 1] - b) >> 5;
    const int32_t len = r < nrow ? rowptr[r + 1] - rowptr[r] : 0;
    for (int j = 0; j < w; ++j) {
        const int32_t c = scol[b + 32 * j + lane];
        const bool real = j < len;
        const int32_t d = c - (int32_t)r;
        // the offset of the first real lane, broadcast
        const unsigned realmask = __ballot_sync(0xffffffffu, real);
        int uniform = 0;
        int32_t d0 = 0;
        if (realmask) {
            d0 = __shfl_sync(0xffffffffu, d, __ffs(realmask) - 1);
            const bool ok = real ? (d == d0) : (r + d0 >= 0 && r + d0 < ncol);     // padding lanes must stay inside x
            uniform = __all_sync(0xffffffffu, ok) ? 1 : 0;
        }
        const int64_t m = (b >> 5) + j;
        if (PASS == 0) {
            if (lane == 0) flag_or_meta[m] = uniform ? 0 : 1;
        } else {
            if (uniform) {
                if (lane == 0) flag_or_meta[m] = (int32_t)(((uint32_t)d0 << 1) | 1u);
            } else {
                const int32_t k = xindex[m];
                if (lane == 0) flag_or_meta[m] = (int32_t)((uint32_t)k << 1);
                xcol[(int64_t)k * 32 + lane] = c;
            }
        }
    }
}

// One CTA owns a contiguous block of rows; its CSR value segment is accumulated in shared memory
// (each thread owns one row => no conflicts, fixed summation order) and written back coalesced.
template <int LC, bool ORDERED>
__global__ void __launch_bounds__(kThreads)
k_scatter_matrix(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc,
                 const uint16_t *__restrict__ slot, const double *__restrict__ emat, double *__restrict__ vals,
                 int32_t nrow, int rows_per_block, int lc_rt) {
    extern __shared__ double sm[];
    const int lc = LC > 0 ? LC : lc_rt;
    int32_t r0 = blockIdx.x * rows_per_block;
    int32_t r1 = min(r0 + rows_per_block, nrow);
    int32_t base = rowptr[r0];
    int32_t seg = rowptr[r1] - base;
    for (int k = threadIdx.x; k < seg; k += blockDim.x) sm[k] = 0.0;
    __syncthreads();
    for (int32_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        double *row = sm + (rowptr[r] - base);
        int32_t e0 = inc_ptr[r], e1 = inc_ptr[r + 1];
        for (int32_t e = e0; e < e1; ++e) {
            const int64_t id = ORDERED ? (int64_t)e : (int64_t)inc[e];
            const double *src = emat + id * lc;
            const uint16_t *sl = slot + (int64_t)e * lc;
            if (LC == 4) {
                const double2 a = *reinterpret_cast<const double2 *>(src);
                const double2 b = *reinterpret_cast<const double2 *>(src + 2);
                const ushort4 s = *reinterpret_cast<const ushort4 *>(sl);
                if (s.x != 0xffff) row[s.x] += a.x;
                if (s.y != 0xffff) row[s.y] += a.y;
                if (s.z != 0xffff) row[s.z] += b.x;
                if (s.w != 0xffff) row[s.w] += b.y;
            } else {
#pragma unroll 4
                for (int j = 0; j < lc; ++j) {
                    uint16_t s = sl[j];
                    if (s != 0xffff) row[s] += src[j];
                }
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < seg; k += blockDim.x) vals[base + k] = sm[k];
}


// LC == 4 (P1 tets), row-ordered input: FOUR lanes per row, lane j owns column j of every incident local row.  A quad reads one
// contiguous 32-byte element row and one 8-byte slot quadruple per step (8 rows -> 8 sectors per warp instruction instead of 32
// scattered lines with a thread per row); per CSR slot the summation order is still ascending incidence index => same bits.
__global__ void __launch_bounds__(kThreads)
k_scatter_matrix_quad(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ inc_ptr, const uint16_t *__restrict__ slot,
                      const double *__restrict__ emat, double *__restrict__ vals, int32_t nrow, int rows_per_block) {
    extern __shared__ double sm[];
    const int32_t r0 = blockIdx.x * rows_per_block;
    const int32_t r1 = min(r0 + rows_per_block, nrow);
    const int32_t base = rowptr[r0];
    const int32_t seg = rowptr[r1] - base;
    for (int k = threadIdx.x; k < seg; k += blockDim.x) sm[k] = 0.0;
    __syncthreads();
    const int j = threadIdx.x & 3;
    for (int32_t r = r0 + (threadIdx.x >> 2); r < r1; r += (blockDim.x >> 2)) {
        double *row = sm + (rowptr[r] - base);
        const int32_t e0 = inc_ptr[r], e1 = inc_ptr[r + 1];
#pragma unroll 4
        for (int32_t e = e0; e < e1; ++e) {
            const double v = emat[(int64_t)e * 4 + j];
            const uint16_t s = slot[(int64_t)e * 4 + j];
            if (s != 0xffff) row[s] += v;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < seg; k += blockDim.x) vals[base + k] = sm[k];
}

template <bool ORDERED>
__global__ void __launch_bounds__(kThreads)
k_scatter_vector(const int32_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc, const double *__restrict__ evec,
                 double beta, double *__restrict__ out, int32_t nrow) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nrow) return;
    int32_t e0 = inc_ptr[r], e1 = inc_ptr[r + 1];
    double s = 0.0;
    for (int32_t e = e0; e < e1; ++e) s += evec[ORDERED ? e : inc[e]];
    out[r] = (beta == 0.0) ? s : beta * out[r] + s;
}

// peer-mode halo: a SELL slice is a BOUNDARY slice when any of its entries references a ghost column (col >= nown)
__global__ void k_slice_kind(const int32_t *__restrict__ sliceptr, const int32_t *__restrict__ scol, int64_t nslice, int64_t nown,
                             uint8_t *__restrict__ kind) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (s >= nslice) return;
    int any = 0;
    for (int32_t k = sliceptr[s] + lane; k < sliceptr[s + 1]; k += 32) any |= scol[k] >= nown ? 1 : 0;
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) kind[s] = any ? 1 : 0;
}

}  // namespace

// Slice lists of `halo` for one operator (cached on the halo by the identity of the SELL column array: the operators of a
// timestep share one plan).  The boundary list is produced by an order-preserving selection => ascending => reproducible sums.
int mp::halo_classify(mp_ctx *ctx, mp_halo *h, int64_t nrow, const int32_t *d_sliceptr, const int32_t *d_scol) {
    if (h->classified_for == (const void *)d_scol) return 0;
    MP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t nslice = (nrow + 31) / 32;
    cudaFree(h->d_slice_kind); cudaFree(h->d_bslices);
    h->d_slice_kind = nullptr; h->d_bslices = nullptr;
    MP_CUDA(cudaMalloc(&h->d_slice_kind, (size_t)(nslice > 0 ? nslice : 1)));
    MP_CUDA(cudaMalloc(&h->d_bslices, sizeof(int32_t) * (size_t)(nslice > 0 ? nslice : 1)));
    int32_t nb = 0;
    if (nslice > 0) {
        k_slice_kind<<<(unsigned)cdiv(nslice * 32, kThreads), kThreads, 0, st>>>(d_sliceptr, d_scol, nslice, h->nown, h->d_slice_kind);
        mp::count_launch(ctx);
        int32_t *d_n = nullptr;
        void *tmp = nullptr;
        MP_CUDA(cudaMalloc(&d_n, sizeof(int32_t)));
        size_t tb = 0;
        cub::CountingInputIterator<int32_t> ids(0);
        cudaError_t e = cub::DeviceSelect::Flagged(nullptr, tb, ids, h->d_slice_kind, h->d_bslices, d_n, (int)nslice, st);
        if (e == cudaSuccess) e = cudaMalloc(&tmp, tb + 16);
        if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(tmp, tb, ids, h->d_slice_kind, h->d_bslices, d_n, (int)nslice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&nb, d_n, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(tmp); cudaFree(d_n);
        if (e != cudaSuccess) { mp::set_error("halo_classify: %s", cudaGetErrorString(e)); return 1; }
    }
    h->h_dev.slice_kind = h->d_slice_kind;
    h->h_dev.bslices = h->d_bslices;
    h->h_dev.nbslice = nb;
    // only the classification fields change: the sequence number lives on the device and must not be reset
    MP_CUDA(cudaMemcpyAsync(&h->d_dev->slice_kind, &h->h_dev.slice_kind, sizeof(void *), cudaMemcpyHostToDevice, st));
    MP_CUDA(cudaMemcpyAsync(&h->d_dev->bslices, &h->h_dev.bslices, sizeof(void *), cudaMemcpyHostToDevice, st));
    MP_CUDA(cudaMemcpyAsync(&h->d_dev->nbslice, &h->h_dev.nbslice, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    MP_CUDA(cudaStreamSynchronize(st));
    h->classified_for = (const void *)d_scol;
    return 0;
}

extern "C" int mp_plan_destroy(mp_plan *p) {
    if (!p) return 0;
    cudaFree(p->d_rowptr); cudaFree(p->d_colidx); cudaFree(p->d_inc_ptr); cudaFree(p->d_inc); cudaFree(p->d_slot);
    cudaFree(p->d_sell_ptr); cudaFree(p->d_sell_col); cudaFree(p->d_pos);
    cudaFree(p->d_sell_meta); cudaFree(p->d_sell_xcol);
    delete p;
    return 0;
}

extern "C" int64_t mp_plan_nnz(const mp_plan *p) { return p ? p->nnz : -1; }
extern "C" int32_t mp_plan_max_row_nnz(const mp_plan *p) { return p ? p->max_row_nnz : -1; }

extern "C" int mp_plan_create(mp_ctx *ctx, const int32_t *d_row_dofs, int32_t lr, const int32_t *d_col_dofs, int32_t lc,
                              int64_t ncell, int64_t nrow, int64_t ncol, mp_plan **out) {
    MP_CHECK(ctx && d_row_dofs && out, "mp_plan_create: null argument");
    MP_CHECK(lr > 0 && ncell >= 0 && nrow > 0, "mp_plan_create: bad sizes");
    MP_CHECK(ncell * lr < (int64_t)INT32_MAX, "mp_plan_create: ncell*lr exceeds int32");
    MP_CHECK(nrow < INT32_MAX && ncol < INT32_MAX, "mp_plan_create: nrow/ncol exceed int32");
    MP_CHECK(lc == 0 || d_col_dofs, "mp_plan_create: lc > 0 but d_col_dofs is null");
    MP_CHECK(ncell * lr * (int64_t)(lc > 0 ? lc : 1) < (int64_t)INT32_MAX, "mp_plan_create: ncell*lr*lc exceeds int32");
    MP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    mp_plan *p = new mp_plan();
    p->lr = lr; p->lc = lc; p->ncell = ncell; p->nrow = nrow; p->ncol = ncol;
    const int64_t ninc_all = ncell * lr;
    int rc = 0;
    int32_t *keys = nullptr, *keys2 = nullptr, *ids = nullptr, *ids2 = nullptr, *cnt = nullptr;
    void *tmp = nullptr;
    uint64_t *pk = nullptr, *pk2 = nullptr, *uk = nullptr;
    int64_t *d_nu = nullptr;
    int32_t *rowcnt = nullptr;
    int *d_max = nullptr;
    auto cleanup = [&]() {
        cudaFree(keys); cudaFree(keys2); cudaFree(ids); cudaFree(ids2); cudaFree(cnt); cudaFree(tmp);
        cudaFree(pk); cudaFree(pk2); cudaFree(uk); cudaFree(d_nu); cudaFree(rowcnt); cudaFree(d_max);
    };
#define PC(call)                                                                                    \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            mp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));    \
            cleanup(); mp_plan_destroy(p); return 1;                                                \
        }                                                                                           \
    } while (0)

    // ---- 1. incidence lists: stable radix sort of (row, id) pairs => ascending cell id inside each row
    PC(cudaMalloc(&keys, sizeof(int32_t) * (ninc_all + 1)));
    PC(cudaMalloc(&keys2, sizeof(int32_t) * (ninc_all + 1)));
    PC(cudaMalloc(&ids, sizeof(int32_t) * (ninc_all + 1)));
    PC(cudaMalloc(&ids2, sizeof(int32_t) * (ninc_all + 1)));
    PC(cudaMalloc(&cnt, sizeof(int32_t) * (nrow + 1)));
    PC(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (nrow + 1), st));
    if (ninc_all > 0) {
        k_inc_keys<<<(unsigned)cdiv(ninc_all, kThreads), kThreads, 0, st>>>(d_row_dofs, ninc_all, (int32_t)nrow, keys, ids, cnt);
        mp::count_launch(ctx);
    }
    {
        int bits = 1;
        while ((1ll << bits) <= nrow) ++bits;  // keys in [0, nrow]
        size_t tb = 0;
        PC(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, ids, ids2, (int)ninc_all, 0, bits, st));
        PC(cudaMalloc(&tmp, tb + 16));
        PC(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, ids, ids2, (int)ninc_all, 0, bits, st));
        cudaFree(tmp); tmp = nullptr;
    }
    PC(cudaMalloc(&p->d_inc_ptr, sizeof(int32_t) * (nrow + 1)));
    {
        size_t tb = 0;
        PC(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, p->d_inc_ptr, (int)(nrow + 1), st));
        PC(cudaMalloc(&tmp, tb + 16));
        PC(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, p->d_inc_ptr, (int)(nrow + 1), st));
        cudaFree(tmp); tmp = nullptr;
    }
    int32_t ninc = 0;
    PC(cudaMemcpyAsync(&ninc, p->d_inc_ptr + nrow, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PC(cudaStreamSynchronize(st));
    p->ninc = ninc;
    PC(cudaMalloc(&p->d_inc, sizeof(int32_t) * (size_t)(ninc > 0 ? ninc : 1)));
    PC(cudaMemcpyAsync(p->d_inc, ids2, sizeof(int32_t) * (size_t)ninc, cudaMemcpyDeviceToDevice, st));
    PC(cudaStreamSynchronize(st));
    cudaFree(keys); cudaFree(keys2); cudaFree(ids); cudaFree(ids2); cudaFree(cnt);
    keys = keys2 = ids = ids2 = cnt = nullptr;
    PC(cudaMalloc(&p->d_pos, sizeof(int32_t) * (size_t)(ninc_all > 0 ? ninc_all : 1)));
    PC(cudaMemsetAsync(p->d_pos, 0xff, sizeof(int32_t) * (size_t)(ninc_all > 0 ? ninc_all : 1), st));   // -1
    if (ninc > 0) {
        k_inc_pos<<<(unsigned)cdiv(ninc, kThreads), kThreads, 0, st>>>(p->d_inc, ninc, p->d_pos);
        mp::count_launch(ctx);
    }

    // ---- 2. CSR pattern: sort 64-bit (row,col) keys of every local entry, unique
    if (lc > 0) {
        const int64_t nkeys = ninc_all * lc;
        if (nkeys >= (int64_t)INT32_MAX) { mp::set_error("mp_plan_create: ncell*lr*lc exceeds int32"); cleanup(); mp_plan_destroy(p); return 2; }
        PC(cudaMalloc(&pk, sizeof(uint64_t) * (nkeys + 1)));
        PC(cudaMalloc(&pk2, sizeof(uint64_t) * (nkeys + 1)));
        k_pattern_keys<<<(unsigned)cdiv(ninc_all, kThreads), kThreads, 0, st>>>(d_row_dofs, d_col_dofs, lr, lc, ncell, (int32_t)nrow, ncol, pk);
        mp::count_launch(ctx);
        {
            size_t tb = 0;
            PC(cub::DeviceRadixSort::SortKeys(nullptr, tb, pk, pk2, (int)nkeys, 0, 64, st));
            PC(cudaMalloc(&tmp, tb + 16));
            PC(cub::DeviceRadixSort::SortKeys(tmp, tb, pk, pk2, (int)nkeys, 0, 64, st));
            cudaFree(tmp); tmp = nullptr;
        }
        uk = pk;  // reuse the unsorted buffer for the unique keys
        pk = nullptr;
        PC(cudaMalloc(&d_nu, sizeof(int64_t)));
        {
            size_t tb = 0;
            PC(cub::DeviceSelect::Unique(nullptr, tb, pk2, uk, d_nu, (int)nkeys, st));
            PC(cudaMalloc(&tmp, tb + 16));
            PC(cub::DeviceSelect::Unique(tmp, tb, pk2, uk, d_nu, (int)nkeys, st));
            cudaFree(tmp); tmp = nullptr;
        }
        int64_t nu = 0;
        PC(cudaMemcpyAsync(&nu, d_nu, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        PC(cudaStreamSynchronize(st));
        cudaFree(pk2); pk2 = nullptr;
        // the sentinel (invalid rows/cols), if present, is the last unique key
        uint64_t lastkey = 0;
        if (nu > 0) PC(cudaMemcpy(&lastkey, uk + (nu - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost));
        int64_t nnz = (nu > 0 && lastkey == ~0ull) ? nu - 1 : nu;
        p->nnz = nnz;
        PC(cudaMalloc(&p->d_colidx, sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1)));
        PC(cudaMalloc(&rowcnt, sizeof(int32_t) * (nrow + 1)));
        PC(cudaMemsetAsync(rowcnt, 0, sizeof(int32_t) * (nrow + 1), st));
        if (nnz > 0) {
            k_unpack_unique<<<(unsigned)cdiv(nnz, kThreads), kThreads, 0, st>>>(uk, nnz, ncol, p->d_colidx, rowcnt);
            mp::count_launch(ctx);
        }
        PC(cudaMalloc(&p->d_rowptr, sizeof(int32_t) * (nrow + 1)));
        {
            size_t tb = 0;
            PC(cub::DeviceScan::ExclusiveSum(nullptr, tb, rowcnt, p->d_rowptr, (int)(nrow + 1), st));
            PC(cudaMalloc(&tmp, tb + 16));
            PC(cub::DeviceScan::ExclusiveSum(tmp, tb, rowcnt, p->d_rowptr, (int)(nrow + 1), st));
            cudaFree(tmp); tmp = nullptr;
        }
        cudaFree(uk); uk = nullptr;
        // ---- 3. slot map + row-block tiling of the scatter kernel
        PC(cudaMalloc(&p->d_slot, sizeof(uint16_t) * (size_t)(ninc > 0 ? (int64_t)ninc * lc : 1)));
        if (ninc > 0) {
            k_slots<<<(unsigned)cdiv(ninc, kThreads), kThreads, 0, st>>>(p->d_inc, ninc, d_row_dofs, d_col_dofs, lr, lc, p->d_rowptr,
                                                                         p->d_colidx, p->d_slot, p->d_pos);
            mp::count_launch(ctx);
        }
        PC(cudaMalloc(&d_max, sizeof(int)));
        PC(cudaMemsetAsync(d_max, 0, sizeof(int), st));
        k_max_rowlen<<<(unsigned)cdiv(nrow, kThreads), kThreads, 0, st>>>(p->d_rowptr, (int32_t)nrow, d_max);
        mp::count_launch(ctx);
        int mx = 0;
        PC(cudaMemcpyAsync(&mx, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
        PC(cudaStreamSynchronize(st));
        p->max_row_nnz = mx;
        if (mx >= 0xffff) { mp::set_error("mp_plan_create: row with %d entries exceeds the 16-bit slot map", mx); cleanup(); mp_plan_destroy(p); return 2; }
        const int cap = 12288;  // doubles of shared memory per CTA (96 KB)
        int rpb = mx > 0 ? cap / mx : kThreads;
        if (rpb > kThreads) rpb = kThreads;
        if (rpb >= 32) rpb &= ~31;
        if (rpb < 1) { mp::set_error("mp_plan_create: row with %d entries does not fit the scatter tile", mx); cleanup(); mp_plan_destroy(p); return 2; }
        p->rows_per_block = rpb;
        p->smem_doubles = rpb * (mx > 0 ? mx : 1);
        // ---- 4. SELL-32 mirror of the pattern (SpMV layout)
        p->nslice = (nrow + 31) / 32;
        int32_t *width = nullptr;
        PC(cudaMalloc(&width, sizeof(int32_t) * (p->nslice + 1)));
        PC(cudaMemsetAsync(width, 0, sizeof(int32_t) * (p->nslice + 1), st));
        k_sell_width<<<(unsigned)cdiv(p->nslice * 32, kThreads), kThreads, 0, st>>>(p->d_rowptr, (int32_t)nrow, p->nslice, width);
        mp::count_launch(ctx);
        PC(cudaMalloc(&p->d_sell_ptr, sizeof(int32_t) * (p->nslice + 1)));
        {
            size_t tb = 0;
            PC(cub::DeviceScan::ExclusiveSum(nullptr, tb, width, p->d_sell_ptr, (int)(p->nslice + 1), st));
            PC(cudaMalloc(&tmp, tb + 16));
            PC(cub::DeviceScan::ExclusiveSum(tmp, tb, width, p->d_sell_ptr, (int)(p->nslice + 1), st));
            cudaFree(tmp); tmp = nullptr;
        }
        int32_t nsell = 0;
        PC(cudaMemcpyAsync(&nsell, p->d_sell_ptr + p->nslice, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        PC(cudaStreamSynchronize(st));
        cudaFree(width);
        p->nsell = nsell;
        PC(cudaMalloc(&p->d_sell_col, sizeof(int32_t) * (size_t)(nsell > 0 ? nsell : 1)));
        k_sell_fill<false><<<(unsigned)cdiv(p->nslice * 32, kThreads), kThreads, 0, st>>>(p->d_rowptr, p->d_colidx, nullptr, (int32_t)nrow,
                                                                                          p->nslice, p->d_sell_ptr, p->d_sell_col, nullptr);
        mp::count_launch(ctx);
        PC(cudaStreamSynchronize(st));
        // ---- 5. slice-uniform column offsets (compressed SELL columns)
        if (nsell > 0) {
            const int64_t nmeta = nsell / 32;
            int32_t *flag = nullptr, *xindex = nullptr;
            PC(cudaMalloc(&flag, sizeof(int32_t) * (nmeta + 1)));
            PC(cudaMalloc(&xindex, sizeof(int32_t) * (nmeta + 1)));
            PC(cudaMemsetAsync(flag, 0, sizeof(int32_t) * (nmeta + 1), st));
            const unsigned gu = (unsigned)cdiv(p->nslice * 32, kThreads);
            k_sell_uniform<0><<<gu, kThreads, 0, st>>>(p->d_rowptr, p->d_sell_ptr, p->d_sell_col, (int32_t)nrow, ncol, p->nslice, flag, nullptr, nullptr);
            mp::count_launch(ctx);
            cudaError_t e = cudaSuccess;
            {
                size_t tb = 0;
                e = cub::DeviceScan::ExclusiveSum(nullptr, tb, flag, xindex, (int)(nmeta + 1), st);
                if (e == cudaSuccess) e = cudaMalloc(&tmp, tb + 16);
                if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tb, flag, xindex, (int)(nmeta + 1), st);
                cudaFree(tmp); tmp = nullptr;
            }
            int32_t nx = 0;
            if (e == cudaSuccess) e = cudaMemcpyAsync(&nx, xindex + nmeta, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            p->nexplicit = nx;
            if (e == cudaSuccess) e = cudaMalloc(&p->d_sell_meta, sizeof(int32_t) * (size_t)nmeta);
            if (e == cudaSuccess) e = cudaMalloc(&p->d_sell_xcol, sizeof(int32_t) * 32 * (size_t)(nx > 0 ? nx : 1));
            if (e == cudaSuccess) {
                k_sell_uniform<1><<<gu, kThreads, 0, st>>>(p->d_rowptr, p->d_sell_ptr, p->d_sell_col, (int32_t)nrow, ncol, p->nslice, p->d_sell_meta,
                                                           xindex, p->d_sell_xcol);
                mp::count_launch(ctx);
                e = cudaStreamSynchronize(st);
            }
            cudaFree(flag); cudaFree(xindex);
            if (e != cudaSuccess) {
                mp::set_error("mp_plan_create (compressed SELL columns): %s", cudaGetErrorString(e));
                cleanup(); mp_plan_destroy(p); return 1;
            }
        }
    }
#undef PC
    cleanup();
    *out = p;
    return 0;
}

extern "C" int mp_plan_export_csr(mp_ctx *ctx, const mp_plan *p, int32_t *d_rowptr, int32_t *d_colidx) {
    MP_CHECK(ctx && p && p->lc > 0, "mp_plan_export_csr: plan has no matrix pattern");
    MP_CUDA(cudaMemcpyAsync(d_rowptr, p->d_rowptr, sizeof(int32_t) * (p->nrow + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    MP_CUDA(cudaMemcpyAsync(d_colidx, p->d_colidx, sizeof(int32_t) * (size_t)p->nnz, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

static int scatter_matrix_impl(mp_ctx *ctx, const mp_plan *p, const double *d_elem_mat, double *d_vals, bool ordered) {
    MP_CHECK(ctx && p && p->lc > 0 && d_elem_mat && d_vals, "mp_scatter_matrix: bad argument");
    if (p->nnz == 0) return 0;
    size_t smem = sizeof(double) * (size_t)p->smem_doubles;
    unsigned grid = (unsigned)cdiv(p->nrow, p->rows_per_block);
    const int kThreadsRt = p->rows_per_block >= kThreads ? kThreads : ((p->rows_per_block + 31) / 32) * 32;
#define LAUNCH_SM(LC_, ORD_)                                                                                                          \
    do {                                                                                                                              \
        MP_CUDA(cudaFuncSetAttribute(k_scatter_matrix<LC_, ORD_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        k_scatter_matrix<LC_, ORD_><<<grid, kThreadsRt, smem, ctx->stream>>>(p->d_rowptr, p->d_inc_ptr, p->d_inc, p->d_slot, d_elem_mat, \
                                                                              d_vals, (int32_t)p->nrow, p->rows_per_block, p->lc);    \
    } while (0)
    if (p->lc == 4 && ordered) {
        MP_CUDA(cudaFuncSetAttribute(k_scatter_matrix_quad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_scatter_matrix_quad<<<grid, kThreads, smem, ctx->stream>>>(p->d_rowptr, p->d_inc_ptr, p->d_slot, d_elem_mat, d_vals,
                                                                      (int32_t)p->nrow, p->rows_per_block);
    } else if (p->lc == 4) { LAUNCH_SM(4, false); }
    else { if (ordered) LAUNCH_SM(0, true); else LAUNCH_SM(0, false); }
#undef LAUNCH_SM
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mp_scatter_matrix(mp_ctx *ctx, const mp_plan *p, const double *d_elem_mat, double *d_vals) {
    return scatter_matrix_impl(ctx, p, d_elem_mat, d_vals, false);
}

extern "C" int mp_scatter_matrix_ordered(mp_ctx *ctx, const mp_plan *p, const double *d_elem_mat, double *d_vals) {
    return scatter_matrix_impl(ctx, p, d_elem_mat, d_vals, true);
}

static int scatter_vector_impl(mp_ctx *ctx, const mp_plan *p, const double *d_elem_vec, double beta, double *d_out, bool ordered) {
    MP_CHECK(ctx && p && d_elem_vec && d_out, "mp_scatter_vector: bad argument");
    const unsigned grid = (unsigned)cdiv(p->nrow, kThreads);
    if (ordered) k_scatter_vector<true><<<grid, kThreads, 0, ctx->stream>>>(p->d_inc_ptr, p->d_inc, d_elem_vec, beta, d_out, (int32_t)p->nrow);
    else k_scatter_vector<false><<<grid, kThreads, 0, ctx->stream>>>(p->d_inc_ptr, p->d_inc, d_elem_vec, beta, d_out, (int32_t)p->nrow);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mp_scatter_vector(mp_ctx *ctx, const mp_plan *p, const double *d_elem_vec, double beta, double *d_out) {
    return scatter_vector_impl(ctx, p, d_elem_vec, beta, d_out, false);
}

extern "C" int mp_scatter_vector_ordered(mp_ctx *ctx, const mp_plan *p, const double *d_elem_vec, double beta, double *d_out) {
    return scatter_vector_impl(ctx, p, d_elem_vec, beta, d_out, true);
}

extern "C" int mp_plan_row_positions(const mp_plan *p, const int32_t **d_pos) {
    MP_CHECK(p && d_pos && p->d_pos, "mp_plan_row_positions: bad argument");
    *d_pos = p->d_pos;
    return 0;
}

extern "C" int mp_plan_incidence(const mp_plan *p, int64_t *nrow, int32_t *lr, int32_t *lc, int32_t *max_row_nnz, const int32_t **d_rowptr,
                                 const int32_t **d_inc_ptr, const int32_t **d_inc, const uint16_t **d_slot) {
    MP_CHECK(p && p->d_inc_ptr && p->d_inc, "mp_plan_incidence: bad argument");
    if (nrow) *nrow = p->nrow;
    if (lr) *lr = p->lr;
    if (lc) *lc = p->lc;
    if (max_row_nnz) *max_row_nnz = p->max_row_nnz;
    if (d_rowptr) *d_rowptr = p->d_rowptr;
    if (d_inc_ptr) *d_inc_ptr = p->d_inc_ptr;
    if (d_inc) *d_inc = p->d_inc;
    if (d_slot) *d_slot = p->d_slot;
    return 0;
}

extern "C" int mp_plan_sell_info(const mp_plan *p, int64_t *nslice, int64_t *nentries, const int32_t **d_sliceptr, const int32_t **d_col) {
    MP_CHECK(p && p->lc > 0 && p->d_sell_ptr, "mp_plan_sell_info: plan has no matrix pattern");
    if (nslice) *nslice = p->nslice;
    if (nentries) *nentries = p->nsell;
    if (d_sliceptr) *d_sliceptr = p->d_sell_ptr;
    if (d_col) *d_col = p->d_sell_col;
    return 0;
}

extern "C" int mp_plan_sellu_info(const mp_plan *p, int64_t *nexplicit, const int32_t **d_meta, const int32_t **d_xcol) {
    MP_CHECK(p && nexplicit, "mp_plan_sellu_info: null argument");
    *nexplicit = p->d_sell_meta ? p->nexplicit : -1;
    if (d_meta) *d_meta = p->d_sell_meta;
    if (d_xcol) *d_xcol = p->d_sell_meta ? p->d_sell_xcol : nullptr;
    return 0;
}

extern "C" int mp_csr_to_sell(mp_ctx *ctx, const mp_plan *p, const double *d_csr_vals, double *d_sell_vals) {
    MP_CHECK(ctx && p && p->d_sell_ptr && d_csr_vals && d_sell_vals, "mp_csr_to_sell: bad argument");
    if (p->nsell == 0) return 0;
    k_sell_fill<true><<<(unsigned)cdiv(p->nslice * 32, kThreads), kThreads, 0, ctx->stream>>>(p->d_rowptr, p->d_colidx, d_csr_vals, (int32_t)p->nrow,
                                                                                               p->nslice, p->d_sell_ptr, nullptr, d_sell_vals);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return 0;
}
