// SpMV kernel laboratory (stand-alone): the 15-point P1 operator of the 6-tet BoxMesh at n^3, several kernel variants,
// timed with CUDA events.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o spmv_lab spmv_lab.cu
// Used to choose the production SpMV in mupopp_b200/csrc/krylov.cu (results in profiles/).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int T = 256;

// V0: stream the matrix arrays only (no gather): lower bound
__global__ void __launch_bounds__(T) k_stream(int64_t nnz, const int32_t *__restrict__ col, const double *__restrict__ val, double *out) {
    double s = 0.0;
    const int64_t stride = gridDim.x * (int64_t)blockDim.x;
    for (int64_t k0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k0 < nnz; k0 += 4 * stride) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { int64_t k = k0 + u * stride; if (k < nnz) s += val[k] * (double)col[k]; }
    }
    if (s == 1.2345e300) out[0] = s;
}

// V1: CTA-stream (products into shared memory, one thread per row sums)
__global__ void __launch_bounds__(T) k_cta_stream(int64_t nrow, int rpb, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                                  const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y) {
    extern __shared__ double prod[];
    const int64_t nblk = (nrow + rpb - 1) / rpb;
    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int64_t r0 = blk * rpb, r1 = min(r0 + (int64_t)rpb, nrow);
        const int32_t k0 = rowptr[r0], k1 = rowptr[r1];
#pragma unroll 4
        for (int32_t k = k0 + threadIdx.x; k < k1; k += T) prod[k - k0] = vals[k] * x[colidx[k]];
        __syncthreads();
        for (int64_t r = r0 + threadIdx.x; r < r1; r += T) {
            const int32_t b = rowptr[r] - k0, e = rowptr[r + 1] - k0;
            double sum = 0.0;
            for (int32_t k = b; k < e; ++k) sum += prod[k];
            y[r] = sum;
        }
        __syncthreads();
    }
}

// V2: CTA-stream staging (col,val) in shared memory, then thread-per-row gathers (row-aligned gather addresses)
__global__ void __launch_bounds__(T) k_cta_stage(int64_t nrow, int rpb, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                                 const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y, int cap) {
    extern __shared__ double sm[];
    double *sval = sm;
    int32_t *scol = reinterpret_cast<int32_t *>(sm + cap);
    const int64_t nblk = (nrow + rpb - 1) / rpb;
    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int64_t r0 = blk * rpb, r1 = min(r0 + (int64_t)rpb, nrow);
        const int32_t k0 = rowptr[r0], k1 = rowptr[r1];
#pragma unroll 4
        for (int32_t k = k0 + threadIdx.x; k < k1; k += T) { sval[k - k0] = vals[k]; scol[k - k0] = colidx[k]; }
        __syncthreads();
        for (int64_t r = r0 + threadIdx.x; r < r1; r += T) {
            const int32_t b = rowptr[r] - k0, e = rowptr[r + 1] - k0;
            double sum = 0.0;
#pragma unroll 4
            for (int32_t k = b; k < e; ++k) sum += sval[k] * x[scol[k]];
            y[r] = sum;
        }
        __syncthreads();
    }
}

// V3: SELL-32: slice s holds rows 32s..32s+31, entries stored [j][lane]; sliceptr in units of entries
__global__ void __launch_bounds__(T) k_sell(int64_t nrow, const int64_t *__restrict__ sliceptr, const int32_t *__restrict__ scol,
                                            const double *__restrict__ sval, const double *__restrict__ x, double *__restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    const int64_t nslice = (nrow + 31) >> 5;
    for (int64_t s = warp; s < nslice; s += nwarp) {
        const int64_t b = sliceptr[s], e = sliceptr[s + 1];
        double sum = 0.0;
#pragma unroll 4
        for (int64_t k = b + lane; k < e; k += 32) sum += sval[k] * x[scol[k]];
        const int64_t row = (s << 5) + lane;
        if (row < nrow) y[row] = sum;
    }
}

// V3b: SELL-32 with fixed slice width W (ELL-like), fully unrolled
template <int W>
__global__ void __launch_bounds__(T) k_sell_fixed(int64_t nrow, const int32_t *__restrict__ scol, const double *__restrict__ sval,
                                                  const double *__restrict__ x, double *__restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    const int64_t nslice = (nrow + 31) >> 5;
    for (int64_t s = warp; s < nslice; s += nwarp) {
        const int64_t b = s * (32 * W) + lane;
        int32_t c[W];
        double v[W];
#pragma unroll
        for (int j = 0; j < W; ++j) { c[j] = scol[b + 32 * j]; v[j] = sval[b + 32 * j]; }
        double sum = 0.0;
#pragma unroll
        for (int j = 0; j < W; ++j) sum += v[j] * x[c[j]];
        const int64_t row = (s << 5) + lane;
        if (row < nrow) y[row] = sum;
    }
}

// V4: sub-warp (8 lanes) per row
__global__ void __launch_bounds__(T) k_vec8(int64_t nrow, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                            const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y) {
    const int lane = threadIdx.x & 31, sub = lane / 8, sl = lane % 8;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t base = warp * 4; base < nrow; base += nwarp * 4) {
        const int64_t row = base + sub;
        double sum = 0.0;
        if (row < nrow) {
            const int32_t b = rowptr[row], e = rowptr[row + 1];
            for (int32_t k = b + sl; k < e; k += 8) sum += vals[k] * x[colidx[k]];
        }
        for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (row < nrow && sl == 0) y[row] = sum;
    }
}

template <typename F>
float timeit(F f, int reps = 20) {
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char **argv) {
    int n = argc > 1 ? atoi(argv[1]) : 200;
    const int64_t n1 = n + 1, N = n1 * n1 * n1;
    const int dirs[7][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 1}};
    std::vector<int32_t> rowptr(N + 1), col;
    std::vector<double> val;
    col.reserve(15 * N); val.reserve(15 * N);
    for (int64_t k = 0; k < n1; ++k) for (int64_t j = 0; j < n1; ++j) for (int64_t i = 0; i < n1; ++i) {
        const int64_t r = (k * n1 + j) * n1 + i;
        rowptr[r] = (int32_t)col.size();
        int64_t cand[15]; int nc = 0;
        for (int s = -1; s <= 1; s += 2) for (int d = 6; d >= 0; --d) { (void)d; }
        // ascending column order: negative directions (largest offset first), self, positive directions
        int64_t offs[15]; int no = 0;
        for (int d = 0; d < 7; ++d) offs[no++] = -(dirs[d][0] + dirs[d][1] * n1 + dirs[d][2] * n1 * n1);
        offs[no++] = 0;
        for (int d = 0; d < 7; ++d) offs[no++] = (dirs[d][0] + dirs[d][1] * n1 + dirs[d][2] * n1 * n1);
        for (int a = 0; a < 15; ++a) {
            int di = 0, dj = 0, dk = 0, sgn = a < 7 ? -1 : 1;
            if (a != 7) { const int d = a < 7 ? a : a - 8; di = sgn * dirs[d][0]; dj = sgn * dirs[d][1]; dk = sgn * dirs[d][2]; }
            const int64_t ii = i + di, jj = j + dj, kk = k + dk;
            if (ii < 0 || ii >= n1 || jj < 0 || jj >= n1 || kk < 0 || kk >= n1) continue;
            cand[nc++] = r + offs[a];
        }
        // sort ascending (small)
        for (int a = 1; a < nc; ++a) { int64_t v = cand[a]; int b = a - 1; while (b >= 0 && cand[b] > v) { cand[b + 1] = cand[b]; --b; } cand[b + 1] = v; }
        for (int a = 0; a < nc; ++a) { col.push_back((int32_t)cand[a]); val.push_back(cand[a] == r ? 14.0 : -1.0); }
    }
    rowptr[N] = (int32_t)col.size();
    const int64_t nnz = col.size();
    printf("n=%d N=%lld nnz=%lld\n", n, (long long)N, (long long)nnz);
    // SELL-32 (variable slice width) and fixed width 15
    const int64_t nslice = (N + 31) / 32;
    std::vector<int64_t> sliceptr(nslice + 1, 0);
    for (int64_t s = 0; s < nslice; ++s) {
        int mx = 0;
        for (int l = 0; l < 32; ++l) { int64_t r = s * 32 + l; if (r < N) mx = std::max(mx, rowptr[r + 1] - rowptr[r]); }
        sliceptr[s + 1] = sliceptr[s] + 32 * (int64_t)mx;
    }
    std::vector<int32_t> scol(sliceptr[nslice], 0), fcol(nslice * 32 * 15, 0);
    std::vector<double> sval(sliceptr[nslice], 0.0), fval(nslice * 32 * 15, 0.0);
    for (int64_t s = 0; s < nslice; ++s) for (int l = 0; l < 32; ++l) {
        int64_t r = s * 32 + l; if (r >= N) continue;
        for (int k = rowptr[r], j = 0; k < rowptr[r + 1]; ++k, ++j) {
            scol[sliceptr[s] + 32 * j + l] = col[k]; sval[sliceptr[s] + 32 * j + l] = val[k];
            fcol[s * 32 * 15 + 32 * j + l] = col[k]; fval[s * 32 * 15 + 32 * j + l] = val[k];
        }
        // pad columns with the row itself (value 0) to keep gathers local
        for (int j = rowptr[r + 1] - rowptr[r]; j < 15; ++j) fcol[s * 32 * 15 + 32 * j + l] = (int32_t)r;
        for (int64_t j = rowptr[r + 1] - rowptr[r]; j < (sliceptr[s + 1] - sliceptr[s]) / 32; ++j) scol[sliceptr[s] + 32 * j + l] = (int32_t)r;
    }
    printf("SELL entries %lld (padding %.2f%%)\n", (long long)sliceptr[nslice], 100.0 * (sliceptr[nslice] - nnz) / nnz);
    int32_t *d_rowptr, *d_col, *d_scol, *d_fcol; double *d_val, *d_sval, *d_fval, *d_x, *d_y, *d_y2; int64_t *d_sliceptr;
    CK(cudaMalloc(&d_rowptr, 4 * (N + 1))); CK(cudaMalloc(&d_col, 4 * nnz)); CK(cudaMalloc(&d_val, 8 * nnz));
    CK(cudaMalloc(&d_scol, 4 * scol.size())); CK(cudaMalloc(&d_sval, 8 * sval.size())); CK(cudaMalloc(&d_sliceptr, 8 * (nslice + 1)));
    CK(cudaMalloc(&d_fcol, 4 * fcol.size())); CK(cudaMalloc(&d_fval, 8 * fval.size()));
    CK(cudaMalloc(&d_x, 8 * N)); CK(cudaMalloc(&d_y, 8 * N)); CK(cudaMalloc(&d_y2, 8 * N));
    CK(cudaMemcpy(d_rowptr, rowptr.data(), 4 * (N + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_col, col.data(), 4 * nnz, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_val, val.data(), 8 * nnz, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_scol, scol.data(), 4 * scol.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_sval, sval.data(), 8 * sval.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_fcol, fcol.data(), 4 * fcol.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_fval, fval.data(), 8 * fval.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sliceptr, sliceptr.data(), 8 * (nslice + 1), cudaMemcpyHostToDevice));
    std::vector<double> hx(N);
    for (int64_t i = 0; i < N; ++i) hx[i] = 1.0 + 1e-3 * (double)(i % 1000);
    CK(cudaMemcpy(d_x, hx.data(), 8 * N, cudaMemcpyHostToDevice));
    const double bytes = 12.0 * nnz + 4.0 * (N + 1) + 16.0 * N;
    auto rep = [&](const char *name, float ms) { printf("%-44s %7.3f ms  %7.1f GB/s\n", name, ms, bytes / ms / 1e6); fflush(stdout); };
    auto check = [&](const char *name) {
        std::vector<double> a(N), b(N);
        CK(cudaMemcpy(a.data(), d_y, 8 * N, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), d_y2, 8 * N, cudaMemcpyDeviceToHost));
        double mx = 0; for (int64_t i = 0; i < N; ++i) mx = std::max(mx, fabs(a[i] - b[i]));
        if (mx > 1e-9) printf("   !! %s differs from reference by %g\n", name, mx);
    };
    const int SM = 148;
    rep("V0 stream matrix arrays only (no gather)", timeit([&] { k_stream<<<SM * 8, T>>>(nnz, d_col, d_val, d_y2); }));
    rep("V4 sub-warp 8 lanes/row", timeit([&] { k_vec8<<<SM * 8, T>>>(N, d_rowptr, d_col, d_val, d_x, d_y); }));
    {
        int rpb = 128; size_t sm = 8 * rpb * 15;
        rep("V1 CTA-stream rpb=128", timeit([&] { k_cta_stream<<<SM * 8, T, sm>>>(N, rpb, d_rowptr, d_col, d_val, d_x, d_y2); }));
        check("V1");
        for (int mult : {4, 6, 8}) {
            char nm[64]; snprintf(nm, 64, "V1 CTA-stream rpb=128 grid=SMx%d", mult);
            rep(nm, timeit([&] { k_cta_stream<<<SM * mult, T, sm>>>(N, rpb, d_rowptr, d_col, d_val, d_x, d_y2); }));
        }
        rpb = 256; sm = 8 * rpb * 15;
        rep("V1 CTA-stream rpb=256", timeit([&] { k_cta_stream<<<SM * 6, T, sm>>>(N, rpb, d_rowptr, d_col, d_val, d_x, d_y2); }));
        rpb = 256; int cap = rpb * 15; sm = 12 * (size_t)cap;
        rep("V2 CTA-stage (row-aligned gather) rpb=256", timeit([&] { k_cta_stage<<<SM * 4, T, sm>>>(N, rpb, d_rowptr, d_col, d_val, d_x, d_y2, cap); }));
        check("V2");
        rpb = 128; cap = rpb * 15; sm = 12 * (size_t)cap;
        rep("V2 CTA-stage (row-aligned gather) rpb=128", timeit([&] { k_cta_stage<<<SM * 8, T, sm>>>(N, rpb, d_rowptr, d_col, d_val, d_x, d_y2, cap); }));
    }
    for (int mult : {4, 8, 16}) {
        char nm[64]; snprintf(nm, 64, "V3 SELL-32 variable width grid=SMx%d", mult);
        rep(nm, timeit([&] { k_sell<<<SM * mult, T>>>(N, d_sliceptr, d_scol, d_sval, d_x, d_y2); }));
    }
    check("V3");
    for (int mult : {3, 4, 6, 8}) {
        char nm[64]; snprintf(nm, 64, "V3b SELL-32 fixed width 15 grid=SMx%d", mult);
        rep(nm, timeit([&] { k_sell_fixed<15><<<SM * mult, T>>>(N, d_fcol, d_fval, d_x, d_y2); }));
    }
    check("V3b");
    return 0;
}
