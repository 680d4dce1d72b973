// Shared internals of libmupopp_b200 (sm_100a).  Not part of the public C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/mupopp_b200.h"

namespace mp {

void set_error(const char *fmt, ...);

#define MP_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            mp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

#define MP_CHECK(cond, ...)            \
    do {                               \
        if (!(cond)) {                 \
            mp::set_error(__VA_ARGS__); \
            return 2;                  \
        }                              \
    } while (0)

#define MP_TRY(expr)           \
    do {                       \
        int _r = (expr);       \
        if (_r) return _r;     \
    } while (0)

// B200: 148 SMs.  Grids of persistent/grid-stride kernels are sized in multiples of this.
constexpr int kNumSM = 148;

struct Workspace {
    void *ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace mp

struct mp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t launches = 0;
    int num_sm = mp::kNumSM;
    // reusable device scratch: Krylov work vectors, reduction partials
    mp::Workspace work[8];
    double *d_scalars = nullptr;   // device scalar block (dot results, alpha/beta ...)
    double *h_scalars = nullptr;   // pinned host mirror
    unsigned int *d_counter = nullptr;  // last-block-done counters for deterministic reductions
    double *d_partials = nullptr;       // [kMaxPartialBlocks * kMaxDots]
    cudaEvent_t ev = nullptr;
    // optional per-launch timing of the dominant kernel (SpMV) with CUDA events on ctx->stream
    int spmv_variant = 0;   // 0 SELL-32 if mirrored else CTA-stream (default), 1 CTA-stream, 2 sub-warp-per-row, 3 warp-stream
    bool timing = false;
    std::vector<cudaEvent_t> tev;   // pairs (start, stop)
    size_t tev_used = 0;
    // NCCL (dlopen'ed); null when single-GPU
    void *nccl_comm = nullptr;
    int nranks = 1, rank = 0;
};

namespace mp {
constexpr int kLocalOff = 16;   // rank-local copy of the scalar block lives at d_scalars + kLocalOff (multi-GPU only)
int allreduce_oop(mp_ctx *ctx, const double *d_send, double *d_recv, int count);
int allreduce_max_host(mp_ctx *ctx, double *h_value);   // max over ranks of one host scalar (no-op on one GPU)
int ws_get(mp_ctx *ctx, int slot, size_t bytes, void **out);
inline void count_launch(mp_ctx *ctx, int n = 1) { ctx->launches += n; }
}  // namespace mp
