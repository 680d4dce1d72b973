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

// ------------------------------------------------------------------ peer-memory collectives (NVLink / NVSwitch, CUDA IPC)
// One process per GPU.  Every rank cudaMalloc's small "windows", exports them with cudaIpcGetMemHandle and maps its peers'
// windows; kernels then exchange Krylov scalars (in-kernel allreduce, grid_reduce in krylov.cu) and ghost dofs (halo push fused
// into the SpMV kernel) with plain stores over NVLink plus sequence-number flags -- no NCCL launch on the iteration's critical path.
namespace mp {
constexpr int kMaxRanks = 16;     // ranks of one NVSwitch domain we size the mailbox for
constexpr int kMaxNeigh = 8;      // halo neighbours per rank
constexpr int kMboxVals = 8;      // doubles per exchange (k_multidot8 is the widest reduction)

struct MboxWindow {                                   // zero-initialised; flags only ever grow
    double val[2][kMaxRanks][kMboxVals];              // [parity of the sequence number][source rank][value]
    unsigned long long flag[2][kMaxRanks];            // sequence number of the exchange that filled val[parity][source]
};

struct CommDev {                                      // device memory, one per context
    int nranks, rank;
    MboxWindow *win[kMaxRanks];                       // win[rank] = my window, the others are IPC mappings of the peers' windows
    unsigned long long seq;                           // next exchange number (starts at 1); touched by ONE thread per exchange
    int error;                                        // set when a peer did not answer within the spin time-out
};

struct ReduceCtl {                                    // device memory, one per context (was: a bare counter array)
    unsigned int count;                               // last-block-done counter of grid_reduce
    unsigned int local_off;                           // NCCL mode only: offset of the rank-local copy of the scalar block
    unsigned int push_count;                          // halo push completion (peer SpMV)
    unsigned int done_count;                          // kernel completion (peer SpMV: bumps the halo sequence number)
    CommDev *comm;                                    // non-null => grid_reduce allreduces through peer memory
};

struct HaloWindow {                                   // header of a rank's ghost window; gwin[2][nghost] doubles follow at +kHaloHdr bytes
    unsigned long long flag[kMaxNeigh][2];            // [my neighbour slot][parity] = sequence number of the push that filled the block
};
constexpr size_t kHaloHdr = 256;

struct HaloDev {                                      // device memory, one per halo plan (peer mode)
    int nneigh, nsend;
    int64_t nown, nghost;                             // local columns >= nown are ghosts, stored at gwin[parity][col - nown]
    const int32_t *send_idx;                          // [nsend] owned entries to push, grouped by neighbour
    int32_t send_off[kMaxNeigh + 1];
    double *peer_gwin[kMaxNeigh];                     // neighbour k's ghost window (mapped)
    int64_t peer_goff[kMaxNeigh], peer_gcap[kMaxNeigh];   // where my block starts in it / its per-parity capacity
    unsigned long long *peer_flag[kMaxNeigh];         // &neighbour's HaloWindow.flag[my slot there][0]
    double *gwin;                                     // my ghost window [2][nghost]
    unsigned long long *flag;                         // &my HaloWindow.flag[0][0]
    unsigned long long seq;                           // next push number (starts at 1)
    // SELL slices of the operator this halo was classified for: boundary slices reference ghost columns
    const uint8_t *slice_kind;                        // [nslice] 1 = boundary
    const int32_t *bslices;                           // [nbslice] ascending
    int32_t nbslice;
};
}  // namespace mp

struct mp_halo {
    int nneigh = 0;
    std::vector<int32_t> rank, send_off, recv_start, recv_count;
    int32_t *d_send_idx = nullptr;  // owned copy
    double *d_sendbuf = nullptr;
    int64_t nsend = 0;
    // peer mode (mp_halo_peer_export / mp_halo_peer_attach)
    bool peer = false;
    int64_t nown = 0, nghost = 0;
    void *window = nullptr;                 // cudaMalloc'ed HaloWindow + gwin
    std::vector<void *> mapped;             // cudaIpcOpenMemHandle results (closed in mp_halo_destroy)
    mp::HaloDev *d_dev = nullptr;           // device copy of the descriptor
    mp::HaloDev h_dev;                      // host mirror
    const void *classified_for = nullptr;   // SELL column array the slice lists were computed for
    uint8_t *d_slice_kind = nullptr;
    int32_t *d_bslices = nullptr;
};

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
    mp::ReduceCtl *d_ctl = nullptr;     // last-block-done counters of the deterministic reductions + peer-allreduce descriptor
    double *d_partials = nullptr;       // [kMaxPartialBlocks * kMaxDots]
    cudaEvent_t ev = nullptr;
    // optional per-launch timing of the dominant kernel (SpMV) with CUDA events on ctx->stream
    int spmv_variant = 0;   // 0 SELL-32 if mirrored else CTA-stream (default), 1 CTA-stream, 2 sub-warp-per-row, 3 warp-stream
    int spmv_c8 = 1;        // use the compressed SELL columns (9 instead of 12 bytes per nonzero) when the operator has them
    int spmv_lpr = 0;       // lanes per row of the SELL kernels: 0 = from the operator size, else 1 / 2 / 4
    int line_prefetch = 0;  // line preconditioner sweeps: L2 prefetch distance in planes (0 = off; measured: no gain once the sweeps run as one wave)
    int minres_graph = 1;   // mp_solve_minres: replay six iterations at a time as a CUDA graph (the block recurrences are launch bound)
    int pcg_variant = -1;   // -1 auto (single-reduction CG on several GPUs, classic on one), 0 classic two-reduction PCG, 1 single-reduction
    bool timing = false;
    std::vector<cudaEvent_t> tev;   // pairs (start, stop)
    size_t tev_used = 0;
    // NCCL (dlopen'ed); null when single-GPU
    void *nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    // peer-memory allreduce (mp_comm_peer_export / mp_comm_peer_attach); `peer` can be switched off for A/B runs (option "peer")
    bool peer = false, peer_attached = false;
    void *mbox = nullptr;                   // my MboxWindow (cudaMalloc)
    std::vector<void *> mbox_mapped;        // peers' windows
    mp::CommDev *d_comm = nullptr;
};

namespace mp {
constexpr int kLocalOff = 16;   // rank-local copy of the scalar block lives at d_scalars + kLocalOff (multi-GPU only)
int allreduce_oop(mp_ctx *ctx, const double *d_send, double *d_recv, int count);
int allreduce_max_host(mp_ctx *ctx, double *h_value);
int set_peer_mode(mp_ctx *ctx, bool on);
int reduce_slots_public(mp_ctx *ctx, int slot, int count);   // NCCL allreduce of rank-local scalar slots (no-op in peer mode / on one GPU)   // max over ranks of one host scalar (no-op on one GPU)
int ws_get(mp_ctx *ctx, int slot, size_t bytes, void **out);
// boundary/interior SELL slice lists of `halo` for the operator whose SELL arrays are given (cached per column array); plan.cu
int halo_classify(mp_ctx *ctx, mp_halo *halo, int64_t nrow, const int32_t *d_sliceptr, const int32_t *d_scol);
inline void count_launch(mp_ctx *ctx, int n = 1) { ctx->launches += n; }
}  // namespace mp
