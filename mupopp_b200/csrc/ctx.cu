// Context, error reporting, workspace, NCCL glue (dlopen'ed) and halo exchange.
//
// Multi-GPU replaces what DOLFIN/PETSc do over MPI in the reference (aprun -n 24|48,
// /root/reference/src/run/sphere_archer/submit.sh:25): VecScatter ghost updates become grouped
// ncclSend/ncclRecv of interface dofs, MPI_Allreduce of inner products becomes one small ncclAllReduce.
#include <dlfcn.h>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mp {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int ws_get(mp_ctx *ctx, int slot, size_t bytes, void **out) {
    Workspace &w = ctx->work[slot];
    if (w.bytes < bytes) {
        if (w.ptr) cudaFree(w.ptr);
        w.ptr = nullptr; w.bytes = 0;
        size_t want = bytes + bytes / 8 + 256;
        MP_CUDA(cudaMalloc(&w.ptr, want));
        w.bytes = want;
    }
    *out = w.ptr;
    return 0;
}

// ---- NCCL, resolved at run time so that the library loads (and the CPU-side symbol test runs) without it
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int (*pfn_GetUniqueId)(ncclUniqueId_t *);
typedef int (*pfn_CommInitRank)(void **, int, ncclUniqueId_t, int);
typedef int (*pfn_CommDestroy)(void *);
typedef int (*pfn_AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*pfn_Send)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*pfn_Recv)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*pfn_Group)(void);
typedef const char *(*pfn_ErrStr)(int);

struct Nccl {
    void *h = nullptr;
    pfn_GetUniqueId GetUniqueId = nullptr;
    pfn_CommInitRank CommInitRank = nullptr;
    pfn_CommDestroy CommDestroy = nullptr;
    pfn_AllReduce AllReduce = nullptr;
    pfn_Send Send = nullptr;
    pfn_Recv Recv = nullptr;
    pfn_Group GroupStart = nullptr, GroupEnd = nullptr;
    pfn_ErrStr ErrStr = nullptr;
};
static Nccl g_nccl;

static int nccl_load() {
    if (g_nccl.h) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !g_nccl.h; ++i) g_nccl.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    MP_CHECK(g_nccl.h, "NCCL: cannot dlopen libnccl.so.2 (%s)", dlerror());
#define LD(sym)                                                                 \
    g_nccl.sym = (decltype(g_nccl.sym))dlsym(g_nccl.h, "nccl" #sym);            \
    MP_CHECK(g_nccl.sym, "NCCL: missing symbol nccl" #sym)
    LD(GetUniqueId); LD(CommInitRank); LD(CommDestroy); LD(AllReduce); LD(Send); LD(Recv); LD(GroupStart); LD(GroupEnd);
#undef LD
    g_nccl.ErrStr = (pfn_ErrStr)dlsym(g_nccl.h, "ncclGetErrorString");
    return 0;
}

#define MP_NCCL(call)                                                                                         \
    do {                                                                                                      \
        int _r = (call);                                                                                      \
        if (_r != 0) {                                                                                        \
            mp::set_error("%s:%d: %s -> NCCL error %d (%s)", __FILE__, __LINE__, #call, _r,                   \
                          mp::g_nccl.ErrStr ? mp::g_nccl.ErrStr(_r) : "?");                                   \
            return 3;                                                                                         \
        }                                                                                                     \
    } while (0)

constexpr int kNcclFloat64 = 8, kNcclSum = 0;

}  // namespace mp

namespace {
__global__ void k_pack(const double *__restrict__ x, const int32_t *__restrict__ idx, double *__restrict__ buf, int64_t n) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) buf[t] = x[idx[t]];
}
}  // namespace

extern "C" {

int mp_version(void) { return 100; }

const char *mp_last_error(void) { return mp::g_err; }

int mp_ctx_create(int device, void *stream, mp_ctx **out) {
    MP_CHECK(out, "mp_ctx_create: out is null");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    MP_CHECK(e == cudaSuccess && ndev > 0, "mp_ctx_create: no CUDA device (%s); this library has no CPU fallback",
             cudaGetErrorString(e));
    MP_CHECK(device >= 0 && device < ndev, "mp_ctx_create: device %d out of range (%d devices)", device, ndev);
    MP_CUDA(cudaSetDevice(device));
    mp_ctx *c = new mp_ctx();
    struct Guard {   // frees the partially built context on any early return below
        mp_ctx *c; bool armed = true;
        ~Guard() { if (armed) mp_ctx_destroy(c); }
    } guard{c};
    c->device = device;
    // NULL = the CUDA default stream (what torch's default stream is): kernels are then ordered with the
    // caller's torch copies on that stream without any extra synchronisation.
    c->stream = (cudaStream_t)stream;
    c->own_stream = false;
    cudaDeviceProp prop;
    MP_CUDA(cudaGetDeviceProperties(&prop, device));
    c->num_sm = prop.multiProcessorCount;
    MP_CUDA(cudaMalloc(&c->d_scalars, sizeof(double) * 64));
    MP_CUDA(cudaMemset(c->d_scalars, 0, sizeof(double) * 64));
    MP_CUDA(cudaMallocHost(&c->h_scalars, sizeof(double) * 64));
    MP_CUDA(cudaMalloc(&c->d_ctl, sizeof(mp::ReduceCtl)));
    MP_CUDA(cudaMemset(c->d_ctl, 0, sizeof(mp::ReduceCtl)));
    MP_CUDA(cudaMalloc(&c->d_partials, sizeof(double) * 4096 * 8));
    MP_CUDA(cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming));
    guard.armed = false;
    *out = c;
    return 0;
}

int mp_ctx_destroy(mp_ctx *c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->nccl_comm && mp::g_nccl.CommDestroy) mp::g_nccl.CommDestroy(c->nccl_comm);
    for (auto &w : c->work) if (w.ptr) cudaFree(w.ptr);
    for (void *m : c->mbox_mapped) if (m) cudaIpcCloseMemHandle(m);
    cudaFree(c->mbox); cudaFree(c->d_comm);
    cudaFree(c->d_scalars); cudaFreeHost(c->h_scalars); cudaFree(c->d_ctl); cudaFree(c->d_partials);
    if (c->ev) cudaEventDestroy(c->ev);
    for (auto e : c->tev) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int mp_ctx_sync(mp_ctx *c) {
    MP_CHECK(c, "mp_ctx_sync: null ctx");
    MP_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int64_t mp_ctx_launch_count(const mp_ctx *c) { return c ? c->launches : -1; }

int mp_ctx_set_option(mp_ctx *c, const char *name, int value) {
    MP_CHECK(c && name, "mp_ctx_set_option: null argument");
    if (!strcmp(name, "spmv_variant")) { c->spmv_variant = value; return 0; }
    if (!strcmp(name, "spmv_c8")) { c->spmv_c8 = value; return 0; }
    if (!strcmp(name, "minres_graph")) { c->minres_graph = value ? 1 : 0; return 0; }
    if (!strcmp(name, "line_prefetch")) { MP_CHECK(value >= 0 && value <= 256, "line_prefetch must be 0..256 planes"); c->line_prefetch = value; return 0; }
    if (!strcmp(name, "pcg_variant")) { MP_CHECK(value >= -1 && value <= 1, "pcg_variant must be -1, 0 or 1"); c->pcg_variant = value; return 0; }
    if (!strcmp(name, "spmv_lpr")) { MP_CHECK(value == 0 || value == 1 || value == 2 || value == 4, "spmv_lpr must be 0, 1, 2 or 4"); c->spmv_lpr = value; return 0; }
    if (!strcmp(name, "peer")) {   // switch the peer-memory collectives on/off (A/B runs against the NCCL path); needs mp_comm_peer_attach
        MP_CHECK(!value || c->peer_attached, "mp_ctx_set_option: peer windows are not attached");
        return mp::set_peer_mode(c, value != 0);
    }
    mp::set_error("mp_ctx_set_option: unknown option '%s'", name);
    return 2;
}

int mp_ctx_timing(mp_ctx *c, int enable, int max_samples) {
    MP_CHECK(c, "mp_ctx_timing: null ctx");
    MP_CUDA(cudaSetDevice(c->device));
    if (enable) {
        size_t want = 2 * (size_t)(max_samples > 0 ? max_samples : 4096);
        while (c->tev.size() < want) {
            cudaEvent_t e;
            MP_CUDA(cudaEventCreate(&e));
            c->tev.push_back(e);
        }
        c->tev_used = 0;
    }
    c->timing = enable != 0;
    return 0;
}

int mp_ctx_timing_read(mp_ctx *c, int *nsamples, double *total_ms) {
    MP_CHECK(c && nsamples && total_ms, "mp_ctx_timing_read: null argument");
    MP_CUDA(cudaStreamSynchronize(c->stream));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < c->tev_used; i += 2) {
        float ms = 0.f;
        MP_CUDA(cudaEventElapsedTime(&ms, c->tev[i], c->tev[i + 1]));
        tot += ms;
    }
    *nsamples = (int)(c->tev_used / 2);
    *total_ms = tot;
    c->tev_used = 0;
    return 0;
}

int mp_nccl_unique_id(uint8_t *out128) {
    MP_CHECK(out128, "mp_nccl_unique_id: null");
    MP_TRY(mp::nccl_load());
    mp::ncclUniqueId_t id;
    MP_NCCL(mp::g_nccl.GetUniqueId(&id));
    memcpy(out128, id.internal, 128);
    return 0;
}

}  // extern "C"

namespace mp {
int allreduce_max_host(mp_ctx *c, double *h_value) {
    if (c->nranks <= 1 || !c->nccl_comm) return 0;
    double *d = c->d_scalars + 60;   // scratch at the end of the 64-double scalar block
    MP_CUDA(cudaMemcpyAsync(d, h_value, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    MP_NCCL(g_nccl.AllReduce(d, d, 1, kNcclFloat64, 2 /* ncclMax */, c->nccl_comm, c->stream));
    MP_CUDA(cudaMemcpyAsync(h_value, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    MP_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// peer mode on: grid_reduce exchanges through the mailbox windows and writes the GLOBAL scalar slots directly (no local copy);
// off: rank-local sums land at +kLocalOff and an out-of-place ncclAllReduce follows (reduce_slots in krylov.cu).
int set_peer_mode(mp_ctx *c, bool on) {
    if (on && !c->peer_attached) return 0;
    MP_CUDA(cudaStreamSynchronize(c->stream));
    mp::CommDev *comm = on ? c->d_comm : nullptr;
    const unsigned int off = (!on && c->nranks > 1) ? (unsigned int)kLocalOff : 0u;
    MP_CUDA(cudaMemcpy(&c->d_ctl->comm, &comm, sizeof(comm), cudaMemcpyHostToDevice));
    MP_CUDA(cudaMemcpy(&c->d_ctl->local_off, &off, sizeof(off), cudaMemcpyHostToDevice));
    c->peer = on;
    return 0;
}

int allreduce_oop(mp_ctx *c, const double *d_send, double *d_recv, int count) {
    if (c->nranks <= 1 || !c->nccl_comm || count == 0) return 0;
    MP_NCCL(g_nccl.AllReduce(d_send, d_recv, (size_t)count, kNcclFloat64, kNcclSum, c->nccl_comm, c->stream));
    return 0;
}
}  // namespace mp

extern "C" {

int mp_comm_init(mp_ctx *c, int nranks, int rank, const uint8_t *id128) {
    MP_CHECK(c && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "mp_comm_init: bad argument");
    MP_TRY(mp::nccl_load());
    MP_CUDA(cudaSetDevice(c->device));
    mp::ncclUniqueId_t id;
    memcpy(id.internal, id128, 128);
    MP_NCCL(mp::g_nccl.CommInitRank(&c->nccl_comm, nranks, id, rank));
    c->nranks = nranks;
    c->rank = rank;
    // reductions now land in the rank-local scalar block (see grid_reduce in krylov.cu)
    const unsigned int off = nranks > 1 ? (unsigned int)mp::kLocalOff : 0u;
    MP_CUDA(cudaMemcpy(&c->d_ctl->local_off, &off, sizeof(off), cudaMemcpyHostToDevice));
    return 0;
}

int mp_comm_destroy(mp_ctx *c) {
    if (c && c->nccl_comm) {
        cudaStreamSynchronize(c->stream);
        mp::g_nccl.CommDestroy(c->nccl_comm);
        c->nccl_comm = nullptr;
        c->nranks = 1; c->rank = 0;
        cudaMemset(&c->d_ctl->local_off, 0, sizeof(unsigned int));
        mp::set_peer_mode(c, false);
    }
    return 0;
}

int mp_allreduce_sum(mp_ctx *c, double *d_buf, int count) {
    MP_CHECK(c && d_buf && count >= 0, "mp_allreduce_sum: bad argument");
    if (c->nranks <= 1 || !c->nccl_comm || count == 0) return 0;
    MP_NCCL(mp::g_nccl.AllReduce(d_buf, d_buf, (size_t)count, mp::kNcclFloat64, mp::kNcclSum, c->nccl_comm, c->stream));
    return 0;
}

int mp_halo_create(mp_ctx *c, int nneigh, const int32_t *neigh_rank, const int32_t *send_off, const int32_t *d_send_idx,
                   const int32_t *recv_start, const int32_t *recv_count, mp_halo **out) {
    MP_CHECK(c && out && nneigh >= 0, "mp_halo_create: bad argument");
    MP_CHECK(nneigh <= mp::kMaxNeigh, "mp_halo_create: more than %d neighbours", mp::kMaxNeigh);
    mp_halo *h = new mp_halo();
    struct Guard {
        mp_halo *h; bool armed = true;
        ~Guard() { if (armed) mp_halo_destroy(h); }
    } guard{h};
    h->nneigh = nneigh;
    h->rank.assign(neigh_rank, neigh_rank + nneigh);
    h->send_off.assign(send_off, send_off + nneigh + 1);
    h->recv_start.assign(recv_start, recv_start + nneigh);
    h->recv_count.assign(recv_count, recv_count + nneigh);
    h->nsend = nneigh ? send_off[nneigh] : 0;
    if (h->nsend > 0) {
        MP_CUDA(cudaMalloc(&h->d_send_idx, sizeof(int32_t) * h->nsend));
        MP_CUDA(cudaMemcpyAsync(h->d_send_idx, d_send_idx, sizeof(int32_t) * h->nsend, cudaMemcpyDeviceToDevice, c->stream));
        MP_CUDA(cudaMalloc(&h->d_sendbuf, sizeof(double) * h->nsend));
    }
    guard.armed = false;
    *out = h;
    return 0;
}

int mp_halo_exchange(mp_ctx *c, mp_halo *h, double *d_x) {
    MP_CHECK(c && d_x, "mp_halo_exchange: bad argument");
    if (!h || h->nneigh == 0 || c->nranks <= 1) return 0;
    MP_CHECK(c->nccl_comm, "mp_halo_exchange: no communicator attached (call mp_comm_init)");
    if (h->nsend > 0) {
        k_pack<<<(unsigned)((h->nsend + 255) / 256), 256, 0, c->stream>>>(d_x, h->d_send_idx, h->d_sendbuf, h->nsend);
        mp::count_launch(c);
    }
    MP_NCCL(mp::g_nccl.GroupStart());
    int first_err = 0;   // the group is always closed, whatever a Send/Recv returns
    for (int k = 0; k < h->nneigh && !first_err; ++k) {
        int32_t ns = h->send_off[k + 1] - h->send_off[k];
        if (ns > 0) first_err = mp::g_nccl.Send(h->d_sendbuf + h->send_off[k], (size_t)ns, mp::kNcclFloat64, h->rank[k], c->nccl_comm, c->stream);
        if (!first_err && h->recv_count[k] > 0)
            first_err = mp::g_nccl.Recv(d_x + h->recv_start[k], (size_t)h->recv_count[k], mp::kNcclFloat64, h->rank[k], c->nccl_comm, c->stream);
    }
    const int end_err = mp::g_nccl.GroupEnd();
    if (first_err || end_err) {
        const int e = first_err ? first_err : end_err;
        mp::set_error("mp_halo_exchange: NCCL error %d (%s)", e, mp::g_nccl.ErrStr ? mp::g_nccl.ErrStr(e) : "?");
        return 3;
    }
    return 0;
}

int mp_halo_destroy(mp_halo *h) {
    if (!h) return 0;
    for (void *m : h->mapped) if (m) cudaIpcCloseMemHandle(m);
    cudaFree(h->window); cudaFree(h->d_dev); cudaFree(h->d_slice_kind); cudaFree(h->d_bslices);
    cudaFree(h->d_send_idx); cudaFree(h->d_sendbuf);
    delete h;
    return 0;
}

/* ---- peer-memory windows (CUDA IPC over NVLink): in-kernel allreduce of the Krylov scalars and halo push fused into SpMV */
int mp_comm_peer_export(mp_ctx *c, uint8_t *handle64) {
    MP_CHECK(c && handle64, "mp_comm_peer_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    MP_CUDA(cudaSetDevice(c->device));
    if (!c->mbox) {
        MP_CUDA(cudaMalloc(&c->mbox, sizeof(mp::MboxWindow)));
        MP_CUDA(cudaMemset(c->mbox, 0, sizeof(mp::MboxWindow)));
        MP_CUDA(cudaDeviceSynchronize());     // the zeroed window must exist before any peer can write into it
    }
    cudaIpcMemHandle_t hd;
    MP_CUDA(cudaIpcGetMemHandle(&hd, c->mbox));
    memcpy(handle64, &hd, 64);
    return 0;
}

int mp_comm_peer_attach(mp_ctx *c, const uint8_t *handles) {
    MP_CHECK(c && handles && c->mbox, "mp_comm_peer_attach: call mp_comm_peer_export first");
    MP_CHECK(c->nranks >= 1 && c->nranks <= mp::kMaxRanks, "mp_comm_peer_attach: 1..%d ranks supported", mp::kMaxRanks);
    MP_CHECK(!c->peer_attached, "mp_comm_peer_attach: already attached");
    MP_CUDA(cudaSetDevice(c->device));
    mp::CommDev hd;
    memset(&hd, 0, sizeof(hd));
    hd.nranks = c->nranks; hd.rank = c->rank; hd.seq = 1; hd.error = 0;
    for (int r = 0; r < c->nranks; ++r) {
        if (r == c->rank) { hd.win[r] = (mp::MboxWindow *)c->mbox; continue; }
        cudaIpcMemHandle_t ih;
        memcpy(&ih, handles + 64 * (size_t)r, 64);
        void *p = nullptr;
        MP_CUDA(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
        c->mbox_mapped.push_back(p);
        hd.win[r] = (mp::MboxWindow *)p;
    }
    MP_CUDA(cudaMalloc(&c->d_comm, sizeof(mp::CommDev)));
    MP_CUDA(cudaMemcpy(c->d_comm, &hd, sizeof(hd), cudaMemcpyHostToDevice));
    c->peer_attached = true;
    return mp::set_peer_mode(c, true);
}

int mp_comm_peer_enabled(const mp_ctx *c) { return c && c->peer ? 1 : 0; }

int mp_halo_peer_export(mp_ctx *c, mp_halo *h, int64_t nown, int64_t nghost, uint8_t *handle64) {
    MP_CHECK(c && h && handle64 && nown >= 0 && nghost >= 0, "mp_halo_peer_export: bad argument");
    MP_CUDA(cudaSetDevice(c->device));
    if (!h->window) {
        const size_t bytes = mp::kHaloHdr + sizeof(double) * 2 * (size_t)(nghost > 0 ? nghost : 1);
        MP_CUDA(cudaMalloc(&h->window, bytes));
        MP_CUDA(cudaMemset(h->window, 0, bytes));
        MP_CUDA(cudaDeviceSynchronize());
        h->nown = nown; h->nghost = nghost;
    }
    cudaIpcMemHandle_t hd;
    MP_CUDA(cudaIpcGetMemHandle(&hd, h->window));
    memcpy(handle64, &hd, 64);
    return 0;
}

int mp_halo_peer_attach(mp_ctx *c, mp_halo *h, const uint8_t *neigh_handles, const int64_t *peer_goff, const int64_t *peer_gcap,
                        const int32_t *peer_slot) {
    MP_CHECK(c && h && h->window, "mp_halo_peer_attach: call mp_halo_peer_export first");
    MP_CHECK(h->nneigh == 0 || (neigh_handles && peer_goff && peer_gcap && peer_slot), "mp_halo_peer_attach: null argument");
    MP_CHECK(!h->peer, "mp_halo_peer_attach: already attached");
    MP_CUDA(cudaSetDevice(c->device));
    mp::HaloDev &d = h->h_dev;
    memset(&d, 0, sizeof(d));
    d.nneigh = h->nneigh; d.nsend = (int)h->nsend; d.nown = h->nown; d.nghost = h->nghost;
    d.send_idx = h->d_send_idx;
    for (int k = 0; k <= h->nneigh; ++k) d.send_off[k] = h->send_off[k];
    d.gwin = (double *)((char *)h->window + mp::kHaloHdr);
    d.flag = (unsigned long long *)h->window;
    d.seq = 1;
    for (int k = 0; k < h->nneigh; ++k) {
        MP_CHECK(peer_slot[k] >= 0 && peer_slot[k] < mp::kMaxNeigh, "mp_halo_peer_attach: bad peer slot");
        MP_CHECK(peer_goff[k] >= 0 && peer_goff[k] + (h->send_off[k + 1] - h->send_off[k]) <= peer_gcap[k],
                 "mp_halo_peer_attach: my block does not fit neighbour %d's ghost window", h->rank[k]);
        cudaIpcMemHandle_t ih;
        memcpy(&ih, neigh_handles + 64 * (size_t)k, 64);
        void *p = nullptr;
        MP_CUDA(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
        h->mapped.push_back(p);
        d.peer_gwin[k] = (double *)((char *)p + mp::kHaloHdr);
        d.peer_goff[k] = peer_goff[k];
        d.peer_gcap[k] = peer_gcap[k];
        d.peer_flag[k] = (unsigned long long *)p + 2 * (size_t)peer_slot[k];
    }
    MP_CUDA(cudaMalloc(&h->d_dev, sizeof(mp::HaloDev)));
    MP_CUDA(cudaMemcpy(h->d_dev, &d, sizeof(d), cudaMemcpyHostToDevice));
    h->peer = true;
    return 0;
}

}  // extern "C"
