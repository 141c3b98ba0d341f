// comm.cu -- the ONE collective of the path: a gather-v of the finished bar frames to one rank, over NCCL, behind the C ABI.
//
// Symbols are independent (the reference holds one symbol per TradesData), so nothing on the data path crosses GPUs;
// each rank finishes its own frame and rank `dst` collects them (BASELINE configs[4]).  NCCL has no
// gather-v, so a step is:   ncclAllGather of the exact byte counts  ->  grouped ncclSend / ncclRecv of exactly those bytes
// (no padding to a common capacity).  The transfer of step k runs on a communication stream and overlaps the kernels of
// step k+1:
//     submit(k):  [ctx stream]  pack the step's segments into staging[k&1] (device-to-device), record `ready`
//                 [comm stream] wait `ready`; all-gather the counts; copy them to pinned host memory; record `sized`
//     submit(k+LAG) / finish(): host waits `sized(k)` (the GPU already has the next step's kernels queued, so it stays busy),
//                               sizes the receive buffers, posts the grouped send/recv of step k on the comm stream.
//                               LAG = 1 (two staging / receive slots) is the default and the configuration validated at
//                               N = 2 and N = 8; FMK_COMM_LAG=2 posts the transfer one step later still (three slots), so that
//                               no NCCL kernel waits on SMs for a slower peer -- experimental: a 2-GPU bench run with it
//                               hung (profiles/README.md), so it is off until that is understood
//     finish():   the ctx stream waits for the last transfer, so a timer stopped on it covers every gather.
// PAYLOAD PATH (default, one node): the destination exports its receive buffer with CUDA IPC and every other rank PUSHES its
// frame into it with a peer-to-peer cudaMemcpyAsync over NVLink -- copy engines, no SM.  A tiny all-gather queued behind the
// copy on every rank's comm stream is the completion fence (it finishes on the destination only after every rank's copy
// has).  Why: ncclSend/ncclRecv kernels hold up to maxCTAs SMs for the length of the transfer, and the dollar task pass is a
// single wave of long-lived blocks -- a block that finds its SM taken starts when the NCCL kernel ends, which cost 0.4 ms of
// a 10 ms step at N = 2 and 1.5 ms at N = 8 (device-timed efficiency 0.87).  The grouped ncclSend/ncclRecv path is kept as
// the fallback (FMK_COMM_P2P=0, or any rank failing to map the buffer -- decided collectively).
// libnccl is loaded with dlopen at fmk_comm_init: single-GPU use of libfmk.so needs no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>
#include <new>
#include "common.cuh"

namespace {
struct NcclApi {
    void *h;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRankConfig)(ncclComm_t *, int, ncclUniqueId, int, ncclConfig_t *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int *);
};
NcclApi g_nccl = {};

const char *nccl_load() {
    if (g_nccl.h) return nullptr;
    const char *names[] = {getenv("FMK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return "libnccl.so.2 not found (set FMK_NCCL_LIB)";
#define FMK_SYM(field, name)                                         \
    *(void **)(&g_nccl.field) = dlsym(h, name);                      \
    if (!g_nccl.field) { dlclose(h); return "libnccl lacks " name; }
    FMK_SYM(GetUniqueId, "ncclGetUniqueId")
    FMK_SYM(CommInitRank, "ncclCommInitRank")
    FMK_SYM(CommDestroy, "ncclCommDestroy")
    FMK_SYM(AllGather, "ncclAllGather")
    FMK_SYM(AllReduce, "ncclAllReduce")
    FMK_SYM(Send, "ncclSend")
    FMK_SYM(Recv, "ncclRecv")
    FMK_SYM(GroupStart, "ncclGroupStart")
    FMK_SYM(GroupEnd, "ncclGroupEnd")
    FMK_SYM(GetErrorString, "ncclGetErrorString")
    FMK_SYM(GetVersion, "ncclGetVersion")
#undef FMK_SYM
    *(void **)(&g_nccl.CommInitRankConfig) = dlsym(h, "ncclCommInitRankConfig");   // optional (NCCL >= 2.14)
    g_nccl.h = h;
    return nullptr;
}
}  // namespace

constexpr int FMK_COMM_MAXSEG = 8;
constexpr int FMK_COMM_SLOTS = 4;

struct fmk_comm {
    fmk_ctx *ctx;
    ncclComm_t comm;
    int rank, world;
    cudaStream_t stream;            // communication stream
    // per pipeline slot
    char *staging[FMK_COMM_SLOTS];               // packed frame of this rank
    int64_t staging_cap[FMK_COMM_SLOTS];
    cudaEvent_t ready[FMK_COMM_SLOTS], sized[FMK_COMM_SLOTS], done[FMK_COMM_SLOTS];
    int has_done[FMK_COMM_SLOTS];
    int64_t *counts_dev[FMK_COMM_SLOTS];         // [world] byte counts after the all-gather
    int64_t *counts_host[FMK_COMM_SLOTS];        // pinned
    int64_t *mine_host[FMK_COMM_SLOTS];          // pinned: this rank's byte count (source of the all-gather input)
    int64_t *mine_dev[FMK_COMM_SLOTS];
    char *recv[FMK_COMM_SLOTS];                  // dst only: frames of all ranks, back to back
    int64_t recv_cap[FMK_COMM_SLOTS];
    int64_t recv_off[FMK_COMM_SLOTS][65];        // dst only: offsets of each rank's frame in recv[s]
    // peer-to-peer push of the payload (CUDA IPC): the mapping of the destination's recv[s] on the other ranks
    int p2p;                                     // 1: push with copy engines; 0: grouped ncclSend / ncclRecv
    char *peer_recv[FMK_COMM_SLOTS];             // non-destination ranks: destination's recv[s] mapped into this process
    int64_t agreed_cap[FMK_COMM_SLOTS];          // capacity of the destination's recv[s], tracked identically on every rank
    int agreed_dst[FMK_COMM_SLOTS];
    unsigned char *ipc_dev, *ipc_host;           // [world + 1][64]: all-gather buffer for the IPC handle (+ this rank's input)
    int pend_slot[FMK_COMM_SLOTS], pend_dst[FMK_COMM_SLOTS];   // FIFO of steps whose counts were exchanged but whose send/recv is not posted yet
    int npending, lag, nslots;
    int last;                       // slot of the last completed gather (-1: none)
    int64_t k;                      // steps submitted
    double *scal_dev;               // small device scratch for barrier / host all-reduce
    double *scal_host;              // pinned
};

#define FMK_NCCL(ctx, call)                                                                                   \
    do {                                                                                                      \
        ncclResult_t r__ = (call);                                                                            \
        if (r__ != ncclSuccess) {                                                                             \
            char b__[400];                                                                                    \
            snprintf(b__, sizeof(b__), "NCCL error %s at %s:%d (%s)", g_nccl.GetErrorString(r__), __FILE__,   \
                     __LINE__, #call);                                                                        \
            return fmk_fail((ctx), FMK_ERR_CUDA, b__);                                                        \
        }                                                                                                     \
    } while (0)

extern "C" {

int fmk_comm_unique_id(void *out128) {
    const char *e = nccl_load();
    if (e) return FMK_ERR_CUDA;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return FMK_ERR_CUDA;
    memcpy(out128, &id, sizeof(id));
    return FMK_OK;
}

int fmk_comm_init(fmk_ctx *ctx, const void *id128, int rank, int world, int max_ctas, fmk_comm **out) {
    FMK_ENTER(ctx);
    *out = nullptr;
    if (world < 1 || world > 64 || rank < 0 || rank >= world) return fmk_fail(ctx, FMK_ERR_ARG, "bad rank / world size");
    const char *e = nccl_load();
    if (e) return fmk_fail(ctx, FMK_ERR_CUDA, e);
    fmk_comm *c = new (std::nothrow) fmk_comm();
    if (!c) return FMK_ERR_ALLOC;
    memset(c, 0, sizeof(*c));
    c->ctx = ctx; c->rank = rank; c->world = world; c->npending = 0; c->last = -1;
    c->lag = 1;
    if (const char *e = getenv("FMK_COMM_LAG")) c->lag = atoi(e) >= 1 && atoi(e) <= FMK_COMM_SLOTS - 1 ? atoi(e) : c->lag;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r;
    if (g_nccl.CommInitRankConfig && max_ctas > 0) {
        // a small CTA budget: the payload is MBs against NVLink 5, and every SM NCCL takes is an SM the step's own
        // kernels lose (the r01 scaling loss at N = 8 was a wave-exact kernel spilling into an extra partial wave)
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.maxCTAs = max_ctas;
        cfg.minCTAs = 1;
        r = g_nccl.CommInitRankConfig(&c->comm, world, id, rank, &cfg);
    } else r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        char b[300];
        snprintf(b, sizeof(b), "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete c;
        return fmk_fail(ctx, FMK_ERR_CUDA, b);
    }
    cudaError_t ce = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    for (int s = 0; s < FMK_COMM_SLOTS && ce == cudaSuccess; s++) {
        ce = cudaEventCreateWithFlags(&c->ready[s], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->sized[s], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->done[s], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaMalloc(&c->counts_dev[s], sizeof(int64_t) * world);
        if (ce == cudaSuccess) ce = cudaMalloc(&c->mine_dev[s], sizeof(int64_t));
        if (ce == cudaSuccess) ce = cudaHostAlloc(&c->counts_host[s], sizeof(int64_t) * world, cudaHostAllocDefault);
        if (ce == cudaSuccess) ce = cudaHostAlloc(&c->mine_host[s], sizeof(int64_t), cudaHostAllocDefault);
    }
    if (ce == cudaSuccess) ce = cudaMalloc(&c->scal_dev, 64 * sizeof(double));
    if (ce == cudaSuccess) ce = cudaHostAlloc(&c->scal_host, 64 * sizeof(double), cudaHostAllocDefault);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->ipc_dev, 64 * (size_t)(world + 1));
    if (ce == cudaSuccess) ce = cudaHostAlloc(&c->ipc_host, 64 * (size_t)(world + 1), cudaHostAllocDefault);
    if (ce != cudaSuccess) return fmk_fail(ctx, FMK_ERR_CUDA, cudaGetErrorString(ce));
    c->p2p = world > 1;
    if (const char *e = getenv("FMK_COMM_P2P")) c->p2p = c->p2p && atoi(e) != 0;
    for (int s = 0; s < FMK_COMM_SLOTS; s++) c->agreed_dst[s] = -1;
    // (Measured and rejected: sizing the dollar task pass's waves without the SMs NCCL may take -- ctx->reserved_sms = maxCTAs --
    //  made the pass 0.3 ms slower at every N and no faster on the receiving rank at N = 8: 3.93 ms either way with 16 CTAs.)
    *out = c;
    return FMK_OK;
}

static int comm_unmap_all(fmk_comm *c);

// COLLECTIVE when frames were gathered with the peer-to-peer push (every rank unmaps before the destination frees).
void fmk_comm_destroy(fmk_comm *c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->comm) comm_unmap_all(c);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    for (int s = 0; s < FMK_COMM_SLOTS; s++) {
        cudaFree(c->staging[s]); cudaFree(c->recv[s]); cudaFree(c->counts_dev[s]); cudaFree(c->mine_dev[s]);
        cudaFreeHost(c->counts_host[s]); cudaFreeHost(c->mine_host[s]);
        cudaEventDestroy(c->ready[s]); cudaEventDestroy(c->sized[s]); cudaEventDestroy(c->done[s]);
    }
    cudaFree(c->scal_dev);
    cudaFreeHost(c->scal_host);
    cudaFree(c->ipc_dev);
    cudaFreeHost(c->ipc_host);
    cudaStreamDestroy(c->stream);
    delete c;
}

int fmk_comm_rank(const fmk_comm *c) { return c->rank; }
int fmk_comm_p2p_active(const fmk_comm *c) { return c->p2p; }
int fmk_comm_world(const fmk_comm *c) { return c->world; }

// Host-value all-reduce (op: 0 = max, 1 = min, 2 = sum) of n <= 64 doubles; also the barrier (n = 0 reduces one dummy).
// Ordered after everything queued on the ctx stream; returns when every rank has contributed.
int fmk_comm_allreduce_f64(fmk_comm *c, double *inout, int n, int op) {
    fmk_ctx *ctx = c->ctx;
    FMK_ENTER(ctx);
    if (n < 0 || n > 64) return fmk_fail(ctx, FMK_ERR_ARG, "allreduce of at most 64 values");
    const int m = n > 0 ? n : 1;
    for (int i = 0; i < m; i++) c->scal_host[i] = n > 0 ? inout[i] : 0.0;
    FMK_CUDA(ctx, cudaMemcpyAsync(c->scal_dev, c->scal_host, m * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const ncclRedOp_t rop = op == 0 ? ncclMax : (op == 1 ? ncclMin : ncclSum);
    FMK_NCCL(ctx, g_nccl.AllReduce(c->scal_dev, c->scal_dev, (size_t)m, ncclDouble, rop, c->comm, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(c->scal_host, c->scal_dev, m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) inout[i] = c->scal_host[i];
    return FMK_OK;
}

int fmk_comm_barrier(fmk_comm *c) {
    FMK_CUDA(c->ctx, cudaStreamSynchronize(c->stream));
    return fmk_comm_allreduce_f64(c, nullptr, 0, 2);
}

static int comm_grow(fmk_ctx *ctx, char **buf, int64_t *cap, int64_t need, cudaStream_t quiesce) {
    if (*cap >= need) return FMK_OK;
    if (*buf) { FMK_CUDA(ctx, cudaStreamSynchronize(quiesce)); FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(*buf); }
    *buf = nullptr; *cap = 0;
    const int64_t want = need + need / 4 + 4096;         // slack: streams of similar size never reallocate again
    FMK_CUDA(ctx, cudaMalloc((void **)buf, (size_t)want));
    *cap = want;
    return FMK_OK;
}

// tiny all-gather of one int64 per rank on the comm stream, waited for on the host: barrier + 1 word of agreement
static int comm_agree(fmk_comm *c, int s, int64_t mine, int64_t *all /* [world] */) {
    fmk_ctx *ctx = c->ctx;
    *c->mine_host[s] = mine;
    FMK_CUDA(ctx, cudaMemcpyAsync(c->mine_dev[s], c->mine_host[s], 8, cudaMemcpyHostToDevice, c->stream));
    FMK_NCCL(ctx, g_nccl.AllGather(c->mine_dev[s], c->counts_dev[s], 1, ncclInt64, c->comm, c->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(all, c->counts_dev[s], 8 * (size_t)c->world, cudaMemcpyDeviceToHost, c->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(c->stream));
    return FMK_OK;
}

// Collective (every rank takes the same decision from the same byte counts): (re)allocate the destination's receive buffer of
// slot s with `want` bytes, export it with CUDA IPC and map it on the other ranks.  Order matters: the others unmap, a barrier,
// the destination frees / allocates / exports, the handle travels in an all-gather, the others map, and a last word of
// agreement turns the peer-to-peer path off everywhere if any rank could not map.
static int comm_p2p_regrow(fmk_comm *c, int s, int dst, int64_t want) {
    fmk_ctx *ctx = c->ctx;
    int64_t *all = reinterpret_cast<int64_t *>(c->ipc_host);       // scratch for comm_agree ([world] int64 <= 64 * (world + 1) bytes)
    FMK_CUDA(ctx, cudaStreamSynchronize(c->stream));
    if (c->peer_recv[s]) { cudaIpcCloseMemHandle(c->peer_recv[s]); c->peer_recv[s] = nullptr; }
    FMK_TRY(comm_agree(c, s, 0, all));                             // every rank has unmapped the old buffer
    unsigned char *mine = c->ipc_host + 64 * (size_t)c->world;
    memset(mine, 0, 64);
    int64_t ok = 1;
    if (c->rank == dst) {
        FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (c->recv[s]) cudaFree(c->recv[s]);
        c->recv[s] = nullptr; c->recv_cap[s] = 0;
        FMK_CUDA(ctx, cudaMalloc((void **)&c->recv[s], (size_t)want));
        c->recv_cap[s] = want;
        cudaIpcMemHandle_t h;
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle travels as 64 bytes");
        if (cudaIpcGetMemHandle(&h, c->recv[s]) == cudaSuccess) memcpy(mine, &h, 64);
        else { ok = 0; cudaGetLastError(); }
    }
    FMK_CUDA(ctx, cudaMemcpyAsync(c->ipc_dev + 64 * (size_t)c->world, mine, 64, cudaMemcpyHostToDevice, c->stream));
    FMK_NCCL(ctx, g_nccl.AllGather(c->ipc_dev + 64 * (size_t)c->world, c->ipc_dev, 64, ncclUint8, c->comm, c->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(c->ipc_host, c->ipc_dev, 64 * (size_t)c->world, cudaMemcpyDeviceToHost, c->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(c->stream));
    if (c->rank != dst) {
        cudaIpcMemHandle_t h;
        memcpy(&h, c->ipc_host + 64 * (size_t)dst, 64);
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) c->peer_recv[s] = (char *)p;
        else { ok = 0; cudaGetLastError(); }
    }
    FMK_TRY(comm_agree(c, s, ok, all));
    for (int r = 0; r < c->world; r++) ok &= all[r];
    if (!ok) {                                                     // some rank cannot map: everybody falls back to send / recv
        if (c->peer_recv[s]) { cudaIpcCloseMemHandle(c->peer_recv[s]); c->peer_recv[s] = nullptr; }
        c->p2p = 0;
    }
    c->agreed_cap[s] = want;
    c->agreed_dst[s] = dst;
    return FMK_OK;
}

// Collective: every rank unmaps the destination's receive buffers, then a barrier -- an exported buffer must not be freed
// while another process still has it mapped.  (No-op when nothing was ever exported.)
static int comm_unmap_all(fmk_comm *c) {
    bool mapped = false;
    for (int s = 0; s < FMK_COMM_SLOTS; s++) {
        if (c->peer_recv[s]) { cudaIpcCloseMemHandle(c->peer_recv[s]); c->peer_recv[s] = nullptr; }
        mapped |= c->agreed_dst[s] >= 0;
        c->agreed_cap[s] = 0; c->agreed_dst[s] = -1;
    }
    if (mapped && c->world > 1) {
        int64_t all[64];
        FMK_TRY(comm_agree(c, 0, 0, all));
    }
    return FMK_OK;
}

// post the transfer of the slot whose byte counts have been exchanged
static int comm_complete_oldest(fmk_comm *c) {
    fmk_ctx *ctx = c->ctx;
    if (c->npending <= 0) return FMK_OK;
    const int s = c->pend_slot[0], dst = c->pend_dst[0];
    for (int q = 1; q < c->npending; q++) { c->pend_slot[q - 1] = c->pend_slot[q]; c->pend_dst[q - 1] = c->pend_dst[q]; }
    c->npending--;
    FMK_CUDA(ctx, cudaEventSynchronize(c->sized[s]));
    int64_t cnt[64];
    for (int r = 0; r < c->world; r++) cnt[r] = c->counts_host[s][r];     // (the pinned block doubles as scratch below)
    int64_t tot = 0;
    for (int r = 0; r < c->world; r++) { c->recv_off[s][r] = tot; tot += (cnt[r] + 255) / 256 * 256; }
    c->recv_off[s][c->world] = tot;
    if (c->p2p && (tot > c->agreed_cap[s] || dst != c->agreed_dst[s])) {
        FMK_TRY(comm_p2p_regrow(c, s, dst, tot + tot / 4 + 4096));
        for (int r = 0; r < c->world; r++) c->counts_host[s][r] = cnt[r];
    }
    if (c->p2p) {
        if (cnt[c->rank] > 0) {
            char *to = (c->rank == dst ? c->recv[s] : c->peer_recv[s]) + c->recv_off[s][c->rank];
            FMK_CUDA(ctx, cudaMemcpyAsync(to, c->staging[s], (size_t)cnt[c->rank], cudaMemcpyDefault, c->stream));
        }
        // completion fence: queued behind the copy on every rank, so it ends on `dst` only after every rank's push has landed
        FMK_NCCL(ctx, g_nccl.AllGather(c->mine_dev[s], c->counts_dev[s], 1, ncclInt64, c->comm, c->stream));
    } else {
        if (c->rank == dst) FMK_TRY(comm_grow(ctx, &c->recv[s], &c->recv_cap[s], tot, c->stream));
        FMK_NCCL(ctx, g_nccl.GroupStart());
        if (c->rank == dst) {
            for (int r = 0; r < c->world; r++) {
                if (r == dst || cnt[r] == 0) continue;
                FMK_NCCL(ctx, g_nccl.Recv(c->recv[s] + c->recv_off[s][r], (size_t)cnt[r], ncclUint8, r, c->comm, c->stream));
            }
        } else if (cnt[c->rank] > 0) {
            FMK_NCCL(ctx, g_nccl.Send(c->staging[s], (size_t)cnt[c->rank], ncclUint8, dst, c->comm, c->stream));
        }
        FMK_NCCL(ctx, g_nccl.GroupEnd());
        if (c->rank == dst && cnt[dst] > 0)   // own frame: device-to-device on the comm stream
            FMK_CUDA(ctx, cudaMemcpyAsync(c->recv[s] + c->recv_off[s][dst], c->staging[s], (size_t)cnt[dst], cudaMemcpyDeviceToDevice, c->stream));
    }
    FMK_CUDA(ctx, cudaEventRecord(c->done[s], c->stream));
    c->has_done[s] = 1;
    c->last = s;
    return FMK_OK;
}
// keep at most `keep` steps waiting for their transfer
static int comm_complete_pending(fmk_comm *c, int keep) {
    while (c->npending > keep) FMK_TRY(comm_complete_oldest(c));
    return FMK_OK;
}

// One gather step: the nseg device segments (pointers valid on the ctx stream) are packed back to back, each padded to 16
// bytes, into this rank's frame; the frame is gathered to `dst`.  Returns immediately (see the pipeline at the top).
int fmk_comm_gather_submit(fmk_comm *c, const void *const *seg_ptrs, const int64_t *seg_bytes, int nseg, int dst) {
    fmk_ctx *ctx = c->ctx;
    FMK_ENTER(ctx);
    if (nseg < 0 || nseg > FMK_COMM_MAXSEG) return fmk_fail(ctx, FMK_ERR_ARG, "too many segments");
    if (dst < 0 || dst >= c->world) return fmk_fail(ctx, FMK_ERR_ARG, "bad destination rank");
    int64_t total = 0;
    for (int q = 0; q < nseg; q++) {
        if (seg_bytes[q] < 0) return fmk_fail(ctx, FMK_ERR_ARG, "negative segment size");
        total += (seg_bytes[q] + 15) / 16 * 16;
    }
    if (c->k == 0) {
        // pipeline depth is fixed by the first frame: lag + 1 staging / receive slots; the destination holds world frames per
        // slot, so multi-GB frames (footprint CSR of 5e8 ticks: 4 GB per rank) fall back to the two-slot pipeline
        if (total * (int64_t)c->world > ((int64_t)4 << 30)) c->lag = 1;
        c->nslots = c->lag + 1;
    }
    FMK_TRY(comm_complete_pending(c, c->lag - 1));
    const int s = (int)(c->k % c->nslots);
    if (c->has_done[s]) {             // the transfer that read this slot two steps ago must be finished before it is rewritten
        FMK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, c->done[s], 0));
    }
    FMK_TRY(comm_grow(ctx, &c->staging[s], &c->staging_cap[s], total, c->stream));
    int64_t off = 0;
    for (int q = 0; q < nseg; q++) {
        if (seg_bytes[q] > 0)
            FMK_CUDA(ctx, cudaMemcpyAsync(c->staging[s] + off, seg_ptrs[q], (size_t)seg_bytes[q], cudaMemcpyDeviceToDevice, ctx->stream));
        off += (seg_bytes[q] + 15) / 16 * 16;
    }
    FMK_CUDA(ctx, cudaEventRecord(c->ready[s], ctx->stream));
    FMK_CUDA(ctx, cudaStreamWaitEvent(c->stream, c->ready[s], 0));
    *c->mine_host[s] = total;
    FMK_CUDA(ctx, cudaMemcpyAsync(c->mine_dev[s], c->mine_host[s], 8, cudaMemcpyHostToDevice, c->stream));
    FMK_NCCL(ctx, g_nccl.AllGather(c->mine_dev[s], c->counts_dev[s], 1, ncclInt64, c->comm, c->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(c->counts_host[s], c->counts_dev[s], 8 * (size_t)c->world, cudaMemcpyDeviceToHost, c->stream));
    FMK_CUDA(ctx, cudaEventRecord(c->sized[s], c->stream));
    c->pend_slot[c->npending] = s;
    c->pend_dst[c->npending] = dst;
    c->npending++;
    c->k++;
    return FMK_OK;
}

// Posts what is still pending and makes the ctx stream wait for every outstanding transfer (stream-ordered, no host sync
// beyond the byte-count wait): a timer stopped on the ctx stream afterwards covers all gathers.
int fmk_comm_gather_finish(fmk_comm *c) {
    fmk_ctx *ctx = c->ctx;
    FMK_ENTER(ctx);
    FMK_TRY(comm_complete_pending(c, 0));
    for (int s = 0; s < FMK_COMM_SLOTS; s++)
        if (c->has_done[s]) FMK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, c->done[s], 0));
    return FMK_OK;
}

// Ends a sequence of gather steps (COLLECTIVE: every rank calls it): waits for everything outstanding, releases the staging / receive buffers and lets the next
// fmk_comm_gather_submit size the pipeline afresh (a host that gathers small OHLCV frames first and multi-GB footprint
// frames later must not keep three receive slots of the large size).
int fmk_comm_gather_reset(fmk_comm *c) {
    fmk_ctx *ctx = c->ctx;
    FMK_ENTER(ctx);
    FMK_TRY(fmk_comm_gather_finish(c));
    FMK_CUDA(ctx, cudaStreamSynchronize(c->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FMK_TRY(comm_unmap_all(c));
    for (int s = 0; s < FMK_COMM_SLOTS; s++) {
        cudaFree(c->staging[s]); c->staging[s] = nullptr; c->staging_cap[s] = 0;
        cudaFree(c->recv[s]); c->recv[s] = nullptr; c->recv_cap[s] = 0;
        c->has_done[s] = 0;
    }
    c->k = 0; c->last = -1; c->lag = 1;
    if (const char *e = getenv("FMK_COMM_LAG")) c->lag = atoi(e) >= 1 && atoi(e) <= FMK_COMM_SLOTS - 1 ? atoi(e) : c->lag;
    return FMK_OK;
}

// On the destination rank, after fmk_comm_gather_finish: where rank r's frame of the LAST step landed (device pointer,
// exact byte count).  Other ranks get bytes = their own count and a null pointer.
int fmk_comm_gather_result(fmk_comm *c, int r, void **dev_ptr, int64_t *bytes) {
    fmk_ctx *ctx = c->ctx;
    *dev_ptr = nullptr; *bytes = 0;
    if (c->last < 0 || r < 0 || r >= c->world) return fmk_fail(ctx, FMK_ERR_ARG, "no finished gather / bad rank");
    const int s = c->last;
    *bytes = c->counts_host[s][r];
    if (c->recv[s]) *dev_ptr = c->recv[s] + c->recv_off[s][r];
    return FMK_OK;
}

// Copies rank r's gathered frame to the host (destination rank only; waits for the transfer).
int fmk_comm_gather_download(fmk_comm *c, int r, void *host, int64_t cap) {
    fmk_ctx *ctx = c->ctx;
    FMK_ENTER(ctx);
    void *p; int64_t b;
    FMK_TRY(fmk_comm_gather_result(c, r, &p, &b));
    if (!p) return fmk_fail(ctx, FMK_ERR_ARG, "not the destination rank");
    if (b > cap) return fmk_fail(ctx, FMK_ERR_CAPACITY, "host buffer too small for the gathered frame");
    FMK_CUDA(ctx, cudaStreamSynchronize(c->stream));
    if (b > 0) FMK_CUDA(ctx, cudaMemcpy(host, p, (size_t)b, cudaMemcpyDeviceToHost));
    return FMK_OK;
}

int fmk_comm_nccl_version(void) {
    if (nccl_load()) return 0;
    int v = 0;
    g_nccl.GetVersion(&v);
    return v;
}

}  // extern "C"
