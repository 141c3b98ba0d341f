// common.cuh -- context, handles and helpers shared by all translation units of libfmk.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/fmk.h"

struct fmk_prof_rec { const char *name; cudaEvent_t a, b; };

struct fmk_ctx {
    int device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    int sm_count;
    int reserved_sms;                     // SMs a concurrent communicator may occupy (fmk_comm_init sets it): wave-exact grids plan without them
    int64_t launches;
    char err[512];
    // ctx-owned result columns of fmk_bar_ohlcv_device (kept so the device-resident bench has real outputs)
    void *res_cols;
    int64_t res_cols_bytes;
    void *flush_buf;
    int64_t flush_bytes;
    int64_t stats[3];
    int64_t cusum_filled;                 // NaNs of sigma forward-filled by the last fmk_cusum_bar_index call
    int owns_stream;
    int prof_on;                          // per-kernel CUDA-event timing (bench.py roofline leg)
    std::vector<fmk_prof_rec> *prof;
    std::vector<cudaEvent_t> *ev_pool;
    int64_t res_nb;                       // number of bars held in res_cols
    // large-block cache (see fmk_dalloc): free blocks and the sizes of the live ones
    std::vector<std::pair<void *, size_t>> *cache_free;
    std::unordered_map<void *, size_t> *cache_live;
    int64_t cache_bytes;                  // bytes held by the cache (free + live)
    struct fmk_copier *copier;            // lazily created staging threads for pageable copies (api.cu)
    int64_t cache_free_bytes, cache_free_cap;   // free-list total and its cap (half the device memory): oldest blocks go first
};

static inline cudaEvent_t fmk_prof_event(fmk_ctx *ctx) {
    cudaEvent_t e;
    if (!ctx->ev_pool->empty()) { e = ctx->ev_pool->back(); ctx->ev_pool->pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

struct fmk_trades {
    int64_t n;
    int64_t *ts;
    double *price;
    double *amount;
    int8_t *side;      // may be null
    double *log_price; // lazily built by triple_barrier
};

struct fmk_index {
    int64_t m;          // n_bars + 1
    int64_t n_ticks;    // length of the trade arrays it indexes
    int64_t *close_ts;  // may be null for fmk_index_from_host without timestamps
    int64_t *close_idx;
    int sorted;         // close_idx is non-decreasing and within [-1, n_ticks): always true for device-built indices
};

struct fmk_buf {
    int64_t bytes;
    void *ptr;
};

struct fmk_footprint {
    int64_t n_bars, n_levels;
    int64_t *level_offsets;  // [n_bars+1]
    int32_t *price_levels;
    float *buy_vol, *sell_vol;
    int32_t *buy_ticks, *sell_ticks;
    uint8_t *buy_imb, *sell_imb;
    uint16_t *buy_imb_sum, *sell_imb_sum;
    int32_t *cot;
    int16_t *run_signed;
    double *vp_skew, *vp_gini;
};

// device-resident bar frame (fmk_bar_features_device): one block of per-bar columns + one block of per-level columns
struct fmk_frame {
    int64_t n_bars, n_levels;
    int flags;
    char *bar_block;
    int64_t bar_bytes;
    char *level_block;
    int64_t level_bytes;
    int64_t col_off[FMK_COL_COUNT];   // byte offset inside its block; -1 = column absent
};

static inline int fmk_fail(fmk_ctx *ctx, int code, const char *msg) {
    if (ctx) {
        snprintf(ctx->err, sizeof(ctx->err), "%s", msg);
    }
    return code;
}

#define FMK_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            char b__[400];                                                                       \
            snprintf(b__, sizeof(b__), "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__),   \
                     __FILE__, __LINE__, #call);                                                 \
            return fmk_fail((ctx), FMK_ERR_CUDA, b__);                                           \
        }                                                                                        \
    } while (0)

// Every extern "C" entry point that takes a ctx starts with this: a process may hold one ctx per GPU, and the CUDA
// "current device" is per host thread, so each call re-binds its own device (a no-op when it already is current).
#define FMK_ENTER(ctx)                                   \
    do {                                                 \
        if (ctx) cudaSetDevice((ctx)->device);           \
    } while (0)

#define FMK_TRY(call)                \
    do {                             \
        int rc__ = (call);           \
        if (rc__ != FMK_OK) return rc__; \
    } while (0)

// Launch on the ctx stream, count it, and surface launch-configuration errors immediately.
#define FMK_LAUNCH(ctx, kernel, grid, block, smem, ...)                                  \
    do {                                                                                 \
        fmk_prof_rec pr__ = {#kernel, nullptr, nullptr};                                 \
        if ((ctx)->prof_on) {                                                            \
            pr__.a = fmk_prof_event(ctx); pr__.b = fmk_prof_event(ctx);                  \
            cudaEventRecord(pr__.a, (ctx)->stream);                                      \
        }                                                                                \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                 \
        if ((ctx)->prof_on) {                                                            \
            cudaEventRecord(pr__.b, (ctx)->stream);                                      \
            (ctx)->prof->push_back(pr__);                                                \
        }                                                                                \
        (ctx)->launches++;                                                               \
        FMK_CUDA((ctx), cudaGetLastError());                                             \
    } while (0)

// Device scratch.  Small blocks come from the stream-ordered pool (cudaMallocAsync, release threshold = infinity).  LARGE
// blocks (column-sized scratch: prefix sums, next[] tables, r / lambda series, CSR blocks -- gigabytes at 1e9 ticks) are kept in
// a per-context cache instead: measured on B200, the pool re-maps physical pages between virtual ranges when differently
// sized multi-GB blocks are freed and re-allocated, which stalled single steps by 40 - 900 ms.  A step asks for the same sizes
// in the same order every time, so after the first step every request is an exact-size hit.  Everything of a context runs on
// one stream, so handing a freed block to a later request needs no synchronisation (stream order is reuse order).
constexpr size_t FMK_CACHE_MIN_BYTES = (size_t)4 << 20;

static inline void fmk_cache_trim(fmk_ctx *ctx) {
    if (!ctx->cache_free || ctx->cache_free->empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto &b : *ctx->cache_free) { cudaFree(b.first); ctx->cache_bytes -= (int64_t)b.second; }
    ctx->cache_free->clear();
    ctx->cache_free_bytes = 0;
}

static inline int fmk_dalloc_bytes(fmk_ctx *ctx, void **p, size_t bytes) {
    *p = nullptr;
    if (bytes < FMK_CACHE_MIN_BYTES || !ctx->cache_free) {
        FMK_CUDA(ctx, cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream));
        return FMK_OK;
    }
    auto &fr = *ctx->cache_free;
    int best = -1;
    for (int k = 0; k < (int)fr.size(); k++)
        if (fr[k].second >= bytes && fr[k].second <= bytes + bytes / 4 + ((size_t)1 << 20) && (best < 0 || fr[k].second < fr[best].second)) best = k;
    if (best >= 0) {
        *p = fr[best].first;
        (*ctx->cache_live)[*p] = fr[best].second;
        ctx->cache_free_bytes -= (int64_t)fr[best].second;
        fr.erase(fr.begin() + best);
        return FMK_OK;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {            // out of memory: give the cached blocks back to the driver and retry once
        cudaGetLastError();
        fmk_cache_trim(ctx);
        e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) {
        char b[200];
        snprintf(b, sizeof(b), "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return fmk_fail(ctx, FMK_ERR_ALLOC, b);
    }
    (*ctx->cache_live)[*p] = bytes;
    ctx->cache_bytes += (int64_t)bytes;
    return FMK_OK;
}

template <typename T>
static inline int fmk_dalloc(fmk_ctx *ctx, T **p, int64_t count) {
    if (count <= 0) count = 1;
    return fmk_dalloc_bytes(ctx, (void **)p, (size_t)count * sizeof(T));
}
static inline void fmk_dfree(fmk_ctx *ctx, void *p) {
    if (!p) return;
    if (ctx->cache_live) {
        auto it = ctx->cache_live->find(p);
        if (it != ctx->cache_live->end()) {
            ctx->cache_free->push_back({p, it->second});
            ctx->cache_free_bytes += (int64_t)it->second;
            ctx->cache_live->erase(it);
            // the free list is capped: workloads that keep changing sizes must not hoard the device (oldest blocks go first)
            while (ctx->cache_free_bytes > ctx->cache_free_cap && ctx->cache_free->size() > 1) {
                auto b = ctx->cache_free->front();
                cudaStreamSynchronize(ctx->stream);
                cudaFree(b.first);
                ctx->cache_bytes -= (int64_t)b.second;
                ctx->cache_free_bytes -= (int64_t)b.second;
                ctx->cache_free->erase(ctx->cache_free->begin());
            }
            return;
        }
    }
    cudaFreeAsync(p, ctx->stream);
}
// RAII scratch buffer freed (stream-ordered) at scope exit
template <typename T>
struct Scratch {
    fmk_ctx *ctx;
    T *p;
    Scratch(fmk_ctx *c) : ctx(c), p(nullptr) {}
    int alloc(int64_t count) { return fmk_dalloc(ctx, &p, count); }
    ~Scratch() { fmk_dfree(ctx, p); }
    operator T *() const { return p; }
};

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Host <-> device copies of column-sized buffers.  Pinned (or registered) host memory goes straight to cudaMemcpyAsync on the
// ctx stream.  PAGEABLE memory -- the NumPy / pandas columns every wrapper call hands over -- is staged by the runtime
// through one internal bounce buffer at ~9 GB/s (measured, e2e_wrapper of round 1); here several host threads copy chunks
// into their own pinned bounce buffers and issue the DMA on their own streams, which takes the copy to the PCIe rate.
// Both return with the copy ordered on the ctx stream (pinned) or complete (pageable), like cudaMemcpy would.
int fmk_copy_h2d(fmk_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int fmk_copy_d2h(fmk_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);

// internal cross-TU entry points
int fmk_dollar_index_impl(fmk_ctx *ctx, const fmk_trades *t, double threshold, fmk_index **out);
int fmk_volume_index_impl(fmk_ctx *ctx, const fmk_trades *t, double threshold, fmk_index **out);
int fmk_cusum_index_impl(fmk_ctx *ctx, const fmk_trades *t, fmk_buf *sigma, double sigma_floor, double sigma_mult, fmk_index **out);
int fmk_gather_close_ts(fmk_ctx *ctx, const fmk_trades *t, fmk_index *ix);
