// common.cuh -- context, handles and helpers shared by all translation units of libfmk.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/fmk.h"

struct fmk_prof_rec { const char *name; cudaEvent_t a, b; };

struct fmk_ctx {
    int device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    int sm_count;
    int64_t launches;
    char err[512];
    // ctx-owned result columns of fmk_bar_ohlcv_device (kept so the device-resident bench has real outputs)
    void *res_cols;
    int64_t res_cols_bytes;
    void *flush_buf;
    int64_t flush_bytes;
    int64_t stats[3];
    int owns_stream;
    int prof_on;                          // per-kernel CUDA-event timing (bench.py roofline leg)
    std::vector<fmk_prof_rec> *prof;
    std::vector<cudaEvent_t> *ev_pool;
    int64_t res_nb;                       // number of bars held in res_cols
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

template <typename T>
static inline int fmk_dalloc(fmk_ctx *ctx, T **p, int64_t count) {
    *p = nullptr;
    if (count <= 0) count = 1;
    FMK_CUDA(ctx, cudaMallocAsync((void **)p, (size_t)count * sizeof(T), ctx->stream));
    return FMK_OK;
}
static inline void fmk_dfree(fmk_ctx *ctx, void *p) {
    if (p) cudaFreeAsync(p, ctx->stream);
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

// internal cross-TU entry points
int fmk_dollar_index_impl(fmk_ctx *ctx, const fmk_trades *t, double threshold, fmk_index **out);
int fmk_volume_index_impl(fmk_ctx *ctx, const fmk_trades *t, double threshold, fmk_index **out);
int fmk_cusum_index_impl(fmk_ctx *ctx, const fmk_trades *t, fmk_buf *sigma, double sigma_floor, double sigma_mult, fmk_index **out);
int fmk_gather_close_ts(fmk_ctx *ctx, const fmk_trades *t, fmk_index *ix);
