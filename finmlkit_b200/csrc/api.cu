// api.cu -- context / handle management, time & tick bar indexers, synthetic stream generator.
#include <math.h>
#include <new>
#include "common.cuh"
#include "scan.cuh"

#include <algorithm>
#include <atomic>
#include <thread>

// ---------------------------------------------------------------------------------------------------------------
// staged multi-threaded copies of pageable host memory
// ---------------------------------------------------------------------------------------------------------------
constexpr size_t COPY_CHUNK = (size_t)8 << 20;       // largest staged chunk (= size of a pinned bounce buffer)
constexpr size_t COPY_DIRECT_BELOW = (size_t)4 << 20;
struct fmk_copier {
    int nthreads;
    std::vector<void *> slots;            // 2 pinned bounce buffers per thread
    std::vector<cudaStream_t> streams;
    std::vector<cudaEvent_t> events;      // one per slot
};

static fmk_copier *copier_get(fmk_ctx *ctx) {
    if (ctx->copier) return ctx->copier;
    fmk_copier *c = new (std::nothrow) fmk_copier();
    if (!c) return nullptr;
    int want = 8;
    if (const char *e = getenv("FMK_COPY_THREADS")) want = atoi(e);
    const int hw = (int)std::thread::hardware_concurrency();
    c->nthreads = std::max(1, std::min(want, hw > 0 ? hw : 4));
    for (int t = 0; t < c->nthreads; t++) {
        cudaStream_t st;
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { c->nthreads = t; break; }
        c->streams.push_back(st);
        for (int b = 0; b < 2; b++) {
            void *p = nullptr;
            cudaEvent_t ev;
            if (cudaHostAlloc(&p, COPY_CHUNK, cudaHostAllocDefault) != cudaSuccess || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                delete c;          // (leaks what was created so far only in this out-of-pinned-memory corner)
                return nullptr;
            }
            c->slots.push_back(p);
            c->events.push_back(ev);
        }
    }
    if (c->nthreads < 1) { delete c; return nullptr; }
    ctx->copier = c;
    return c;
}

static void copier_destroy(fmk_ctx *ctx) {
    fmk_copier *c = ctx->copier;
    if (!c) return;
    for (auto st : c->streams) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (auto p : c->slots) cudaFreeHost(p);
    for (auto e : c->events) cudaEventDestroy(e);
    delete c;
    ctx->copier = nullptr;
}

static bool host_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static int copy_staged(fmk_ctx *ctx, char *dev, char *host, size_t bytes, bool to_device) {
    fmk_copier *c = copier_get(ctx);
    if (!c) return -1;
    // the device range may still be in use by work queued on the ctx stream (cached blocks are reused in stream order)
    cudaEvent_t ready = fmk_prof_event(ctx);
    cudaEventRecord(ready, ctx->stream);
    // medium-sized columns (a 1e6-tick wrapper call moves 8 MB per column) still want every thread busy: shrink the chunk
    size_t chunk = bytes / (2 * (size_t)c->nthreads);
    chunk = std::max((size_t)1 << 20, std::min(COPY_CHUNK, (chunk + 4095) / 4096 * 4096));
    const size_t nchunks = (bytes + chunk - 1) / chunk;
    std::atomic<size_t> next(0);
    std::atomic<int> failed(0);
    const int device = ctx->device;
    auto worker = [&](int t) {
        cudaSetDevice(device);
        cudaStream_t st = c->streams[t];
        cudaStreamWaitEvent(st, ready, 0);
        int used[2] = {0, 0};
        for (int b = 0;; b ^= 1) {
            const size_t k = next.fetch_add(1);
            if (k >= nchunks) break;
            const size_t off = k * chunk, len = std::min(chunk, bytes - off);
            void *slot = c->slots[2 * t + b];
            cudaEvent_t ev = c->events[2 * t + b];
            if (used[b] && cudaEventSynchronize(ev) != cudaSuccess) { failed = 1; break; }   // the DMA that used this slot is done
            if (to_device) {
                memcpy(slot, host + off, len);
                if (cudaMemcpyAsync(dev + off, slot, len, cudaMemcpyHostToDevice, st) != cudaSuccess) { failed = 1; break; }
                cudaEventRecord(ev, st);
                used[b] = 1;
            } else {
                if (cudaMemcpyAsync(slot, dev + off, len, cudaMemcpyDeviceToHost, st) != cudaSuccess) { failed = 1; break; }
                if (cudaStreamSynchronize(st) != cudaSuccess) { failed = 1; break; }
                memcpy(host + off, slot, len);
            }
        }
        if (cudaStreamSynchronize(st) != cudaSuccess) failed = 1;
    };
    std::vector<std::thread> th;
    const int nt = (int)std::min<size_t>((size_t)c->nthreads, nchunks);
    for (int t = 1; t < nt; t++) th.emplace_back(worker, t);
    worker(0);
    for (auto &x : th) x.join();
    ctx->ev_pool->push_back(ready);
    if (failed) { cudaGetLastError(); return -1; }
    return 0;
}

int fmk_copy_h2d(fmk_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    if (bytes == 0) return FMK_OK;
    if (bytes >= COPY_DIRECT_BELOW && !host_is_pinned(src_host) && copy_staged(ctx, (char *)dst_dev, (char *)src_host, bytes, true) == 0)
        return FMK_OK;
    FMK_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return FMK_OK;
}

int fmk_copy_d2h(fmk_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    if (bytes == 0) return FMK_OK;
    if (bytes >= COPY_DIRECT_BELOW && !host_is_pinned(dst_host) && copy_staged(ctx, (char *)src_dev, (char *)dst_host, bytes, false) == 0)
        return FMK_OK;
    FMK_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return FMK_OK;
}

extern "C" {

const char *fmk_version(void) { return "finmlkit_b200 0.1 (sm_100a)"; }

int fmk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static int ctx_create(int device, void *ext_stream, int use_ext, fmk_ctx **out);
int fmk_ctx_create(int device, fmk_ctx **out) { return ctx_create(device, nullptr, 0, out); }
int fmk_ctx_create_on_stream(int device, void *stream, fmk_ctx **out) { return ctx_create(device, stream, 1, out); }

static int ctx_create(int device, void *ext_stream, int use_ext, fmk_ctx **out) {
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device >= n) return FMK_ERR_CUDA;
    fmk_ctx *ctx = new (std::nothrow) fmk_ctx();
    if (!ctx) return FMK_ERR_ALLOC;
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->prof = new std::vector<fmk_prof_rec>();
    ctx->ev_pool = new std::vector<cudaEvent_t>();
    ctx->cache_free = new std::vector<std::pair<void *, size_t>>();
    ctx->cache_live = new std::unordered_map<void *, size_t>();
    ctx->cache_free_cap = (int64_t)48 << 30;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return FMK_ERR_CUDA; }
    if (use_ext) { ctx->stream = (cudaStream_t)ext_stream; ctx->owns_stream = 0; }
    else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return FMK_ERR_CUDA; }
        ctx->owns_stream = 1;
    }
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    {
        size_t fr = 0, tot = 0;
        if (cudaMemGetInfo(&fr, &tot) == cudaSuccess && tot > 0) ctx->cache_free_cap = (int64_t)(tot / 2);
    }
    // keep freed blocks in the stream-ordered pool: scratch allocation inside a step must not hit the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = ctx;
    return FMK_OK;
}

void fmk_ctx_destroy(fmk_ctx *ctx) {
    if (!ctx) return;
    int prev_dev = -1;               // run from a finalizer at an arbitrary time: leave the caller's current device alone
    cudaGetDevice(&prev_dev);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->res_cols) cudaFree(ctx->res_cols);
    if (ctx->flush_buf) cudaFree(ctx->flush_buf);
    copier_destroy(ctx);
    fmk_cache_trim(ctx);
    for (auto &kv : *ctx->cache_live) cudaFree(kv.first);      // handles the caller never freed
    delete ctx->cache_free;
    delete ctx->cache_live;
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    for (auto &r : *ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto &e : *ctx->ev_pool) cudaEventDestroy(e);
    delete ctx->prof;
    delete ctx->ev_pool;
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
}

const char *fmk_last_error(fmk_ctx *ctx) { return ctx ? ctx->err : "no context (CUDA device missing?)"; }

int fmk_ctx_sync(fmk_ctx *ctx) {
    FMK_ENTER(ctx);
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

int fmk_timer_start(fmk_ctx *ctx) {
    FMK_ENTER(ctx);
    FMK_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return FMK_OK;
}

int fmk_timer_stop(fmk_ctx *ctx, float *ms_out) {
    FMK_ENTER(ctx);
    FMK_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    FMK_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    FMK_CUDA(ctx, cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1));
    return FMK_OK;
}

int64_t fmk_launch_count(fmk_ctx *ctx) { return ctx->launches; }

int fmk_prof_enable(fmk_ctx *ctx, int on) {
    FMK_ENTER(ctx);
    ctx->prof_on = on;
    return FMK_OK;
}

// Drains the recorded launches: writes up to cap rows of (name, launches, total ms); returns the number of distinct
// kernels.  names_out receives cap * 64 bytes of NUL-terminated names.
int fmk_prof_report(fmk_ctx *ctx, char *names_out, int64_t *counts_out, float *ms_out, int cap) {
    FMK_ENTER(ctx);
    cudaStreamSynchronize(ctx->stream);
    int nk = 0;
    for (auto &r : *ctx->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        int k = 0;
        for (; k < nk; k++) if (strncmp(names_out + 64 * k, r.name, 63) == 0) break;
        if (k == nk) {
            if (nk >= cap) { ctx->ev_pool->push_back(r.a); ctx->ev_pool->push_back(r.b); continue; }
            strncpy(names_out + 64 * k, r.name, 63);
            names_out[64 * k + 63] = 0;
            counts_out[k] = 0; ms_out[k] = 0.f;
            nk++;
        }
        counts_out[k]++;
        ms_out[k] += ms;
        ctx->ev_pool->push_back(r.a);
        ctx->ev_pool->push_back(r.b);
    }
    ctx->prof->clear();
    return nk;
}

// device pointer / bar count of the columns written by fmk_bar_ohlcv_device (layout: 6 x f64[nb] open, high, low,
// close, vwap, median ; i64[nb] trades ; f32[nb] volume)
int fmk_result_cols(fmk_ctx *ctx, void **ptr, int64_t *n_bars, int64_t *bytes) {
    FMK_ENTER(ctx);
    *ptr = ctx->res_cols; *n_bars = ctx->res_nb; *bytes = ctx->res_nb * (6 * 8 + 8 + 4);
    return FMK_OK;
}

int fmk_index_stats(fmk_ctx *ctx, int64_t *s) {
    FMK_ENTER(ctx);
    s[0] = ctx->stats[0]; s[1] = ctx->stats[1]; s[2] = ctx->stats[2];
    return FMK_OK;
}

// gives the cached large scratch blocks back to the driver (they are otherwise kept for the next call of the same size)
int fmk_ctx_trim(fmk_ctx *ctx) {
    FMK_ENTER(ctx);
    fmk_cache_trim(ctx);
    return FMK_OK;
}

int fmk_host_alloc(void **out, int64_t bytes) {
    return cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? FMK_OK : FMK_ERR_ALLOC;
}
void fmk_host_free(void *p) { if (p) cudaFreeHost(p); }

}  // extern "C"

__global__ void k_flush(uint4 *buf, int64_t n16, unsigned v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x)
        buf[i] = make_uint4(v, v, v, v);
}

extern "C" int fmk_flush_l2(fmk_ctx *ctx) {
    FMK_ENTER(ctx);
    const int64_t bytes = 512ll << 20;  // 4x the 126 MB L2
    if (!ctx->flush_buf) {
        FMK_CUDA(ctx, cudaMalloc(&ctx->flush_buf, (size_t)bytes));
        ctx->flush_bytes = bytes;
    }
    static unsigned tick = 0;
    FMK_LAUNCH(ctx, k_flush, ctx->sm_count * 8, 256, 0, (uint4 *)ctx->flush_buf, bytes / 16, ++tick);
    return FMK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// buffers / trades
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

int fmk_buf_alloc(fmk_ctx *ctx, int64_t bytes, fmk_buf **out) {
    FMK_ENTER(ctx);
    *out = nullptr;
    fmk_buf *b = new (std::nothrow) fmk_buf();
    if (!b) return FMK_ERR_ALLOC;
    b->bytes = bytes;
    char *p;
    int rc = fmk_dalloc(ctx, &p, bytes);
    if (rc) { delete b; return rc; }
    b->ptr = p;
    *out = b;
    return FMK_OK;
}

int fmk_buf_upload(fmk_ctx *ctx, const void *host, int64_t bytes, fmk_buf **out) {
    FMK_ENTER(ctx);
    FMK_TRY(fmk_buf_alloc(ctx, bytes, out));
    if (bytes > 0) FMK_TRY(fmk_copy_h2d(ctx, (*out)->ptr, host, (size_t)bytes));
    return FMK_OK;
}

int fmk_buf_download(fmk_ctx *ctx, const fmk_buf *b, void *host, int64_t bytes) {
    FMK_ENTER(ctx);
    if (bytes > b->bytes) return fmk_fail(ctx, FMK_ERR_CAPACITY, "download larger than buffer");
    if (bytes > 0) FMK_TRY(fmk_copy_d2h(ctx, host, b->ptr, (size_t)bytes));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

int64_t fmk_buf_bytes(const fmk_buf *b) { return b->bytes; }
void *fmk_buf_devptr(const fmk_buf *b) { return b->ptr; }
void fmk_buf_free(fmk_ctx *ctx, fmk_buf *b) {
    FMK_ENTER(ctx);
    if (!b) return;
    fmk_dfree(ctx, b->ptr);
    delete b;
}

static int trades_alloc(fmk_ctx *ctx, int64_t n, int with_ts, int with_side, fmk_trades **out) {
    *out = nullptr;
    fmk_trades *t = new (std::nothrow) fmk_trades();
    if (!t) return FMK_ERR_ALLOC;
    memset(t, 0, sizeof(*t));
    t->n = n;
    int rc = FMK_OK;
    if (with_ts) rc = fmk_dalloc(ctx, &t->ts, n);
    if (!rc) rc = fmk_dalloc(ctx, &t->price, n);
    if (!rc) rc = fmk_dalloc(ctx, &t->amount, n);
    if (!rc && with_side) rc = fmk_dalloc(ctx, &t->side, n);
    if (rc) { fmk_trades_free(ctx, t); return rc; }
    *out = t;
    return FMK_OK;
}

int fmk_trades_refill(fmk_ctx *ctx, fmk_trades *t, const int64_t *ts, const double *price, const double *amount,
                      const int8_t *side, int64_t n) {
    FMK_ENTER(ctx);
    if (n != t->n) return fmk_fail(ctx, FMK_ERR_ARG, "refill length differs from handle length");
    if (n == 0) return FMK_OK;
    if (ts && t->ts) FMK_TRY(fmk_copy_h2d(ctx, t->ts, ts, (size_t)n * 8));
    FMK_TRY(fmk_copy_h2d(ctx, t->price, price, (size_t)n * 8));
    FMK_TRY(fmk_copy_h2d(ctx, t->amount, amount, (size_t)n * 8));
    if (side && t->side) FMK_TRY(fmk_copy_h2d(ctx, t->side, side, (size_t)n));
    if (t->log_price) { fmk_dfree(ctx, t->log_price); t->log_price = nullptr; }
    return FMK_OK;
}

int fmk_trades_upload(fmk_ctx *ctx, const int64_t *ts, const double *price, const double *amount, const int8_t *side,
                      int64_t n, fmk_trades **out) {
    FMK_ENTER(ctx);
    if (n < 0) return fmk_fail(ctx, FMK_ERR_ARG, "negative length");
    FMK_TRY(trades_alloc(ctx, n, ts != nullptr, side != nullptr, out));
    int rc = fmk_trades_refill(ctx, *out, ts, price, amount, side, n);
    if (rc) { fmk_trades_free(ctx, *out); *out = nullptr; return rc; }
    // pageable host memory: the copies above are staged synchronously by the runtime, so the caller's arrays can be
    // released on return; pinned memory callers must keep them alive until the next sync.
    return FMK_OK;
}

int fmk_trades_download(fmk_ctx *ctx, const fmk_trades *t, int64_t *ts, double *price, double *amount, int8_t *side) {
    FMK_ENTER(ctx);
    const size_t n = (size_t)t->n;
    if (ts && t->ts) FMK_TRY(fmk_copy_d2h(ctx, ts, t->ts, n * 8));
    if (price) FMK_TRY(fmk_copy_d2h(ctx, price, t->price, n * 8));
    if (amount) FMK_TRY(fmk_copy_d2h(ctx, amount, t->amount, n * 8));
    if (side && t->side) FMK_TRY(fmk_copy_d2h(ctx, side, t->side, n));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

}  // extern "C"

// float32 amounts (what the reference's split-trade merge yields, bar/data_model.py:326-344) travel over PCIe as 4 B/tick
// and are widened on the device: float -> double is exact, so every kernel sees the value Numba's float64 arithmetic sees.
__global__ void k_widen_f32(const float *__restrict__ in, int64_t n, double *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (double)in[i];
}

// out[k] = src[idx[k]] for 8-byte elements (e.g. CUSUMBarKit.get_sigma, bar/kit.py:176-181: sigma[bar_close_indices])
__global__ void k_gather8(const unsigned long long *__restrict__ src, const int64_t *__restrict__ idx, int64_t m, int64_t n,
                          unsigned long long *__restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    int64_t j = idx[k];
    if (j < 0) j += n;
    out[k] = (j >= 0 && j < n) ? src[j] : 0ull;
}

extern "C" {

int fmk_trades_upload_f32amt(fmk_ctx *ctx, const int64_t *ts, const double *price, const float *amount, const int8_t *side,
                             int64_t n, fmk_trades **out) {
    FMK_ENTER(ctx);
    if (n < 0) return fmk_fail(ctx, FMK_ERR_ARG, "negative length");
    FMK_TRY(trades_alloc(ctx, n, ts != nullptr, side != nullptr, out));
    fmk_trades *t = *out;
    auto body = [&]() -> int {
        if (n == 0) return FMK_OK;
        Scratch<float> f(ctx);
        FMK_TRY(f.alloc(n));
        if (ts) FMK_TRY(fmk_copy_h2d(ctx, t->ts, ts, (size_t)n * 8));
        FMK_TRY(fmk_copy_h2d(ctx, t->price, price, (size_t)n * 8));
        FMK_TRY(fmk_copy_h2d(ctx, f.p, amount, (size_t)n * 4));
        if (side) FMK_TRY(fmk_copy_h2d(ctx, t->side, side, (size_t)n));
        FMK_LAUNCH(ctx, k_widen_f32, ctx->sm_count * 8, 256, 0, (const float *)f.p, n, t->amount);
        return FMK_OK;
    };
    const int rc = body();
    if (rc) { fmk_trades_free(ctx, t); *out = nullptr; }
    return rc;
}

// Adds a column to a handle that was uploaded without it (which: 0 = timestamps int64, 1 = side int8): the wrappers upload
// lazily -- DollarBarKit.build_ohlcv needs neither, a later TBMLabel / build_directional_features on the same trades does.
int fmk_trades_add_column(fmk_ctx *ctx, fmk_trades *t, int which, const void *host) {
    FMK_ENTER(ctx);
    if (which == 0) {
        if (!t->ts) FMK_TRY(fmk_dalloc(ctx, &t->ts, t->n));
        FMK_TRY(fmk_copy_h2d(ctx, t->ts, host, (size_t)t->n * 8));
    } else if (which == 1) {
        if (!t->side) FMK_TRY(fmk_dalloc(ctx, &t->side, t->n));
        FMK_TRY(fmk_copy_h2d(ctx, t->side, host, (size_t)t->n));
    } else return fmk_fail(ctx, FMK_ERR_ARG, "which must be 0 (timestamps) or 1 (side)");
    return FMK_OK;
}

// Month-store loader (the reference keeps /trades/YYYY-MM tables, bar/data_model.py:420-574, and concatenates
// them on the host): one device SoA handle is allocated for the whole range and every month's columns are written straight
// into it at their offset -- no concatenated host frame is ever built.
int fmk_trades_alloc(fmk_ctx *ctx, int64_t n, int with_ts, int with_side, fmk_trades **out) {
    FMK_ENTER(ctx);
    if (n < 0) return fmk_fail(ctx, FMK_ERR_ARG, "negative length");
    return trades_alloc(ctx, n, with_ts, with_side, out);
}

int fmk_trades_write(fmk_ctx *ctx, fmk_trades *t, int64_t offset, int64_t count, const int64_t *ts, const double *price,
                     const void *amount, int amount_is_f32, const int8_t *side) {
    FMK_ENTER(ctx);
    if (offset < 0 || count < 0 || offset + count > t->n) return fmk_fail(ctx, FMK_ERR_ARG, "segment outside the handle");
    if (count == 0) return FMK_OK;
    if (ts && t->ts) FMK_TRY(fmk_copy_h2d(ctx, t->ts + offset, ts, (size_t)count * 8));
    if (price) FMK_TRY(fmk_copy_h2d(ctx, t->price + offset, price, (size_t)count * 8));
    if (amount) {
        if (amount_is_f32) {
            Scratch<float> f(ctx);
            FMK_TRY(f.alloc(count));
            FMK_TRY(fmk_copy_h2d(ctx, f.p, amount, (size_t)count * 4));
            FMK_LAUNCH(ctx, k_widen_f32, ctx->sm_count * 8, 256, 0, (const float *)f.p, count, t->amount + offset);
        } else FMK_TRY(fmk_copy_h2d(ctx, t->amount + offset, amount, (size_t)count * 8));
    }
    if (side && t->side) FMK_TRY(fmk_copy_h2d(ctx, t->side + offset, side, (size_t)count));
    if (t->log_price) { fmk_dfree(ctx, t->log_price); t->log_price = nullptr; }
    return FMK_OK;
}

int fmk_buf_gather8(fmk_ctx *ctx, const fmk_buf *src, const int64_t *idx, int64_t m, void *out) {
    FMK_ENTER(ctx);
    if (m <= 0) return FMK_OK;
    Scratch<int64_t> di(ctx);
    Scratch<unsigned long long> d(ctx);
    FMK_TRY(di.alloc(m)); FMK_TRY(d.alloc(m));
    FMK_CUDA(ctx, cudaMemcpyAsync(di.p, idx, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_LAUNCH(ctx, k_gather8, (unsigned)cdiv(m, 256), 256, 0, (const unsigned long long *)src->ptr, (const int64_t *)di.p, m,
               src->bytes / 8, d.p);
    FMK_CUDA(ctx, cudaMemcpyAsync(out, d.p, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

int64_t fmk_trades_size(const fmk_trades *t) { return t->n; }

void fmk_trades_free(fmk_ctx *ctx, fmk_trades *t) {
    FMK_ENTER(ctx);
    if (!t) return;
    fmk_dfree(ctx, t->ts);
    fmk_dfree(ctx, t->price);
    fmk_dfree(ctx, t->amount);
    fmk_dfree(ctx, t->side);
    fmk_dfree(ctx, t->log_price);
    delete t;
}

int64_t fmk_index_size(const fmk_index *ix) { return ix->m; }

int fmk_index_download(fmk_ctx *ctx, const fmk_index *ix, int64_t *close_ts, int64_t *close_idx) {
    FMK_ENTER(ctx);
    if (close_ts && ix->close_ts)
        FMK_CUDA(ctx, cudaMemcpyAsync(close_ts, ix->close_ts, (size_t)ix->m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (close_idx)
        FMK_CUDA(ctx, cudaMemcpyAsync(close_idx, ix->close_idx, (size_t)ix->m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

void fmk_index_free(fmk_ctx *ctx, fmk_index *ix) {
    FMK_ENTER(ctx);
    if (!ix) return;
    fmk_dfree(ctx, ix->close_ts);
    fmk_dfree(ctx, ix->close_idx);
    delete ix;
}

}  // extern "C"

// close_ts[k] = ts[close_idx[k]]  (bar/kit.py:67,101,135,173: `close_ts = timestamps[close_indices]`)
__global__ void k_gather_ts(const int64_t *ts, const int64_t *ci, int64_t m, int64_t n, int64_t *out) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) {
        int64_t j = ci[k];
        if (j < 0) j += n;
        out[k] = (j >= 0 && j < n) ? ts[j] : 0;
    }
}

int fmk_gather_close_ts(fmk_ctx *ctx, const fmk_trades *t, fmk_index *ix) {
    FMK_ENTER(ctx);
    if (!t->ts) return FMK_OK;   // timestamps were not uploaded: the host gathers ts[close_idx] itself
    if (!ix->close_ts) FMK_TRY(fmk_dalloc(ctx, &ix->close_ts, ix->m));
    if (ix->m > 0)
        FMK_LAUNCH(ctx, k_gather_ts, (unsigned)cdiv(ix->m, 256), 256, 0, t->ts, ix->close_idx, ix->m, t->n, ix->close_ts);
    return FMK_OK;
}

extern "C" int fmk_index_from_host(fmk_ctx *ctx, const fmk_trades *t, const int64_t *close_idx, int64_t m, fmk_index **out) {
    FMK_ENTER(ctx);
    *out = nullptr;
    fmk_index *ix = new (std::nothrow) fmk_index();
    if (!ix) return FMK_ERR_ALLOC;
    memset(ix, 0, sizeof(*ix));
    ix->m = m;
    ix->n_ticks = t->n;
    ix->sorted = 1;     // caller-provided indices: verify (the tiled OHLCV kernel needs monotone in-range indices)
    for (int64_t k = 0; k < m; k++)
        if (close_idx[k] < -1 || close_idx[k] >= t->n || (k > 0 && close_idx[k] < close_idx[k - 1])) { ix->sorted = 0; break; }
    int rc = fmk_dalloc(ctx, &ix->close_idx, m);
    if (rc) { delete ix; return rc; }
    if (m > 0) {
        cudaError_t e = cudaMemcpyAsync(ix->close_idx, close_idx, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { fmk_index_free(ctx, ix); return fmk_fail(ctx, FMK_ERR_CUDA, cudaGetErrorString(e)); }
    }
    rc = fmk_gather_close_ts(ctx, t, ix);
    if (rc) { fmk_index_free(ctx, ix); return rc; }
    *out = ix;
    return FMK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// a1: time bars (bar/logic.py:12-51)
// ---------------------------------------------------------------------------------------------------------------
// Python float floor division (numba lowers int64 // float64 to this).
static double py_floordiv(double vx, double wx) {
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    if (mod != 0.0 && ((wx < 0) != (mod < 0))) div -= 1.0;
    double fd;
    if (div != 0.0) {
        fd = floor(div);
        if (div - fd > 0.5) fd += 1.0;
    } else {
        fd = copysign(0.0, vx / wx);
    }
    return fd;
}

// clock[i] = int64(start + i*step) evaluated in float64 exactly like Numba's np.arange;
// idx[i] = searchsorted(ts, clock[i], 'right') - 1 with exact int64 compares.
__global__ void k_time_bar(const int64_t *__restrict__ ts, int64_t n, double start, double step, int64_t m,
                           int64_t *__restrict__ clock, int64_t *__restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double c = __dadd_rn(start, __dmul_rn((double)i, step));
    const int64_t key = (int64_t)c;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(ts + mid) <= key) lo = mid + 1; else hi = mid;
    }
    clock[i] = key;
    idx[i] = lo - 1;
}

extern "C" int fmk_time_bar_index(fmk_ctx *ctx, const fmk_trades *t, double interval_seconds, fmk_index **out) {
    FMK_ENTER(ctx);
    *out = nullptr;
    if (t->n <= 0) return fmk_fail(ctx, FMK_ERR_ARG, "empty trades");
    if (!t->ts) return fmk_fail(ctx, FMK_ERR_ARG, "time bars need the timestamp column on the device");
    if (!(interval_seconds > 0)) return fmk_fail(ctx, FMK_ERR_ARG, "interval must be positive");
    int64_t ends[2];
    FMK_CUDA(ctx, cudaMemcpyAsync(&ends[0], t->ts, 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(&ends[1], t->ts + (t->n - 1), 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // scalar clock set-up in float64, term by term as logic.py:30-39
    const double iv = interval_seconds * 1e9;
    const double start = py_floordiv((double)ends[0], iv) * iv;
    const double last = ceil((double)ends[1] / iv) * iv;
    const double stop = last + iv + 1.0;
    int64_t m = (int64_t)ceil((stop - start) / iv);
    if (m < 0) m = 0;
    fmk_index *ix = new (std::nothrow) fmk_index();
    if (!ix) return FMK_ERR_ALLOC;
    memset(ix, 0, sizeof(*ix));
    ix->m = m;
    ix->n_ticks = t->n;
    ix->sorted = 1;
    int rc = fmk_dalloc(ctx, &ix->close_ts, m);
    if (!rc) rc = fmk_dalloc(ctx, &ix->close_idx, m);
    if (rc) { fmk_index_free(ctx, ix); return rc; }
    if (m > 0) {
        auto launch = [&]() -> int {
            FMK_LAUNCH(ctx, k_time_bar, (unsigned)cdiv(m, 256), 256, 0, (const int64_t *)t->ts, t->n, start, iv, m, ix->close_ts, ix->close_idx);
            return FMK_OK;
        };
        rc = launch();
        if (rc) { fmk_index_free(ctx, ix); return rc; }
    }
    *out = ix;
    return FMK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// a2: tick bars (bar/logic.py:54-84) -- closed form of the counter recurrence:
//   threshold <= 1 : every index 0..n-1;   threshold >= 2 : {0} U {k*threshold - 1 : k >= 1}
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_tick_bar(int64_t m, int64_t thr, int64_t *idx) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    idx[k] = (thr <= 1) ? k : (k == 0 ? 0 : k * thr - 1);
}

extern "C" int fmk_tick_bar_index(fmk_ctx *ctx, const fmk_trades *t, int64_t threshold, fmk_index **out) {
    FMK_ENTER(ctx);
    *out = nullptr;
    const int64_t n = t->n;
    if (n <= 0) return fmk_fail(ctx, FMK_ERR_ARG, "empty trades");
    const int64_t m = threshold <= 1 ? n : 1 + n / threshold;
    fmk_index *ix = new (std::nothrow) fmk_index();
    if (!ix) return FMK_ERR_ALLOC;
    memset(ix, 0, sizeof(*ix));
    ix->m = m;
    ix->n_ticks = n;
    ix->sorted = 1;
    int rc = fmk_dalloc(ctx, &ix->close_idx, m);
    if (rc) { fmk_index_free(ctx, ix); return rc; }
    auto launch = [&]() -> int {
        FMK_LAUNCH(ctx, k_tick_bar, (unsigned)cdiv(m, 256), 256, 0, m, threshold, ix->close_idx);
        return FMK_OK;
    };
    rc = launch();
    if (!rc) rc = fmk_gather_close_ts(ctx, t, ix);
    if (rc) { fmk_index_free(ctx, ix); return rc; }
    *out = ix;
    return FMK_OK;
}

extern "C" int fmk_volume_bar_index(fmk_ctx *ctx, const fmk_trades *t, double threshold, fmk_index **out) {
    FMK_ENTER(ctx);
    return fmk_volume_index_impl(ctx, t, threshold, out);
}
extern "C" int fmk_dollar_bar_index(fmk_ctx *ctx, const fmk_trades *t, double threshold, fmk_index **out) {
    FMK_ENTER(ctx);
    return fmk_dollar_index_impl(ctx, t, threshold, out);
}
extern "C" int fmk_cusum_bar_index(fmk_ctx *ctx, const fmk_trades *t, fmk_buf *sigma, double sigma_floor,
                                   double sigma_mult, fmk_index **out) {
    FMK_ENTER(ctx);
    return fmk_cusum_index_impl(ctx, t, sigma, sigma_floor, sigma_mult, out);
}

// ---------------------------------------------------------------------------------------------------------------
// synthetic stream on the device (counter-based RNG; SURVEY 8d shape)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t seed, uint64_t i, uint64_t stream) {
    uint64_t r = mix64(seed * 0x9e3779b97f4a7c15ull + mix64(i * 4 + stream + 0x632be59bd9b4e019ull));
    return ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);  // (0,1)
}
__device__ __forceinline__ double normal01(uint64_t seed, uint64_t i, uint64_t s0) {
    double u1 = u01(seed, i, s0), u2 = u01(seed, i, s0 + 1);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

struct GapIn {
    uint64_t seed;
    __device__ int64_t operator()(int64_t i) const {
        return 1 + (int64_t)floor(-50e6 * log(u01(seed, (uint64_t)i, 0)));
    }
};
struct TsOut {
    int64_t *ts;
    __device__ void operator()(int64_t i, int64_t cs) const {
        int64_t t = 1700000000000000000ll + cs;
        ts[i] = t / 1000000 * 1000000;  // floor to ms -> duplicate timestamps
    }
};
struct RetIn {
    uint64_t seed;
    __device__ double operator()(int64_t i) const { return 2e-5 * normal01(seed, (uint64_t)i, 1); }
};
struct PxOut {
    double *px;
    __device__ void operator()(int64_t i, double cs) const { px[i] = rint(30000.0 * exp(cs) * 10.0) / 10.0; }
};
struct FlipIn {
    uint64_t seed;
    __device__ int64_t operator()(int64_t i) const { return u01(seed, (uint64_t)i, 3) < 0.3 ? 1 : 0; }
};
struct SideOut {
    int8_t *side;
    __device__ void operator()(int64_t i, int64_t cs) const { side[i] = (cs & 1) ? -1 : 1; }
};

__global__ void k_synth_amount(uint64_t seed, int64_t n, double *amt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // lognormal(-4, 1.2) + 0.001 rounded to 3 decimals; normal from an independent stream pair (4,5 -> use stream 2 base)
    double u1 = u01(seed, (uint64_t)i, 2), u2 = u01(seed ^ 0xabcdef1234567ull, (uint64_t)i, 2);
    double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    amt[i] = rint((exp(-4.0 + 1.2 * z) + 0.001) * 1000.0) / 1000.0;
}

extern "C" int fmk_trades_synth(fmk_ctx *ctx, int64_t n, uint64_t seed, fmk_trades **out) {
    FMK_ENTER(ctx);
    if (n <= 0) return fmk_fail(ctx, FMK_ERR_ARG, "n must be positive");
    FMK_TRY(trades_alloc(ctx, n, 1, 1, out));
    fmk_trades *t = *out;
    int rc = device_inclusive_scan<int64_t>(ctx, GapIn{seed}, TsOut{t->ts}, n, (int64_t *)nullptr);
    if (!rc) rc = device_inclusive_scan<double>(ctx, RetIn{seed}, PxOut{t->price}, n, (double *)nullptr);
    if (!rc) rc = device_inclusive_scan<int64_t>(ctx, FlipIn{seed}, SideOut{t->side}, n, (int64_t *)nullptr);
    if (!rc) {
        auto launch = [&]() -> int {
            FMK_LAUNCH(ctx, k_synth_amount, (unsigned)cdiv(n, 256), 256, 0, seed, n, t->amount);
            return FMK_OK;
        };
        rc = launch();
    }
    if (rc) { fmk_trades_free(ctx, t); *out = nullptr; }
    return rc;
}
