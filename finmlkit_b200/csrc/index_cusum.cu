// index_cusum.cu -- CUSUM bar indexer on the device (reference: finmlkit/bar/logic.py:152-221).
//
// Reference recurrence per tick i (after forward-filling sigma): r = log(p_i/p_{i-1});
//   s+ = max(0, s+ + r); s- = min(0, s- + r); if ts_i == ts_{i+1}: no test; lam = max(mult*sigma_i, floor);
//   if s+ >= lam: emit, s+ = 0  elif s- <= -lam: emit, s- = 0.
// The two accumulators are clamped at exactly 0.0 and reset to exactly 0.0, so trajectories started from different
// states COALESCE bit-for-bit once both have been clamped/reset at the same tick.  The pipeline exploits that:
//
//   C1  k_cusum_fill / k_cusum_prep : forward-fill sigma (scan of "last non-NaN"), then per tick r_i, lam_i and the
//                                     "may close" flag (fully parallel; the libm calls live here)
//   C2  k_cusum_tasks   : one lane per chunk; warm-up over the PREVIOUS chunk from the zero state, record the state
//                         reached at the chunk start (speculative), then replay the chunk marking closes in a bitmap
//   C3  k_cusum_check / k_cusum_tasks(work list) : parallel fix-point.  Chunk k is consistent when the state it started
//                         from is bit-identical to the end state of chunk k-1; every inconsistent chunk is replayed (all of
//                         them in parallel, one lane each) from its predecessor's current end state, and the check is
//                         repeated until nothing changes.  Chunk 0 starts from the true initial state, so by induction the
//                         fixed point is the sequential trajectory; the number of rounds is the longest run of chunks
//                         over which the state fails to coalesce (a handful on market data; n_chunks in the worst case)
//   C4  bitmap -> index list (popcount scan + ordered write)
// Given identical r_i the result is the reference's, bit for bit; r_i itself uses CUDA's log (<= 1 ulp from glibc's),
// which can only matter at an exact tie of a ~1e-19-wide band (documented in DESIGN.md).
#include <math.h>
#include <stdlib.h>
#include <new>
#include "common.cuh"
#include "scan.cuh"

// ---- forward fill: scan with the operator "right if right is not NaN else left" over (value) ---------------------
__device__ __forceinline__ double ff_nan() { return __longlong_as_double(0x7ff8000000000000ll); }
struct FF {
    double v;
    __device__ FF() { v = ff_nan(); }
    __device__ explicit FF(int) { v = ff_nan(); }   // identity (T(0))
    __device__ explicit FF(double x) { v = x; }
};
__device__ __forceinline__ FF operator+(const FF &a, const FF &b) { return (b.v == b.v) ? b : a; }
__device__ __forceinline__ FF __shfl_up_sync(unsigned m, const FF &x, int o) {
    FF r;
    r.v = ::__shfl_up_sync(m, x.v, o);
    return r;
}
struct FFIn {
    const double *s;
    __device__ FF operator()(int64_t i) const { return FF(s[i]); }
};
struct FFOut {
    double *s;
    int64_t first;
    __device__ void operator()(int64_t i, const FF &cs) const {
        if (i >= first) s[i] = cs.v;   // logic.py:186-189 fills from the first non-NaN index on
    }
};

__global__ void k_cusum_first(const double *__restrict__ sigma, int64_t n, unsigned long long *first) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // one atomic per warp at most, and none once a smaller index is already recorded
    const unsigned ok = __ballot_sync(0xffffffffu, i < n && sigma[i] == sigma[i]);
    if (ok && (threadIdx.x & 31) == (unsigned)(__ffs(ok) - 1) && (unsigned long long)i < *(volatile unsigned long long *)first)
        atomicMin(first, (unsigned long long)i);
}

__global__ void k_cusum_prep(const int64_t *__restrict__ ts, const double *__restrict__ p,
                             const double *__restrict__ sigma, int64_t n, double floor_, double mult,
                             double *__restrict__ r, double *__restrict__ lam, uint8_t *__restrict__ allowed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    r[i] = i > 0 ? log(__ddiv_rn(p[i], p[i - 1])) : 0.0;
    const double l = __dmul_rn(mult, sigma[i]);
    lam[i] = (floor_ > l) ? floor_ : l;            // python max(l, floor): floor only if floor > l (NaN l stays NaN)
    allowed[i] = !(i + 1 < n && ts[i] == ts[i + 1]);
}

// cusum_filter inputs: r_i = log(x_i / x_{i-1}), thr_i (constant or per element); every tick may fire
__global__ void k_cusum_filter_prep(const double *__restrict__ x, const double *__restrict__ thr, int64_t nthr, int64_t n,
                                    double *__restrict__ r, double *__restrict__ lam, uint8_t *__restrict__ allowed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    r[i] = i > 0 ? log(__ddiv_rn(x[i], x[i - 1])) : 0.0;
    lam[i] = nthr == 1 ? thr[0] : thr[i];
    allowed[i] = 1;
}

struct CusumState { double sp, sn; };

// one reference step; returns true when the bar closes / the event fires at this tick.
// MODE 0: _cusum_bar_indexer (bar/logic.py:203-218): s+ >= lam first, then s- <= -lam, only where `allowed`.
// MODE 1: cusum_filter (sampling/filters.py:54-67): s- < -thr first, then s+ > thr (strict), every tick.
template <int MODE>
__device__ __forceinline__ bool cusum_step(CusumState &s, double r, double lam, bool allowed) {
    const double a = __dadd_rn(s.sp, r), b = __dadd_rn(s.sn, r);
    s.sp = (a > 0.0) ? a : 0.0;      // python max(0.0, a): a only if a > 0.0 (NaN -> 0.0)
    s.sn = (b < 0.0) ? b : 0.0;
    if (MODE == 0) {
        if (!allowed) return false;
        if (s.sp >= lam) { s.sp = 0.0; return true; }
        if (s.sn <= -lam) { s.sn = 0.0; return true; }
    } else {
        if (s.sn < -lam) { s.sn = 0.0; return true; }
        if (s.sp > lam) { s.sp = 0.0; return true; }
    }
    return false;
}

constexpr int CT_WARPS = 4;
constexpr int CT_R = 16;

// lane = chunk.  Replay [warm_start, chunk_start) silently from the zero state, then [chunk_start, chunk_end) with
// closes recorded in the bitmap (chunk bounds are multiples of 32 ticks, so words are never shared between lanes).
template <int MODE>
__global__ void __launch_bounds__(CT_WARPS * 32) k_cusum_tasks(const double *__restrict__ r, const double *__restrict__ lam,
                                                               const uint8_t *__restrict__ allowed, int64_t n,
                                                               int64_t first, int64_t CH, int64_t nchunks,
                                                               unsigned *__restrict__ bitmap,
                                                               CusumState *__restrict__ spec_start,
                                                               CusumState *__restrict__ spec_end,
                                                               const int64_t *__restrict__ work, int64_t nwork,
                                                               const CusumState *__restrict__ prev_end, int tpw) {
    // tpw = tasks (lanes that replay a chunk) per warp.  A replay is one long dependent chain, so throughput comes from the
    // number of resident WARPS, not lanes: when there are few tasks (the repair rounds) they are spread one or two per warp.
    __shared__ double sr[CT_WARPS][32][CT_R + 1];
    __shared__ double sl[CT_WARPS][32][CT_R + 1];
    __shared__ uint8_t sa[CT_WARPS][32][CT_R];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t tid = ((int64_t)blockIdx.x * CT_WARPS + w) * tpw + lane;
    // work == nullptr: speculative pass over every chunk; otherwise replay of the listed chunks from prev_end[k-1]
    int64_t k = lane < tpw ? tid : nchunks;
    if (work) k = (lane < tpw && tid < nwork) ? work[tid] : nchunks;
    const int64_t lo = first + 1 + k * CH;              // first tick of the chunk (ticks <= first are never tested)
    int64_t hi = lo + CH;
    if (hi > n) hi = n;
    bool active = k < nchunks && lo < n;
    int64_t pos = lo - CH;                              // warm-up over the previous chunk
    if (pos < first + 1) pos = first + 1;
    CusumState s{0.0, 0.0};
    if (work && active) { pos = lo; s = prev_end[k - 1]; }   // listed chunks have k >= 1
    unsigned word = 0;
    const int half = lane >> 4, col = lane & 15;
    if (active && pos == lo) spec_start[k] = s;        // chunk 0: the true initial state
    while (__any_sync(0xffffffffu, active)) {
        // two phases: all 48 loads of the round are issued before the first shared-memory store waits on one of them
        // (interleaved load/store batches serialised four HBM round trips per 16 ticks -- 0.29 us per tick in the repair
        // rounds, where a single warp per SM has nothing else to overlap with)
        double ga[16], gb[16];
        uint8_t gc[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int row = 2 * q + half;
            const int64_t rp = __shfl_sync(0xffffffffu, pos, row);
            const int ra = __shfl_sync(0xffffffffu, (int)active, row);
            const int64_t idx = rp + col;
            ga[q] = 0.0; gb[q] = 0.0; gc[q] = 0;
            if (2 * q < tpw && ra && idx < n) { ga[q] = __ldg(r + idx); gb[q] = __ldg(lam + idx); gc[q] = __ldg(allowed + idx); }
        }
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int row = 2 * q + half;
            if (2 * q < tpw) { sr[w][row][col] = ga[q]; sl[w][row][col] = gb[q]; sa[w][row][col] = gc[q]; }
        }
        __syncwarp();
        if (active) {
#pragma unroll 1
            for (int tt = 0; tt < CT_R; tt++) {
                const int64_t i = pos + tt;
                if (i >= hi) { active = false; break; }
                if (i == lo) spec_start[k] = s;
                const bool close = cusum_step<MODE>(s, sr[w][lane][tt], sl[w][lane][tt], sa[w][lane][tt] != 0);
                if (i >= lo) {
                    const int64_t rel = i - (first + 1);
                    if (close) word |= 1u << (rel & 31);
                    if ((rel & 31) == 31 || i + 1 == hi) { bitmap[rel >> 5] = word; word = 0; }
                }
            }
            pos += CT_R;
            if (pos >= hi) active = false;
        }
        __syncwarp();
    }
    if (k < nchunks && lo < n) spec_end[k] = s;
}

// C3: consistency check of the chunk chain; inconsistent chunks are appended to the work list
__global__ void k_cusum_check(const CusumState *__restrict__ start, const CusumState *__restrict__ end, int64_t nchunks,
                              int64_t *__restrict__ work, unsigned long long *count) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (k >= nchunks) return;
    const CusumState a = start[k], b = end[k - 1];
    if (__double_as_longlong(a.sp) != __double_as_longlong(b.sp) || __double_as_longlong(a.sn) != __double_as_longlong(b.sn))
        work[atomicAdd(count, 1ull)] = k;
}
__global__ void k_cusum_commit(const int64_t *__restrict__ work, int64_t nwork, const CusumState *__restrict__ end_next,
                               CusumState *__restrict__ end) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nwork) end[work[t]] = end_next[work[t]];
}

// C4: bitmap -> ordered index list
struct PopIn {
    const unsigned *bm;
    __device__ int64_t operator()(int64_t w) const { return __popc(bm[w]); }
};
struct PopOut {
    const unsigned *bm;
    int64_t *out;
    int64_t base;     // tick index of bit 0
    __device__ void operator()(int64_t w, int64_t incl) const {
        unsigned x = bm[w];
        int64_t pos = incl - __popc(x) + 1;   // out[0] is the open marker
        while (x) {
            const int b = __ffs(x) - 1;
            out[pos++] = base + w * 32 + b;
            x &= x - 1;
        }
    }
};

struct CountOut {
    __device__ void operator()(int64_t, int64_t) const {}
};

__global__ void k_set_first(int64_t *out, int64_t v) { out[0] = v; }

// Shared driver of the chunk chain: speculative pass, parallel fix-point, bitmap -> ordered index list.
// idx_out[0] is left for the caller (open marker); idx_out[1..total] are the closing / event ticks.
#define CUSUM_TASKS(grid, block, smem, ...)                                                      \
    do {                                                                                         \
        if (mode == 0) FMK_LAUNCH(ctx, k_cusum_tasks<0>, grid, block, smem, __VA_ARGS__);        \
        else FMK_LAUNCH(ctx, k_cusum_tasks<1>, grid, block, smem, __VA_ARGS__);                  \
    } while (0)

static int cusum_chain(fmk_ctx *ctx, const Scratch<double> &r, const Scratch<double> &lam, const Scratch<uint8_t> &allowed,
                       int64_t n, int64_t first, int mode, int64_t **idx_out, int64_t *total_out) {
    const int64_t m_ticks = n - (first + 1);   // ticks that can close a bar
    int64_t total = 0;
    int64_t *idx = nullptr;
    if (m_ticks > 0) {
        // ~64k chunks keep every SM busy in the speculative pass and make a repair round cheap; chunks are multiples of 32
        // ticks so that bitmap words are never shared between lanes
        int64_t CH = cdiv(m_ticks, 65536);
        if (CH < 4096) CH = 4096;
        if (const char *e = getenv("FMK_CUSUM_CH")) CH = atoll(e) > 0 ? atoll(e) : CH;   // test hook: tiny chunks, many rounds
        CH = cdiv(CH, 32) * 32;
        const int64_t nchunks = cdiv(m_ticks, CH);
        const int64_t nwords = cdiv(m_ticks, 32);
        Scratch<unsigned> bitmap(ctx);
        Scratch<CusumState> ss(ctx), se(ctx), se_next(ctx);
        Scratch<int64_t> work(ctx), dtotal(ctx);
        Scratch<unsigned long long> dcount(ctx);
        FMK_TRY(bitmap.alloc(nwords)); FMK_TRY(ss.alloc(nchunks)); FMK_TRY(se.alloc(nchunks)); FMK_TRY(se_next.alloc(nchunks));
        FMK_TRY(work.alloc(nchunks)); FMK_TRY(dcount.alloc(1));
        FMK_TRY(dtotal.alloc(1));
        // tasks per warp: aim at >= 16 resident warps per SM
        // (measured at 1e9 ticks: spreading the few tasks of a repair round over more warps -- 2 per warp instead of 32 --
        //  is slower, 242 vs 199 ms for the whole chain: the per-warp staging cost dominates.  Kept at 32.)
        auto pick_tpw = [&](int64_t) { return 32; };
        int tpw = pick_tpw(nchunks);
        CUSUM_TASKS((unsigned)cdiv(cdiv(nchunks, tpw), CT_WARPS), CT_WARPS * 32, 0, (const double *)r.p,
                    (const double *)lam.p, (const uint8_t *)allowed.p, n, first, CH, nchunks, bitmap.p, ss.p, se.p,
                    (const int64_t *)nullptr, (int64_t)0, (const CusumState *)nullptr, tpw);
        int64_t hrep = 0, rounds = 0;
        for (;;) {
            unsigned long long hcount = 0;
            FMK_CUDA(ctx, cudaMemsetAsync(dcount.p, 0, 8, ctx->stream));
            if (nchunks > 1)
                FMK_LAUNCH(ctx, k_cusum_check, (unsigned)cdiv(nchunks - 1, 256), 256, 0, (const CusumState *)ss.p,
                           (const CusumState *)se.p, nchunks, work.p, dcount.p);
            FMK_CUDA(ctx, cudaMemcpyAsync(&hcount, dcount.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
            FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (hcount == 0) break;
            const int64_t nw = (int64_t)hcount;
            tpw = pick_tpw(nw);
            CUSUM_TASKS((unsigned)cdiv(cdiv(nw, tpw), CT_WARPS), CT_WARPS * 32, 0, (const double *)r.p,
                        (const double *)lam.p, (const uint8_t *)allowed.p, n, first, CH, nchunks, bitmap.p, ss.p, se_next.p,
                        (const int64_t *)work.p, nw, (const CusumState *)se.p, tpw);
            FMK_LAUNCH(ctx, k_cusum_commit, (unsigned)cdiv(nw, 256), 256, 0, (const int64_t *)work.p, nw,
                       (const CusumState *)se_next.p, se.p);
            hrep += nw; rounds++;
        }
        // count, allocate, write
        Scratch<int64_t> wsum(ctx);
        FMK_TRY(wsum.alloc(1));
        FMK_TRY((device_inclusive_scan<int64_t>(ctx, PopIn{bitmap.p}, CountOut{}, nwords, dtotal.p)));
        FMK_CUDA(ctx, cudaMemcpyAsync(&total, dtotal.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->stats[0] = nchunks; ctx->stats[1] = hrep; ctx->stats[2] = rounds;
        FMK_TRY(fmk_dalloc(ctx, &idx, total + 1));
        int rc = device_inclusive_scan<int64_t>(ctx, PopIn{bitmap.p}, PopOut{bitmap.p, idx, first + 1}, nwords, (int64_t *)nullptr);
        if (rc) { fmk_dfree(ctx, idx); return rc; }
    } else {
        FMK_TRY(fmk_dalloc(ctx, &idx, 1));
    }
    *idx_out = idx;
    *total_out = total;
    return FMK_OK;
}

int fmk_cusum_index_impl(fmk_ctx *ctx, const fmk_trades *t, fmk_buf *sigma, double sigma_floor, double sigma_mult,
                         fmk_index **out_ix) {
    FMK_ENTER(ctx);
    *out_ix = nullptr;
    const int64_t n = t->n;
    if (n <= 0) return fmk_fail(ctx, FMK_ERR_ARG, "empty trades");
    if (!t->ts) return fmk_fail(ctx, FMK_ERR_ARG, "CUSUM bars need the timestamp column on the device");
    if (sigma->bytes < n * 8) return fmk_fail(ctx, FMK_ERR_ARG, "Prices, timestamps, and sigma arrays must have the same length.");
    double *sg = (double *)sigma->ptr;
    ctx->stats[0] = ctx->stats[1] = ctx->stats[2] = 0;

    // first non-NaN sigma (logic.py:175-179; 0 when every element is NaN)
    Scratch<unsigned long long> dfirst(ctx);
    FMK_TRY(dfirst.alloc(1));
    FMK_CUDA(ctx, cudaMemsetAsync(dfirst.p, 0xff, 8, ctx->stream));
    FMK_LAUNCH(ctx, k_cusum_first, (unsigned)cdiv(n, 256), 256, 0, (const double *)sg, n, dfirst.p);
    unsigned long long hfirst = 0;
    FMK_CUDA(ctx, cudaMemcpyAsync(&hfirst, dfirst.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t first = hfirst == ~0ull ? 0 : (int64_t)hfirst;
    if (hfirst != ~0ull) FMK_TRY((device_inclusive_scan<FF>(ctx, FFIn{sg}, FFOut{sg, first}, n, (FF *)nullptr)));

    Scratch<double> r(ctx), lam(ctx);
    Scratch<uint8_t> allowed(ctx);
    FMK_TRY(r.alloc(n)); FMK_TRY(lam.alloc(n)); FMK_TRY(allowed.alloc(n));
    FMK_LAUNCH(ctx, k_cusum_prep, (unsigned)cdiv(n, 256), 256, 0, (const int64_t *)t->ts, (const double *)t->price,
               (const double *)sg, n, sigma_floor, sigma_mult, r.p, lam.p, allowed.p);

    int64_t total = 0;
    int64_t *idx = nullptr;
    FMK_TRY(cusum_chain(ctx, r, lam, allowed, n, first, 0, &idx, &total));
    {
        auto launch = [&]() -> int { FMK_LAUNCH(ctx, k_set_first, 1, 1, 0, idx, first); return FMK_OK; };
        const int lrc = launch();
        if (lrc) { fmk_dfree(ctx, idx); return lrc; }
    }
    fmk_index *ix = new (std::nothrow) fmk_index();
    if (!ix) { fmk_dfree(ctx, idx); return FMK_ERR_ALLOC; }
    memset(ix, 0, sizeof(*ix));
    ix->m = total + 1;
    ix->n_ticks = n;
    ix->sorted = 1;
    ix->close_idx = idx;
    int rc = fmk_gather_close_ts(ctx, t, ix);
    if (rc) { fmk_index_free(ctx, ix); return rc; }
    *out_ix = ix;
    return FMK_OK;
}


// cusum_filter (sampling/filters.py:6-70): host series / thresholds in, device buffer of int64 event indices out.
extern "C" int fmk_cusum_filter(fmk_ctx *ctx, const double *series, int64_t n, const double *threshold, int64_t n_thr,
                                fmk_buf **events_out, int64_t *n_events) {
    FMK_ENTER(ctx);
    *events_out = nullptr; *n_events = 0;
    if (n <= 1) return fmk_fail(ctx, FMK_ERR_ARG, "Input time series must have at least 2 elements.");
    if (n_thr != 1 && n_thr != n)
        return fmk_fail(ctx, FMK_ERR_ARG, "Threshold array must either contain 1 const. element or len(raw_time_series) elements.");
    Scratch<double> x(ctx), th(ctx), r(ctx), lam(ctx);
    Scratch<uint8_t> allowed(ctx);
    FMK_TRY(x.alloc(n)); FMK_TRY(th.alloc(n_thr)); FMK_TRY(r.alloc(n)); FMK_TRY(lam.alloc(n)); FMK_TRY(allowed.alloc(n));
    FMK_CUDA(ctx, cudaMemcpyAsync(x.p, series, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(th.p, threshold, (size_t)n_thr * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_LAUNCH(ctx, k_cusum_filter_prep, (unsigned)cdiv(n, 256), 256, 0, (const double *)x.p, (const double *)th.p, n_thr, n,
               r.p, lam.p, allowed.p);
    int64_t *idx = nullptr, total = 0;
    FMK_TRY(cusum_chain(ctx, r, lam, allowed, n, 0, 1, &idx, &total));
    fmk_buf *b = new (std::nothrow) fmk_buf();
    if (!b) { fmk_dfree(ctx, idx); return FMK_ERR_ALLOC; }
    // hand out the events without the leading marker slot: copy down by one element (stream ordered)
    b->bytes = total * 8;
    int rc = fmk_dalloc(ctx, (int64_t **)&b->ptr, total);
    if (rc) { delete b; fmk_dfree(ctx, idx); return rc; }
    if (total > 0) FMK_CUDA(ctx, cudaMemcpyAsync(b->ptr, idx + 1, (size_t)total * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    fmk_dfree(ctx, idx);
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *events_out = b;
    *n_events = total;
    return FMK_OK;
}
