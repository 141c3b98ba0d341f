// series.cu -- tick-level series on the device: comp_lagged_returns (feature/core/utils.py:12-64), ewmst
// (feature/core/volatility.py:139-219) and triple_barrier (label/tbm.py:11-158).
#include <math.h>
#include <new>
#include "common.cuh"

#define FULL 0xffffffffu

// np.searchsorted(int64 ts, float64 key, 'right') with Numba's promotion of both sides to float64
__device__ __forceinline__ int64_t ss_right_f64(const int64_t *__restrict__ ts, int64_t lo, int64_t hi, double key) {
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if ((double)__ldg(ts + mid) <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------------------------
// a13: lagged returns.  lag(i) is monotone in i, so the first and last tick of a block bracket every other tick's
// answer.  The brackets come from a pre-pass with one thread per bracket (k_lagged_brackets), and the bracketed slice of
// timestamps is staged in shared memory (as float64) so the per-tick searches are branch-free and never touch global
// memory.  (r01: two full binary searches by two threads of every 256-tick block stalled the block for ~15 us -> 44 ms at
// 1e9 ticks; a block-cooperative 256-way search cut the latency but issued 8x more scattered probes -> 37 ms.)
// ---------------------------------------------------------------------------------------------------------------
constexpr int LR_THREADS = 256;
constexpr int LR_ITEMS = 4;                       // ticks per thread
constexpr int LR_TILE = LR_THREADS * LR_ITEMS;    // ticks per block
constexpr int LR_STAGE = 4096;

// Bracket pre-pass: one THREAD per block bracket (two per 1024-tick block).  A full binary search is ~30 dependent loads,
// but 2 n / 1024 of them run concurrently, so the latency is hidden by occupancy instead of stalling a whole block.
__global__ void k_lagged_brackets(const int64_t *__restrict__ ts, int64_t n, double w, int64_t nblk,
                                  int64_t *__restrict__ br) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= 2 * nblk) return;
    const int64_t blk = q >> 1;
    int64_t i = blk * LR_TILE + ((q & 1) ? LR_TILE - 1 : 0);
    if (i > n - 1) i = n - 1;
    br[q] = ss_right_f64(ts, 0, i + 1, __dadd_rn((double)ts[i], -w));
}

__global__ void __launch_bounds__(LR_THREADS) k_lagged_returns(const int64_t *__restrict__ ts,
                                                               const double *__restrict__ close, int64_t n, double w,
                                                               int is_log, const int64_t *__restrict__ br,
                                                               double *__restrict__ out) {
    __shared__ double ts_s[LR_STAGE];          // staged as float64: the comparisons are float64, convert once
    const int64_t i0 = (int64_t)blockIdx.x * LR_TILE;
    const int64_t b0 = br[2 * (int64_t)blockIdx.x], b1 = br[2 * (int64_t)blockIdx.x + 1];
    const int64_t span = b1 - b0;                   // per-tick insertion points lie in [b0, b1]: they read ts[b0 .. b1 - 1]
    // the halving steps below sum to LR_STAGE - 1, so a staged search can advance at most LR_STAGE - 1 positions: a span of
    // exactly LR_STAGE (insertion point b1 = b0 + LR_STAGE for the block's last tick) must take the global-memory search
    const bool staged = span < LR_STAGE;
    if (staged)
        for (int64_t q = threadIdx.x; q < span; q += LR_THREADS) ts_s[q] = (double)__ldg(ts + b0 + q);
    __syncthreads();
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    // start_idx = searchsorted(ts, ts[0] + w, 'left'):  i < start_idx  <=>  (double)ts[i] < (double)ts[0] + w
    const double first_key = __dadd_rn((double)__ldg(ts), w);
#pragma unroll
    for (int k = 0; k < LR_ITEMS; k++) {
        const int64_t i = i0 + threadIdx.x + (int64_t)k * LR_THREADS;
        if (i >= n) break;
        const int64_t tsi = ts[i];
        if ((double)tsi < first_key) { out[i] = nan; continue; }
        const double target = __dadd_rn((double)tsi, -w);
        int64_t hi = b1;
        if (hi > i + 1) hi = i + 1;
        int64_t lo = b0;
        if (lo > hi) lo = hi;
        // the bracket holds searchsorted results (insertion points); the answer lies in [lo, hi]
        int64_t ip;
        if (staged) {
            // branch-free: number of staged timestamps in [lo, hi) that are <= target (they form a prefix)
            int l = (int)(lo - b0);
            const int h = (int)(hi - b0);
#pragma unroll
            for (int step = LR_STAGE / 2; step > 0; step >>= 1)
                if (l + step <= h && ts_s[l + step - 1] <= target) l += step;
            ip = b0 + l;
        } else ip = ss_right_f64(ts, lo, hi, target);
        const int64_t lag = ip - 1;
        double r = nan;
        if (lag >= 0 && lag < i) {
            const double cl = close[lag];
            if (cl != 0.0) {
                const double q = __ddiv_rn(close[i], cl);
                r = is_log ? log(q) : __dadd_rn(q, -1.0);
            } else r = __longlong_as_double(0x7ff0000000000000ll);  // +inf
        }
        out[i] = r;
    }
}

static int run_lagged_returns(fmk_ctx *ctx, const int64_t *ts, const double *close, int64_t n, double window_sec,
                              int is_log, double *out) {
    if (!(window_sec > 0)) return fmk_fail(ctx, FMK_ERR_ARG, "The return window must be greater than zero.");
    if (n <= 0) return FMK_OK;
    const int64_t nblk = cdiv(n, LR_TILE);
    Scratch<int64_t> br(ctx);
    FMK_TRY(br.alloc(2 * nblk));
    FMK_LAUNCH(ctx, k_lagged_brackets, (unsigned)cdiv(2 * nblk, 256), 256, 0, ts, n, window_sec * 1e9, nblk, br.p);
    FMK_LAUNCH(ctx, k_lagged_returns, (unsigned)nblk, LR_THREADS, 0, ts, close, n, window_sec * 1e9, is_log,
               (const int64_t *)br.p, out);
    return FMK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// a14: ewmst.  The four accumulators obey first-order linear recurrences with a shared decay:
//     V' = om V + a ; V2' = om^2 V2 + a^2 ; Sy' = om Sy + a y ; Syy' = om Syy + a y y      (om = 1 - a)
// i.e. each tick is an affine map, and affine maps compose associatively -> chunked scan (reduce, scan of chunk
// composites, apply).  Within a thread the ticks are applied in order with the reference's expression order.
// ---------------------------------------------------------------------------------------------------------------
struct Ewm { double A, A2, bV, bV2, bSy, bSyy; };
__device__ __forceinline__ Ewm ewm_identity() { return Ewm{1.0, 1.0, 0.0, 0.0, 0.0, 0.0}; }
// apply f first, then g
__device__ __forceinline__ Ewm ewm_compose(const Ewm &f, const Ewm &g) {
    Ewm h;
    h.A = g.A * f.A;
    h.A2 = g.A2 * f.A2;
    h.bV = g.A * f.bV + g.bV;
    h.bV2 = g.A2 * f.bV2 + g.bV2;
    h.bSy = g.A * f.bSy + g.bSy;
    h.bSyy = g.A * f.bSyy + g.bSyy;
    return h;
}
__device__ __forceinline__ Ewm ewm_shfl_up(const Ewm &x, int o) {
    Ewm y;
    y.A = __shfl_up_sync(FULL, x.A, o); y.A2 = __shfl_up_sync(FULL, x.A2, o);
    y.bV = __shfl_up_sync(FULL, x.bV, o); y.bV2 = __shfl_up_sync(FULL, x.bV2, o);
    y.bSy = __shfl_up_sync(FULL, x.bSy, o); y.bSyy = __shfl_up_sync(FULL, x.bSyy, o);
    return y;
}

constexpr int EW_THREADS = 256;
constexpr int EW_ITEMS = 8;
constexpr int EW_TILE = EW_THREADS * EW_ITEMS;

struct EwTick { double alpha, om, y; };
__device__ __forceinline__ EwTick ew_tick(const int64_t *__restrict__ ts, const double *__restrict__ y, int64_t i,
                                          double half_life) {
    // i >= 1.  dt = (ts[i] - ts[i-1]) / 1e9 ; alpha = 1 - exp(-dt / half_life)
    const double dt = __ddiv_rn((double)(ts[i] - ts[i - 1]), 1e9);
    const double alpha = __dadd_rn(1.0, -exp(__ddiv_rn(-dt, half_life)));
    return EwTick{alpha, __dadd_rn(1.0, -alpha), y[i]};
}
__device__ __forceinline__ void ew_step(double &V, double &V2, double &Sy, double &Syy, const EwTick &t) {
    V = t.alpha + t.om * V;
    V2 = t.alpha * t.alpha + (t.om * t.om) * V2;
    if (t.y != t.y) { Sy = t.om * Sy; Syy = t.om * Syy; }
    else { Sy = t.alpha * t.y + t.om * Sy; Syy = t.alpha * t.y * t.y + t.om * Syy; }
}
__device__ __forceinline__ Ewm ew_elem(const EwTick &t) {
    Ewm e;
    e.A = t.om; e.A2 = t.om * t.om; e.bV = t.alpha; e.bV2 = t.alpha * t.alpha;
    if (t.y != t.y) { e.bSy = 0.0; e.bSyy = 0.0; }
    else { e.bSy = t.alpha * t.y; e.bSyy = t.alpha * t.y * t.y; }
    return e;
}

// composite of the thread's EW_ITEMS ticks (ticks i = 1..n-1 only; tick 0 is the identity)
// alpha_io: the reduce pass stores every tick's alpha there and the apply pass reads it back instead of evaluating exp and the
// two IEEE divisions a second time (8 B/tick more traffic on an fp64-compute-bound kernel; same bits either way)
template <bool CACHED>
__device__ __forceinline__ Ewm ew_thread_composite(const int64_t *ts, const double *y, int64_t base, int64_t n,
                                                   double half_life, double *alpha_io, EwTick (&ticks)[EW_ITEMS]) {
    Ewm f = ewm_identity();
#pragma unroll
    for (int k = 0; k < EW_ITEMS; k++) {
        const int64_t i = base + k;
        if (i >= 1 && i < n) {
            EwTick t;
            if (CACHED) {
                const double alpha = alpha_io[i];
                t = EwTick{alpha, __dadd_rn(1.0, -alpha), y[i]};
            } else {
                t = ew_tick(ts, y, i, half_life);
                alpha_io[i] = t.alpha;
            }
            ticks[k] = t;
            f = ewm_compose(f, ew_elem(t));
        }
    }
    return f;
}

// block-wide exclusive scan of composites; returns the composite of all earlier threads, block total in *tot
__device__ __forceinline__ Ewm ew_block_excl(const Ewm &x, Ewm *tot) {
    __shared__ Ewm wtot[EW_THREADS / 32];
    __shared__ Ewm btot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    Ewm inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Ewm yv = ewm_shfl_up(inc, o);
        if (lane >= o) inc = ewm_compose(yv, inc);
    }
    if (lane == 31) wtot[w] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        Ewm run = ewm_identity();
        for (int k = 0; k < EW_THREADS / 32; k++) {
            Ewm t = wtot[k];
            wtot[k] = run;
            run = ewm_compose(run, t);
        }
        btot = run;
    }
    __syncthreads();
    Ewm ex = ewm_shfl_up(inc, 1);
    if (lane == 0) ex = ewm_identity();
    Ewm r = ewm_compose(wtot[w], ex);
    *tot = btot;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(EW_THREADS) k_ewm_reduce(const int64_t *__restrict__ ts, const double *__restrict__ y,
                                                           int64_t n, double half_life, Ewm *tile_f, double *alpha_out) {
    const int64_t base = (int64_t)blockIdx.x * EW_TILE + (int64_t)threadIdx.x * EW_ITEMS;
    EwTick ticks[EW_ITEMS];
    Ewm f = ew_thread_composite<false>(ts, y, base, n, half_life, alpha_out, ticks);
    Ewm tot;
    ew_block_excl(f, &tot);
    if (threadIdx.x == 0) tile_f[blockIdx.x] = tot;
}

// Exclusive scan of the tile composites in three small grid-wide steps (1e9 ticks are 488k tiles; a single block walking them
// took 4 - 6 ms): per group of EW_THREADS tiles a block scan that also emits the group composite, one block scanning the
// group composites, and a pass that prepends the group prefix.
__global__ void __launch_bounds__(EW_THREADS) k_ewm_tiles_local(Ewm *tile_f, int64_t ntiles, Ewm *group_f) {
    const int64_t k = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x;
    const Ewm x = k < ntiles ? tile_f[k] : ewm_identity();
    Ewm tot;
    const Ewm ex = ew_block_excl(x, &tot);
    if (k < ntiles) tile_f[k] = ex;
    if (threadIdx.x == 0) group_f[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(EW_THREADS) k_ewm_tiles_top(Ewm *group_f, int64_t ngroups) {
    __shared__ Ewm carry;
    if (threadIdx.x == 0) carry = ewm_identity();
    __syncthreads();
    for (int64_t b = 0; b < ngroups; b += EW_THREADS) {
        const int64_t k = b + threadIdx.x;
        Ewm x = k < ngroups ? group_f[k] : ewm_identity();
        Ewm tot;
        Ewm ex = ew_block_excl(x, &tot);
        Ewm c = carry;
        if (k < ngroups) group_f[k] = ewm_compose(c, ex);
        __syncthreads();
        if (threadIdx.x == 0) carry = ewm_compose(c, tot);
        __syncthreads();
    }
}
__global__ void __launch_bounds__(EW_THREADS) k_ewm_tiles_add(Ewm *tile_f, int64_t ntiles, const Ewm *__restrict__ group_f) {
    const int64_t k = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x;
    if (k < ntiles && blockIdx.x > 0) tile_f[k] = ewm_compose(group_f[blockIdx.x], tile_f[k]);
}

__global__ void __launch_bounds__(EW_THREADS) k_ewm_apply(const int64_t *__restrict__ ts, const double *__restrict__ y,
                                                          int64_t n, double half_life, double sigma_floor,
                                                          const Ewm *__restrict__ tile_f, double *__restrict__ out,
                                                          double *alpha_in) {
    const int64_t base = (int64_t)blockIdx.x * EW_TILE + (int64_t)threadIdx.x * EW_ITEMS;
    EwTick ticks[EW_ITEMS];
    Ewm f = ew_thread_composite<true>(ts, y, base, n, half_life, alpha_in, ticks);
    Ewm tot;
    Ewm ex = ew_block_excl(f, &tot);
    Ewm pre = ewm_compose(tile_f[blockIdx.x], ex);
    // state before this thread's first tick (initial state is all zeros, so only the offsets matter)
    double V = pre.bV, V2 = pre.bV2, Sy = pre.bSy, Syy = pre.bSyy;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
    for (int k = 0; k < EW_ITEMS; k++) {
        const int64_t i = base + k;
        if (i >= n) break;
        if (i == 0) { out[0] = nan; continue; }
        ew_step(V, V2, Sy, Syy, ticks[k]);
        double o;
        if (V > 0.0) {
            const double mean = __ddiv_rn(Sy, V), e2 = __ddiv_rn(Syy, V);
            const double var_raw = e2 - mean * mean;
            const double denom = V - __ddiv_rn(V2, V);
            double var = 0.0;
            if (denom > 0.0 && var_raw > 0.0) var = var_raw * __ddiv_rn(V, denom);
            double sg = __dsqrt_rn(var);
            if (sg < sigma_floor) sg = sigma_floor;
            o = sg;
        } else o = nan;
        out[i] = o;
    }
}

static int run_ewmst(fmk_ctx *ctx, const int64_t *ts, const double *y, int64_t n, double half_life, double sigma_floor,
                     double *out) {
    if (n <= 0) return FMK_OK;
    const int64_t ntiles = cdiv(n, EW_TILE);
    Scratch<Ewm> tiles(ctx);
    FMK_TRY(tiles.alloc(ntiles));
    Scratch<double> alpha(ctx);
    FMK_TRY(alpha.alloc(n));
    FMK_LAUNCH(ctx, k_ewm_reduce, (unsigned)ntiles, EW_THREADS, 0, ts, y, n, half_life, tiles.p, alpha.p);
    const int64_t ngroups = cdiv(ntiles, EW_THREADS);
    Scratch<Ewm> groups(ctx);
    FMK_TRY(groups.alloc(ngroups));
    FMK_LAUNCH(ctx, k_ewm_tiles_local, (unsigned)ngroups, EW_THREADS, 0, tiles.p, ntiles, groups.p);
    FMK_LAUNCH(ctx, k_ewm_tiles_top, 1, EW_THREADS, 0, groups.p, ngroups);
    FMK_LAUNCH(ctx, k_ewm_tiles_add, (unsigned)ngroups, EW_THREADS, 0, tiles.p, ntiles, (const Ewm *)groups.p);
    FMK_LAUNCH(ctx, k_ewm_apply, (unsigned)ntiles, EW_THREADS, 0, ts, y, n, half_life, sigma_floor, (const Ewm *)tiles.p, out, alpha.p);
    return FMK_OK;
}

extern "C" int fmk_lagged_returns(fmk_ctx *ctx, const int64_t *ts, const double *close, int64_t n, double window_sec,
                                  int is_log, double *out) {
    FMK_ENTER(ctx);
    if (!(window_sec > 0)) return fmk_fail(ctx, FMK_ERR_ARG, "The return window must be greater than zero.");
    Scratch<int64_t> dts(ctx);
    Scratch<double> dc(ctx), dout(ctx);
    FMK_TRY(dts.alloc(n)); FMK_TRY(dc.alloc(n)); FMK_TRY(dout.alloc(n));
    if (n > 0) {
        FMK_TRY(fmk_copy_h2d(ctx, dts.p, ts, (size_t)n * 8));
        FMK_TRY(fmk_copy_h2d(ctx, dc.p, close, (size_t)n * 8));
        FMK_TRY(run_lagged_returns(ctx, dts.p, dc.p, n, window_sec, is_log, dout.p));
        FMK_TRY(fmk_copy_d2h(ctx, out, dout.p, (size_t)n * 8));
    }
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

extern "C" int fmk_ewmst(fmk_ctx *ctx, const int64_t *ts, const double *y, int64_t n, double half_life,
                         double sigma_floor, double *out) {
    FMK_ENTER(ctx);
    Scratch<int64_t> dts(ctx);
    Scratch<double> dy(ctx), dout(ctx);
    FMK_TRY(dts.alloc(n)); FMK_TRY(dy.alloc(n)); FMK_TRY(dout.alloc(n));
    if (n > 0) {
        FMK_TRY(fmk_copy_h2d(ctx, dts.p, ts, (size_t)n * 8));
        FMK_TRY(fmk_copy_h2d(ctx, dy.p, y, (size_t)n * 8));
        FMK_TRY(run_ewmst(ctx, dts.p, dy.p, n, half_life, sigma_floor, dout.p));
        FMK_TRY(fmk_copy_d2h(ctx, out, dout.p, (size_t)n * 8));
    }
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

extern "C" int fmk_lagged_returns_dev(fmk_ctx *ctx, const fmk_trades *t, double window_sec, int is_log, fmk_buf **out) {
    FMK_ENTER(ctx);
    if (!t->ts) return fmk_fail(ctx, FMK_ERR_ARG, "lagged returns need the timestamp column on the device");
    FMK_TRY(fmk_buf_alloc(ctx, t->n * 8, out));
    int rc = run_lagged_returns(ctx, t->ts, t->price, t->n, window_sec, is_log, (double *)(*out)->ptr);
    if (rc) { fmk_buf_free(ctx, *out); *out = nullptr; }
    return rc;
}

extern "C" int fmk_ewmst_dev(fmk_ctx *ctx, const fmk_trades *t, const fmk_buf *y, double half_life, double sigma_floor,
                             fmk_buf **out) {
    FMK_ENTER(ctx);
    if (!t->ts) return fmk_fail(ctx, FMK_ERR_ARG, "ewmst needs the timestamp column on the device");
    if (y->bytes < t->n * 8) return fmk_fail(ctx, FMK_ERR_ARG, "y is shorter than the trades");
    FMK_TRY(fmk_buf_alloc(ctx, t->n * 8, out));
    int rc = run_ewmst(ctx, t->ts, (const double *)y->ptr, t->n, half_life, sigma_floor, (double *)(*out)->ptr);
    if (rc) { fmk_buf_free(ctx, *out); *out = nullptr; }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// a16: triple_barrier.  One warp per event; the forward path is scanned 32 ticks per step with a ballot for the
// first touch, so the early `break` of the reference becomes "lowest set lane".  log(close) is computed once per
// trades handle (the reference precomputes np.log(close), tbm.py:69).
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_log(const double *__restrict__ x, int64_t n, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = log(x[i]);
}

__global__ void __launch_bounds__(256) k_triple_barrier(const int64_t *__restrict__ ts, const double *__restrict__ lc,
                                                        int64_t n, const int64_t *__restrict__ ev,
                                                        const double *__restrict__ tg, int64_t ne, double bottom,
                                                        double top, double vert_ns, double minc_ns,
                                                        const int8_t *__restrict__ side, double min_ret,
                                                        int8_t *labels, int64_t *touch_idx, double *rets, double *ratios) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < ne; e += nwarps) {
        const int64_t t0i = ev[e];
        const double tgt = tg[e];
        const double upper = __dmul_rn(tgt, top), lower = __dmul_rn(-tgt, bottom);
        const bool uv = isfinite(upper) && upper != 0.0, lv = isfinite(lower) && lower != 0.0;
        const int64_t t0 = ts[t0i];
        const double t1 = __dadd_rn((double)t0, vert_ns);
        const int64_t t1i = ss_right_f64(ts, 0, n, t1) - 1;
        if (t1i <= t0i) {   // skipped event (tbm.py:97-100)
            if (lane == 0) { labels[e] = 0; touch_idx[e] = t0i; rets[e] = nan; ratios[e] = nan; }
            continue;
        }
        const bool is_meta = side != nullptr;
        const double sm = is_meta ? (double)side[e] : 1.0;
        const double base = lc[t0i];
        double mu = 0.0, ml = 0.0;       // per-lane maxima, reduced at the end
        int64_t touch = t1i;
        double ret_final = 0.0;
        for (int64_t j0 = t0i + 1; j0 <= t1i; j0 += 32) {
            const int64_t j = j0 + lane;
            const bool act = j <= t1i;
            bool ok = false;
            double ret = 0.0;
            if (act) {
                const int64_t dur = ts[j] - t0;
                ok = !((double)dur < minc_ns);
                ret = __dmul_rn(__dadd_rn(lc[j], -base), sm);
            }
            const bool hit = ok && (ret >= upper || ret <= lower);
            const unsigned hm = __ballot_sync(FULL, hit);
            const int first = hm ? (__ffs(hm) - 1) : 32;
            if (ok && lane <= first) {
                if (ret > 0.0 && uv) { const double r = __ddiv_rn(ret, upper); if (r > mu) mu = r; }
                else if (ret < 0.0 && lv) { const double r = __ddiv_rn(ret, lower); if (r > ml) ml = r; }
            }
            if (hm) {
                touch = j0 + first;
                ret_final = __shfl_sync(FULL, ret, first);
                break;
            }
            // no touch in this step: remember ret of the last evaluated tick (ok is monotone in j)
            const unsigned om = __ballot_sync(FULL, ok);
            if (om) {
                const int lastl = 31 - __clz(om);
                ret_final = __shfl_sync(FULL, ret, lastl);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mu = fmax(mu, __shfl_xor_sync(FULL, mu, o));
            ml = fmax(ml, __shfl_xor_sync(FULL, ml, o));
        }
        if (lane == 0) {
            const double ret = ret_final;
            touch_idx[e] = touch;
            rets[e] = ret;
            if (is_meta) labels[e] = ret >= min_ret ? 1 : 0;
            else labels[e] = ret > 0.0 ? 1 : (ret < 0.0 ? -1 : 1);
            if (touch == t1i) {   // tbm.py:145 tests the index, not whether a barrier was hit
                double r;
                if (ret > 0.0) { r = __ddiv_rn(mu, __dadd_rn(1.0, ml)); if (!uv) r = nan; }
                else { r = __ddiv_rn(ml, __dadd_rn(1.0, mu)); if (!lv) r = nan; }
                ratios[e] = (1.0 < r) ? 1.0 : r;
            } else ratios[e] = 1.0;
        }
    }
}

extern "C" int fmk_triple_barrier(fmk_ctx *ctx, const fmk_trades *t, const int64_t *event_idx, const double *targets,
                                  int64_t ne, int64_t n_targets, double bottom_mult, double top_mult,
                                  double vertical_barrier_s, double min_close_time_s, const int8_t *side,
                                  int64_t n_side, double min_ret, int8_t *labels, int64_t *touch_idx, double *rets,
                                  double *ratios) {
    FMK_ENTER(ctx);
    if (vertical_barrier_s <= 0) return fmk_fail(ctx, FMK_ERR_ARG, "The vertical barrier must be greater than zero.");
    if (min_ret < 0) return fmk_fail(ctx, FMK_ERR_ARG, "The minimum return must be non-negative.");
    if (ne != n_targets) return fmk_fail(ctx, FMK_ERR_ARG, "The lengths of event_idxs and targets must match.");
    if (ne == 0) return fmk_fail(ctx, FMK_ERR_ARG, "The event_idxs array must not be empty.");
    if (side && n_side != ne) return fmk_fail(ctx, FMK_ERR_ARG, "The length of event_idxs must match the length of side.");
    const int64_t n = t->n;
    if (!t->ts) return fmk_fail(ctx, FMK_ERR_ARG, "triple_barrier needs the timestamp column on the device");
    // event indices must address the trade arrays (the reference would read out of bounds / wrap)
    for (int64_t e = 0; e < ne; e++)
        if (event_idx[e] < 0 || event_idx[e] >= n) return fmk_fail(ctx, FMK_ERR_ARG, "event index out of range");
    fmk_trades *tm = const_cast<fmk_trades *>(t);
    if (!tm->log_price) {
        FMK_TRY(fmk_dalloc(ctx, &tm->log_price, n));
        FMK_LAUNCH(ctx, k_log, (unsigned)cdiv(n, 256), 256, 0, t->price, n, tm->log_price);
    }
    Scratch<int64_t> dev(ctx), dtouch(ctx);
    Scratch<double> dtg(ctx), dret(ctx);
    Scratch<int8_t> dside(ctx), dlab(ctx);
    FMK_TRY(dev.alloc(ne)); FMK_TRY(dtouch.alloc(ne)); FMK_TRY(dtg.alloc(ne)); FMK_TRY(dret.alloc(2 * ne));
    FMK_TRY(dlab.alloc(ne));
    FMK_CUDA(ctx, cudaMemcpyAsync(dev.p, event_idx, (size_t)ne * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(dtg.p, targets, (size_t)ne * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (side) {
        FMK_TRY(dside.alloc(ne));
        FMK_CUDA(ctx, cudaMemcpyAsync(dside.p, side, (size_t)ne, cudaMemcpyHostToDevice, ctx->stream));
    }
    int64_t blocks = cdiv(ne, 8);
    const int64_t maxb = (int64_t)ctx->sm_count * 32;
    if (blocks > maxb) blocks = maxb;
    FMK_LAUNCH(ctx, k_triple_barrier, (unsigned)blocks, 256, 0, t->ts, (const double *)tm->log_price, n,
               (const int64_t *)dev.p, (const double *)dtg.p, ne, bottom_mult, top_mult, vertical_barrier_s * 1e9,
               min_close_time_s * 1e9, side ? (const int8_t *)dside.p : (const int8_t *)nullptr, min_ret, dlab.p,
               dtouch.p, dret.p, dret.p + ne);
    FMK_CUDA(ctx, cudaMemcpyAsync(labels, dlab.p, (size_t)ne, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(touch_idx, dtouch.p, (size_t)ne * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(rets, dret.p, (size_t)ne * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(ratios, dret.p + ne, (size_t)ne * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}
