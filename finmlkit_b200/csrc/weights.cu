// weights.cu -- sample weights on ticks: average_uniqueness (label/weights.py:7-49) and
// return_attribution (label/weights.py:52-103) of the reference.
//
// The reference does E range-increments over an int16 concurrency array and then E serial sums over the label paths
// (mean path 61 k ticks at the headline config -> O(sum of path lengths)).  Here both are O(N + E):
//
//   W1  k_w_scatter    +1 at event_idx, -1 at touch_idx+1 into an int32 difference array (atomics; order-free: integer)
//   W2  scan           concurrency = prefix sum, stored as int16 with the reference's silent wrap-around (mod 2^16)
//   W3  k_w_tile_sums  per 256-tick tile: sum of 1/c (c != 0), sum of log(p_j/p_{j-1})/c (c > 0, finite), count of
//                      "special" ticks (c == 0, or an infinite log return) that need the reference's literal loop
//   W4  scan           double-double prefix over the tiles (N/256 elements)
//   W5  k_w_events     warp per event: partial head tile + (double-double difference of tile prefixes) + partial tail
//                      tile.  Events whose range holds a special tick are summed serially, in the reference's order.
//
// Values agree with the reference's sequential sums to ~1e-13 relative (<< the 1e-9 bar); concurrency is bit-exact.
// Indices must lie in [0, n): the reference reads out of bounds / wraps there.  touch < event gives an empty path:
// uniqueness NaN (Numba would raise ZeroDivisionError on the empty mean), attribution 0.
#include <math.h>
#include <new>
#include "common.cuh"
#include "dollar_core.h"
#include "scan.cuh"

constexpr int WT = 256;          // ticks per tile

struct WAcc {
    dd_t u, r;
    long long sp;
    __device__ WAcc() {}
    __device__ explicit WAcc(int) { u.hi = u.lo = r.hi = r.lo = 0.0; sp = 0; }
};
__device__ __forceinline__ WAcc operator+(const WAcc &a, const WAcc &b) {
    WAcc c;
    c.u = dd_add(a.u, b.u); c.r = dd_add(a.r, b.r); c.sp = a.sp + b.sp;
    return c;
}
__device__ __forceinline__ WAcc __shfl_up_sync(unsigned m, const WAcc &x, int o) {
    WAcc r;
    r.u.hi = ::__shfl_up_sync(m, x.u.hi, o); r.u.lo = ::__shfl_up_sync(m, x.u.lo, o);
    r.r.hi = ::__shfl_up_sync(m, x.r.hi, o); r.r.lo = ::__shfl_up_sync(m, x.r.lo, o);
    r.sp = ::__shfl_up_sync(m, x.sp, o);
    return r;
}
struct WTile { double su, sr; long long sp; };

__global__ void k_w_scatter(const int64_t *__restrict__ ev, const int64_t *__restrict__ touch, int64_t ne,
                            int32_t *__restrict__ diff) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ne) return;
    const int64_t s = ev[i], e = touch[i];
    if (e < s) return;
    atomicAdd(diff + s, 1);
    atomicAdd(diff + e + 1, -1);
}

struct ConcIn {
    const int32_t *diff;
    __device__ int32_t operator()(int64_t i) const { return diff[i]; }
};
struct ConcOut {
    int16_t *c16;
    __device__ void operator()(int64_t i, int32_t cs) const { c16[i] = (int16_t)(uint16_t)(uint32_t)cs; }
};

// the two per-tick terms; `special` marks ticks the tile sums cannot represent
__device__ __forceinline__ void w_terms(const int16_t *__restrict__ c16, const double *__restrict__ close, int64_t j,
                                        double *tu, double *tr, int *special) {
    const int c = c16[j];
    *special = 0;
    *tu = 0.0; *tr = 0.0;
    if (c == 0) { *special = 1; return; }
    *tu = __ddiv_rn(1.0, (double)c);
    if (close && c > 0 && j > 0) {
        const double p0 = close[j - 1];
        if (p0 != 0.0) {
            const double lr = log(__ddiv_rn(close[j], p0));
            if (lr == lr) {
                if (isinf(lr)) *special = 1;
                else *tr = __ddiv_rn(lr, (double)c);
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_w_tile_sums(const int16_t *__restrict__ c16, const double *__restrict__ close,
                                                     int64_t n, int64_t ntiles, WTile *__restrict__ tiles) {
    const int lane = threadIdx.x & 31;
    const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (tile >= ntiles) return;
    const int64_t base = tile * WT;
    double su = 0.0, sr = 0.0;
    int sp = 0;
#pragma unroll
    for (int k = 0; k < WT / 32; k++) {
        const int64_t j = base + lane + 32 * k;
        if (j < n) {
            double tu, tr; int s;
            w_terms(c16, close, j, &tu, &tr, &s);
            su += tu; sr += tr; sp += s;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        su += __shfl_xor_sync(0xffffffffu, su, o);
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        sp += __shfl_xor_sync(0xffffffffu, sp, o);
    }
    if (lane == 0) { tiles[tile].su = su; tiles[tile].sr = sr; tiles[tile].sp = sp; }
}

struct TileIn {
    const WTile *t;
    __device__ WAcc operator()(int64_t i) const {
        WAcc a;
        a.u.hi = t[i].su; a.u.lo = 0.0; a.r.hi = t[i].sr; a.r.lo = 0.0; a.sp = t[i].sp;
        return a;
    }
};
struct TileOut {   // exclusive prefix table P[0..ntiles]: P[i+1] = inclusive prefix of tile i
    WAcc *P;
    __device__ void operator()(int64_t i, const WAcc &cs) const {
        P[i + 1] = cs;
        if (i == 0) P[0] = WAcc(0);
    }
};

// warp-cooperative direct sum of the terms over [a, b)
__device__ __forceinline__ void w_range(const int16_t *__restrict__ c16, const double *__restrict__ close, int64_t a,
                                        int64_t b, int lane, double *su, double *sr, int *sp) {
    double u = 0.0, r = 0.0;
    int s = 0;
    for (int64_t j = a + lane; j < b; j += 32) {
        double tu, tr; int q;
        w_terms(c16, close, j, &tu, &tr, &q);
        u += tu; r += tr; s += q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u += __shfl_xor_sync(0xffffffffu, u, o);
        r += __shfl_xor_sync(0xffffffffu, r, o);
        s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    *su = u; *sr = r; *sp = s;
}

__global__ void __launch_bounds__(256) k_w_events(const int16_t *__restrict__ c16, const double *__restrict__ close,
                                                  const WAcc *__restrict__ P, const int64_t *__restrict__ ev,
                                                  const int64_t *__restrict__ touch, int64_t ne,
                                                  double *__restrict__ w_u, double *__restrict__ w_r) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < ne; i += nwarps) {
        const int64_t s = ev[i], e1 = touch[i] + 1;   // [s, e1)
        if (e1 <= s) {
            if (lane == 0) { if (w_u) w_u[i] = __longlong_as_double(0x7ff8000000000000ll); if (w_r) w_r[i] = 0.0; }
            continue;
        }
        const int64_t t0 = (s + WT - 1) / WT, t1 = e1 / WT;   // full tiles [t0, t1)
        double su, sr; int sp;
        double tot_u, tot_r;
        long long special;
        if (t0 >= t1) {
            w_range(c16, close, s, e1, lane, &su, &sr, &sp);
            tot_u = su; tot_r = sr; special = sp;
        } else {
            double hu, hr, tu, tr; int hs, ts;
            w_range(c16, close, s, t0 * WT, lane, &hu, &hr, &hs);
            w_range(c16, close, t1 * WT, e1, lane, &tu, &tr, &ts);
            const WAcc a = P[t0], b = P[t1];
            dd_t na = {-a.u.hi, -a.u.lo}, nr = {-a.r.hi, -a.r.lo};
            const dd_t du = dd_add(b.u, na), dr = dd_add(b.r, nr);
            tot_u = (du.hi + du.lo) + (hu + tu);
            tot_r = (dr.hi + dr.lo) + (hr + tr);
            special = (b.sp - a.sp) + hs + ts;
        }
        if (special == 0) {
            if (lane == 0) {
                if (w_u) w_u[i] = __ddiv_rn(tot_u, (double)(e1 - s));
                if (w_r) w_r[i] = fabs(tot_r);
            }
        } else if (lane == 0) {
            // the reference's literal loops (weights.py:43-46, 88-95): zeros of the wrapped concurrency give inf terms,
            // a zero close gives an infinite log return
            double c = 0.0, w = 0.0;
            for (int64_t j = s; j < e1; j++) {
                const int cc = c16[j];
                c = __dadd_rn(c, __ddiv_rn(1.0, (double)cc));
                if (close && cc > 0 && j > 0) {
                    const double p0 = close[j - 1];
                    if (p0 != 0.0) {
                        const double lr = log(__ddiv_rn(close[j], p0));
                        if (lr == lr) w = __dadd_rn(w, __ddiv_rn(lr, (double)cc));
                    }
                }
            }
            if (w_u) w_u[i] = __ddiv_rn(c, (double)(e1 - s));
            if (w_r) w_r[i] = fabs(w);
        }
    }
}

// shared driver: device event arrays + device close (may be null: uniqueness only) -> host outputs
static int run_weights(fmk_ctx *ctx, int64_t n, const double *d_close, const int64_t *event_idx, const int64_t *touch_idx,
                       int64_t ne, const int16_t *h_conc_in, double *w_u, double *w_r, int16_t *h_conc_out) {
    if (n <= 0) return fmk_fail(ctx, FMK_ERR_ARG, "empty series");
    for (int64_t i = 0; i < ne; i++)
        if (event_idx[i] < 0 || event_idx[i] >= n || touch_idx[i] < 0 || touch_idx[i] >= n)
            return fmk_fail(ctx, FMK_ERR_ARG, "event / touch indices must lie in [0, len(timestamps))");
    Scratch<int64_t> dev(ctx), dtouch(ctx);
    Scratch<int16_t> c16(ctx);
    FMK_TRY(dev.alloc(ne)); FMK_TRY(dtouch.alloc(ne)); FMK_TRY(c16.alloc(n));
    FMK_CUDA(ctx, cudaMemcpyAsync(dev.p, event_idx, (size_t)ne * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(dtouch.p, touch_idx, (size_t)ne * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (h_conc_in) {
        FMK_CUDA(ctx, cudaMemcpyAsync(c16.p, h_conc_in, (size_t)n * 2, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        Scratch<int32_t> diff(ctx);
        FMK_TRY(diff.alloc(n + 1));
        FMK_CUDA(ctx, cudaMemsetAsync(diff.p, 0, (size_t)(n + 1) * 4, ctx->stream));
        if (ne > 0) FMK_LAUNCH(ctx, k_w_scatter, (unsigned)cdiv(ne, 256), 256, 0, dev.p, dtouch.p, ne, diff.p);
        FMK_TRY((device_inclusive_scan<int32_t>(ctx, ConcIn{diff.p}, ConcOut{c16.p}, n, (int32_t *)nullptr)));
    }
    if (h_conc_out) FMK_CUDA(ctx, cudaMemcpyAsync(h_conc_out, c16.p, (size_t)n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    if (ne > 0 && (w_u || w_r)) {
        const int64_t ntiles = cdiv(n, WT);
        Scratch<WTile> tiles(ctx);
        Scratch<WAcc> P(ctx);
        Scratch<double> dw(ctx);
        FMK_TRY(tiles.alloc(ntiles)); FMK_TRY(P.alloc(ntiles + 1)); FMK_TRY(dw.alloc(2 * ne));
        FMK_LAUNCH(ctx, k_w_tile_sums, (unsigned)cdiv(ntiles, 8), 256, 0, c16.p, d_close, n, ntiles, tiles.p);
        FMK_TRY((device_inclusive_scan<WAcc>(ctx, TileIn{tiles.p}, TileOut{P.p}, ntiles, (WAcc *)nullptr)));
        int64_t blocks = cdiv(ne, 8);
        const int64_t maxb = (int64_t)ctx->sm_count * 32;
        if (blocks > maxb) blocks = maxb;
        FMK_LAUNCH(ctx, k_w_events, (unsigned)blocks, 256, 0, c16.p, d_close, P.p, dev.p, dtouch.p, ne,
                   w_u ? dw.p : (double *)nullptr, w_r ? dw.p + ne : (double *)nullptr);
        if (w_u) FMK_CUDA(ctx, cudaMemcpyAsync(w_u, dw.p, (size_t)ne * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (w_r) FMK_CUDA(ctx, cudaMemcpyAsync(w_r, dw.p + ne, (size_t)ne * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

// weights.py:97-101: weights *= n_events / sum(weights); ValueError if the sum is <= 0
static int normalize_weights(fmk_ctx *ctx, double *w, int64_t ne) {
    double s = 0.0;
    for (int64_t i = 0; i < ne; i++) s += w[i];
    if (s <= 0.) return fmk_fail(ctx, FMK_ERR_ARG, "Sum of weights is zero or negative, cannot normalize.");
    const double f = (double)ne / s;
    for (int64_t i = 0; i < ne; i++) w[i] *= f;
    return FMK_OK;
}

extern "C" int fmk_average_uniqueness(fmk_ctx *ctx, int64_t n, const int64_t *event_idx, const int64_t *touch_idx,
                                      int64_t n_events, int64_t n_touch, double *weights, int16_t *concurrency) {
    FMK_ENTER(ctx);
    if (n_events != n_touch)
        return fmk_fail(ctx, FMK_ERR_ARG, "Timestamps and lookahead indices must have the same length.");
    return run_weights(ctx, n, nullptr, event_idx, touch_idx, n_events, nullptr, weights, nullptr, concurrency);
}

extern "C" int fmk_return_attribution(fmk_ctx *ctx, const int64_t *event_idx, const int64_t *touch_idx, int64_t n_events,
                                      const double *close, const int16_t *concurrency, int64_t n, int normalize,
                                      double *weights) {
    FMK_ENTER(ctx);
    Scratch<double> dclose(ctx);
    FMK_TRY(dclose.alloc(n));
    FMK_CUDA(ctx, cudaMemcpyAsync(dclose.p, close, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_TRY(run_weights(ctx, n, dclose.p, event_idx, touch_idx, n_events, concurrency, nullptr, weights, nullptr));
    return normalize ? normalize_weights(ctx, weights, n_events) : FMK_OK;
}

extern "C" int fmk_sample_weights(fmk_ctx *ctx, const fmk_trades *t, const int64_t *event_idx, const int64_t *touch_idx,
                                  int64_t n_events, int normalize, double *avg_uniqueness, double *return_attribution,
                                  int16_t *concurrency) {
    FMK_ENTER(ctx);
    FMK_TRY(run_weights(ctx, t->n, t->price, event_idx, touch_idx, n_events, nullptr, avg_uniqueness, return_attribution,
                        concurrency));
    return (normalize && return_attribution) ? normalize_weights(ctx, return_attribution, n_events) : FMK_OK;
}
