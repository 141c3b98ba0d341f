// volprofile.cu -- rolling volume profile: volume_profile_rolling with aggregate_footprint,
// bucket_price_levels, comp_poc_hva_lva and calc_volume_percentage_above_poc (feature/core/volume.py:133-456) on the CSR
// footprint that fmk_bar_footprints already holds on the device.
//
//   P1  k_vp_window  : per output bar i the window [s, e) of bars with ts in [ts_i - W, ts_i] (two binary searches), the
//                      window's lowest / highest price level (round-half-even of an IEEE division, like the reference)
//   P2  k_vp_profile : one block per output bar.  Thread per price level, looping over the window's bars IN BAR ORDER, so
//                      every level's float32 buy / sell sums are accumulated exactly like `aligned[indices] += volumes[t]`;
//                      thread per bucket for the (sequential float32) bucketing; thread 0 walks the value area.
// All decisions are taken on float32 sums formed in the reference's order -> POC / HVA / LVA are bit-exact.
#include <math.h>
#include <new>
#include "common.cuh"

struct VpBar { int64_t off; int32_t cnt; int32_t lvl0; int32_t contig; int32_t pad; };

__global__ void k_vp_bars(const int64_t *__restrict__ off, const int32_t *__restrict__ levels, int64_t nb,
                          VpBar *__restrict__ bars) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb) return;
    VpBar b;
    b.off = off[t]; b.cnt = (int32_t)(off[t + 1] - off[t]); b.lvl0 = b.cnt > 0 ? levels[b.off] : 0; b.pad = 0;
    int c = 1;
    for (int32_t k = 1; k < b.cnt; k++) if (levels[b.off + k] != b.lvl0 + k) { c = 0; break; }
    b.contig = c;
    bars[t] = b;
}

__device__ __forceinline__ int64_t vp_search(const int64_t *__restrict__ a, int64_t n, int64_t key, bool right) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        const int64_t x = a[mid];
        if (right ? (x <= key) : (x < key)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct VpWin { int64_t s, e, lo; int32_t L; int32_t pad; };

__global__ void k_vp_window(const int64_t *__restrict__ ts, const double *__restrict__ highs,
                            const double *__restrict__ lows, int64_t nb, int64_t first, int64_t win_ns, double tick,
                            VpWin *__restrict__ w, int *__restrict__ maxL) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + first;
    if (i >= nb) return;
    const int64_t end_ts = ts[i], start_ts = end_ts - win_ns;
    int64_t s = vp_search(ts, nb, start_ts, false);
    const int64_t e = vp_search(ts, nb, end_ts, true);
    if (s == e) s = s - 1 > 0 ? s - 1 : 0;
    double mn = lows[s], mx = highs[s];
    for (int64_t t = s; t < e; t++) {       // np.min / np.max: NaN-free inputs
        const double l = lows[t], h = highs[t];
        if (l < mn) mn = l;
        if (h > mx) mx = h;
    }
    const int64_t lo = (int64_t)rint(__ddiv_rn(mn, tick)), hi = (int64_t)rint(__ddiv_rn(mx, tick));
    VpWin r;
    r.s = s; r.e = e; r.lo = lo; r.pad = 0;
    const int64_t L = hi - lo + 1;
    r.L = L > 0 && L < 0x7fffffff ? (int32_t)L : 0;
    w[i] = r;
    atomicMax(maxL, r.L);
}

// value-area walk of comp_poc_hva_lva / calc_volume_percentage_above_poc on (level(k), vol[k]), k < n; run by one thread.
// level(k) = lv ? lv[k] : lo + k.  The pair sums are float32 (Numba types each assignment separately), the running total
// of the walk is float64, np.sum of the float32 profile has a float32 accumulator.
__device__ void vp_poc_hva_lva(const int32_t *lv, int64_t lo, const float *vol, int64_t n, double va_pct, int32_t *poc_o,
                               int32_t *hva_o, int32_t *lva_o, float *pct_o) {
#define VP_LV(k) (lv ? lv[(k)] : (int32_t)(lo + (k)))
#define VP_PAIR(a, b) ((double)__fadd_rn(vol[(a)], vol[(b)]))
    float total = 0.0f;
    for (int64_t i = 0; i < n; i++) total = __fadd_rn(total, vol[i]);
    int64_t pi = 0;
    float best = vol[0];
    for (int64_t i = 1; i < n; i++) if (vol[i] > best) { best = vol[i]; pi = i; }
    const int32_t poc = VP_LV(pi);
    int32_t hva = poc, lva = poc;
    const double va_thrs = __dmul_rn((double)total, __ddiv_rn(va_pct, 100.0));
    double cum = vol[pi];
    int64_t up = pi + 1, dn = pi - 1;
    double cu = 0.0, cd = 0.0;
    if (up < n) cu = up + 1 < n ? VP_PAIR(up, up + 1) : (double)vol[up];
    if (dn >= 0) cd = dn - 1 >= 0 ? VP_PAIR(dn, dn - 1) : (double)vol[dn];
    while (cum < va_thrs) {
        if (cu > cd) {
            cum = __dadd_rn(cum, cu); hva = VP_LV(up + 1 < n - 1 ? up + 1 : n - 1); up += 2;
            cu = -1.0;
            if (up < n) cu = up + 1 < n ? VP_PAIR(up, up + 1) : (double)vol[up];
        } else if (cu < cd) {
            cum = __dadd_rn(cum, cd); lva = VP_LV(dn - 1 > 0 ? dn - 1 : 0); dn -= 2;
            cd = -1.0;
            if (dn >= 0) cd = dn - 1 >= 0 ? VP_PAIR(dn, dn - 1) : (double)vol[dn];
        } else if (cu == cd && cd != -1.0) {
            cum = __dadd_rn(cum, __dadd_rn(cu, cd));
            hva = VP_LV(up + 1 < n - 1 ? up + 1 : n - 1); lva = VP_LV(dn - 1 > 0 ? dn - 1 : 0);
            up += 2; dn -= 2;
            cu = -1.0;
            if (up < n) cu = up + 1 < n ? VP_PAIR(up, up + 1) : (double)vol[up];
            cd = -1.0;
            if (dn >= 0) cd = dn - 1 >= 0 ? VP_PAIR(dn, dn - 1) : (double)vol[dn];
        } else break;
    }
    *poc_o = poc; *hva_o = hva; *lva_o = lva;
    float pct = 0.0f;
    if (!(total <= 0.0f)) {
        double above = 0.0;
        for (int64_t i = 0; i < n; i++) if (VP_LV(i) > poc) above = __dadd_rn(above, (double)vol[i]);
        if (!(above <= 0.0)) pct = (float)__ddiv_rn(above, (double)total);
    }
    *pct_o = pct;
#undef VP_LV
#undef VP_PAIR
}

constexpr int VP_THREADS = 256;

__global__ void __launch_bounds__(VP_THREADS) k_vp_profile(const VpWin *__restrict__ win, const VpBar *__restrict__ bars,
                                                           const int32_t *__restrict__ levels,
                                                           const float *__restrict__ buy, const float *__restrict__ sell,
                                                           int64_t nb, int64_t first, int64_t n_bins, int maxL,
                                                           int max_binned, double va_pct, float *__restrict__ scratch,
                                                           int32_t *__restrict__ poc, int32_t *__restrict__ hva,
                                                           int32_t *__restrict__ lva, float *__restrict__ pct) {
    extern __shared__ unsigned char vp_raw[];
    float *bv = reinterpret_cast<float *>(vp_raw);                       // [max_binned]
    int32_t *blv = reinterpret_cast<int32_t *>(bv + max_binned);         // [max_binned]
    float *tot = scratch + (size_t)blockIdx.x * maxL;
    for (int64_t i = first + blockIdx.x; i < nb; i += gridDim.x) {
        const VpWin w = win[i];
        const int L = w.L;
        if (L <= 0) continue;
        // aggregate_footprint: per level, float32 sums over the window's bars in bar order
        for (int l = threadIdx.x; l < L; l += VP_THREADS) {
            const int64_t level = w.lo + l;
            float ab = 0.0f, as = 0.0f;
            for (int64_t t = w.s; t < w.e; t++) {
                const VpBar b = bars[t];
                int64_t q = level - b.lvl0;
                if (!b.contig) {            // arbitrary ascending levels: binary search for an exact match
                    int lo2 = 0, hi2 = b.cnt;
                    while (lo2 < hi2) { const int m = (lo2 + hi2) >> 1; if (levels[b.off + m] < level) lo2 = m + 1; else hi2 = m; }
                    q = (lo2 < b.cnt && levels[b.off + lo2] == level) ? lo2 : -1;
                }
                if (q >= 0 && q < b.cnt) { ab = __fadd_rn(ab, buy[b.off + q]); as = __fadd_rn(as, sell[b.off + q]); }
            }
            tot[l] = __fadd_rn(ab, as);
        }
        __syncthreads();
        const float *pv = tot;
        const int32_t *plv = nullptr;
        int64_t n = L;
        if (n_bins > 0) {
            // bucket_price_levels (volume.py:208-280)
            const int64_t lo = w.lo, hi = w.lo + L - 1, range = hi - lo;
            int64_t bw = range / n_bins;
            if (bw < 1) bw = 1;
            if ((bw & 1) == 0) bw += 1;
            const int64_t nedges = (range + bw) / bw + (((range + bw) % bw) ? 1 : 0);   // len(arange(lo, hi + bw, bw))
            const int64_t nbin = nedges - 1;
            const bool single = nedges < 2;
            const int64_t e_last = single ? hi + 1 : lo + (nedges - 1) * bw;
            const int64_t last_idx = single ? (hi >= e_last ? 1 : 0) : (hi >= e_last ? nbin : (hi - lo) / bw);
            const bool left = last_idx == nbin;
            const int64_t nout = nbin + (left ? 1 : 0);
            for (int64_t b = threadIdx.x; b < nout; b += VP_THREADS) {
                float acc = 0.0f;
                if (b < nbin) {
                    const int64_t a = lo + b * bw, c = lo + (b + 1) * bw;
                    const int64_t m2 = a + c - 1;
                    blv[b] = (int32_t)(m2 >= 0 ? m2 / 2 : -((-m2 + 1) / 2));        // python floor division
                    int64_t k0 = a - lo, k1 = c - lo;                                 // levels [a, c) and < e_last
                    if (k1 > e_last - lo) k1 = e_last - lo;
                    if (k1 > L) k1 = L;
                    for (int64_t k = k0; k < k1; k++) acc = __fadd_rn(acc, tot[k]);
                } else {                                                             // leftover bin: levels >= e_last
                    blv[b] = (int32_t)hi;
                    for (int64_t k = single ? 0 : e_last - lo; k < L; k++) acc = __fadd_rn(acc, tot[k]);   // single bin: every level
                }
                bv[b] = acc;
            }
            __syncthreads();
            pv = bv; plv = blv; n = nout;
        }
        if (threadIdx.x == 0 && n > 0) vp_poc_hva_lva(plv, w.lo, pv, n, va_pct, &poc[i], &hva[i], &lva[i], &pct[i]);
        __syncthreads();
    }
}

static int vp_run(fmk_ctx *ctx, const int64_t *d_off, const int32_t *d_levels, const float *d_buy, const float *d_sell,
                  int64_t nb, const int64_t *bar_ts, const double *highs, const double *lows, double window_sec, int64_t n_bins,
                  double price_tick, double va_pct, int32_t *poc, int32_t *hva, int32_t *lva, float *pct) {
    if (nb <= 0) return fmk_fail(ctx, FMK_ERR_ARG, "Input arrays should have the same length and be non-empty.");
    for (int64_t i = 0; i < nb; i++) { poc[i] = hva[i] = lva[i] = 0; pct[i] = 0.0f; }
    const int64_t win_ns = (int64_t)(window_sec * 1e9);
    // first_interval_idx = searchsorted(ts, ts[0] + window) on the host copy
    int64_t first;
    {
        int64_t lo = 0, hi = nb;
        const int64_t key = bar_ts[0] + win_ns;
        while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if (bar_ts[mid] < key) lo = mid + 1; else hi = mid; }
        first = lo;
    }
    if (first >= nb) return FMK_OK;
    Scratch<int64_t> dts(ctx);
    Scratch<double> dh(ctx), dl(ctx);
    Scratch<VpBar> bars(ctx);
    Scratch<VpWin> win(ctx);
    Scratch<int> dmax(ctx);
    Scratch<int32_t> dout(ctx);
    Scratch<float> dpct(ctx);
    FMK_TRY(dts.alloc(nb)); FMK_TRY(dh.alloc(nb)); FMK_TRY(dl.alloc(nb)); FMK_TRY(bars.alloc(nb)); FMK_TRY(win.alloc(nb));
    FMK_TRY(dmax.alloc(1)); FMK_TRY(dout.alloc(3 * nb)); FMK_TRY(dpct.alloc(nb));
    FMK_CUDA(ctx, cudaMemcpyAsync(dts.p, bar_ts, (size_t)nb * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(dh.p, highs, (size_t)nb * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(dl.p, lows, (size_t)nb * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemsetAsync(dmax.p, 0, 4, ctx->stream));
    FMK_CUDA(ctx, cudaMemsetAsync(dout.p, 0, (size_t)nb * 12, ctx->stream));
    FMK_CUDA(ctx, cudaMemsetAsync(dpct.p, 0, (size_t)nb * 4, ctx->stream));
    FMK_LAUNCH(ctx, k_vp_bars, (unsigned)cdiv(nb, 256), 256, 0, d_off, d_levels, nb, bars.p);
    FMK_LAUNCH(ctx, k_vp_window, (unsigned)cdiv(nb - first, 256), 256, 0, (const int64_t *)dts.p, (const double *)dh.p,
               (const double *)dl.p, nb, first, win_ns, price_tick, win.p, dmax.p);
    int maxL = 0;
    FMK_CUDA(ctx, cudaMemcpyAsync(&maxL, dmax.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (maxL <= 0) return FMK_OK;
    int64_t blocks = nb - first;
    const int64_t maxb = (int64_t)ctx->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    Scratch<float> scratch(ctx);
    FMK_TRY(scratch.alloc(blocks * maxL));
    const int max_binned = n_bins > 0 ? (int)(2 * n_bins + 8) : 1;
    const size_t smem = (size_t)max_binned * 8;
    if (smem > 48 * 1024) return fmk_fail(ctx, FMK_ERR_ARG, "n_bins too large");
    FMK_LAUNCH(ctx, k_vp_profile, (unsigned)blocks, VP_THREADS, smem, (const VpWin *)win.p, (const VpBar *)bars.p, d_levels,
               d_buy, d_sell, nb, first, n_bins, maxL, max_binned, va_pct, scratch.p, dout.p, dout.p + nb, dout.p + 2 * nb,
               dpct.p);
    FMK_CUDA(ctx, cudaMemcpyAsync(poc, dout.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(hva, dout.p + nb, (size_t)nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(lva, dout.p + 2 * nb, (size_t)nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(pct, dpct.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

// device-resident footprint (the handle fmk_bar_footprints returns)
extern "C" int fmk_volume_profile_rolling_fp(fmk_ctx *ctx, const fmk_footprint *fp, const int64_t *bar_ts,
                                             const double *highs, const double *lows, double window_sec, int64_t n_bins,
                                             double price_tick, double va_pct, int32_t *poc, int32_t *hva, int32_t *lva,
                                             float *pct_above_poc) {
    FMK_ENTER(ctx);
    return vp_run(ctx, fp->level_offsets, fp->price_levels, fp->buy_vol, fp->sell_vol, fp->n_bars, bar_ts, highs, lows,
                  window_sec, n_bins, price_tick, va_pct, poc, hva, lva, pct_above_poc);
}

// host CSR (a FootprintData built elsewhere): level_offsets[n_bars + 1] + flat per-level arrays
extern "C" int fmk_volume_profile_rolling(fmk_ctx *ctx, const int64_t *level_offsets, const int32_t *price_levels,
                                          const float *buy_volumes, const float *sell_volumes, int64_t n_bars,
                                          const int64_t *bar_ts, const double *highs, const double *lows, double window_sec,
                                          int64_t n_bins, double price_tick, double va_pct, int32_t *poc, int32_t *hva,
                                          int32_t *lva, float *pct_above_poc) {
    FMK_ENTER(ctx);
    if (n_bars <= 0) return fmk_fail(ctx, FMK_ERR_ARG, "Input arrays should have the same length and be non-empty.");
    const int64_t total = level_offsets[n_bars];
    Scratch<int64_t> doff(ctx);
    Scratch<int32_t> dlv(ctx);
    Scratch<float> db(ctx), ds(ctx);
    FMK_TRY(doff.alloc(n_bars + 1)); FMK_TRY(dlv.alloc(total)); FMK_TRY(db.alloc(total)); FMK_TRY(ds.alloc(total));
    FMK_CUDA(ctx, cudaMemcpyAsync(doff.p, level_offsets, (size_t)(n_bars + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(dlv.p, price_levels, (size_t)total * 4, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(db.p, buy_volumes, (size_t)total * 4, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(ds.p, sell_volumes, (size_t)total * 4, cudaMemcpyHostToDevice, ctx->stream));
    return vp_run(ctx, doff.p, dlv.p, db.p, ds.p, n_bars, bar_ts, highs, lows, window_sec, n_bins, price_tick, va_pct, poc,
                  hva, lva, pct_above_poc);
}
