/*
 * fmk_oracle.c -- CPU restatement of finmlkit's Numba tick-data hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (finmlkit_b200) never links, imports or calls anything in oracle/.
 *
 * Every function restates, in plain C with the same evaluation order, one function of the
 * reference (quantscious/finmlkit v0.1.11, paths relative to /root/reference).  Parity is
 * PINNED: tests/golden/make_golden.py imports the reference in the build container, runs
 * both, and commits the reference outputs as fixtures; tests/test_oracle_golden.py checks this
 * file against those fixtures and against the reference's own hand-written test vectors.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -fno-fast-math  (see oracle/Makefile)
 *   -ffp-contract=off : Numba/LLVM does not fuse `cum += p*v` (SURVEY hazard B).
 * libm: log/exp/log1p/sqrt come from the same glibc Numba's LLVM intrinsics resolve to.
 *
 * `prange` loops of the reference are `#pragma omp parallel for` here so that the CPU baseline
 * uses all host cores like Numba's `parallel=True` does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FMKO_OK 0
#define FMKO_ERR_LEN (-1)
#define FMKO_ERR_FEW_INDICES (-2)
#define FMKO_ERR_CAP (-3)
#define FMKO_ERR_LEVEL (-4)
#define FMKO_ERR_VERTICAL (-5)
#define FMKO_ERR_MINRET (-6)
#define FMKO_ERR_EMPTY (-7)
#define FMKO_ERR_WINDOW (-8)
#define FMKO_ERR_THETA (-9)

int fmko_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void fmko_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* Python/Numba float floor division a // b (b > 0 here). */
static double py_floordiv(double vx, double wx) {
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    if (mod != 0.0 && ((wx < 0) != (mod < 0))) {
        div -= 1.0;
    }
    double floordiv;
    if (div != 0.0) {
        floordiv = floor(div);
        if (div - floordiv > 0.5) floordiv += 1.0;
    } else {
        floordiv = copysign(0.0, vx / wx);
    }
    return floordiv;
}

/* np.searchsorted(int64 a, int64 key, side='right') */
static int64_t ss_right_i64(const int64_t *a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
/* np.searchsorted(int64 a, float64 key, ...) : numba promotes both sides to float64 */
static int64_t ss_right_f64(const int64_t *a, int64_t n, double key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (isnan(key) || (double)a[mid] <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
static int64_t ss_left_f64(const int64_t *a, int64_t n, double key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if ((double)a[mid] < key) lo = mid + 1; else hi = mid; /* NaN key: a<NaN false -> numba less_than_float treats nan as +inf; not on this path */
    }
    return lo;
}

/* ---- a1: finmlkit/bar/logic.py:12-51 _time_bar_indexer -------------------------------------
 * Two-phase: call with clock==NULL to get the bar-clock length, then with buffers. */
int64_t fmko_time_bar_indexer(const int64_t *ts, int64_t n, double interval_seconds,
                              int64_t *clock, int64_t *idx, int64_t cap) {
    if (n <= 0) return 0;
    double iv = interval_seconds * 1e9;                                  /* logic.py:30 */
    double start = py_floordiv((double)ts[0], iv) * iv;                  /* logic.py:33 */
    double last = ceil((double)ts[n - 1] / iv) * iv;                     /* logic.py:36 */
    double stop = last + iv + 1.0;                                       /* logic.py:39 */
    /* numba np.arange(start, stop, step): nitems = ceil((stop-start)/step); arr[i] = int64(start + i*step) */
    double nitems_c = (stop - start) / iv;
    int64_t nitems = (int64_t)ceil(nitems_c);
    if (nitems < 0) nitems = 0;
    if (clock == NULL) return nitems;
    if (cap < nitems) return FMKO_ERR_CAP;
    for (int64_t i = 0; i < nitems; i++) {
        clock[i] = (int64_t)(start + (double)i * iv);
    }
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nitems; i++) {
        idx[i] = ss_right_i64(ts, n, clock[i]) - 1;                      /* logic.py:42 */
    }
    return nitems;
}

/* ---- a2: logic.py:54-84 _tick_bar_indexer ---------------------------------------------------
 * All threshold indexers: returns number of indices; writes at most cap of them (idx may be NULL). */
int64_t fmko_tick_bar_indexer(int64_t n, int64_t threshold, int64_t *idx, int64_t cap) {
    int64_t m = 0;
    if (n <= 0) return 0; /* reference would still emit 0; callers never pass empty arrays */
    if (idx && m < cap) idx[m] = 0;
    m++;
    int64_t cum = 1;
    for (int64_t i = 1; i < n; i++) {
        cum += 1;
        if (cum >= threshold) {
            if (idx && m < cap) idx[m] = i;
            m++;
            cum = 0;
        }
    }
    return m;
}

/* ---- a3: logic.py:87-115 _volume_bar_indexer ------------------------------------------------ */
int64_t fmko_volume_bar_indexer(const double *v, int64_t n, double threshold, int64_t *idx, int64_t cap) {
    int64_t m = 0;
    if (n <= 0) return 0;
    if (idx && m < cap) idx[m] = 0;
    m++;
    double cum = v[0];
    for (int64_t i = 1; i < n; i++) {
        cum += v[i];
        if (cum >= threshold) {
            if (idx && m < cap) idx[m] = i;
            m++;
            cum = 0.0;
        }
    }
    return m;
}

/* ---- a4: logic.py:118-149 _dollar_bar_indexer ----------------------------------------------- */
int64_t fmko_dollar_bar_indexer(const double *p, const double *v, int64_t n, double threshold,
                                int64_t *idx, int64_t cap) {
    int64_t m = 0;
    if (n <= 0) return 0;
    if (idx && m < cap) idx[m] = 0;
    m++;
    double cum = p[0] * v[0];
    for (int64_t i = 1; i < n; i++) {
        double d = p[i] * v[i];
        cum = cum + d;
        if (cum >= threshold) {
            if (idx && m < cap) idx[m] = i;
            m++;
            cum = cum - threshold;
        }
    }
    return m;
}

/* ---- a5: logic.py:152-221 _cusum_bar_indexer  (sigma is forward-filled IN PLACE like the reference) */
int64_t fmko_cusum_bar_indexer(const int64_t *ts, const double *p, double *sigma, int64_t n,
                               double sigma_floor, double sigma_mult, int64_t *idx, int64_t cap) {
    int64_t m = 0;
    if (n <= 0) return 0;
    int64_t first = 0;
    for (int64_t i = 0; i < n; i++) {
        if (!isnan(sigma[i])) { first = i; break; }
    }
    for (int64_t i = first; i < n; i++) {
        if (isnan(sigma[i])) sigma[i] = sigma[i - 1 >= 0 ? i - 1 : n - 1];
    }
    if (idx && m < cap) idx[m] = first;
    m++;
    double s_pos = 0.0, s_neg = 0.0;
    int64_t i = first + 1;
    while (i < n) {
        double ret = log(p[i] / p[i - 1]);
        s_pos = fmax(0.0, s_pos + ret);
        s_neg = fmin(0.0, s_neg + ret);
        if (i + 1 < n && ts[i] == ts[i + 1]) { i++; continue; }
        double lam = sigma_mult * sigma[i];
        if (!(lam > sigma_floor)) lam = (lam != lam) ? lam : sigma_floor; /* python max(a,b): b if b>a else a */
        if (s_pos >= lam) {
            if (idx && m < cap) idx[m] = i;
            m++;
            s_pos = 0.0;
        } else if (s_neg <= -lam) {
            if (idx && m < cap) idx[m] = i;
            m++;
            s_neg = 0.0;
        }
        i++;
    }
    return m;
}

/* nth_element (quickselect, median-of-three pivot, insertion sort on short ranges): after the call a[k] holds the k-th
 * order statistic, a[0..k) <= a[k] <= a(k..n).  Numba's np.median / np.percentile are selections too (introselect in
 * numba/np/arraymath.py); an order statistic does not depend on the algorithm, so any exact selection is a faithful
 * restatement -- and, unlike the qsort this replaced, it costs what the reference's costs (bench cpu_baseline). */
static void nth_element_f64(double *a, int64_t n, int64_t k) {
    int64_t lo = 0, hi = n - 1;
    while (hi - lo > 16) {
        int64_t mid = lo + ((hi - lo) >> 1);
        double x = a[lo], y = a[mid], z = a[hi], t;
        if (y < x) { t = x; x = y; y = t; }
        if (z < y) { t = y; y = z; z = t; if (y < x) { t = x; x = y; y = t; } }
        a[lo] = x; a[mid] = y; a[hi] = z;
        const double pv = y;
        int64_t i = lo, j = hi;
        for (;;) {
            do i++; while (a[i] < pv);
            do j--; while (a[j] > pv);
            if (i >= j) break;
            t = a[i]; a[i] = a[j]; a[j] = t;
        }
        /* a[lo..j] <= pv <= a[j+1..hi] */
        if (k <= j) hi = j; else lo = j + 1;
    }
    for (int64_t i = lo + 1; i <= hi; i++) {
        double x = a[i];
        int64_t j = i - 1;
        while (j >= lo && a[j] > x) { a[j + 1] = a[j]; j--; }
        a[j + 1] = x;
    }
}

/* ---- a7: bar/base.py:306-407 comp_bar_ohlcv ------------------------------------------------- */
int fmko_bar_ohlcv(const double *p, const double *v, int64_t n, int64_t nv, const int64_t *ci, int64_t nci,
                   double *o, double *h, double *l, double *c, float *vol, double *vwap,
                   int64_t *trades, double *median) {
    if (n != nv) return FMKO_ERR_LEN;
    if (nci < 2) return FMKO_ERR_FEW_INDICES;
    int64_t nb = nci - 1;
    #pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < nb; i++) {
        int64_t start = ci[i], end = ci[i + 1];
        if (start == end) {
            int64_t e = end < 0 ? end + n : end;
            o[i] = c[i] = h[i] = l[i] = p[e];
            vol[i] = 0.0f; vwap[i] = 0.0; trades[i] = 0; median[i] = 0.0;
            continue;
        }
        start += 1;
        double hi = p[start], lo = p[start], tv = 0.0, td = 0.0;
        int64_t cnt = end - start + 1;
        double *sizes = (double *)malloc(sizeof(double) * (size_t)(cnt > 0 ? cnt : 1));
        int64_t k = 0;
        for (int64_t j = start; j <= end; j++) {
            double pr = p[j], vv = v[j];
            sizes[k++] = vv;
            if (pr > hi) hi = pr;
            if (pr < lo) lo = pr;
            tv += vv;
            td += pr * vv;
        }
        o[i] = p[start]; c[i] = p[end]; h[i] = hi; l[i] = lo;
        vol[i] = (float)tv;
        vwap[i] = tv > 0 ? td / tv : 0.0;
        trades[i] = cnt;
        if (cnt > 0) {
            nth_element_f64(sizes, cnt, cnt / 2);
            if ((cnt & 1) == 0) {                                   /* numba _median_inner: (a[n/2-1] + a[n/2]) / 2 */
                double lowmid = sizes[0];
                for (int64_t q = 1; q < cnt / 2; q++) if (sizes[q] > lowmid) lowmid = sizes[q];
                median[i] = (lowmid + sizes[cnt / 2]) / 2;
            } else median[i] = sizes[cnt / 2];
        } else median[i] = 0.0;
        free(sizes);
    }
    return FMKO_OK;
}

/* ---- a8: bar/base.py:409-546 comp_bar_directional_features ---------------------------------- */
int fmko_bar_directional(const double *p, const double *v, int64_t n, const int64_t *ci, int64_t nci,
                         const int8_t *side,
                         int64_t *ticks_buy, int64_t *ticks_sell, float *volume_buy, float *volume_sell,
                         float *dollars_buy, float *dollars_sell, float *mean_spread, float *max_spread,
                         int64_t *cum_ticks_min, int64_t *cum_ticks_max, float *cum_volume_min,
                         float *cum_volume_max, float *cum_dollars_min, float *cum_dollars_max) {
    int64_t nb = nci - 1;
    #pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < nb; i++) {
        int64_t start = ci[i] + 1, end = ci[i + 1];
        int64_t tb = 0, tsell = 0, cticks = 0;
        double vb = 0, vs = 0, db = 0, ds = 0, cvol = 0, cdol = 0, maxsp = 0, cumsp = 0;
        int64_t ctmin = 1000000000LL, ctmax = -1000000000LL;
        float cvmin = 1e9f, cvmax = -1e9f, cdmin = 1e9f, cdmax = -1e9f;
        int prev = 0;
        if (end > start) {
            int64_t q = start - 1; if (q < 0) q += n;       /* negative index wraps (H3) */
            prev = side[q];
        }
        for (int64_t j = start; j <= end; j++) {
            int cur = side[j];
            if (cur != prev) {
                int64_t q = j - 1; if (q < 0) q += n;
                double sp = fabs(p[j] - p[q]);
                if (sp > maxsp) maxsp = sp;
                cumsp += sp;
            }
            prev = cur;
            if (cur == 1) {
                tb++; vb += v[j]; db += p[j] * v[j];
                cticks += 1; cvol += v[j]; cdol += p[j] * v[j];
            } else if (cur == -1) {
                tsell++; vs += v[j]; ds += p[j] * v[j];
                cticks -= 1; cvol -= v[j]; cdol -= p[j] * v[j];
            } else continue;
            if (cticks > ctmax) ctmax = cticks;
            if (cticks < ctmin) ctmin = cticks;
            /* the reference does max(float32 array element, float64) and stores back as float32 */
            { double a = (double)cvmax; double r = a > cvol ? a : cvol; cvmax = (float)r; }
            { double a = (double)cvmin; double r = a < cvol ? a : cvol; cvmin = (float)r; }
            { double a = (double)cdmax; double r = a > cdol ? a : cdol; cdmax = (float)r; }
            { double a = (double)cdmin; double r = a < cdol ? a : cdol; cdmin = (float)r; }
        }
        ticks_buy[i] = tb; ticks_sell[i] = tsell;
        volume_buy[i] = (float)vb; volume_sell[i] = (float)vs;
        dollars_buy[i] = (float)db; dollars_sell[i] = (float)ds;
        max_spread[i] = (float)maxsp;
        mean_spread[i] = (float)(cumsp / (double)(tb + tsell));   /* 0/0 -> NaN (measured, no exception) */
        cum_ticks_min[i] = ctmin; cum_ticks_max[i] = ctmax;
        cum_volume_min[i] = cvmin; cum_volume_max[i] = cvmax;
        cum_dollars_min[i] = cdmin; cum_dollars_max[i] = cdmax;
    }
    return FMKO_OK;
}

/* ---- a9: bar/base.py:549-612 comp_bar_trade_size_features ------------------------------------ */
int fmko_bar_trade_size(const double *a, int64_t n, const double *theta, int64_t ntheta,
                        const int64_t *ci, int64_t nci, double theta_mult,
                        float *mean_size_rel, float *size_95_rel, float *pct_block, float *size_gini) {
    (void)n;
    if (ntheta != nci - 1) return FMKO_ERR_THETA;
    int64_t nb = nci - 1;
    #pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < nb; i++) {
        mean_size_rel[i] = size_95_rel[i] = pct_block[i] = size_gini[i] = NAN;
        int64_t start = ci[i] + 1, end = ci[i + 1];
        if (start > end) continue;
        if (theta[i] == 0.0) continue;
        double thr = theta[i] * theta_mult;
        int64_t cnt = end - start + 1;
        const double *ab = a + start;
        double s = 0.0;
        for (int64_t j = 0; j < cnt; j++) s += ab[j];
        mean_size_rel[i] = (float)log1p((s / (double)cnt) / thr);
        /* numba np.percentile: rank = 1 + (n-1)*q/100; linear interpolation between closest ranks */
        double pv;
        if (cnt == 1) pv = ab[0];
        else {
            double *tmp = (double *)malloc(sizeof(double) * (size_t)cnt);
            memcpy(tmp, ab, sizeof(double) * (size_t)cnt);
            double rank = 1 + (double)(cnt - 1) * (95.0 / 100.0);
            double f = floor(rank), m = rank - f;
            int64_t k = (int64_t)(f - 1);
            nth_element_f64(tmp, cnt, k);
            double lower = tmp[k], upper = tmp[k];
            if (k + 1 < cnt) {                                      /* (k+1)-th statistic = min of the right part */
                upper = tmp[k + 1];
                for (int64_t q = k + 2; q < cnt; q++) if (tmp[q] < upper) upper = tmp[q];
            }
            pv = lower * (1 - m) + upper * m;
            free(tmp);
        }
        size_95_rel[i] = (float)log1p(pv / thr);
        double total = s;
        if (total == 0) continue;
        double block = 0.0;
        for (int64_t j = 0; j < cnt; j++) if (ab[j] > thr) block += ab[j];
        pct_block[i] = (float)(block / total);
        if (cnt == 1) size_gini[i] = 0.0f;
        else {
            double g = 0.0;
            for (int64_t j = 0; j < cnt; j++) { double r = ab[j] / total; g += r * r; }
            size_gini[i] = (float)(1.0 - g);
        }
    }
    return FMKO_OK;
}

/* Python round(): round-half-even of a double -> rint() in the default rounding mode. */
static inline int64_t py_round_i64(double x) { return (int64_t)rint(x); }

/* ---- a11: bar/base.py:755-850 comp_footprint_features (one bar, L levels) -------------------- */
static void footprint_features(const int32_t *levels, const float *buy, const float *sell, int64_t L, double factor,
                               uint8_t *buy_imb, uint8_t *sell_imb, int16_t *run_signed, int32_t *cot,
                               double *vp_skew, double *vp_gini) {
    for (int64_t k = 0; k < L; k++) { buy_imb[k] = 0; sell_imb[k] = 0; }
    if (L > 1) {
        for (int64_t k = 0; k + 1 < L; k++) sell_imb[k] = (double)sell[k] > ((double)buy[k + 1] * factor);
        for (int64_t k = 1; k < L; k++) buy_imb[k] = (double)buy[k] > ((double)sell[k - 1] * factor);
    }
    int64_t max_run = 0, max_sign = 0, run = 0, run_sign = 0;
    for (int64_t k = 0; k < L; k++) {
        int sign = buy_imb[k] ? 1 : (sell_imb[k] ? -1 : 0);
        if (sign != 0 && sign == run_sign) run += 1;
        else if (sign != 0) { run = 1; run_sign = sign; }
        else { run = 0; run_sign = 0; }
        if (run > max_run) { max_run = run; max_sign = run_sign; }
    }
    *run_signed = (int16_t)(max_run * max_sign);
    /* total_volumes = buy + sell in float32; .sum() accumulates in float32 (numba: accumulator of the array dtype) */
    float sumtot = 0.0f;
    int64_t arg = 0; float best = 0.0f;
    for (int64_t k = 0; k < L; k++) {
        float t = buy[k] + sell[k];
        sumtot += t;
        if (k == 0 || t > best) { best = t; arg = k; }
    }
    *cot = levels[arg];
    *vp_skew = 0.0; *vp_gini = 0.0;
    if (sumtot > 0 && L > 0) {
        /* Numba types every intermediate here as float32 (int32 * float32 -> float32; measured with
         * nopython_signatures in the build container): vwap, deviations, dot and the gini terms are float32. */
        float num = 0.0f;
        for (int64_t k = 0; k < L; k++) num += (float)levels[k] * (float)(buy[k] + sell[k]);
        float vw = num / sumtot;
        /* np.dot(float32, float32) is BLAS sdot: its accumulation order is implementation-defined, and the
         * mathematical value is 0 -- vp_skew is rounding noise, compared with an absolute tolerance. */
        float dot = 0.0f;
        for (int64_t k = 0; k < L; k++) dot += ((float)levels[k] - vw) * (float)(buy[k] + sell[k]);
        *vp_skew = (double)(dot / sumtot);
        float g = 0.0f;
        for (int64_t k = 0; k < L; k++) { float r = (float)(buy[k] + sell[k]) / sumtot; g += r * r; }
        *vp_gini = 1.0 - (double)g;
    }
}

/* ---- a10: bar/base.py:615-752 comp_bar_footprints -> CSR --------------------------------------
 * Phase 1 (levels==NULL): fills level_offsets[nb+1] and returns total number of levels.
 * Phase 2: fills the flat per-level arrays and the per-bar statistics. */
int64_t fmko_bar_footprints(const double *p, const double *a, int64_t n, const int64_t *ci, int64_t nci,
                            const int8_t *side, double tick, const double *lows, const double *highs,
                            double factor, int64_t *level_offsets,
                            int32_t *levels, float *buy_vol, float *sell_vol, int32_t *buy_ticks, int32_t *sell_ticks,
                            uint8_t *buy_imb, uint8_t *sell_imb,
                            uint16_t *buy_imb_sum, uint16_t *sell_imb_sum, int32_t *cot, int16_t *run_signed,
                            double *vp_skew, double *vp_gini) {
    (void)n;
    int64_t nb = nci - 1;
    level_offsets[0] = 0;
    for (int64_t i = 0; i < nb; i++) {
        int64_t low = py_round_i64(lows[i] / tick), high = py_round_i64(highs[i] / tick);
        level_offsets[i + 1] = level_offsets[i] + (high - low + 1);
    }
    if (levels == NULL) return level_offsets[nb];
    int err = 0;
    #pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < nb; i++) {
        int64_t start = ci[i] + 1, end = ci[i + 1];
        int64_t low = py_round_i64(lows[i] / tick);
        int64_t off = level_offsets[i], L = level_offsets[i + 1] - off;
        for (int64_t k = 0; k < L; k++) {
            levels[off + k] = (int32_t)(low + k);
            buy_vol[off + k] = sell_vol[off + k] = 0.0f;
            buy_ticks[off + k] = sell_ticks[off + k] = 0;
        }
        for (int64_t j = start; j <= end; j++) {
            int64_t lv = py_round_i64(p[j] / tick) - low;
            if (lv >= 0 && lv < L) {
                /* float32 element += float64 amount: computed in float64, stored as float32 */
                if (side[j] == 1) { buy_vol[off + lv] = (float)((double)buy_vol[off + lv] + a[j]); buy_ticks[off + lv] += 1; }
                else if (side[j] == -1) { sell_vol[off + lv] = (float)((double)sell_vol[off + lv] + a[j]); sell_ticks[off + lv] += 1; }
            } else {
                #pragma omp atomic write
                err = 1;
            }
        }
        footprint_features(levels + off, buy_vol + off, sell_vol + off, L, factor,
                           buy_imb + off, sell_imb + off, &run_signed[i], &cot[i], &vp_skew[i], &vp_gini[i]);
        uint16_t bs = 0, ss = 0;
        for (int64_t k = 0; k < L; k++) { bs += buy_imb[off + k]; ss += sell_imb[off + k]; }
        buy_imb_sum[i] = bs; sell_imb_sum[i] = ss;
    }
    if (err) return FMKO_ERR_LEVEL;
    return level_offsets[nb];
}

/* ---- a13: feature/core/utils.py:12-64 comp_lagged_returns ------------------------------------ */
int fmko_lagged_returns(const int64_t *ts, const double *close, int64_t n, double window_sec, int is_log, double *out) {
    if (window_sec <= 0) return FMKO_ERR_WINDOW;
    for (int64_t i = 0; i < n; i++) out[i] = NAN;
    if (n == 0) return FMKO_OK;
    double w = window_sec * 1e9;
    int64_t start = ss_left_f64(ts, n, (double)ts[0] + w);
    #pragma omp parallel for schedule(static)
    for (int64_t i = start; i < n; i++) {
        double target = (double)ts[i] - w;
        int64_t lag = ss_right_f64(ts, n, target) - 1;
        if (lag >= 0 && lag < i) {
            if (close[lag] != 0.0) {
                out[i] = is_log ? log(close[i] / close[lag]) : close[i] / close[lag] - 1.0;
            } else out[i] = INFINITY;
        } else out[i] = NAN;
    }
    return FMKO_OK;
}

/* ---- a14: feature/core/volatility.py:139-219 ewmst ------------------------------------------- */
int fmko_ewmst(const int64_t *ts, const double *y, int64_t n, double half_life, double sigma_floor, double *out) {
    if (n == 0) return FMKO_OK;
    double V = 0, V2 = 0, Sy = 0, Syy = 0;
    int64_t last = ts[0];
    out[0] = NAN;
    for (int64_t i = 1; i < n; i++) {
        double dt = (double)(ts[i] - last) / 1e9;
        last = ts[i];
        double alpha = 1.0 - exp(-dt / half_life);
        double om = 1.0 - alpha;
        double yi = y[i];
        V = alpha + om * V;
        V2 = alpha * alpha + (om * om) * V2;
        if (isnan(yi)) { Sy = om * Sy; Syy = om * Syy; }
        else { Sy = alpha * yi + om * Sy; Syy = alpha * yi * yi + om * Syy; }
        if (V > 0.0) {
            double mean = Sy / V, e2 = Syy / V;
            double var_raw = e2 - mean * mean;
            double denom = V - (V2 / V);
            double var = (denom > 0.0 && var_raw > 0.0) ? var_raw * (V / denom) : 0.0;
            double sg = sqrt(var);
            if (sg < sigma_floor) sg = sigma_floor;
            out[i] = sg;
        } else out[i] = NAN;
    }
    return FMKO_OK;
}

/* ---- a16: label/tbm.py:11-158 triple_barrier --------------------------------------------------
 * side == NULL -> side prediction (labels -1/+1); else meta labels (0/1).
 * Skipped events (t1_idx <= t0_idx): label 0, ret NaN, ratio NaN, touch_idx = t0_idx (the reference leaves it
 * uninitialised -- SURVEY H10; excluded from bit-compares). */
int fmko_triple_barrier(const int64_t *ts, const double *close, int64_t n, int64_t nclose,
                        const int64_t *event_idx, const double *targets, int64_t ne, int64_t ntargets,
                        double bottom_mult, double top_mult, double vertical_s, double min_close_s,
                        const int8_t *side, int64_t nside, double min_ret,
                        int8_t *labels, int64_t *touch_idx, double *rets, double *ratios) {
    if (vertical_s <= 0) return FMKO_ERR_VERTICAL;
    if (min_ret < 0) return FMKO_ERR_MINRET;
    if (n != nclose) return FMKO_ERR_LEN;
    if (ne != ntargets) return FMKO_ERR_LEN;
    if (ne == 0) return FMKO_ERR_EMPTY;
    int is_meta = side != NULL;
    if (is_meta && nside != ne) return FMKO_ERR_LEN;
    double vert_ns = vertical_s * 1e9, minc_ns = min_close_s * 1e9;
    double *lc = (double *)malloc(sizeof(double) * (size_t)n);
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) lc[i] = log(close[i]);
    #pragma omp parallel for schedule(dynamic, 4)
    for (int64_t e = 0; e < ne; e++) {
        labels[e] = 0; rets[e] = NAN; ratios[e] = NAN;
        int64_t t0i = event_idx[e];
        touch_idx[e] = t0i;
        double tgt = targets[e];
        double upper = tgt * top_mult, lower = -tgt * bottom_mult;
        int uv = isfinite(upper) && upper != 0.0, lv = isfinite(lower) && lower != 0.0;
        int64_t t0 = ts[t0i];
        double t1 = (double)t0 + vert_ns;
        int64_t t1i = ss_right_f64(ts, n, t1) - 1;
        if (t1i <= t0i) continue;
        double sm = is_meta ? (double)side[e] : 1.0;
        int64_t touch = t1i;
        double mu = 0.0, ml = 0.0, base = lc[t0i], ret = 0.0;
        for (int64_t j = t0i + 1; j <= t1i; j++) {
            int64_t dur = ts[j] - t0;
            if ((double)dur < minc_ns) continue;
            ret = (lc[j] - base) * sm;
            if (ret > 0.0 && uv) { double r = ret / upper; if (r > mu) mu = r; }
            else if (ret < 0.0 && lv) { double r = ret / lower; if (r > ml) ml = r; }
            if (ret >= upper) { touch = j; break; }
            if (ret <= lower) { touch = j; break; }
        }
        touch_idx[e] = touch;
        rets[e] = ret;
        if (is_meta) labels[e] = ret >= min_ret ? 1 : 0;
        else labels[e] = ret > 0 ? 1 : (ret < 0 ? -1 : 1);
        if (touch == t1i) {
            double r;
            if (ret > 0.) { r = mu / (1 + ml); if (!uv) r = NAN; }
            else { r = ml / (1 + mu); if (!lv) r = NAN; }
            /* python min(r, 1.): returns 1. only if 1. < r, so NaN propagates */
            ratios[e] = (1. < r) ? 1. : r;
        } else ratios[e] = 1.;
    }
    free(lc);
    return FMKO_OK;
}

/* ---- a15: bar-level volatility / order-flow features ------------------------------------------------------------ */
/* feature/core/volatility.py:256-286 realized_vol */
int fmko_realized_vol(const double *r, int64_t n, int64_t window, int is_sample, double *out) {
    for (int64_t i = 0; i < n; i++) out[i] = NAN;
    if (window < 1) return FMKO_OK;
    #pragma omp parallel for schedule(static)
    for (int64_t i = window - 1; i < n; i++) {
        int64_t valid = 0;
        for (int64_t j = i - window + 1; j <= i; j++) if (!isnan(r[j])) valid++;
        if (valid > 1) {
            double s = 0.0;                      /* np.nansum(r_window ** 2): sequential, NaN counted as 0 */
            for (int64_t j = i - window + 1; j <= i; j++) { double x = r[j] * r[j]; if (!isnan(x)) s += x; }
            int64_t div = is_sample ? valid - 1 : valid;
            out[i] = sqrt(s / (double)div);
        }
    }
    return FMKO_OK;
}

/* feature/core/volatility.py:9-69 ewms */
int fmko_ewms(const double *y, int64_t n, int64_t span, double *out) {
    if (span <= 1) { for (int64_t i = 0; i < n; i++) out[i] = NAN; return FMKO_OK; }
    double alpha = 2.0 / ((double)span + 1.0), om = 1.0 - alpha;
    double Sw = 0, Sw2 = 0, Sy = 0, Sy2 = 0;
    for (int64_t t = 0; t < n; t++) {
        double yt = y[t];
        int nan = isnan(yt);
        Sw = om * Sw + (nan ? 0.0 : 1.0);
        Sw2 = (om * om) * Sw2 + (nan ? 0.0 : 1.0);
        if (!nan) { Sy = om * Sy + yt; Sy2 = om * Sy2 + yt * yt; }
        else { Sy = om * Sy; Sy2 = om * Sy2; }
        if (Sw > 0.0) {
            double mean = Sy / Sw;
            double den = Sw - (Sw2 / Sw);
            if (den > 0.0) {
                double var = (Sy2 / Sw - mean * mean) * Sw / den;
                if (!(var > 0.0)) var = (var != var) ? var : 0.0;   /* python max(var, 0.0): 0.0 only if 0.0 > var */
                out[t] = sqrt(var);
            } else out[t] = NAN;
        } else out[t] = NAN;
    }
    return FMKO_OK;
}

/* feature/core/volume.py:610-641 vpin (float32 output) */
int fmko_vpin(const double *vb, const double *vs, int64_t n, int64_t window, float *out) {
    double *bc = (double *)malloc(sizeof(double) * (size_t)(n + 1)), *sc = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *ac = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    int64_t *nf = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    bc[0] = sc[0] = ac[0] = 0.0; nf[0] = 0;
    for (int64_t i = 0; i < n; i++) {
        out[i] = NAN;
        double b = vb[i], s = vs[i];
        int nan = isnan(b) || isnan(s);
        bc[i + 1] = bc[i] + (nan ? 0.0 : b);
        sc[i + 1] = sc[i] + (nan ? 0.0 : s);
        ac[i + 1] = ac[i] + (nan ? 0.0 : fabs(b - s));
        nf[i + 1] = nf[i] + nan;
        if (i >= window - 1 && nf[i + 1] - nf[i + 1 - window] == 0) {
            double tot = (bc[i + 1] - bc[i + 1 - window]) + (sc[i + 1] - sc[i + 1 - window]);
            if (tot > 1e-9) out[i] = (float)((ac[i + 1] - ac[i + 1 - window]) / tot);
        }
    }
    free(bc); free(sc); free(ac); free(nf);
    return FMKO_OK;
}

/* feature/core/volume.py:572-607 comp_flow_acceleration */
int fmko_flow_acceleration(const double *vol, int64_t n, int64_t window, int64_t recent, double *out) {
    const double eps = 1e-12;
    for (int64_t i = 0; i < n; i++) out[i] = NAN;
    if (n < window || recent >= window) return FMKO_OK;
    double *S = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    S[0] = 0.0;
    for (int64_t i = 0; i < n; i++) S[i + 1] = S[i] + vol[i];
    #pragma omp parallel for schedule(static)
    for (int64_t i = window - 1; i < n; i++) {
        double rs = S[i + 1] - S[i + 1 - recent], ps = S[i + 1 - recent] - S[i + 1 - window];
        out[i] = log((rs + eps) / (ps + eps));
    }
    free(S);
    return FMKO_OK;
}

/* ---- sample weights on ticks (label/weights.py) ------------------------------------------------------ */
/* Python slice a[start:stop] normalisation for a length-n array */
static void py_slice(int64_t start, int64_t stop, int64_t n, int64_t *s, int64_t *e) {
    if (start < 0) { start += n; if (start < 0) start = 0; } else if (start > n) start = n;
    if (stop < 0) { stop += n; if (stop < 0) stop = 0; } else if (stop > n) stop = n;
    *s = start; *e = stop > start ? stop : start;
}

/* label/weights.py:7-49 average_uniqueness.  concurrency is int16 and wraps silently like the reference's
 * `concurrency[start:end+1] += 1`; the weight is np.mean(1.0 / slice) = sequential sum / size (inf on a zero entry,
 * NaN for an empty slice where Numba's python error model would raise ZeroDivisionError). */
int fmko_average_uniqueness(int64_t n, const int64_t *ev, const int64_t *touch, int64_t ne, int64_t ntouch,
                            double *weights, int16_t *conc) {
    if (ne != ntouch) return FMKO_ERR_LEN;
    memset(conc, 0, sizeof(int16_t) * (size_t)n);
    for (int64_t i = 0; i < ne; i++) {
        int64_t s, e;
        py_slice(ev[i], touch[i] + 1, n, &s, &e);
        for (int64_t j = s; j < e; j++) conc[j] = (int16_t)(uint16_t)((uint16_t)conc[j] + 1u);
    }
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < ne; i++) {
        int64_t s, e;
        py_slice(ev[i], touch[i] + 1, n, &s, &e);
        double c = 0.0;
        for (int64_t j = s; j < e; j++) c += 1.0 / (double)conc[j];
        weights[i] = (e > s) ? c / (double)(e - s) : NAN;
    }
    return FMKO_OK;
}

/* label/weights.py:52-103 return_attribution.  Returns 1 when normalize is set and the sum of weights is <= 0
 * (the reference raises ValueError("Sum of weights is zero or negative, cannot normalize.")). */
int fmko_return_attribution(const int64_t *ev, const int64_t *touch, int64_t ne, const double *close, const int16_t *conc,
                            int64_t n, int normalize, double *weights) {
    double *lr = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (n > 0) lr[0] = NAN;
    #pragma omp parallel for schedule(static)
    for (int64_t i = 1; i < n; i++) lr[i] = close[i - 1] != 0.0 ? log(close[i] / close[i - 1]) : NAN;
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < ne; i++) {
        double w = 0.0;
        for (int64_t j = ev[i]; j <= touch[i]; j++)
            if (conc[j] > 0 && !isnan(lr[j])) w += lr[j] / (double)conc[j];
        weights[i] = fabs(w);
    }
    free(lr);
    if (normalize) {
        double s = 0.0;
        for (int64_t i = 0; i < ne; i++) s += weights[i];
        if (s <= 0.) return 1;
        double f = (double)ne / s;
        for (int64_t i = 0; i < ne; i++) weights[i] *= f;
    }
    return FMKO_OK;
}

/* ---- event sampler and ingest scans ------------------------------------------------------------------- */
/* sampling/filters.py:6-70 cusum_filter.  thr has 1 or n elements.  Two-phase (out == NULL counts).  Returns -1 / -2 for
 * the reference's two ValueErrors. */
int64_t fmko_cusum_filter(const double *x, int64_t n, const double *thr, int64_t nthr, int64_t *out, int64_t cap) {
    if (n <= 1) return -1;
    if (nthr != 1 && nthr != n) return -2;
    double sp = 0.0, sn = 0.0;
    int64_t m = 0;
    for (int64_t i = 1; i < n; i++) {
        double ret = log(x[i] / x[i - 1]);
        double t = nthr == 1 ? thr[0] : thr[i];
        double a = sp + ret, b = sn + ret;
        sp = (a > 0.0) ? a : 0.0;            /* python max(0.0, a) */
        sn = (b < 0.0) ? b : 0.0;            /* python min(0.0, b) */
        if (sn < -t) { sn = 0.0; if (out && m < cap) out[m] = i; m++; }
        else if (sp > t) { sp = 0.0; if (out && m < cap) out[m] = i; m++; }
    }
    return m;
}

/* bar/utils.py:12-46 comp_trade_side / comp_trade_side_vector (tick rule) */
int fmko_trade_side_vector(const double *p, int64_t n, int8_t *out) {
    if (n <= 0) return FMKO_OK;
    out[0] = 0;
    int prev = 0;
    double pp = p[0];
    for (int64_t i = 1; i < n; i++) {
        double dp = p[i] - pp;
        if (fabs(dp) > 1e-12) prev = dp > 0 ? 1 : (dp < 0 ? -1 : 0);
        out[i] = (int8_t)prev;
        pp = p[i];
    }
    return FMKO_OK;
}

/* bar/utils.py:263-329 merge_split_trades: inputs ordered by (timestamp, price, side); returns the merged count */
int64_t fmko_merge_split_trades(const int64_t *ts, const double *p, const float *a, const uint8_t *ibm, int64_t n,
                                int64_t *ots, double *op, float *oa, int8_t *oside) {
    if (n <= 0) return 0;
    int64_t m = 0;
    ots[0] = ts[0]; op[0] = p[0]; oa[0] = a[0];
    if (ibm) oside[0] = ibm[0] ? -1 : 1;
    for (int64_t i = 1; i < n; i++) {
        int same = ts[i] == ots[m] && fabs(p[i] - op[m]) < 1e-8;
        if (ibm) same = same && ((ibm[i] != 0) == (oside[m] == -1));
        if (same) oa[m] = oa[m] + a[i];       /* float32 accumulation in arrival order */
        else {
            m++;
            ots[m] = ts[i]; op[m] = p[i]; oa[m] = a[i];
            if (ibm) oside[m] = ibm[i] ? -1 : 1;
        }
    }
    return m + 1;
}

/* ---- rolling volume profile (feature/core/volume.py:133-456) on the CSR footprint ---------------------- */
static int64_t ss_i64(const int64_t *a, int64_t n, int64_t key, int right) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (right ? (a[mid] <= key) : (a[mid] < key)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* comp_poc_hva_lva (volume.py:283-366) + calc_volume_percentage_above_poc (:369-393) on one profile */
static void poc_hva_lva(const int32_t *lv, const float *vol, int64_t n, double va_pct, int32_t *poc_o, int32_t *hva_o,
                        int32_t *lva_o, float *pct_o) {
    float total = 0.0f;                               /* np.sum of a float32 array under Numba: float32 accumulator */
    for (int64_t i = 0; i < n; i++) total += vol[i];
    int64_t pi = 0;
    for (int64_t i = 1; i < n; i++) if (vol[i] > vol[pi]) pi = i;      /* np.argmax: first maximum */
    int32_t poc = lv[pi], hva = poc, lva = poc;
    double va_thrs = (double)total * (va_pct / 100.0);
    double cum = vol[pi];
    int64_t up = pi + 1, dn = pi - 1;
    /* Numba types each re-assignment separately (SSA): the two-level sums are float32 + float32 = float32 and only the
     * merged variable is float64, so the up/down comparison sees float32-rounded sums (measured against the JIT). */
    double cu = 0.0, cd = 0.0;
    if (up < n) { cu = vol[up]; if (up + 1 < n) cu = (float)(vol[up] + vol[up + 1]); }
    if (dn >= 0) { cd = vol[dn]; if (dn - 1 >= 0) cd = (float)(vol[dn] + vol[dn - 1]); }
    while (cum < va_thrs) {
        if (cu > cd) {
            cum += cu; hva = lv[up + 1 < n - 1 ? up + 1 : n - 1]; up += 2;
            cu = -1.0;
            if (up < n) { cu = vol[up]; if (up + 1 < n) cu = (float)(vol[up] + vol[up + 1]); }
        } else if (cu < cd) {
            cum += cd; lva = lv[dn - 1 > 0 ? dn - 1 : 0]; dn -= 2;
            cd = -1.0;
            if (dn >= 0) { cd = vol[dn]; if (dn - 1 >= 0) cd = (float)(vol[dn] + vol[dn - 1]); }
        } else if (cu == cd && cd != -1.0) {
            cum += cu + cd;
            hva = lv[up + 1 < n - 1 ? up + 1 : n - 1]; lva = lv[dn - 1 > 0 ? dn - 1 : 0];
            up += 2; dn -= 2;
            cu = -1.0;
            if (up < n) { cu = vol[up]; if (up + 1 < n) cu = (float)(vol[up] + vol[up + 1]); }
            cd = -1.0;
            if (dn >= 0) { cd = vol[dn]; if (dn - 1 >= 0) cd = (float)(vol[dn] + vol[dn - 1]); }
        } else break;                                   /* "BUG! Stuck in loop" branch of the reference */
    }
    *poc_o = poc; *hva_o = hva; *lva_o = lva;
    double above = 0.0;
    if (!(total <= 0)) {
        for (int64_t i = 0; i < n; i++) if (lv[i] > poc) above += vol[i];
        *pct_o = (above <= 0.0) ? 0.0f : (float)(above / (double)total);
    } else *pct_o = 0.0f;
}

/* volume_profile_rolling (volume.py:396-456).  n_bins <= 0 means None (no bucketing). */
int fmko_volume_profile_rolling(const int64_t *ts, const double *highs, const double *lows, int64_t nb,
                                const int64_t *off, const int32_t *levels, const float *buy, const float *sell,
                                double window_sec, int64_t n_bins, double tick, double va_pct,
                                int32_t *poc, int32_t *hva, int32_t *lva, float *pct) {
    for (int64_t i = 0; i < nb; i++) { poc[i] = hva[i] = lva[i] = 0; pct[i] = 0.0f; }
    if (nb <= 0) return FMKO_OK;
    const int64_t win = (int64_t)(window_sec * 1e9);
    const int64_t first = ss_i64(ts, nb, ts[0] + win, 0);
    #pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = first; i < nb; i++) {
        const int64_t end_ts = ts[i], start_ts = end_ts - win;
        int64_t s = ss_i64(ts, nb, start_ts, 0), e = ss_i64(ts, nb, end_ts, 1);
        if (s == e) s = s - 1 > 0 ? s - 1 : 0;
        double mn = lows[s], mx = highs[s];
        for (int64_t t = s; t < e; t++) { if (lows[t] < mn) mn = lows[t]; if (highs[t] > mx) mx = highs[t]; }
        const int64_t lo = (int64_t)rint(mn / tick), hi = (int64_t)rint(mx / tick);
        const int64_t L = hi - lo + 1;
        if (L <= 0) continue;
        float *ab = (float *)calloc((size_t)L, sizeof(float)), *as = (float *)calloc((size_t)L, sizeof(float));
        int32_t *lv = (int32_t *)malloc(sizeof(int32_t) * (size_t)L);
        for (int64_t k = 0; k < L; k++) lv[k] = (int32_t)(lo + k);
        for (int64_t t = s; t < e; t++)
            for (int64_t k = off[t]; k < off[t + 1]; k++) {
                const int64_t q = (int64_t)levels[k] - lo;      /* searchsorted on the complete integer grid */
                if (q < 0 || q >= L) continue;
                ab[q] = ab[q] + buy[k]; as[q] = as[q] + sell[k];
            }
        float *tot = ab;
        for (int64_t k = 0; k < L; k++) tot[k] = ab[k] + as[k];
        int64_t n = L;
        int32_t *plv = lv; float *pv = tot;
        int32_t *blv = NULL; float *bv = NULL;
        if (n_bins > 0) {
            /* bucket_price_levels (volume.py:208-280) */
            const int64_t range = hi - lo;
            int64_t bw = range / n_bins; if (bw < 1) bw = 1;
            if (bw % 2 == 0) bw += 1;
            int64_t nedges = 0;
            for (int64_t x = lo; x < hi + bw; x += bw) nedges++;
            int64_t nbin = nedges - 1;
            int64_t e0 = lo, e_last = lo + (nedges - 1) * bw;
            int single = 0;
            if (nedges < 2) { single = 1; e_last = hi + 1; }   /* edges = [min, max + 1], n_bins stays len - 1 = 0 */
            /* digitize(max) - 1: leftovers iff the last level falls at/after the last edge */
            int64_t last_idx;
            if (single) last_idx = (hi >= e_last) ? 1 : 0;
            else last_idx = (hi >= e_last) ? nbin : (hi - e0) / bw;
            const int left = last_idx == nbin;
            const int64_t nout = nbin + (left ? 1 : 0);
            blv = (int32_t *)calloc((size_t)(nout > 0 ? nout : 1), sizeof(int32_t));
            bv = (float *)calloc((size_t)(nout > 0 ? nout : 1), sizeof(float));
            for (int64_t b = 0; b < nbin; b++) {
                const int64_t a = e0 + b * bw, c = e0 + (b + 1) * bw;
                int64_t m2 = a + c - 1;                          /* python floor division by 2 */
                blv[b] = (int32_t)(m2 >= 0 ? m2 / 2 : -((-m2 + 1) / 2));
            }
            if (left) blv[nbin] = (int32_t)hi;
            for (int64_t k = 0; k < L; k++) {
                const int64_t x = lo + k;
                int64_t bi;
                if (single) bi = (x >= e_last) ? 1 : (x >= e0 ? 0 : -1);
                else bi = (x >= e_last) ? nbin : (x - e0) / bw;
                if (bi >= 0 && bi < nbin) bv[bi] = bv[bi] + tot[k];
                else if (bi == nbin) bv[nbin] = bv[nbin] + tot[k];
            }
            n = nout; plv = blv; pv = bv;
        }
        if (n > 0) poc_hva_lva(plv, pv, n, va_pct, &poc[i], &hva[i], &lva[i], &pct[i]);
        free(ab); free(as); free(lv); free(blv); free(bv);
    }
    return FMKO_OK;
}

/* ---- a6: tick-imbalance / tick-run bars -- OWN SEMANTICS, PARITY UNPINNED -------------------------------------------
 * The reference only has stubs (bar/logic.py:224-261 raise NotImplementedError), so there is nothing to pin these against:
 * they restate the definitions of DESIGN.md section 5.5 (AFML 2.3.2.1 / 2.3.2.2 in the reference's conventions: int8 +-1
 * sides as b_t, close-index list starting with 0 like logic.py:73-84, accumulators restarting from 0 after a close) as plain
 * sequential loops, and the GPU path is checked against THEM.
 * kind 0: theta += b_t, close when |theta| >= threshold.  kind 1: count buys / sells, close when max >= threshold. */
int64_t fmko_imbalance_bar_indexer(const int8_t *b, int64_t n, double threshold, int kind, int64_t *idx, int64_t cap) {
    int64_t m = 0;
    if (n <= 0) return 0;
    if (idx) { if (m < cap) idx[m] = 0; }
    m++;
    int64_t theta = 0, nb = 0, ns = 0;
    for (int64_t i = 1; i < n; i++) {
        int hit;
        if (kind == 0) {
            theta += b[i] > 0 ? 1 : (b[i] < 0 ? -1 : 0);
            hit = (double)(theta < 0 ? -theta : theta) >= threshold;
        } else {
            if (b[i] > 0) nb++; else if (b[i] < 0) ns++;
            hit = (double)(nb > ns ? nb : ns) >= threshold;
        }
        if (hit) {
            if (idx) { if (m < cap) idx[m] = i; else return FMKO_ERR_CAP; }
            m++;
            theta = 0; nb = 0; ns = 0;
        }
    }
    return m;
}

/* EMA-adaptive tick-imbalance bars (AFML 2.3.2.1): the bar closes at the first tick with |theta| >= E[T] * |E[b]|, where
 * E[T] is the EWMA (adjust=True form of feature/core/ma.py:7-43, span = span_bars) of the lengths of the completed bars and
 * E[b] the same EWMA of their mean tick sign theta_close / T; before the first close E[T] = expected_ticks_init and
 * |E[b]| = expected_imbalance_init.  The threshold is clamped to [thr_min, thr_max] (AFML's definition is known to run away).
 * ewma state (u, v): u = x + (1 - alpha) u, v = 1 + (1 - alpha) v, value u / v. */
int64_t fmko_imbalance_bar_indexer_ema(const int8_t *b, int64_t n, double expected_ticks_init, double expected_imbalance_init,
                                       int64_t span_bars, double thr_min, double thr_max, int64_t *idx, double *thr_out,
                                       int64_t cap) {
    int64_t m = 0;
    if (n <= 0) return 0;
    if (idx) { if (m < cap) idx[m] = 0; }
    m++;
    const double alpha = 2.0 / ((double)span_bars + 1.0), om = 1.0 - alpha;
    double uT = 0.0, vT = 0.0, uB = 0.0, vB = 0.0;
    double ET = expected_ticks_init, EB = expected_imbalance_init;
    double thr = ET * fabs(EB);
    if (thr < thr_min) thr = thr_min;
    if (thr > thr_max) thr = thr_max;
    int64_t theta = 0, last = 0;
    for (int64_t i = 1; i < n; i++) {
        theta += b[i] > 0 ? 1 : (b[i] < 0 ? -1 : 0);
        if ((double)(theta < 0 ? -theta : theta) >= thr) {
            if (idx) { if (m < cap) { idx[m] = i; if (thr_out) thr_out[m] = thr; } else return FMKO_ERR_CAP; }
            m++;
            const double T = (double)(i - last);
            const double mb = (double)theta / T;
            if (vT == 0.0) { uT = T; vT = 1.0; uB = mb; vB = 1.0; }
            else { uT = T + om * uT; vT = 1.0 + om * vT; uB = mb + om * uB; vB = 1.0 + om * vB; }
            ET = uT / vT; EB = uB / vB;
            thr = ET * fabs(EB);
            if (thr < thr_min) thr = thr_min;
            if (thr > thr_max) thr = thr_max;
            theta = 0; last = i;
        }
    }
    return m;
}
