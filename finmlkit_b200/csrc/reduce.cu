// reduce.cu -- per-bar reductions keyed on the close-index array (bar/base.py:303-850 of the reference).
//
// Layout: bar i covers ticks (ci[i], ci[i+1]] of the SoA trade columns.  One warp owns one bar: lanes stride the
// bar's contiguous tick range with coalesced 8-byte loads (4 independent loads in flight per lane per column),
// partial results are combined with warp shuffles.  Bars are independent, so there are no atomics and results are
// deterministic.  When bars are very short (avg < 8 ticks) a thread-per-bar variant is used instead.
// HBM traffic: price + amount read once (16 B/tick) for OHLCV; amount once more (8 B/tick) for the median.
#include <math.h>
#include <stdlib.h>
#include <new>
#include "common.cuh"
#include "scan.cuh"

#define FULL 0xffffffffu

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}
__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(FULL, x, o));
    return x;
}
__device__ __forceinline__ double warp_min(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmin(x, __shfl_xor_sync(FULL, x, o));
    return x;
}
__device__ __forceinline__ int64_t warp_sum_i64(int64_t x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}
__device__ __forceinline__ int64_t wrap_idx(int64_t j, int64_t n) { return j < 0 ? j + n : j; }

// ---------------------------------------------------------------------------------------------------------------
// a7: comp_bar_ohlcv (bar/base.py:306-407) without the median (see k_bar_order_stats)
// ---------------------------------------------------------------------------------------------------------------
struct OhlcvOut {
    double *open, *high, *low, *close, *vwap;
    float *volume;
    int64_t *trades;
};

__device__ __forceinline__ void ohlcv_empty(const OhlcvOut &o, int64_t i, const double *p, int64_t e, int64_t n) {
    const double pe = p[wrap_idx(e, n)];  // base.py:352-361: O=H=L=C=prices[end], rest 0
    o.open[i] = pe; o.high[i] = pe; o.low[i] = pe; o.close[i] = pe;
    o.volume[i] = 0.0f; o.vwap[i] = 0.0; o.trades[i] = 0;
}

__global__ void __launch_bounds__(256) k_bar_ohlcv_warp(const double *__restrict__ p, const double *__restrict__ v,
                                                        const int64_t *__restrict__ ci, int64_t nb, int64_t n,
                                                        OhlcvOut o) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nb; i += nwarps) {
        const int64_t s = ci[i], e = ci[i + 1];
        if (s == e) {
            if (lane == 0) ohlcv_empty(o, i, p, e, n);
            continue;
        }
        const int64_t start = s + 1;
        double hi = -INFINITY, lo = INFINITY, sv = 0.0, sd = 0.0;
        int64_t j = start + lane;
        // 4 independent (price, amount) pairs in flight per lane
        for (; j + 96 <= e; j += 128) {
            double p0 = __ldg(p + j), p1 = __ldg(p + j + 32), p2 = __ldg(p + j + 64), p3 = __ldg(p + j + 96);
            double v0 = __ldg(v + j), v1 = __ldg(v + j + 32), v2 = __ldg(v + j + 64), v3 = __ldg(v + j + 96);
            hi = fmax(fmax(hi, fmax(p0, p1)), fmax(p2, p3));
            lo = fmin(fmin(lo, fmin(p0, p1)), fmin(p2, p3));
            sv += (v0 + v1) + (v2 + v3);
            sd += (p0 * v0 + p1 * v1) + (p2 * v2 + p3 * v3);
        }
        for (; j <= e; j += 32) {
            double pj = __ldg(p + j), vj = __ldg(v + j);
            hi = fmax(hi, pj); lo = fmin(lo, pj);
            sv += vj; sd += pj * vj;
        }
        hi = warp_max(hi); lo = warp_min(lo); sv = warp_sum(sv); sd = warp_sum(sd);
        if (lane == 0) {
            o.open[i] = p[start]; o.close[i] = p[e]; o.high[i] = hi; o.low[i] = lo;
            o.volume[i] = (float)sv;
            o.vwap[i] = sv > 0 ? sd / sv : 0.0;
            o.trades[i] = e - start + 1;
        }
    }
}

// thread-per-bar variant for very short bars (sequential sums: same order as the reference)
__global__ void __launch_bounds__(256) k_bar_ohlcv_thread(const double *__restrict__ p, const double *__restrict__ v,
                                                          const int64_t *__restrict__ ci, int64_t nb, int64_t n,
                                                          OhlcvOut o) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const int64_t s = ci[i], e = ci[i + 1];
    if (s == e) { ohlcv_empty(o, i, p, e, n); return; }
    const int64_t start = s + 1;
    double hi = p[start], lo = p[start], sv = 0.0, sd = 0.0;
    for (int64_t j = start; j <= e; j++) {
        double pj = p[j], vj = v[j];
        if (pj > hi) hi = pj;
        if (pj < lo) lo = pj;
        sv += vj; sd += pj * vj;
    }
    o.open[i] = p[start]; o.close[i] = p[e]; o.high[i] = hi; o.low[i] = lo;
    o.volume[i] = (float)sv;
    o.vwap[i] = sv > 0 ? sd / sv : 0.0;
    o.trades[i] = e - start + 1;
}

// ---------------------------------------------------------------------------------------------------------------
// order statistics per bar: np.median (base.py:401-405) and np.percentile(., 95) (base.py:593) with Numba's
// definitions (numba/np/arraymath.py _median_inner / _collect_percentiles_inner).
//
// One WARP per bar, adaptive MSB radix select on the order-preserving 64-bit key of each amount:
//   diff pass  : OR/AND of the keys still in play -> first bit where they differ (common prefixes and all-equal
//                groups -- heavy on exchange-quantised sizes -- cost one pass, not eight)
//   hist pass  : 1024-bin shared-memory histogram of the 10 bits from that bit down, pick the bucket holding rank k
//   gather     : once <= 32 candidates remain they are compacted into the lanes and ranked by counting
// Every pass also tracks the smallest key ABOVE the selected range, which is the (k+1)-th order statistic when the
// k-th is the largest candidate (median of an even count, percentile interpolation).  The bar's amounts are re-read
// from L1/L2 on each pass (a 1000-tick bar is 8 KB); HBM sees them once.
// ---------------------------------------------------------------------------------------------------------------
constexpr int OS_WARPS = 8;
constexpr int OS_RBITS = 10;                 // radix digit width
constexpr int OS_NBIN = 1 << OS_RBITS;
constexpr int OS_HIST = OS_NBIN + OS_NBIN / 32; // padded histogram words per warp

__device__ __forceinline__ unsigned long long dkey(double x) {  // order-preserving map double -> uint64
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ unsigned long long warp_or64(unsigned long long x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x |= __shfl_xor_sync(FULL, x, o);
    return x;
}
__device__ __forceinline__ unsigned long long warp_and64(unsigned long long x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x &= __shfl_xor_sync(FULL, x, o);
    return x;
}
__device__ __forceinline__ unsigned long long warp_min64(unsigned long long x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long y = __shfl_xor_sync(FULL, x, o);
        x = y < x ? y : x;
    }
    return x;
}

// k-th (0-based) and (k+1)-th smallest of a[0..cnt): warp-cooperative.  hist: OS_HIST words of this warp's shared memory,
// cand: 32 u64 of this warp's shared memory.  If k+1 == cnt, *r1 = *r0.
// have_first: the OR / AND of all keys is already known (the fused OHLCV pass computes it while streaming the bar),
// which saves the first diff pass.
__device__ void warp_select_two(const double *__restrict__ a, int64_t cnt, int64_t k, unsigned *hist,
                                unsigned long long *cand, double *r0, double *r1, bool have_first = false,
                                unsigned long long orv0 = 0ull, unsigned long long andv0 = ~0ull) {
    const int lane = threadIdx.x & 31;
    unsigned long long mask = 0ull, prefix = 0ull;   // keys in play: (key & mask) == prefix
    int64_t kk = k, c = cnt;
    for (;;) {
        const unsigned long long hi_bound = prefix | ~mask;   // largest key of the selected range
        if (c <= 32) {
            // gather the (<= 32) candidates into shared memory -- unordered: equal keys are interchangeable for rank
            // selection -- and, only when the (k+1)-th statistic may lie outside the bucket, the smallest key above it
            const bool need_above = (kk + 1 >= c);
            unsigned long long amin = ~0ull;
            unsigned *cnt_s = hist;             // hist[0] doubles as the append counter
            if (lane == 0) *cnt_s = 0u;
            __syncwarp();
            int64_t j = lane;
            for (; j + 96 < cnt; j += 128) {    // 4 independent loads in flight per lane
                const double x0 = __ldg(a + j), x1 = __ldg(a + j + 32), x2 = __ldg(a + j + 64), x3 = __ldg(a + j + 96);
                const unsigned long long ks[4] = {dkey(x0), dkey(x1), dkey(x2), dkey(x3)};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if ((ks[q] & mask) == prefix) cand[atomicAdd(cnt_s, 1u) & 31u] = ks[q];
                    else if (need_above && ks[q] > hi_bound && ks[q] < amin) amin = ks[q];
                }
            }
            for (; j < cnt; j += 32) {
                const unsigned long long key = dkey(__ldg(a + j));
                if ((key & mask) == prefix) cand[atomicAdd(cnt_s, 1u) & 31u] = key;
                else if (need_above && key > hi_bound && key < amin) amin = key;
            }
            __syncwarp();
            if (need_above) amin = warp_min64(amin);
            const unsigned long long mine = lane < c ? cand[lane] : ~0ull;
            int rank = 0;
            for (int q = 0; q < (int)c; q++) {
                const unsigned long long o = __shfl_sync(FULL, mine, q);
                rank += (o < mine) || (o == mine && q < lane);
            }
            const unsigned b0 = __ballot_sync(FULL, lane < c && rank == (int)kk);
            const unsigned b1 = __ballot_sync(FULL, lane < c && rank == (int)kk + 1);
            const unsigned long long k0 = __shfl_sync(FULL, mine, __ffs(b0) - 1);
            unsigned long long k1 = b1 ? __shfl_sync(FULL, mine, __ffs(b1) - 1) : amin;
            if (k1 == ~0ull && !b1) k1 = k0;   // k is the last element of the bar
            *r0 = dunkey(k0); *r1 = dunkey(k1);
            __syncwarp();
            return;
        }
        // diff pass
        unsigned long long orv = 0ull, andv = ~0ull, amin = ~0ull;
        if (have_first && mask == 0ull) { orv = orv0; andv = andv0; }
        else {
            int64_t j = lane;
            for (; j + 96 < cnt; j += 128) {   // 4 independent loads in flight per lane
                const double x0 = __ldg(a + j), x1 = __ldg(a + j + 32), x2 = __ldg(a + j + 64), x3 = __ldg(a + j + 96);
                const unsigned long long ks[4] = {dkey(x0), dkey(x1), dkey(x2), dkey(x3)};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if ((ks[q] & mask) == prefix) { orv |= ks[q]; andv &= ks[q]; }
                    else if (ks[q] > hi_bound && ks[q] < amin) amin = ks[q];
                }
            }
            for (; j < cnt; j += 32) {
                const unsigned long long key = dkey(__ldg(a + j));
                if ((key & mask) == prefix) { orv |= key; andv &= key; }
                else if (key > hi_bound && key < amin) amin = key;
            }
        }
        orv = warp_or64(orv); andv = warp_and64(andv);
        const unsigned long long diff = orv ^ andv;
        if (diff == 0ull) {     // every key in play is identical
            amin = warp_min64(amin);
            *r0 = dunkey(orv);
            *r1 = (kk + 1 < c) ? dunkey(orv) : (amin == ~0ull ? dunkey(orv) : dunkey(amin));
            return;
        }
        const int hb = 63 - __clzll((long long)diff);
        const int shift = hb >= OS_RBITS - 1 ? hb - (OS_RBITS - 1) : 0;
        for (int b = lane; b < OS_HIST; b += 32) hist[b] = 0u;
        __syncwarp();
        // bin d lives at hist[d + (d >> 5)] (one pad word per 32 bins) so that the per-lane group sums below are
        // bank-conflict free
        {
            int64_t j = lane;
            for (; j + 96 < cnt; j += 128) {
                const double x0 = __ldg(a + j), x1 = __ldg(a + j + 32), x2 = __ldg(a + j + 64), x3 = __ldg(a + j + 96);
                const unsigned long long ks[4] = {dkey(x0), dkey(x1), dkey(x2), dkey(x3)};
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if ((ks[q] & mask) == prefix) {
                        const unsigned d = (unsigned)(ks[q] >> shift) & (OS_NBIN - 1u);
                        atomicAdd(&hist[d + (d >> 5)], 1u);
                    }
            }
            for (; j < cnt; j += 32) {
                const unsigned long long key = dkey(__ldg(a + j));
                if ((key & mask) == prefix) {
                    const unsigned d = (unsigned)(key >> shift) & (OS_NBIN - 1u);
                    atomicAdd(&hist[d + (d >> 5)], 1u);
                }
            }
        }
        __syncwarp();
        // level 1: lane L sums bins [32L, 32L+32); level 2: the owner group's 32 bins, one per lane
        unsigned s = 0;
#pragma unroll 8
        for (int q = 0; q < 32; q++) s += hist[33 * lane + q];
        unsigned inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += y;
        }
        const unsigned exc = inc - s;
        const bool mineb = (unsigned long long)kk >= exc && (unsigned long long)kk < inc;
        const int owner = __ffs(__ballot_sync(FULL, mineb)) - 1;
        const unsigned exc_owner = __shfl_sync(FULL, exc, owner);
        const unsigned bv = hist[33 * owner + lane];
        unsigned inc2 = bv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(FULL, inc2, o);
            if (lane >= o) inc2 += y;
        }
        const unsigned exc2 = inc2 - bv;
        const unsigned long long k2 = (unsigned long long)kk - exc_owner;
        const int sel = __ffs(__ballot_sync(FULL, k2 >= exc2 && k2 < inc2)) - 1;
        const int bsel = owner * 32 + sel;
        const unsigned below = exc_owner + __shfl_sync(FULL, exc2, sel);
        const unsigned csel = __shfl_sync(FULL, bv, sel);
        kk -= below;
        c = csel;
        // new range: common bits above the digit come from any key in play (orv), the digit is bsel
        const unsigned long long above_mask = (shift + OS_RBITS >= 64) ? 0ull : (~0ull << (shift + OS_RBITS));
        mask = above_mask | ((unsigned long long)(OS_NBIN - 1) << shift);
        prefix = (orv & above_mask) | ((unsigned long long)bsel << shift);
        __syncwarp();
    }
}

// mode bit 0: median -> median_out ; bit 1: 95th percentile -> p95_out
__global__ void __launch_bounds__(OS_WARPS * 32) k_bar_order_stats(const double *__restrict__ a,
                                                                   const int64_t *__restrict__ ci, int64_t nb, int mode,
                                                                   double *__restrict__ median_out,
                                                                   double *__restrict__ p95_out) {
    __shared__ unsigned hist_s[OS_WARPS][OS_HIST];
    __shared__ unsigned long long cand_s[OS_WARPS][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nb; i += nwarps) {
        const int64_t start = ci[i] + 1, e = ci[i + 1];
        const int64_t cnt = e - start + 1;
        if (cnt <= 0) {  // empty bar: median 0.0 (base.py:360); the percentile is not evaluated for empty bars
            if (lane == 0 && (mode & 1)) median_out[i] = 0.0;
            continue;
        }
        const double *seg = a + start;
        double r0, r1;
        if (mode & 1) {
            // numba _median_inner: odd -> a[n>>1]; even -> (a[n/2-1] + a[n/2]) / 2
            const int64_t mk = (cnt & 1) ? (cnt >> 1) : (cnt >> 1) - 1;
            warp_select_two(seg, cnt, mk, hist_s[w], cand_s[w], &r0, &r1);
            if (lane == 0) median_out[i] = (cnt & 1) ? r0 : (r0 + r1) / 2;
        }
        if (mode & 2) {
            if (cnt == 1) { if (lane == 0) p95_out[i] = seg[0]; }
            else {
                // numba _collect_percentiles_inner: rank = 1 + (n-1)*q/100; lower*(1-m) + upper*m
                const double rank = 1.0 + (double)(cnt - 1) * (95.0 / 100.0);
                const double f = floor(rank), mfrac = rank - f;
                warp_select_two(seg, cnt, (int64_t)(f - 1.0), hist_s[w], cand_s[w], &r0, &r1);
                if (lane == 0) p95_out[i] = r0 * (1 - mfrac) + r1 * mfrac;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Order statistics on RAW bit patterns.  Trade sizes are never negative, and for non-negative doubles the IEEE bit
// pattern orders like the value, so the per-element key transform (14 % of the fused kernel's instructions) can be
// dropped; the histogram passes then only need the HIGH 32-bit word of each amount (LDG.32, 32-bit integer ops), and
// full 64-bit values are touched only for the <= 32 final candidates.  A bar that holds a negative size / -0.0 (sign
// bit in the OR of the patterns) or whose sizes differ only below the high word takes warp_select_two instead.
//   h1  1024-bin histogram (two 16-bit counters per word, bank-transposed for a conflict-free two-level scan) of the
//       10 bits below the highest differing high-word bit; repeated on the selected bucket while it holds > 32 distinct
//       candidates
//   h2  bucket of <= 32: members fetched by index and ranked in registers; bucket of > 32: OR/AND of the members --
//       exchange-quantised sizes make them identical, which ends the search
//   h3  (k+1)-th statistic above the bucket, when needed: min over the members of the higher buckets
// Returns false when the caller has to fall back to the generic select.  cnt < 65536 (packed counters).
// ---------------------------------------------------------------------------------------------------------------
constexpr int RS_WORDS = 512;
__device__ __forceinline__ unsigned rs_word(unsigned d) {   // lane L of the scan owns bins [32L, 32L+32) = words q*32 + L
    const unsigned w = d >> 1;
    return ((w & 15u) << 5) | (w >> 4);
}

// fused_shift >= 0: the caller already built the first-level histogram while streaming the bar, with that shift (see
// k_bar_ohlcv_median); it is used when it is the shift this bar needs.  *shift_out receives the first-level shift.
__device__ bool warp_select_raw(const double *__restrict__ seg, int ncnt, int kk, bool want1, unsigned long long ro,
                                unsigned long long ra, unsigned *hist, double *r0, double *r1, int fused_shift = -1,
                                int *shift_out = nullptr) {
    const int lane = threadIdx.x & 31;
    const unsigned *segw = reinterpret_cast<const unsigned *>(seg);   // high word of element q: segw[2q + 1]
    unsigned *cidx = hist + RS_WORDS;          // 32 candidate indices
    unsigned *ccnt = hist + RS_WORDS + 32;     // append counter
    unsigned diff_hi = (unsigned)((ro ^ ra) >> 32);
    if (diff_hi == 0u) return false;
    unsigned mask = 0u, prefix = 0u, upper = (unsigned)(ra >> 32);
    for (;;) {
        const int hb = 31 - __clz(diff_hi);
        const int shift = hb >= 9 ? hb - 9 : 0;
        if (mask == 0u && shift_out) *shift_out = shift;
        const bool have_hist = mask == 0u && fused_shift == shift;      // first level already counted during the streaming pass
        __syncwarp();                          // earlier readers / writers of the histogram (previous level, streaming pass) are done
        if (!have_hist) {
#pragma unroll
            for (int q = 0; q < RS_WORDS / 32; q++) hist[q * 32 + lane] = 0u;
            __syncwarp();
        }
        if (!have_hist) {
            int q = lane;
            if (mask == 0u) {                  // first level: every element is in play
                for (; q + 224 < ncnt; q += 256) {      // eight loads in flight per lane: half as many L2 round trips
                    unsigned h[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) h[u] = __ldg(segw + 2 * (q + 32 * u) + 1);
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const unsigned d = (h[u] >> shift) & 1023u;
                        atomicAdd(&hist[rs_word(d)], (d & 1u) ? 65536u : 1u);
                    }
                }
                for (; q + 96 < ncnt; q += 128) {
                    const unsigned h0 = __ldg(segw + 2 * q + 1), h1 = __ldg(segw + 2 * q + 65),
                                   h2 = __ldg(segw + 2 * q + 129), h3 = __ldg(segw + 2 * q + 193);
                    const unsigned d0 = (h0 >> shift) & 1023u, d1 = (h1 >> shift) & 1023u, d2 = (h2 >> shift) & 1023u,
                                   d3 = (h3 >> shift) & 1023u;
                    atomicAdd(&hist[rs_word(d0)], (d0 & 1u) ? 65536u : 1u);
                    atomicAdd(&hist[rs_word(d1)], (d1 & 1u) ? 65536u : 1u);
                    atomicAdd(&hist[rs_word(d2)], (d2 & 1u) ? 65536u : 1u);
                    atomicAdd(&hist[rs_word(d3)], (d3 & 1u) ? 65536u : 1u);
                }
                for (; q < ncnt; q += 32) {
                    const unsigned d = (__ldg(segw + 2 * q + 1) >> shift) & 1023u;
                    atomicAdd(&hist[rs_word(d)], (d & 1u) ? 65536u : 1u);
                }
            } else {
#pragma unroll 4
                for (; q < ncnt; q += 32) {
                    const unsigned h = __ldg(segw + 2 * q + 1);
                    if ((h & mask) == prefix) {
                        const unsigned d = (h >> shift) & 1023u;
                        atomicAdd(&hist[rs_word(d)], (d & 1u) ? 65536u : 1u);
                    }
                }
            }
        }
        __syncwarp();
        unsigned ps = 0;
#pragma unroll
        for (int q = 0; q < RS_WORDS / 32; q++) ps += hist[q * 32 + lane];
        const unsigned ssum = (ps & 0xffffu) + (ps >> 16);
        unsigned inc = ssum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned y = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += y;
        }
        const unsigned exc = inc - ssum;
        const unsigned kq = (unsigned)kk;
        const int owner = __ffs(__ballot_sync(FULL, kq >= exc && kq < inc)) - 1;
        const unsigned exc_owner = __shfl_sync(FULL, exc, owner);
        const unsigned wv = hist[((lane >> 1) << 5) | owner];
        const unsigned bv = (lane & 1) ? (wv >> 16) : (wv & 0xffffu);
        unsigned inc2 = bv;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned y = __shfl_up_sync(FULL, inc2, d);
            if (lane >= d) inc2 += y;
        }
        const unsigned exc2 = inc2 - bv;
        const unsigned k2 = kq - exc_owner;
        const int sel = __ffs(__ballot_sync(FULL, k2 >= exc2 && k2 < inc2)) - 1;
        const unsigned bsel = (unsigned)(owner * 32 + sel);
        const int c1 = (int)__shfl_sync(FULL, bv, sel);
        kk = (int)(k2 - __shfl_sync(FULL, exc2, sel));          // rank inside the bucket
        const unsigned above_mask = (shift + 10 >= 32) ? 0u : (~0u << (shift + 10));
        mask = above_mask | (1023u << shift);
        prefix = (upper & above_mask) | (bsel << shift);
        const bool inside1 = kk + 1 < c1;
        unsigned long long key0 = 0ull, key1 = 0ull;
        bool resolved = false;
        if (c1 <= 32) {
            if (lane == 0) *ccnt = 0u;
            __syncwarp();
            {
                int q = lane;
                for (; q + 224 < ncnt; q += 256) {
                    unsigned h[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) h[u] = __ldg(segw + 2 * (q + 32 * u) + 1);
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if ((h[u] & mask) == prefix) cidx[atomicAdd(ccnt, 1u) & 31u] = (unsigned)(q + 32 * u);
                }
                for (; q + 96 < ncnt; q += 128) {
                    const unsigned h0 = __ldg(segw + 2 * q + 1), h1 = __ldg(segw + 2 * q + 65),
                                   h2 = __ldg(segw + 2 * q + 129), h3 = __ldg(segw + 2 * q + 193);
                    if ((h0 & mask) == prefix) cidx[atomicAdd(ccnt, 1u) & 31u] = (unsigned)q;
                    if ((h1 & mask) == prefix) cidx[atomicAdd(ccnt, 1u) & 31u] = (unsigned)(q + 32);
                    if ((h2 & mask) == prefix) cidx[atomicAdd(ccnt, 1u) & 31u] = (unsigned)(q + 64);
                    if ((h3 & mask) == prefix) cidx[atomicAdd(ccnt, 1u) & 31u] = (unsigned)(q + 96);
                }
                for (; q < ncnt; q += 32)
                    if ((__ldg(segw + 2 * q + 1) & mask) == prefix) cidx[atomicAdd(ccnt, 1u) & 31u] = (unsigned)q;
            }
            __syncwarp();
            const unsigned long long mine =
                lane < c1 ? (unsigned long long)__double_as_longlong(__ldg(seg + cidx[lane])) : ~0ull;
            int rank = 0;
            for (int q = 0; q < c1; q++) {
                const unsigned long long x = __shfl_sync(FULL, mine, q);
                rank += (x < mine) || (x == mine && q < lane);
            }
            const unsigned b0 = __ballot_sync(FULL, lane < c1 && rank == kk);
            key0 = __shfl_sync(FULL, mine, __ffs(b0) - 1);
            if (want1 && inside1) {
                const unsigned b1 = __ballot_sync(FULL, lane < c1 && rank == kk + 1);
                key1 = __shfl_sync(FULL, mine, __ffs(b1) - 1);
            }
            resolved = true;
        } else {
            unsigned long long bo = 0ull, ba = ~0ull;
#pragma unroll 4
            for (int q = lane; q < ncnt; q += 32)
                if ((__ldg(segw + 2 * q + 1) & mask) == prefix) {
                    const unsigned long long key = (unsigned long long)__double_as_longlong(__ldg(seg + q));
                    bo |= key; ba &= key;
                }
            bo = warp_or64(bo); ba = warp_and64(ba);
            if (bo == ba) { key0 = bo; key1 = bo; resolved = true; }
            else {
                upper = (unsigned)(ba >> 32);                   // common bits of the members
                diff_hi = (unsigned)((bo ^ ba) >> 32);
                if (diff_hi == 0u || shift == 0) return false;  // members differ only below the high word
            }
        }
        if (resolved) {
            if (want1 && !inside1) {              // the (k+1)-th statistic is the smallest value above the bucket
                const unsigned hi_bound = prefix | ~mask;
                unsigned long long amin = ~0ull;
#pragma unroll 4
                for (int q = lane; q < ncnt; q += 32)
                    if (__ldg(segw + 2 * q + 1) > hi_bound) {
                        const unsigned long long key = (unsigned long long)__double_as_longlong(__ldg(seg + q));
                        amin = key < amin ? key : amin;
                    }
                key1 = warp_min64(amin);
            }
            *r0 = __longlong_as_double((long long)key0);
            *r1 = __longlong_as_double((long long)key1);
            return true;
        }
    }
}

// Fused comp_bar_ohlcv (the kernel run_ohlcv launches): one warp streams its bar once from HBM (O/H/L/C, sums, OR/AND
// of the raw size patterns), then selects the median on the sizes it has just pulled through L1/L2.
// Occupancy: measured on B200 at 1e9 ticks -- 3 / 4-5 / 6 resident blocks per SM give 5.06 / 4.01 / 4.95 ms: fewer warps
// cannot hide the L2 latency of the select passes, more warps shrink L1 (shared-memory carve-out) and thrash it.
__global__ void __launch_bounds__(OS_WARPS * 32, 5) k_bar_ohlcv_median(const double *__restrict__ p,
                                                                    const double *__restrict__ v,
                                                                    const int64_t *__restrict__ ci, int64_t nb, int64_t n,
                                                                    OhlcvOut o, double *__restrict__ median_out) {
    __shared__ unsigned hist_s[OS_WARPS][OS_HIST];
    __shared__ unsigned long long cand_s[OS_WARPS][32];
    static_assert(RS_WORDS + 33 <= OS_HIST, "raw-select scratch must fit the generic histogram");
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // The first-level histogram of the median select is built DURING the streaming pass, with the digit position the previous
    // bar of this warp needed (sizes of neighbouring bars live in the same binades): when the guess is right -- almost always --
    // the select starts with its bucket scan instead of another pass over the bar through L2 (the kernel is latency-bound on
    // exactly those re-reads: ncu long_scoreboard 7.65 warps per issue in round 1).  A wrong guess costs nothing but the wasted
    // shared-memory atomics: the select then builds the histogram itself, as before.
    int pred_shift = 15;
    unsigned *const fh = hist_s[w];
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nb; i += nwarps) {
        const int64_t s = ci[i], e = ci[i + 1];
        if (s == e) {
            if (lane == 0) { ohlcv_empty(o, i, p, e, n); median_out[i] = 0.0; }
            continue;
        }
        const int64_t start = s + 1;
        const bool fuse = e - start + 1 < 65536;       // packed 16-bit counters
        __syncwarp();
        if (fuse) {
#pragma unroll
            for (int q = 0; q < RS_WORDS / 32; q++) fh[q * 32 + lane] = 0u;
        }
        __syncwarp();
#define FMK_FUSED_COUNT(x)                                                                    \
        {                                                                                     \
            const unsigned d__ = ((unsigned)__double2hiint(x) >> pred_shift) & 1023u;         \
            atomicAdd(&fh[rs_word(d__)], (d__ & 1u) ? 65536u : 1u);                           \
        }
        double hi = -INFINITY, lo = INFINITY, sv = 0.0, sd = 0.0;
        unsigned long long ro = 0ull, ra = ~0ull;
        int64_t j = start + lane;
        for (; j + 96 <= e; j += 128) {
            const double p0 = __ldg(p + j), p1 = __ldg(p + j + 32), p2 = __ldg(p + j + 64), p3 = __ldg(p + j + 96);
            const double v0 = __ldg(v + j), v1 = __ldg(v + j + 32), v2 = __ldg(v + j + 64), v3 = __ldg(v + j + 96);
            if (fuse) { FMK_FUSED_COUNT(v0) FMK_FUSED_COUNT(v1) FMK_FUSED_COUNT(v2) FMK_FUSED_COUNT(v3) }
            // strict compares like the reference (base.py:381-384): a NaN price never replaces the running high / low
            hi = p0 > hi ? p0 : hi; hi = p1 > hi ? p1 : hi; hi = p2 > hi ? p2 : hi; hi = p3 > hi ? p3 : hi;
            lo = p0 < lo ? p0 : lo; lo = p1 < lo ? p1 : lo; lo = p2 < lo ? p2 : lo; lo = p3 < lo ? p3 : lo;
            sv += (v0 + v1) + (v2 + v3);
            sd += (p0 * v0 + p1 * v1) + (p2 * v2 + p3 * v3);
            const unsigned long long k0 = (unsigned long long)__double_as_longlong(v0), k1 = (unsigned long long)__double_as_longlong(v1),
                                     k2 = (unsigned long long)__double_as_longlong(v2), k3 = (unsigned long long)__double_as_longlong(v3);
            ro |= (k0 | k1) | (k2 | k3);
            ra &= (k0 & k1) & (k2 & k3);
        }
        for (; j <= e; j += 32) {
            const double pj = __ldg(p + j), vj = __ldg(v + j);
            if (fuse) FMK_FUSED_COUNT(vj)
            hi = pj > hi ? pj : hi; lo = pj < lo ? pj : lo;
            sv += vj; sd += pj * vj;
            const unsigned long long kj = (unsigned long long)__double_as_longlong(vj);
            ro |= kj; ra &= kj;
        }
#undef FMK_FUSED_COUNT
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double h2 = __shfl_xor_sync(FULL, hi, d), l2 = __shfl_xor_sync(FULL, lo, d);
            hi = h2 > hi ? h2 : hi;
            lo = l2 < lo ? l2 : lo;
        }
        sv = warp_sum(sv); sd = warp_sum(sd);
        ro = warp_or64(ro); ra = warp_and64(ra);
        const int64_t cnt = e - start + 1;
        if (lane == 0) {
            o.open[i] = p[start]; o.close[i] = p[e]; o.high[i] = hi; o.low[i] = lo;
            o.volume[i] = (float)sv;
            o.vwap[i] = sv > 0 ? sd / sv : 0.0;
            o.trades[i] = cnt;
        }
        const bool odd = cnt & 1;
        const int64_t mk = odd ? (cnt >> 1) : (cnt >> 1) - 1;
        double r0, r1;
        bool done = false;
        if (ro == ra) { r0 = r1 = __longlong_as_double((long long)ro); done = true; }      // all sizes identical
        else if (!(ro >> 63) && cnt < 65536) {
            int used = pred_shift;
            done = warp_select_raw(v + start, (int)cnt, (int)mk, !odd, ro, ra, hist_s[w], &r0, &r1, pred_shift, &used);
            pred_shift = used;
        }
        if (!done) {
            __syncwarp();
            warp_select_two(v + start, cnt, mk, hist_s[w], cand_s[w], &r0, &r1);
        }
        if (lane == 0) median_out[i] = odd ? r0 : (r0 + r1) / 2;
    }
}

static int launch_order_stats(fmk_ctx *ctx, const double *a, const int64_t *ci, int64_t nb, int mode, double *med,
                              double *p95) {
    if (nb <= 0) return FMK_OK;
    int64_t blocks = cdiv(nb, OS_WARPS);
    const int64_t maxb = (int64_t)ctx->sm_count * 16;
    if (blocks > maxb) blocks = maxb;
    FMK_LAUNCH(ctx, k_bar_order_stats, (unsigned)blocks, OS_WARPS * 32, 0, a, ci, nb, mode, med, p95);
    return FMK_OK;
}

static int check_index(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix) {
    if (ix->m < 2) return fmk_fail(ctx, FMK_ERR_ARG, "Bar close indices must contain at least two elements.");
    if (ix->n_ticks != t->n) return fmk_fail(ctx, FMK_ERR_ARG, "index was built for a different trades handle");
    // caller-supplied indices (fmk_index_from_host) were validated on upload: an index outside [-1, n) or a decreasing pair
    // would make every per-bar kernel read out of bounds, and an illegal address poisons the whole CUDA context
    if (!ix->sorted)
        return fmk_fail(ctx, FMK_ERR_ARG, "Bar close indices must be non-decreasing and lie in [-1, len(prices)).");
    return FMK_OK;
}

static int run_ohlcv(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, OhlcvOut o, double *median) {
    const int64_t nb = ix->m - 1;
    if (median && t->n / nb >= 8) {
        int64_t blocks = cdiv(nb, OS_WARPS);
        const int64_t maxb = (int64_t)ctx->sm_count * 16;
        if (blocks > maxb) blocks = maxb;
        FMK_LAUNCH(ctx, k_bar_ohlcv_median, (unsigned)blocks, OS_WARPS * 32, 0, t->price, t->amount, ix->close_idx, nb, t->n, o, median);
        return FMK_OK;
    }
    if (t->n / nb >= 8) {
        int64_t warps = nb;
        int64_t blocks = cdiv(warps, 8);
        int64_t maxb = (int64_t)ctx->sm_count * 64;
        if (blocks > maxb) blocks = maxb;
        FMK_LAUNCH(ctx, k_bar_ohlcv_warp, (unsigned)blocks, 256, 0, t->price, t->amount, ix->close_idx, nb, t->n, o);
    } else {
        FMK_LAUNCH(ctx, k_bar_ohlcv_thread, (unsigned)cdiv(nb, 256), 256, 0, t->price, t->amount, ix->close_idx, nb, t->n, o);
    }
    if (median) FMK_TRY(launch_order_stats(ctx, t->amount, ix->close_idx, nb, 1, median, nullptr));
    return FMK_OK;
}

template <typename T>
static int d2h(fmk_ctx *ctx, T *host, const T *dev, int64_t count) {
    if (host && count > 0) FMK_TRY(fmk_copy_d2h(ctx, host, dev, (size_t)count * sizeof(T)));
    return FMK_OK;
}

extern "C" int fmk_bar_ohlcv(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, double *open, double *high,
                             double *low, double *close, float *volume, double *vwap, int64_t *trades, double *median) {
    FMK_ENTER(ctx);
    FMK_TRY(check_index(ctx, t, ix));
    const int64_t nb = ix->m - 1;
    Scratch<double> d(ctx);   // open, high, low, close, vwap, median
    Scratch<float> f(ctx);
    Scratch<int64_t> l(ctx);
    FMK_TRY(d.alloc(6 * nb));
    FMK_TRY(f.alloc(nb));
    FMK_TRY(l.alloc(nb));
    OhlcvOut o{d.p, d.p + nb, d.p + 2 * nb, d.p + 3 * nb, d.p + 4 * nb, f.p, l.p};
    FMK_TRY(run_ohlcv(ctx, t, ix, o, median ? d.p + 5 * nb : nullptr));
    FMK_TRY(d2h(ctx, open, o.open, nb));
    FMK_TRY(d2h(ctx, high, o.high, nb));
    FMK_TRY(d2h(ctx, low, o.low, nb));
    FMK_TRY(d2h(ctx, close, o.close, nb));
    FMK_TRY(d2h(ctx, vwap, o.vwap, nb));
    FMK_TRY(d2h(ctx, median, d.p + 5 * nb, nb));
    FMK_TRY(d2h(ctx, volume, o.volume, nb));
    FMK_TRY(d2h(ctx, trades, o.trades, nb));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

extern "C" int fmk_bar_ohlcv_device(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, int with_median) {
    FMK_ENTER(ctx);
    FMK_TRY(check_index(ctx, t, ix));
    const int64_t nb = ix->m - 1;
    const int64_t need = nb * (6 * 8 + 4 + 8);
    if (ctx->res_cols_bytes < need) {
        // grow with 25 % slack: repeated builds over streams of similar size never reallocate inside a step
        const int64_t want = need + need / 4 + 4096;
        if (ctx->res_cols) { FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->res_cols); }
        ctx->res_cols = nullptr;
        ctx->res_cols_bytes = 0;
        FMK_CUDA(ctx, cudaMalloc(&ctx->res_cols, (size_t)want));
        ctx->res_cols_bytes = want;
    }
    ctx->res_nb = nb;
    double *d = (double *)ctx->res_cols;
    int64_t *l = (int64_t *)(d + 6 * nb);
    float *f = (float *)(l + nb);
    OhlcvOut o{d, d + nb, d + 2 * nb, d + 3 * nb, d + 4 * nb, f, l};
    return run_ohlcv(ctx, t, ix, o, with_median ? d + 5 * nb : nullptr);
}

// ---------------------------------------------------------------------------------------------------------------
// a8: comp_bar_directional_features (bar/base.py:409-546)
// Ordered warp pass: 32 ticks per step, inclusive warp scans of the signed tick/volume/dollar flows with a carry,
// running min/max taken only at ticks whose side is +-1 (side 0 ticks are skipped exactly like the reference).
// ---------------------------------------------------------------------------------------------------------------
struct DirOut {
    int64_t *ticks_buy, *ticks_sell;
    float *volume_buy, *volume_sell, *dollars_buy, *dollars_sell, *mean_spread, *max_spread;
    int64_t *cum_ticks_min, *cum_ticks_max;
    float *cum_volume_min, *cum_volume_max, *cum_dollars_min, *cum_dollars_max;
};

__global__ void __launch_bounds__(256) k_bar_directional(const double *__restrict__ p, const double *__restrict__ v,
                                                         const int8_t *__restrict__ side,
                                                         const int64_t *__restrict__ ci, int64_t nb, int64_t n,
                                                         DirOut o) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nb; i += nwarps) {
        const int64_t start = ci[i] + 1, e = ci[i + 1];
        int64_t tb = 0, tsell = 0;
        double vb = 0, vs = 0, db = 0, ds = 0, maxsp = 0, cumsp = 0;
        int64_t ctmin = 1000000000ll, ctmax = -1000000000ll;
        double cvmin = INFINITY, cvmax = -INFINITY, cdmin = INFINITY, cdmax = -INFINITY;
        int64_t carry_t = 0;
        double carry_v = 0, carry_d = 0;
        int prev_init = 0;
        if (e > start) prev_init = side[wrap_idx(start - 1, n)];
        for (int64_t base = start; base <= e; base += 32) {
            const int64_t j = base + lane;
            const bool act = j <= e;
            int cur = 0, prv = 0;
            double pj = 0, vj = 0, pprev = 0;
            if (act) {
                cur = side[j];
                pj = p[j]; vj = v[j];
                const int64_t q = wrap_idx(j - 1, n);
                prv = (j == start) ? prev_init : (int)side[q];
                if (cur != prv) pprev = p[q];
            }
            if (act && cur != prv) {
                double sp = fabs(pj - pprev);
                maxsp = fmax(maxsp, sp);
                cumsp += sp;
            }
            const int st = (cur == 1) ? 1 : (cur == -1 ? -1 : 0);
            const double dol = pj * vj;
            if (st == 1) { tb++; vb += vj; db += dol; }
            else if (st == -1) { tsell++; vs += vj; ds += dol; }
            // inclusive scans of signed flows
            int64_t it = st;
            double iv = (double)st * vj, id = (double)st * dol;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                int64_t yt = __shfl_up_sync(FULL, it, off);
                double yv = __shfl_up_sync(FULL, iv, off);
                double yd = __shfl_up_sync(FULL, id, off);
                if (lane >= off) { it += yt; iv += yv; id += yd; }
            }
            it += carry_t; iv += carry_v; id += carry_d;
            if (st != 0) {
                ctmax = it > ctmax ? it : ctmax; ctmin = it < ctmin ? it : ctmin;
                cvmax = fmax(cvmax, iv); cvmin = fmin(cvmin, iv);
                cdmax = fmax(cdmax, id); cdmin = fmin(cdmin, id);
            }
            carry_t = __shfl_sync(FULL, it, 31);
            carry_v = __shfl_sync(FULL, iv, 31);
            carry_d = __shfl_sync(FULL, id, 31);
        }
        tb = warp_sum_i64(tb); tsell = warp_sum_i64(tsell);
        vb = warp_sum(vb); vs = warp_sum(vs); db = warp_sum(db); ds = warp_sum(ds);
        maxsp = warp_max(maxsp); cumsp = warp_sum(cumsp);
        cvmax = warp_max(cvmax); cvmin = warp_min(cvmin); cdmax = warp_max(cdmax); cdmin = warp_min(cdmin);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            int64_t a = __shfl_xor_sync(FULL, ctmax, off), b = __shfl_xor_sync(FULL, ctmin, off);
            ctmax = a > ctmax ? a : ctmax; ctmin = b < ctmin ? b : ctmin;
        }
        if (lane == 0) {
            o.ticks_buy[i] = tb; o.ticks_sell[i] = tsell;
            o.volume_buy[i] = (float)vb; o.volume_sell[i] = (float)vs;
            o.dollars_buy[i] = (float)db; o.dollars_sell[i] = (float)ds;
            o.max_spread[i] = (float)maxsp;
            o.mean_spread[i] = (float)(cumsp / (double)(tb + tsell));  // 0/0 -> NaN like the reference
            o.cum_ticks_min[i] = ctmin; o.cum_ticks_max[i] = ctmax;
            // float32(+-1e9) sentinels survive when no +-1 tick was seen (base.py:457-462)
            o.cum_volume_min[i] = (tb + tsell) ? (float)fmin(cvmin, (double)1e9f) : 1e9f;
            o.cum_volume_max[i] = (tb + tsell) ? (float)fmax(cvmax, (double)-1e9f) : -1e9f;
            o.cum_dollars_min[i] = (tb + tsell) ? (float)fmin(cdmin, (double)1e9f) : 1e9f;
            o.cum_dollars_max[i] = (tb + tsell) ? (float)fmax(cdmax, (double)-1e9f) : -1e9f;
        }
    }
}

extern "C" int fmk_bar_directional(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, int64_t *ticks_buy,
                                   int64_t *ticks_sell, float *volume_buy, float *volume_sell, float *dollars_buy,
                                   float *dollars_sell, float *mean_spread, float *max_spread, int64_t *cum_ticks_min,
                                   int64_t *cum_ticks_max, float *cum_volume_min, float *cum_volume_max,
                                   float *cum_dollars_min, float *cum_dollars_max) {
    FMK_ENTER(ctx);
    FMK_TRY(check_index(ctx, t, ix));
    if (!t->side) return fmk_fail(ctx, FMK_ERR_ARG, "trades have no 'side' column");
    const int64_t nb = ix->m - 1;
    Scratch<int64_t> l(ctx);
    Scratch<float> f(ctx);
    FMK_TRY(l.alloc(4 * nb));
    FMK_TRY(f.alloc(10 * nb));
    DirOut o{l.p, l.p + nb, f.p, f.p + nb, f.p + 2 * nb, f.p + 3 * nb, f.p + 4 * nb, f.p + 5 * nb,
             l.p + 2 * nb, l.p + 3 * nb, f.p + 6 * nb, f.p + 7 * nb, f.p + 8 * nb, f.p + 9 * nb};
    int64_t blocks = cdiv(nb, 8);
    int64_t maxb = (int64_t)ctx->sm_count * 64;
    if (blocks > maxb) blocks = maxb;
    FMK_LAUNCH(ctx, k_bar_directional, (unsigned)blocks, 256, 0, t->price, t->amount, t->side, ix->close_idx, nb, t->n, o);
    FMK_TRY(d2h(ctx, ticks_buy, o.ticks_buy, nb)); FMK_TRY(d2h(ctx, ticks_sell, o.ticks_sell, nb));
    FMK_TRY(d2h(ctx, volume_buy, o.volume_buy, nb)); FMK_TRY(d2h(ctx, volume_sell, o.volume_sell, nb));
    FMK_TRY(d2h(ctx, dollars_buy, o.dollars_buy, nb)); FMK_TRY(d2h(ctx, dollars_sell, o.dollars_sell, nb));
    FMK_TRY(d2h(ctx, mean_spread, o.mean_spread, nb)); FMK_TRY(d2h(ctx, max_spread, o.max_spread, nb));
    FMK_TRY(d2h(ctx, cum_ticks_min, o.cum_ticks_min, nb)); FMK_TRY(d2h(ctx, cum_ticks_max, o.cum_ticks_max, nb));
    FMK_TRY(d2h(ctx, cum_volume_min, o.cum_volume_min, nb)); FMK_TRY(d2h(ctx, cum_volume_max, o.cum_volume_max, nb));
    FMK_TRY(d2h(ctx, cum_dollars_min, o.cum_dollars_min, nb)); FMK_TRY(d2h(ctx, cum_dollars_max, o.cum_dollars_max, nb));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// a9: comp_bar_trade_size_features (bar/base.py:549-612)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bar_trade_size(const double *__restrict__ a, const double *__restrict__ theta,
                                                        const int64_t *__restrict__ ci, int64_t nb, double theta_mult,
                                                        const double *__restrict__ p95, float *mean_size_rel,
                                                        float *size_95_rel, float *pct_block, float *size_gini) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float fnan = __int_as_float(0x7fc00000);
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nb; i += nwarps) {
        const int64_t start = ci[i] + 1, e = ci[i + 1];
        float o0 = fnan, o1 = fnan, o2 = fnan, o3 = fnan;
        const double th = theta[i];
        if (start <= e && th != 0.0) {
            const double thr = th * theta_mult;
            const int64_t cnt = e - start + 1;
            double s = 0.0, blk = 0.0;
            for (int64_t j = start + lane; j <= e; j += 32) {
                double x = __ldg(a + j);
                s += x;
                if (x > thr) blk += x;
            }
            s = warp_sum(s); blk = warp_sum(blk);
            o0 = (float)log1p((s / (double)cnt) / thr);
            o1 = (float)log1p(p95[i] / thr);
            if (s != 0.0) {
                o2 = (float)(blk / s);
                if (cnt == 1) o3 = 0.0f;
                else {
                    double g = 0.0;
                    for (int64_t j = start + lane; j <= e; j += 32) {
                        double r = __ldg(a + j) / s;
                        g += r * r;
                    }
                    g = warp_sum(g);
                    o3 = (float)(1.0 - g);
                }
            }
        }
        if (lane == 0) { mean_size_rel[i] = o0; size_95_rel[i] = o1; pct_block[i] = o2; size_gini[i] = o3; }
    }
}

extern "C" int fmk_bar_trade_size(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, const double *theta,
                                  int64_t n_theta, double theta_mult, float *mean_size_rel, float *size_95_rel,
                                  float *pct_block, float *size_gini) {
    FMK_ENTER(ctx);
    FMK_TRY(check_index(ctx, t, ix));
    const int64_t nb = ix->m - 1;
    if (n_theta != nb)
        return fmk_fail(ctx, FMK_ERR_ARG, "Theta should match the the number of bars (len(bar_close_indices) - 1).");
    Scratch<double> d(ctx);
    Scratch<float> f(ctx);
    FMK_TRY(d.alloc(2 * nb));
    FMK_TRY(f.alloc(4 * nb));
    FMK_CUDA(ctx, cudaMemcpyAsync(d.p, theta, (size_t)nb * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_TRY(launch_order_stats(ctx, t->amount, ix->close_idx, nb, 2, nullptr, d.p + nb));
    int64_t blocks = cdiv(nb, 8);
    int64_t maxb = (int64_t)ctx->sm_count * 64;
    if (blocks > maxb) blocks = maxb;
    FMK_LAUNCH(ctx, k_bar_trade_size, (unsigned)blocks, 256, 0, t->amount, d.p, ix->close_idx, nb, theta_mult, d.p + nb,
               f.p, f.p + nb, f.p + 2 * nb, f.p + 3 * nb);
    FMK_TRY(d2h(ctx, mean_size_rel, f.p, nb)); FMK_TRY(d2h(ctx, size_95_rel, f.p + nb, nb));
    FMK_TRY(d2h(ctx, pct_block, f.p + 2 * nb, nb)); FMK_TRY(d2h(ctx, size_gini, f.p + 3 * nb, nb));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// a10/a11: comp_bar_footprints + comp_footprint_features (bar/base.py:615-850) -> CSR on the device
// ---------------------------------------------------------------------------------------------------------------
struct LevelsIn {
    const double *lows, *highs;
    double tick;
    __device__ int64_t operator()(int64_t i) const {
        // int(round(x / tick)): IEEE division then round-half-even
        long long lo = __double2ll_rn(__ddiv_rn(lows[i], tick)), hi = __double2ll_rn(__ddiv_rn(highs[i], tick));
        long long L = hi - lo + 1;
        return L > 0 ? L : 0;
    }
};
struct OffOut {
    int64_t *off;
    __device__ void operator()(int64_t i, int64_t cs) const { off[i + 1] = cs; }
};

// One warp per bar.  Level volumes are float32 accumulated in tick order (base.py:694-717): inside a 32-tick step,
// lanes that hit the same (level, side) bin are serialised in lane (= tick) order by the lowest lane of the group,
// so every bin sees exactly the reference's sequence of `float32(float64(bin) + amount)` updates.
// Bars whose price range spans <= FP_CAP levels (practically all of them) accumulate in per-warp SHARED memory and are
// flushed once: the float32 sums are order-dependent, so every 32-tick step is a read-modify-write of the (level, side)
// cells it touches, and doing that through global memory cost two L2 round trips per step (17.8 ms at 1e9 ticks,
// latency-bound at 32 % issue activity).  Wider bars keep the global-memory path.
// The capacity is a launch parameter (dynamic shared memory): it is sized from the average number of levels per bar of the
// call (a $1M dollar bar of the synthetic stream spans ~300 levels, a 50-BTC volume bar ~440: the fixed 192 of round 1 sent
// most of those bars down the global-memory path).
__global__ void __launch_bounds__(256) k_bar_footprint(const double *__restrict__ p, const double *__restrict__ a,
                                                       const int8_t *__restrict__ side,
                                                       const int64_t *__restrict__ ci, int64_t nb,
                                                       const double *__restrict__ lows, double tick,
                                                       const int64_t *__restrict__ off, int32_t *levels, float *bvol,
                                                       float *svol, int32_t *bt, int32_t *st, int *err, int FP_CAP) {
    extern __shared__ float fp_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float *const vol_w = fp_smem + (size_t)w * 4 * FP_CAP;                 // [2 * level + side]
    int32_t *const cnt_w = reinterpret_cast<int32_t *>(vol_w + 2 * FP_CAP);
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nb; i += nwarps) {
        const int64_t start = ci[i] + 1, e = ci[i + 1];
        const int64_t o = off[i], L = off[i + 1] - o;
        const long long low = __double2ll_rn(__ddiv_rn(lows[i], tick));
        const bool in_smem = L <= FP_CAP;
        __syncwarp();
        if (in_smem) {
            for (int k = lane; k < 2 * (int)L; k += 32) { vol_w[k] = 0.0f; cnt_w[k] = 0; }
        } else {
            for (int64_t k = lane; k < L; k += 32) { bvol[o + k] = 0.0f; svol[o + k] = 0.0f; bt[o + k] = 0; st[o + k] = 0; }
        }
        for (int64_t k = lane; k < L; k += 32) levels[o + k] = (int32_t)(low + k);
        __syncwarp();
        for (int64_t base = start; base <= e; base += 32) {
            const int64_t j = base + lane;
            long long bin = -1;  // 2*level + (sell ? 1 : 0); -1 = no update
            double amt = 0.0;
            if (j <= e) {
                const int sd = side[j];
                const long long lv = __double2ll_rn(__ddiv_rn(p[j], tick)) - low;
                if (lv < 0 || lv >= L) atomicExch(err, 1);
                else if (sd == 1) bin = 2 * lv;
                else if (sd == -1) bin = 2 * lv + 1;
                amt = a[j];
            }
            const unsigned grp = __match_any_sync(FULL, bin);
            const int leader = __ffs(grp) - 1;
            const bool lead = (lane == leader) && bin >= 0;
            float acc = 0.0f;
            int cntk = 0;
            float *vp = nullptr;
            int32_t *tp = nullptr;
            if (lead) {
                if (in_smem) { vp = &vol_w[bin]; tp = &cnt_w[bin]; }
                else {
                    const int64_t lv = o + (bin >> 1);
                    vp = (bin & 1) ? svol + lv : bvol + lv;
                    tp = (bin & 1) ? st + lv : bt + lv;
                }
                acc = *vp;
            }
            // every lane walks the largest group size; shuffles are warp-wide
            unsigned rem = grp;
            const unsigned any_multi = __any_sync(FULL, bin >= 0);
            if (any_multi) {
                int maxg = __popc(grp);
                for (int off2 = 16; off2 > 0; off2 >>= 1) maxg = max(maxg, __shfl_xor_sync(FULL, maxg, off2));
                for (int r = 0; r < maxg; r++) {
                    int src = rem ? (__ffs(rem) - 1) : lane;
                    double x = __shfl_sync(FULL, amt, src);
                    if (lead && rem) { acc = (float)((double)acc + x); cntk++; }
                    if (rem) rem &= rem - 1;
                }
                if (lead) { *vp = acc; *tp += cntk; }
            }
            __syncwarp();
        }
        if (in_smem) {
            for (int k = lane; k < (int)L; k += 32) {
                bvol[o + k] = vol_w[2 * k]; svol[o + k] = vol_w[2 * k + 1];
                bt[o + k] = cnt_w[2 * k]; st[o + k] = cnt_w[2 * k + 1];
            }
        }
    }
}

// one thread per bar: comp_footprint_features (base.py:755-850) with Numba's float32 typing of every intermediate
// comp_footprint_features (base.py:755-850).  The float32 sums (np.sum / np.dot of float32 arrays) are sequential per bar,
// so a THREAD owns a bar -- but a bar has hundreds of levels, and 32 threads walking 32 different CSR segments touch 32
// sectors per load.  The warp therefore stages 32 levels of each of its 32 bars at a time through shared memory with
// coalesced row loads (row r = bar of lane r), and every lane consumes its own row.  Two staged passes: (1) imbalance flags,
// longest signed run, total volume, argmax, sum(level * volume); (2) the vwap-centred dot product and the Gini sum.
constexpr int FF_WARPS = 4;
__global__ void __launch_bounds__(FF_WARPS * 32) k_footprint_features(const int64_t *__restrict__ off, int64_t nb,
                                                                      const int32_t *__restrict__ levels,
                                                                      const float *__restrict__ bvol,
                                                                      const float *__restrict__ svol, double factor,
                                                                      uint8_t *bimb, uint8_t *simb, uint16_t *bsum,
                                                                      uint16_t *ssum, int32_t *cot, int16_t *run_signed,
                                                                      double *vp_skew, double *vp_gini) {
    __shared__ float sb[FF_WARPS][32][35];      // 34 used: one level of look-behind and one of look-ahead for the flags
    __shared__ float ss_[FF_WARPS][32][35];
    __shared__ uint8_t sfl[FF_WARPS][32][36];   // flags staged for coalesced stores: bit 0 buy, bit 1 sell
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i = ((int64_t)blockIdx.x * FF_WARPS + w) * 32 + lane;
    const bool have = i < nb;
    const int64_t o = have ? off[i] : 0;
    const int64_t L = have ? off[i + 1] - o : 0;
    int64_t Lmax = L;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const int64_t y = __shfl_xor_sync(FULL, Lmax, d); Lmax = y > Lmax ? y : Lmax; }
    const int32_t lv0 = (have && L > 0) ? levels[o] : 0;     // levels are lv0 + k (k_bar_footprint writes an arange)
    long long max_run = 0, max_sign = 0, run = 0, run_sign = 0;
    unsigned bs = 0, ssn = 0;
    float sumtot = 0.0f, best = 0.0f, num = 0.0f;
    int64_t arg = 0;
    // ---- pass 1 ----
    for (int64_t k0 = 0; k0 < Lmax; k0 += 32) {
        __syncwarp();
        for (int r = 0; r < 32; r++) {          // row r: levels k0-1 .. k0+32 of lane r's bar -> columns 0 .. 33
            const int64_t orr = __shfl_sync(FULL, o, r), Lr = __shfl_sync(FULL, L, r);
            for (int c = lane; c < 34; c += 32) {
                const int64_t k = k0 - 1 + c;
                const bool ok = k >= 0 && k < Lr;
                sb[w][r][c] = ok ? __ldg(bvol + orr + k) : 0.0f;
                ss_[w][r][c] = ok ? __ldg(svol + orr + k) : 0.0f;
            }
        }
        __syncwarp();
        for (int c = 0; c < 32; c++) {
            const int64_t k = k0 + c;
            unsigned char fl = 0;
            if (k < L) {
                const float bk = sb[w][lane][c + 1], sk = ss_[w][lane][c + 1];
                const bool si = (L > 1 && k + 1 < L) ? ((double)sk > __dmul_rn((double)sb[w][lane][c + 2], factor)) : false;
                const bool bi = (L > 1 && k >= 1) ? ((double)bk > __dmul_rn((double)ss_[w][lane][c], factor)) : false;
                fl = (unsigned char)((bi ? 1 : 0) | (si ? 2 : 0));
                bs += bi; ssn += si;
                const int sign = bi ? 1 : (si ? -1 : 0);
                if (sign != 0 && sign == run_sign) run += 1;
                else if (sign != 0) { run = 1; run_sign = sign; }
                else { run = 0; run_sign = 0; }
                if (run > max_run) { max_run = run; max_sign = run_sign; }
                const float t = __fadd_rn(bk, sk);
                sumtot = __fadd_rn(sumtot, t);
                if (k == 0 || t > best) { best = t; arg = k; }
                num = __fadd_rn(num, __fmul_rn((float)(lv0 + (int32_t)k), t));
            }
            sfl[w][lane][c] = fl;
        }
        __syncwarp();
        for (int r = 0; r < 32; r++) {          // coalesced flag stores, row by row
            const int64_t orr = __shfl_sync(FULL, o, r), Lr = __shfl_sync(FULL, L, r);
            const int64_t k = k0 + lane;
            if (k < Lr) { const unsigned char fl = sfl[w][r][lane]; bimb[orr + k] = fl & 1; simb[orr + k] = (fl >> 1) & 1; }
        }
    }
    if (have) {
        bsum[i] = (uint16_t)bs; ssum[i] = (uint16_t)ssn;
        run_signed[i] = (int16_t)(max_run * max_sign);
        cot[i] = L > 0 ? lv0 + (int32_t)arg : 0;
    }
    // ---- pass 2 ----
    const bool stats = have && sumtot > 0 && L > 0;
    const float vw = stats ? __fdiv_rn(num, sumtot) : 0.0f;
    float dot = 0.0f, g = 0.0f;
    for (int64_t k0 = 0; k0 < Lmax; k0 += 32) {
        __syncwarp();
        for (int r = 0; r < 32; r++) {
            const int64_t orr = __shfl_sync(FULL, o, r), Lr = __shfl_sync(FULL, L, r);
            const int64_t k = k0 + lane;
            const bool ok = k < Lr;
            sb[w][r][lane] = ok ? __ldg(bvol + orr + k) : 0.0f;
            ss_[w][r][lane] = ok ? __ldg(svol + orr + k) : 0.0f;
        }
        __syncwarp();
        if (stats)
            for (int c = 0; c < 32; c++) {
                const int64_t k = k0 + c;
                if (k < L) {
                    const float t = __fadd_rn(sb[w][lane][c], ss_[w][lane][c]);
                    dot = __fadd_rn(dot, __fmul_rn(__fsub_rn((float)(lv0 + (int32_t)k), vw), t));
                    const float rr = __fdiv_rn(t, sumtot);
                    g = __fadd_rn(g, __fmul_rn(rr, rr));
                }
            }
    }
    if (have) {
        vp_skew[i] = stats ? (double)__fdiv_rn(dot, sumtot) : 0.0;
        vp_gini[i] = stats ? 1.0 - (double)g : 0.0;
    }
}

extern "C" void fmk_footprint_free(fmk_ctx *ctx, fmk_footprint *fp) {
    FMK_ENTER(ctx);
    if (!fp) return;
    fmk_dfree(ctx, fp->level_offsets); fmk_dfree(ctx, fp->price_levels);
    fmk_dfree(ctx, fp->buy_vol); fmk_dfree(ctx, fp->sell_vol);
    fmk_dfree(ctx, fp->buy_ticks); fmk_dfree(ctx, fp->sell_ticks);
    fmk_dfree(ctx, fp->buy_imb); fmk_dfree(ctx, fp->sell_imb);
    fmk_dfree(ctx, fp->buy_imb_sum); fmk_dfree(ctx, fp->sell_imb_sum);
    fmk_dfree(ctx, fp->cot); fmk_dfree(ctx, fp->run_signed);
    fmk_dfree(ctx, fp->vp_skew); fmk_dfree(ctx, fp->vp_gini);
    delete fp;
}

extern "C" int64_t fmk_footprint_levels(const fmk_footprint *fp) { return fp->n_levels; }

// level_offsets[0..nb] from device lows / highs (one scan); *total_out = number of (bar, level) rows (host sync)
static int fp_offsets(fmk_ctx *ctx, int64_t nb, const double *dlows, const double *dhighs, double tick,
                      int64_t *level_offsets, int64_t *total_out) {
    FMK_CUDA(ctx, cudaMemsetAsync(level_offsets, 0, 8, ctx->stream));
    FMK_TRY(device_inclusive_scan<int64_t>(ctx, LevelsIn{dlows, dhighs, tick}, OffOut{level_offsets}, nb, (int64_t *)nullptr));
    FMK_CUDA(ctx, cudaMemcpyAsync(total_out, level_offsets + nb, 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

// fills every array of `fp` (pointers already set, level_offsets computed); err_dev: device int, zeroed by the caller
static int fp_fill(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, const double *dlows, double tick, double factor,
                   const fmk_footprint *fp, int *err_dev) {
    const int64_t nb = fp->n_bars;
    int64_t blocks = cdiv(nb, 8);
    int64_t maxb = (int64_t)ctx->sm_count * 64;
    if (blocks > maxb) blocks = maxb;
    // shared-memory capacity (levels per warp), 16 bytes per level and warp.  Measured on B200 at 1e9 ticks (volume bars of ~440
    // levels / dollar bars of ~300): 192 -> 18.1 / 18.1 ms, 320 -> 17.7 / 16.4, 448 -> 16.5 / 15.7, 576 -> 18.4 / 18.3,
    // 832 -> 23.8 / 24.2 ms: beyond 448 the lost occupancy (3 -> 2 blocks per SM) costs more than the shared-memory path saves
    int cap = 448;
    if (const char *e = getenv("FMK_FP_CAP")) cap = atoi(e) > 0 ? atoi(e) : cap;      // profiling override
    const size_t fp_smem_bytes = (size_t)8 * 4 * cap * sizeof(float);
    FMK_CUDA(ctx, cudaFuncSetAttribute(k_bar_footprint, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp_smem_bytes));
    FMK_LAUNCH(ctx, k_bar_footprint, (unsigned)blocks, 256, fp_smem_bytes, t->price, t->amount, t->side, ix->close_idx, nb, dlows, tick,
               fp->level_offsets, fp->price_levels, fp->buy_vol, fp->sell_vol, fp->buy_ticks, fp->sell_ticks, err_dev, cap);
    FMK_LAUNCH(ctx, k_footprint_features, (unsigned)cdiv(nb, FF_WARPS * 32), FF_WARPS * 32, 0, fp->level_offsets, nb, fp->price_levels,
               fp->buy_vol, fp->sell_vol, factor, fp->buy_imb, fp->sell_imb, fp->buy_imb_sum, fp->sell_imb_sum, fp->cot,
               fp->run_signed, fp->vp_skew, fp->vp_gini);
    return FMK_OK;
}

static int fp_check_err(fmk_ctx *ctx, const int *err_dev) {
    int herr = 0;
    FMK_CUDA(ctx, cudaMemcpyAsync(&herr, err_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (herr) return fmk_fail(ctx, FMK_ERR_LEVEL, "Something went wrong! Invalid price level index!");
    return FMK_OK;
}

extern "C" int fmk_bar_footprints(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, double tick,
                                  const double *bar_lows, const double *bar_highs, double factor, fmk_footprint **out) {
    FMK_ENTER(ctx);
    *out = nullptr;
    FMK_TRY(check_index(ctx, t, ix));
    if (!t->side) return fmk_fail(ctx, FMK_ERR_ARG, "trades have no 'side' column");
    if (!(tick > 0)) return fmk_fail(ctx, FMK_ERR_ARG, "price_tick_size must be positive");
    const int64_t nb = ix->m - 1;
    Scratch<double> lh(ctx);
    Scratch<int> err(ctx);
    FMK_TRY(lh.alloc(2 * nb));
    FMK_TRY(err.alloc(1));
    FMK_CUDA(ctx, cudaMemcpyAsync(lh.p, bar_lows, (size_t)nb * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(lh.p + nb, bar_highs, (size_t)nb * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemsetAsync(err.p, 0, sizeof(int), ctx->stream));
    fmk_footprint *fp = new (std::nothrow) fmk_footprint();
    if (!fp) return FMK_ERR_ALLOC;
    memset(fp, 0, sizeof(*fp));
    fp->n_bars = nb;
    int rc = fmk_dalloc(ctx, &fp->level_offsets, nb + 1);
    int64_t total = 0;
    if (!rc) rc = fp_offsets(ctx, nb, lh.p, lh.p + nb, tick, fp->level_offsets, &total);
    fp->n_levels = total;
    if (!rc) rc = fmk_dalloc(ctx, &fp->price_levels, total);
    if (!rc) rc = fmk_dalloc(ctx, &fp->buy_vol, total);
    if (!rc) rc = fmk_dalloc(ctx, &fp->sell_vol, total);
    if (!rc) rc = fmk_dalloc(ctx, &fp->buy_ticks, total);
    if (!rc) rc = fmk_dalloc(ctx, &fp->sell_ticks, total);
    if (!rc) rc = fmk_dalloc(ctx, &fp->buy_imb, total);
    if (!rc) rc = fmk_dalloc(ctx, &fp->sell_imb, total);
    if (!rc) rc = fmk_dalloc(ctx, &fp->buy_imb_sum, nb);
    if (!rc) rc = fmk_dalloc(ctx, &fp->sell_imb_sum, nb);
    if (!rc) rc = fmk_dalloc(ctx, &fp->cot, nb);
    if (!rc) rc = fmk_dalloc(ctx, &fp->run_signed, nb);
    if (!rc) rc = fmk_dalloc(ctx, &fp->vp_skew, nb);
    if (!rc) rc = fmk_dalloc(ctx, &fp->vp_gini, nb);
    if (!rc) rc = fp_fill(ctx, t, ix, lh.p, tick, factor, fp, err.p);
    if (!rc) rc = fp_check_err(ctx, err.p);
    if (rc) { fmk_footprint_free(ctx, fp); return rc; }
    *out = fp;
    return FMK_OK;
}

extern "C" int fmk_footprint_download(fmk_ctx *ctx, const fmk_footprint *fp, int64_t *level_offsets,
                                      int32_t *price_levels, float *buy_volumes, float *sell_volumes,
                                      int32_t *buy_ticks, int32_t *sell_ticks, uint8_t *buy_imbalances,
                                      uint8_t *sell_imbalances, uint16_t *buy_imb_sum, uint16_t *sell_imb_sum,
                                      int32_t *cot_price_level, int16_t *imb_max_run_signed, double *vp_skew,
                                      double *vp_gini) {
    FMK_ENTER(ctx);
    const int64_t nb = fp->n_bars, nl = fp->n_levels;
    FMK_TRY(d2h(ctx, level_offsets, fp->level_offsets, nb + 1));
    FMK_TRY(d2h(ctx, price_levels, fp->price_levels, nl));
    FMK_TRY(d2h(ctx, buy_volumes, fp->buy_vol, nl)); FMK_TRY(d2h(ctx, sell_volumes, fp->sell_vol, nl));
    FMK_TRY(d2h(ctx, buy_ticks, fp->buy_ticks, nl)); FMK_TRY(d2h(ctx, sell_ticks, fp->sell_ticks, nl));
    FMK_TRY(d2h(ctx, buy_imbalances, fp->buy_imb, nl)); FMK_TRY(d2h(ctx, sell_imbalances, fp->sell_imb, nl));
    FMK_TRY(d2h(ctx, buy_imb_sum, fp->buy_imb_sum, nb)); FMK_TRY(d2h(ctx, sell_imb_sum, fp->sell_imb_sum, nb));
    FMK_TRY(d2h(ctx, cot_price_level, fp->cot, nb)); FMK_TRY(d2h(ctx, imb_max_run_signed, fp->run_signed, nb));
    FMK_TRY(d2h(ctx, vp_skew, fp->vp_skew, nb)); FMK_TRY(d2h(ctx, vp_gini, fp->vp_gini, nb));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Device-resident bar frame: build_ohlcv + build_directional_features + build_trade_size_features + build_footprints
// (bar/base.py:132-300) against one index, every output column kept on the device (fmk.h: fmk_frame).
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_copy_theta(const double *__restrict__ med, int64_t nb, double *__restrict__ theta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb) theta[i] = med[i];
}

extern "C" void fmk_frame_free(fmk_ctx *ctx, fmk_frame *f) {
    if (!f) return;
    FMK_ENTER(ctx);
    fmk_dfree(ctx, f->bar_block);
    fmk_dfree(ctx, f->level_block);
    delete f;
}

extern "C" int fmk_frame_info(const fmk_frame *f, int64_t *n_bars, int64_t *n_levels, int64_t *bar_block_bytes,
                              int64_t *level_block_bytes, int64_t *col_offsets) {
    if (n_bars) *n_bars = f->n_bars;
    if (n_levels) *n_levels = f->n_levels;
    if (bar_block_bytes) *bar_block_bytes = f->bar_bytes;
    if (level_block_bytes) *level_block_bytes = f->level_bytes;
    if (col_offsets) for (int k = 0; k < FMK_COL_COUNT; k++) col_offsets[k] = f->col_off[k];
    return FMK_OK;
}

extern "C" int fmk_frame_devptrs(const fmk_frame *f, void **bar_block, void **level_block) {
    if (bar_block) *bar_block = f->bar_block;
    if (level_block) *level_block = f->level_block;
    return FMK_OK;
}

extern "C" int fmk_frame_download(fmk_ctx *ctx, const fmk_frame *f, void *bar_block_host, void *level_block_host) {
    FMK_ENTER(ctx);
    if (bar_block_host && f->bar_bytes > 0) FMK_TRY(fmk_copy_d2h(ctx, bar_block_host, f->bar_block, (size_t)f->bar_bytes));
    if (level_block_host && f->level_bytes > 0) FMK_TRY(fmk_copy_d2h(ctx, level_block_host, f->level_block, (size_t)f->level_bytes));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

static inline int64_t frame_put(int64_t &cur, int64_t bytes) {
    const int64_t o = cur;
    cur += (bytes + 15) / 16 * 16;
    return o;
}

extern "C" int fmk_bar_features_device(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, int flags,
                                       const double *theta, int64_t n_theta, double theta_mult, double tick,
                                       double factor, fmk_frame **out) {
    FMK_ENTER(ctx);
    *out = nullptr;
    FMK_TRY(check_index(ctx, t, ix));
    const int64_t nb = ix->m - 1;
    const bool want_ohlcv = flags & FMK_F_OHLCV, want_med = flags & FMK_F_MEDIAN, want_dir = flags & FMK_F_DIRECTIONAL,
               want_ts = flags & FMK_F_TRADE_SIZE, want_fp = flags & FMK_F_FOOTPRINT;
    if (want_med && !want_ohlcv) return fmk_fail(ctx, FMK_ERR_ARG, "FMK_F_MEDIAN needs FMK_F_OHLCV");
    if ((want_dir || want_fp) && !t->side) return fmk_fail(ctx, FMK_ERR_ARG, "trades have no 'side' column");
    if (want_fp && !want_ohlcv) return fmk_fail(ctx, FMK_ERR_ARG, "FMK_F_FOOTPRINT needs FMK_F_OHLCV (bar lows / highs)");
    if (want_fp && !(tick > 0)) return fmk_fail(ctx, FMK_ERR_ARG, "price_tick_size must be positive");
    if (want_ts && theta && n_theta != nb)
        return fmk_fail(ctx, FMK_ERR_ARG, "Theta should match the the number of bars (len(bar_close_indices) - 1).");
    if (want_ts && !theta && !want_med) return fmk_fail(ctx, FMK_ERR_ARG, "theta == NULL needs FMK_F_MEDIAN");

    fmk_frame *f = new (std::nothrow) fmk_frame();
    if (!f) return FMK_ERR_ALLOC;
    memset(f, 0, sizeof(*f));
    f->n_bars = nb; f->flags = flags;
    for (int k = 0; k < FMK_COL_COUNT; k++) f->col_off[k] = -1;
    // the per-bar block starts with a 512-byte header (magic, n_bars, n_levels, block sizes, flags, column offsets), so a
    // frame that crossed NVLink / PCIe as raw bytes describes itself (fmk.h: FMK_FRAME_HEADER_BYTES)
    int64_t cur = FMK_FRAME_HEADER_BYTES;
    if (ix->close_ts) f->col_off[FMK_COL_CLOSE_TS] = frame_put(cur, nb * 8);
    f->col_off[FMK_COL_CLOSE_IDX] = frame_put(cur, nb * 8);
    if (want_ohlcv) {
        for (int k = FMK_COL_OPEN; k <= FMK_COL_VWAP; k++) f->col_off[k] = frame_put(cur, nb * 8);
        if (want_med) f->col_off[FMK_COL_MEDIAN] = frame_put(cur, nb * 8);
        f->col_off[FMK_COL_TRADES] = frame_put(cur, nb * 8);
        f->col_off[FMK_COL_VOLUME] = frame_put(cur, nb * 4);
    }
    if (want_dir) {
        for (int k = FMK_COL_TICKS_BUY; k <= FMK_COL_CUM_TICKS_MAX; k++) f->col_off[k] = frame_put(cur, nb * 8);
        for (int k = FMK_COL_VOLUME_BUY; k <= FMK_COL_CUM_DOLLARS_MAX; k++) f->col_off[k] = frame_put(cur, nb * 4);
    }
    if (want_ts) for (int k = FMK_COL_MEAN_SIZE_REL; k <= FMK_COL_SIZE_GINI; k++) f->col_off[k] = frame_put(cur, nb * 4);
    if (want_fp) {
        f->col_off[FMK_COL_FP_LEVEL_OFFSETS] = frame_put(cur, (nb + 1) * 8);
        f->col_off[FMK_COL_FP_VP_SKEW] = frame_put(cur, nb * 8);
        f->col_off[FMK_COL_FP_VP_GINI] = frame_put(cur, nb * 8);
        f->col_off[FMK_COL_FP_COT] = frame_put(cur, nb * 4);
        f->col_off[FMK_COL_FP_BUY_IMB_SUM] = frame_put(cur, nb * 2);
        f->col_off[FMK_COL_FP_SELL_IMB_SUM] = frame_put(cur, nb * 2);
        f->col_off[FMK_COL_FP_RUN_SIGNED] = frame_put(cur, nb * 2);
    }
    f->bar_bytes = cur;
    int rc = fmk_dalloc(ctx, &f->bar_block, cur);
    if (rc) { delete f; return rc; }
    char *B = f->bar_block;
#define COLP(T, id) ((T *)(B + f->col_off[id]))
    auto body = [&]() -> int {
        if (ix->close_ts)
            FMK_CUDA(ctx, cudaMemcpyAsync(COLP(int64_t, FMK_COL_CLOSE_TS), ix->close_ts + 1, (size_t)nb * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        FMK_CUDA(ctx, cudaMemcpyAsync(COLP(int64_t, FMK_COL_CLOSE_IDX), ix->close_idx + 1, (size_t)nb * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (want_ohlcv) {
            OhlcvOut o{COLP(double, FMK_COL_OPEN), COLP(double, FMK_COL_HIGH), COLP(double, FMK_COL_LOW), COLP(double, FMK_COL_CLOSE),
                       COLP(double, FMK_COL_VWAP), COLP(float, FMK_COL_VOLUME), COLP(int64_t, FMK_COL_TRADES)};
            FMK_TRY(run_ohlcv(ctx, t, ix, o, want_med ? COLP(double, FMK_COL_MEDIAN) : nullptr));
        }
        if (want_dir) {
            DirOut o{COLP(int64_t, FMK_COL_TICKS_BUY), COLP(int64_t, FMK_COL_TICKS_SELL), COLP(float, FMK_COL_VOLUME_BUY),
                     COLP(float, FMK_COL_VOLUME_SELL), COLP(float, FMK_COL_DOLLARS_BUY), COLP(float, FMK_COL_DOLLARS_SELL),
                     COLP(float, FMK_COL_MEAN_SPREAD), COLP(float, FMK_COL_MAX_SPREAD), COLP(int64_t, FMK_COL_CUM_TICKS_MIN),
                     COLP(int64_t, FMK_COL_CUM_TICKS_MAX), COLP(float, FMK_COL_CUM_VOLUME_MIN), COLP(float, FMK_COL_CUM_VOLUME_MAX),
                     COLP(float, FMK_COL_CUM_DOLLARS_MIN), COLP(float, FMK_COL_CUM_DOLLARS_MAX)};
            int64_t blocks = cdiv(nb, 8);
            const int64_t maxb = (int64_t)ctx->sm_count * 64;
            if (blocks > maxb) blocks = maxb;
            FMK_LAUNCH(ctx, k_bar_directional, (unsigned)blocks, 256, 0, t->price, t->amount, t->side, ix->close_idx, nb, t->n, o);
        }
        if (want_ts) {
            Scratch<double> d(ctx);
            FMK_TRY(d.alloc(2 * nb));
            if (theta) FMK_CUDA(ctx, cudaMemcpyAsync(d.p, theta, (size_t)nb * 8, cudaMemcpyHostToDevice, ctx->stream));
            else FMK_LAUNCH(ctx, k_copy_theta, (unsigned)cdiv(nb, 256), 256, 0, (const double *)COLP(double, FMK_COL_MEDIAN), nb, d.p);
            FMK_TRY(launch_order_stats(ctx, t->amount, ix->close_idx, nb, 2, nullptr, d.p + nb));
            int64_t blocks = cdiv(nb, 8);
            const int64_t maxb = (int64_t)ctx->sm_count * 64;
            if (blocks > maxb) blocks = maxb;
            FMK_LAUNCH(ctx, k_bar_trade_size, (unsigned)blocks, 256, 0, t->amount, (const double *)d.p, ix->close_idx, nb, theta_mult,
                       (const double *)(d.p + nb), COLP(float, FMK_COL_MEAN_SIZE_REL), COLP(float, FMK_COL_SIZE_95_REL),
                       COLP(float, FMK_COL_PCT_BLOCK), COLP(float, FMK_COL_SIZE_GINI));
        }
        if (want_fp) {
            Scratch<int> err(ctx);
            FMK_TRY(err.alloc(1));
            FMK_CUDA(ctx, cudaMemsetAsync(err.p, 0, sizeof(int), ctx->stream));
            int64_t total = 0;
            FMK_TRY(fp_offsets(ctx, nb, COLP(double, FMK_COL_LOW), COLP(double, FMK_COL_HIGH), tick, COLP(int64_t, FMK_COL_FP_LEVEL_OFFSETS), &total));
            f->n_levels = total;
            int64_t lc = 0;
            f->col_off[FMK_COL_FP_PRICE_LEVELS] = frame_put(lc, total * 4);
            f->col_off[FMK_COL_FP_BUY_VOL] = frame_put(lc, total * 4);
            f->col_off[FMK_COL_FP_SELL_VOL] = frame_put(lc, total * 4);
            f->col_off[FMK_COL_FP_BUY_TICKS] = frame_put(lc, total * 4);
            f->col_off[FMK_COL_FP_SELL_TICKS] = frame_put(lc, total * 4);
            f->col_off[FMK_COL_FP_BUY_IMB] = frame_put(lc, total);
            f->col_off[FMK_COL_FP_SELL_IMB] = frame_put(lc, total);
            f->level_bytes = lc;
            FMK_TRY(fmk_dalloc(ctx, &f->level_block, lc));
            char *Lb = f->level_block;
            fmk_footprint v;
            memset(&v, 0, sizeof(v));
            v.n_bars = nb; v.n_levels = total;
            v.level_offsets = COLP(int64_t, FMK_COL_FP_LEVEL_OFFSETS);
            v.price_levels = (int32_t *)(Lb + f->col_off[FMK_COL_FP_PRICE_LEVELS]);
            v.buy_vol = (float *)(Lb + f->col_off[FMK_COL_FP_BUY_VOL]);
            v.sell_vol = (float *)(Lb + f->col_off[FMK_COL_FP_SELL_VOL]);
            v.buy_ticks = (int32_t *)(Lb + f->col_off[FMK_COL_FP_BUY_TICKS]);
            v.sell_ticks = (int32_t *)(Lb + f->col_off[FMK_COL_FP_SELL_TICKS]);
            v.buy_imb = (uint8_t *)(Lb + f->col_off[FMK_COL_FP_BUY_IMB]);
            v.sell_imb = (uint8_t *)(Lb + f->col_off[FMK_COL_FP_SELL_IMB]);
            v.buy_imb_sum = COLP(uint16_t, FMK_COL_FP_BUY_IMB_SUM);
            v.sell_imb_sum = COLP(uint16_t, FMK_COL_FP_SELL_IMB_SUM);
            v.cot = COLP(int32_t, FMK_COL_FP_COT);
            v.run_signed = COLP(int16_t, FMK_COL_FP_RUN_SIGNED);
            v.vp_skew = COLP(double, FMK_COL_FP_VP_SKEW);
            v.vp_gini = COLP(double, FMK_COL_FP_VP_GINI);
            FMK_TRY(fp_fill(ctx, t, ix, COLP(double, FMK_COL_LOW), tick, factor, &v, err.p));
            FMK_TRY(fp_check_err(ctx, err.p));
        }
        return FMK_OK;
    };
#undef COLP
    rc = body();
    if (!rc) {
        int64_t hdr[FMK_FRAME_HEADER_BYTES / 8];
        memset(hdr, 0, sizeof(hdr));
        hdr[0] = FMK_FRAME_MAGIC; hdr[1] = f->n_bars; hdr[2] = f->n_levels; hdr[3] = f->bar_bytes; hdr[4] = f->level_bytes;
        hdr[5] = f->flags; hdr[6] = FMK_COL_COUNT;
        for (int k = 0; k < FMK_COL_COUNT; k++) hdr[8 + k] = f->col_off[k];
        cudaError_t e = cudaMemcpyAsync(f->bar_block, hdr, sizeof(hdr), cudaMemcpyHostToDevice, ctx->stream);   // pageable: staged before return
        if (e != cudaSuccess) rc = fmk_fail(ctx, FMK_ERR_CUDA, cudaGetErrorString(e));
    }
    if (rc) { fmk_frame_free(ctx, f); return rc; }
    *out = f;
    return FMK_OK;
}
