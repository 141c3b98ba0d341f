// ohlcv_conveyor.cuh -- fused comp_bar_ohlcv (bar/base.py:306-407) as a TMA-fed shared-memory conveyor.
// Included by reduce.cu after the order-statistic helpers (dkey, warp_select_two, OhlcvOut, ...).
//
// Why: with a warp per bar reading global memory directly, every bar pays several dependent HBM round trips (index pair,
// five or six 128-tick rounds, candidate fetch) and the order-statistic passes re-read the bar through L1/L2; the kernel
// was issue/latency bound at 44 % of the HBM roofline.  Here the HBM side is decoupled from the bar structure:
//
//   * one persistent CTA per SM owns a contiguous range of ticks (n / gridDim.x) and the bars that END in it;
//   * a producer warp streams that range through a shared-memory ring (12 chunks x 1024 ticks x {price, amount} = 192 KB)
//     with cp.async.bulk (TMA, 8 KB per copy) completing on one mbarrier per chunk -- ~190 KB in flight per SM, no
//     registers, no per-bar round trips;
//   * consumer warps take the CTA's bars round-robin.  A bar is processed ENTIRELY from the ring: the O/H/L/C + sums
//     pass, then the median's radix-select passes on the high 32 bits of the amount keys (LDS.32 + integer ops), the
//     candidate ranking, everything.  HBM sees every tick once; L2 sees it once.
//   * ring slots are recycled in order: each consumer publishes the first tick it still needs; the producer re-arms
//     the slot of chunk g-NCH once min(published) has moved past it.
//
// Exactness: identical arithmetic and summation grouping to k_bar_ohlcv_median_v1 (lane-strided partial sums in groups
// of four, warp tree), so every output column is bit-identical to it; the median is exact by construction.
#pragma once

constexpr int CV_CONSUMERS = 12;
constexpr int CV_THREADS = (CV_CONSUMERS + 1) * 32;
constexpr int CV_CHUNK = 1024;                   // ticks per TMA chunk (8 KB per column)
constexpr int CV_NCH = 12;                       // chunks in the ring
constexpr int CV_RING = CV_CHUNK * CV_NCH;       // 12288 ticks
constexpr int CV_MAXBAR = (CV_NCH - 1) * CV_CHUNK;  // longest bar that is guaranteed to fit
constexpr int CV_WORDS = 512;                    // 1024 bins, two 16-bit counters per word

struct CvWarp {
    unsigned hist[CV_WORDS];
    unsigned cidx[32];
    unsigned ccnt;
    unsigned pad[3];
};
struct CvShared {
    double ring_p[CV_RING];
    double ring_v[CV_RING];
    CvWarp w[CV_CONSUMERS];
    unsigned g_hist[OS_HIST];                    // scratch of the generic fallback (one warp at a time, see g_lock)
    unsigned long long g_cand[32];
    unsigned long long full[CV_NCH];             // mbarriers
    volatile long long pos[CV_CONSUMERS];        // first tick each consumer still needs
    volatile long long issued;                   // chunks armed so far (local chunk numbers < issued)
    long long range[2];                          // bars [bfirst, blast) of this CTA
    unsigned long long next_bar;                 // dynamic assignment: next unprocessed bar of this CTA
    int g_lock;
};

__device__ __forceinline__ unsigned cv_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cv_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cv_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void cv_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cv_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cv_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cv_smem(bar)) : "memory");
}
__device__ __forceinline__ bool cv_mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(cv_smem(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void cv_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(cv_smem(dst)), "l"(src), "r"(bytes), "r"(cv_smem(bar)) : "memory");
}

__device__ __forceinline__ unsigned cv_word(unsigned d) {   // see ms_word: bank-transposed packed histogram
    const unsigned w = d >> 1;
    return ((w & 15u) << 5) | (w >> 4);
}
__device__ __forceinline__ unsigned cv_hkey(unsigned hw) {  // high word of dkey()
    return (hw & 0x80000000u) ? ~hw : (hw | 0x80000000u);
}

// first index i in [0, m) with a[i] >= key (m if none); a is non-decreasing
__device__ __forceinline__ int64_t cv_lower_bound(const int64_t *__restrict__ a, int64_t m, int64_t key) {
    int64_t lo = 0, hi = m;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Bars that do not fit the ring (or whose keys defeat the 32-bit select): the v1 body on global memory, one warp at a
// time through the CTA's generic scratch.
__device__ void cv_generic_select(CvShared *S, const double *seg, int64_t cnt, int64_t mk, double *r0, double *r1,
                                  bool have_first, unsigned long long orv, unsigned long long andv) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) while (atomicCAS(&S->g_lock, 0, 1) != 0) __nanosleep(100);
    __syncwarp();
    warp_select_two(seg, cnt, mk, S->g_hist, S->g_cand, r0, r1, have_first, orv, andv);
    __syncwarp();
    if (lane == 0) { __threadfence_block(); atomicExch(&S->g_lock, 0); }
}

__global__ void __launch_bounds__(CV_THREADS, 1) k_bar_ohlcv_conveyor(const double *__restrict__ p,
                                                                       const double *__restrict__ v,
                                                                       const int64_t *__restrict__ ci, int64_t nb,
                                                                       int64_t n, OhlcvOut o,
                                                                       double *__restrict__ median_out) {
    extern __shared__ __align__(128) unsigned char cv_raw[];
    CvShared *S = reinterpret_cast<CvShared *>(cv_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

    // ---- partition: this CTA owns the bars whose end index lies in [r0, r1) (ends = ci[1..nb], non-decreasing) ----
    {
        int64_t per = (n + gridDim.x - 1) / gridDim.x;
        per = ((per + CV_CHUNK - 1) / CV_CHUNK) * CV_CHUNK;
        if (threadIdx.x < 2) {
            const int64_t b = (int64_t)blockIdx.x + threadIdx.x;
            int64_t r;
            if (b == 0) r = 0;
            else if (b >= gridDim.x || b * per >= n) r = nb;
            else r = cv_lower_bound(ci + 1, nb, b * per);
            S->range[threadIdx.x] = r;
        }
        if (threadIdx.x == 0) {
            for (int k = 0; k < CV_NCH; k++) cv_mbar_init(&S->full[k], 1);
            S->issued = 0;
            S->g_lock = 0;
            S->next_bar = 0ull;
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    const int64_t bfirst = S->range[0], blast = S->range[1];
    if (bfirst >= blast) return;
    const int64_t first_tick = __ldg(ci + bfirst) + 1 < 0 ? 0 : __ldg(ci + bfirst) + 1;
    const int64_t g0 = first_tick / CV_CHUNK;                 // first global chunk of this CTA
    const int64_t base_tick = g0 * CV_CHUNK;
    if (wid < CV_CONSUMERS && lane == 0) S->pos[wid] = first_tick;
    __syncthreads();

    if (wid == CV_CONSUMERS) {
        // =================================== producer ===================================
        if (lane != 0) return;
        const int64_t last_tick = __ldg(ci + blast);          // end of the last owned bar (>= first_tick - 1)
        if (last_tick < first_tick) { S->issued = 0; return; } // only empty bars
        const int64_t g1 = last_tick / CV_CHUNK;
        for (int64_t g = g0; g <= g1; g++) {
            const int64_t gl = g - g0;
            const int slot = (int)(gl % CV_NCH);
            if (gl >= CV_NCH) {
                const long long need = (long long)(g - CV_NCH + 1) * CV_CHUNK;   // every consumer must be past this
                for (;;) {
                    long long mn = 0x7fffffffffffffffll;
#pragma unroll
                    for (int k = 0; k < CV_CONSUMERS; k++) { const long long x = S->pos[k]; mn = x < mn ? x : mn; }
                    if (mn >= need) break;
                }
                __threadfence_block();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            const int64_t t0 = g * CV_CHUNK;
            int64_t cntk = n - t0;
            if (cntk > CV_CHUNK) cntk = CV_CHUNK;
            const unsigned even = (unsigned)(cntk & ~1ll);
            if (cntk & 1) {   // odd last element of the column: plain copy (a bulk copy must be a multiple of 16 bytes)
                S->ring_p[slot * CV_CHUNK + cntk - 1] = p[t0 + cntk - 1];
                S->ring_v[slot * CV_CHUNK + cntk - 1] = v[t0 + cntk - 1];
            }
            if (even) {
                cv_mbar_expect_tx(&S->full[slot], even * 16u);
                cv_bulk_g2s(&S->ring_p[slot * CV_CHUNK], p + t0, even * 8u, &S->full[slot]);
                cv_bulk_g2s(&S->ring_v[slot * CV_CHUNK], v + t0, even * 8u, &S->full[slot]);
            } else {
                cv_mbar_arrive(&S->full[slot]);
            }
            __threadfence_block();
            S->issued = gl + 1;
        }
        // do not leave the CTA with copies in flight: wait for the last ring's worth of chunks
        const int64_t total = g1 - g0 + 1;
        for (int64_t gl = total > CV_NCH ? total - CV_NCH : 0; gl < total; gl++)
            while (!cv_mbar_try_wait(&S->full[gl % CV_NCH], (unsigned)((gl / CV_NCH) & 1))) {}
        return;
    }

    // =================================== consumers ===================================
    CvWarp *W = &S->w[wid];
    const unsigned *rvw = reinterpret_cast<const unsigned *>(S->ring_v);
    // Bars are handed out dynamically (a finished warp takes the next unprocessed bar), which keeps the in-flight window
    // compact: CV_CONSUMERS consecutive bars, so the ring's look-ahead is not eaten by a warp that ran ahead.
    for (;;) {
        unsigned long long tk = 0ull;
        if (lane == 0) tk = atomicAdd(&S->next_bar, 1ull);
        const int64_t i = bfirst + (int64_t)__shfl_sync(FULL, tk, 0);
        if (i >= blast) break;
        const int64_t s = __ldg(ci + i), e = __ldg(ci + i + 1);
        if (s == e) {
            if (lane == 0) { ohlcv_empty(o, i, p, e, n); median_out[i] = 0.0; }
            continue;
        }
        const int64_t start = s + 1;
        const int64_t cnt = e - s;
        const bool odd = cnt & 1;
        const int64_t mk = odd ? (cnt >> 1) : (cnt >> 1) - 1;
        double hi = -INFINITY, lo = INFINITY, sv = 0.0, sd = 0.0;
        unsigned long long orv = 0ull, andv = ~0ull;
        double r0, r1;

        if (cnt > CV_MAXBAR) {
            // ---- too long for the ring: stream it from global memory (v1 body) ----
            __syncwarp();
            if (lane == 0) { __threadfence_block(); S->pos[wid] = e + 1; }
            int64_t j = start + lane;
            for (; j + 96 <= e; j += 128) {
                double p0 = __ldg(p + j), p1 = __ldg(p + j + 32), p2 = __ldg(p + j + 64), p3 = __ldg(p + j + 96);
                double v0 = __ldg(v + j), v1 = __ldg(v + j + 32), v2 = __ldg(v + j + 64), v3 = __ldg(v + j + 96);
                hi = fmax(fmax(hi, fmax(p0, p1)), fmax(p2, p3));
                lo = fmin(fmin(lo, fmin(p0, p1)), fmin(p2, p3));
                sv += (v0 + v1) + (v2 + v3);
                sd += (p0 * v0 + p1 * v1) + (p2 * v2 + p3 * v3);
                const unsigned long long k0 = dkey(v0), k1 = dkey(v1), k2 = dkey(v2), k3 = dkey(v3);
                orv |= (k0 | k1) | (k2 | k3);
                andv &= (k0 & k1) & (k2 & k3);
            }
            for (; j <= e; j += 32) {
                double pj = __ldg(p + j), vj = __ldg(v + j);
                hi = fmax(hi, pj); lo = fmin(lo, pj);
                sv += vj; sd += pj * vj;
                const unsigned long long kj = dkey(vj);
                orv |= kj; andv &= kj;
            }
            hi = warp_max(hi); lo = warp_min(lo); sv = warp_sum(sv); sd = warp_sum(sd);
            orv = warp_or64(orv); andv = warp_and64(andv);
            if (lane == 0) {
                o.open[i] = p[start]; o.close[i] = p[e]; o.high[i] = hi; o.low[i] = lo;
                o.volume[i] = (float)sv;
                o.vwap[i] = sv > 0 ? sd / sv : 0.0;
                o.trades[i] = cnt;
            }
            if (orv == andv) r0 = r1 = dunkey(orv);
            else cv_generic_select(S, v + start, cnt, mk, &r0, &r1, true, orv, andv);
            if (lane == 0) median_out[i] = odd ? r0 : (r0 + r1) / 2;
            continue;
        }

        // ---- publish the low-water mark, then wait for the chunks that cover (s, e] ----
        __syncwarp();
        if (lane == 0) { __threadfence_block(); S->pos[wid] = start; }
        {
            const int64_t gl_hi = e / CV_CHUNK - g0;
            while (S->issued <= gl_hi) {}
            __threadfence_block();
            for (int64_t gl = start / CV_CHUNK - g0; gl <= gl_hi; gl++)
                while (!cv_mbar_try_wait(&S->full[gl % CV_NCH], (unsigned)((gl / CV_NCH) & 1))) {}
        }
        const int ncnt = (int)cnt;
        const int off0 = (int)((start - base_tick) % CV_RING);
#define CV_AT(q) ((off0 + (q)) >= CV_RING ? (off0 + (q)) - CV_RING : (off0 + (q)))
        if (ncnt <= 1024) {
            // ---- register-resident path: each lane keeps the high words of its <= 32 amounts, so the ring is read
            //      once per tick and the select passes run on registers.  Sizes are compared as raw IEEE bit patterns,
            //      which order like the values when none is negative (any sign bit -> generic path below). ----
            unsigned hv[32];
            double hi0 = -INFINITY, lo0 = INFINITY, sva = 0.0, svb = 0.0, sda = 0.0, sdb = 0.0;
            unsigned long long ro = 0ull, ra = ~0ull;
#pragma unroll
            for (int t = 0; t < 32; t++) {
                hv[t] = 0u;
                if (t * 32 < ncnt) {
                    const int q = t * 32 + lane;
                    if (q < ncnt) {
                        const int a = CV_AT(q);
                        const double pj = S->ring_p[a], vj = S->ring_v[a];
                        hi0 = pj > hi0 ? pj : hi0;
                        lo0 = pj < lo0 ? pj : lo0;
                        if (t & 1) { svb += vj; sdb += pj * vj; } else { sva += vj; sda += pj * vj; }
                        const unsigned long long raw = (unsigned long long)__double_as_longlong(vj);
                        ro |= raw; ra &= raw;
                        hv[t] = (unsigned)(raw >> 32);
                    }
                }
            }
            sv = sva + svb; sd = sda + sdb;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const double h2 = __shfl_xor_sync(FULL, hi0, d), l2 = __shfl_xor_sync(FULL, lo0, d);
                hi0 = h2 > hi0 ? h2 : hi0;
                lo0 = l2 < lo0 ? l2 : lo0;
                sv += __shfl_xor_sync(FULL, sv, d);
                sd += __shfl_xor_sync(FULL, sd, d);
            }
            ro = warp_or64(ro); ra = warp_and64(ra);
            if (lane == 0) {
                o.open[i] = S->ring_p[CV_AT(0)]; o.close[i] = S->ring_p[CV_AT(ncnt - 1)]; o.high[i] = hi0; o.low[i] = lo0;
                o.volume[i] = (float)sv;
                o.vwap[i] = sv > 0 ? sd / sv : 0.0;
                o.trades[i] = cnt;
            }
            bool done = false;
            if (ro == ra) { r0 = r1 = __longlong_as_double((long long)ro); done = true; }   // all sizes identical
            unsigned diff_hi = (unsigned)((ro ^ ra) >> 32);
            if (!done && !(ro >> 63) && diff_hi != 0u) {
                unsigned mask = 0u, prefix = 0u, upper = (unsigned)(ra >> 32);
                int kk = (int)mk;
                const bool want1 = !odd;
                for (;;) {
                    const int hb = 31 - __clz(diff_hi);
                    const int shift = hb >= 9 ? hb - 9 : 0;
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < CV_WORDS / 32; q++) W->hist[q * 32 + lane] = 0u;
                    __syncwarp();
#pragma unroll
                    for (int t = 0; t < 32; t++)
                        if (t * 32 < ncnt) {
                            const unsigned h = hv[t];
                            if (t * 32 + lane < ncnt && (h & mask) == prefix) {
                                const unsigned d = (h >> shift) & 1023u;
                                atomicAdd(&W->hist[cv_word(d)], (d & 1u) ? 65536u : 1u);
                            }
                        }
                    __syncwarp();
                    unsigned ps = 0;
#pragma unroll
                    for (int q = 0; q < CV_WORDS / 32; q++) ps += W->hist[q * 32 + lane];
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
                    const unsigned wv = W->hist[((lane >> 1) << 5) | owner];
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
                    kk = (int)(k2 - __shfl_sync(FULL, exc2, sel));
                    const unsigned above_mask = (shift + 10 >= 32) ? 0u : (~0u << (shift + 10));
                    mask = above_mask | (1023u << shift);
                    prefix = (upper & above_mask) | (bsel << shift);
                    const bool inside1 = kk + 1 < c1;
                    unsigned long long key0 = 0ull, key1 = 0ull;
                    bool resolved = false, give_up = false;
                    if (c1 <= 32) {
                        if (lane == 0) W->ccnt = 0u;
                        __syncwarp();
#pragma unroll
                        for (int t = 0; t < 32; t++)
                            if (t * 32 < ncnt) {
                                if (t * 32 + lane < ncnt && (hv[t] & mask) == prefix)
                                    W->cidx[atomicAdd(&W->ccnt, 1u) & 31u] = (unsigned)(t * 32 + lane);
                            }
                        __syncwarp();
                        const unsigned long long mine =
                            lane < c1 ? (unsigned long long)__double_as_longlong(S->ring_v[CV_AT((int)W->cidx[lane])]) : ~0ull;
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
#pragma unroll
                        for (int t = 0; t < 32; t++)
                            if (t * 32 < ncnt) {
                                if (t * 32 + lane < ncnt && (hv[t] & mask) == prefix) {
                                    const unsigned long long key =
                                        (unsigned long long)__double_as_longlong(S->ring_v[CV_AT(t * 32 + lane)]);
                                    bo |= key; ba &= key;
                                }
                            }
                        bo = warp_or64(bo); ba = warp_and64(ba);
                        if (bo == ba) { key0 = bo; key1 = bo; resolved = true; }
                        else {
                            upper = (unsigned)(ba >> 32);
                            diff_hi = (unsigned)((bo ^ ba) >> 32);
                            if (diff_hi == 0u || shift == 0) give_up = true;
                        }
                    }
                    if (resolved) {
                        if (want1 && !inside1) {
                            const unsigned hi_bound = prefix | ~mask;
                            unsigned long long amin = ~0ull;
#pragma unroll
                            for (int t = 0; t < 32; t++)
                                if (t * 32 < ncnt) {
                                    if (t * 32 + lane < ncnt && hv[t] > hi_bound) {
                                        const unsigned long long key =
                                            (unsigned long long)__double_as_longlong(S->ring_v[CV_AT(t * 32 + lane)]);
                                        amin = key < amin ? key : amin;
                                    }
                                }
                            key1 = warp_min64(amin);
                        }
                        r0 = __longlong_as_double((long long)key0); r1 = __longlong_as_double((long long)key1);
                        done = true;
                        break;
                    }
                    if (give_up) break;
                }
            }
            if (!done) cv_generic_select(S, v + start, cnt, mk, &r0, &r1, false, 0ull, ~0ull);
            if (lane == 0) median_out[i] = odd ? r0 : (r0 + r1) / 2;
            continue;
        }
        {   // S0: O/H/L/C, sums, OR/AND of the amount keys -- from the ring
            int q = lane;
            for (; q + 96 < ncnt; q += 128) {
                const int a0 = CV_AT(q), a1 = CV_AT(q + 32), a2 = CV_AT(q + 64), a3 = CV_AT(q + 96);
                const double p0 = S->ring_p[a0], p1 = S->ring_p[a1], p2 = S->ring_p[a2], p3 = S->ring_p[a3];
                const double v0 = S->ring_v[a0], v1 = S->ring_v[a1], v2 = S->ring_v[a2], v3 = S->ring_v[a3];
                hi = fmax(fmax(hi, fmax(p0, p1)), fmax(p2, p3));
                lo = fmin(fmin(lo, fmin(p0, p1)), fmin(p2, p3));
                sv += (v0 + v1) + (v2 + v3);
                sd += (p0 * v0 + p1 * v1) + (p2 * v2 + p3 * v3);
                const unsigned long long k0 = dkey(v0), k1 = dkey(v1), k2 = dkey(v2), k3 = dkey(v3);
                orv |= (k0 | k1) | (k2 | k3);
                andv &= (k0 & k1) & (k2 & k3);
            }
            for (; q < ncnt; q += 32) {
                const int a = CV_AT(q);
                const double pj = S->ring_p[a], vj = S->ring_v[a];
                hi = fmax(hi, pj); lo = fmin(lo, pj);
                sv += vj; sd += pj * vj;
                const unsigned long long kj = dkey(vj);
                orv |= kj; andv &= kj;
            }
        }
        hi = warp_max(hi); lo = warp_min(lo); sv = warp_sum(sv); sd = warp_sum(sd);
        orv = warp_or64(orv); andv = warp_and64(andv);
        if (lane == 0) {
            o.open[i] = S->ring_p[CV_AT(0)]; o.close[i] = S->ring_p[CV_AT(ncnt - 1)]; o.high[i] = hi; o.low[i] = lo;
            o.volume[i] = (float)sv;
            o.vwap[i] = sv > 0 ? sd / sv : 0.0;
            o.trades[i] = cnt;
        }

        // ---- median: adaptive radix select on the high 32 key bits, all in shared memory ----
        bool done = false;
        if (orv == andv) { r0 = r1 = dunkey(orv); done = true; }      // every size in the bar is identical
        unsigned diff_hi = (unsigned)((orv ^ andv) >> 32);
        if (!done && diff_hi != 0u) {
            unsigned mask = 0u, prefix = 0u;                          // keys in play: (hkey & mask) == prefix
            unsigned upper = (unsigned)(andv >> 32);                  // common bits above the current window
            int kk = (int)mk;
            const bool want1 = !odd;
            for (;;) {
                const int hb = 31 - __clz(diff_hi);
                const int shift = hb >= 9 ? hb - 9 : 0;
                __syncwarp();
#pragma unroll
                for (int q = 0; q < CV_WORDS / 32; q++) W->hist[q * 32 + lane] = 0u;
                __syncwarp();
                {
                    int q = lane;
                    for (; q + 96 < ncnt; q += 128) {
                        const unsigned h0 = cv_hkey(rvw[2 * CV_AT(q) + 1]), h1 = cv_hkey(rvw[2 * CV_AT(q + 32) + 1]),
                                       h2 = cv_hkey(rvw[2 * CV_AT(q + 64) + 1]), h3 = cv_hkey(rvw[2 * CV_AT(q + 96) + 1]);
                        if ((h0 & mask) == prefix) { const unsigned d = (h0 >> shift) & 1023u; atomicAdd(&W->hist[cv_word(d)], (d & 1u) ? 65536u : 1u); }
                        if ((h1 & mask) == prefix) { const unsigned d = (h1 >> shift) & 1023u; atomicAdd(&W->hist[cv_word(d)], (d & 1u) ? 65536u : 1u); }
                        if ((h2 & mask) == prefix) { const unsigned d = (h2 >> shift) & 1023u; atomicAdd(&W->hist[cv_word(d)], (d & 1u) ? 65536u : 1u); }
                        if ((h3 & mask) == prefix) { const unsigned d = (h3 >> shift) & 1023u; atomicAdd(&W->hist[cv_word(d)], (d & 1u) ? 65536u : 1u); }
                    }
                    for (; q < ncnt; q += 32) {
                        const unsigned h = cv_hkey(rvw[2 * CV_AT(q) + 1]);
                        if ((h & mask) == prefix) { const unsigned d = (h >> shift) & 1023u; atomicAdd(&W->hist[cv_word(d)], (d & 1u) ? 65536u : 1u); }
                    }
                }
                __syncwarp();
                // two-level scan of the 1024 bins (lane L owns bins [32L, 32L+32)); packed halves cannot overflow
                unsigned ps = 0;
#pragma unroll
                for (int q = 0; q < CV_WORDS / 32; q++) ps += W->hist[q * 32 + lane];
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
                const unsigned wv = W->hist[((lane >> 1) << 5) | owner];
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
                kk = (int)(k2 - __shfl_sync(FULL, exc2, sel));         // rank inside the bucket
                // narrow the range to the bucket
                const unsigned above_mask = (shift + 10 >= 32) ? 0u : (~0u << (shift + 10));
                mask = above_mask | (1023u << shift);
                prefix = (upper & above_mask) | (bsel << shift);
                upper = prefix;
                const bool inside1 = kk + 1 < c1;
                unsigned long long key0 = 0ull, key1 = 0ull;
                bool resolved = false, give_up = false;
                if (c1 <= 32) {
                    if (lane == 0) W->ccnt = 0u;
                    __syncwarp();
#pragma unroll 4
                    for (int q = lane; q < ncnt; q += 32)
                        if ((cv_hkey(rvw[2 * CV_AT(q) + 1]) & mask) == prefix) W->cidx[atomicAdd(&W->ccnt, 1u) & 31u] = (unsigned)q;
                    __syncwarp();
                    const unsigned long long mine = lane < c1 ? dkey(S->ring_v[CV_AT((int)W->cidx[lane])]) : ~0ull;
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
                    for (int q = lane; q < ncnt; q += 32) {
                        const int a = CV_AT(q);
                        if ((cv_hkey(rvw[2 * a + 1]) & mask) == prefix) {
                            const unsigned long long key = dkey(S->ring_v[a]);
                            bo |= key; ba &= key;
                        }
                    }
                    bo = warp_or64(bo); ba = warp_and64(ba);
                    if (bo == ba) { key0 = bo; key1 = bo; resolved = true; }   // exchange-quantised sizes: one value
                    else {
                        upper = (unsigned)(ba >> 32);                           // common bits of the members
                        diff_hi = (unsigned)((bo ^ ba) >> 32);
                        if (diff_hi == 0u || shift == 0) give_up = true;        // differ only below the high word
                    }
                }
                if (resolved) {
                    if (want1 && !inside1) {          // the (k+1)-th statistic is the smallest key above the bucket
                        const unsigned hi_bound = prefix | ~mask;
                        unsigned long long amin = ~0ull;
                        for (int q = lane; q < ncnt; q += 32) {
                            const int a = CV_AT(q);
                            if (cv_hkey(rvw[2 * a + 1]) > hi_bound) {
                                const unsigned long long key = dkey(S->ring_v[a]);
                                amin = key < amin ? key : amin;
                            }
                        }
                        key1 = warp_min64(amin);
                    }
                    r0 = dunkey(key0); r1 = dunkey(key1);
                    done = true;
                    break;
                }
                if (give_up) break;
            }
        }
#undef CV_AT
        if (!done) cv_generic_select(S, v + start, cnt, mk, &r0, &r1, true, orv, andv);
        if (lane == 0) median_out[i] = odd ? r0 : (r0 + r1) / 2;
    }
    __syncwarp();
    if (lane == 0) { __threadfence_block(); S->pos[wid] = 0x7fffffffffffffffll; }
}
