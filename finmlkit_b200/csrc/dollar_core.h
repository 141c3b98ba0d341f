// dollar_core.h -- bit-exact parallel dollar-bar indexer: host/device shared core (logic.py:118-149).
//
// Reference recurrence (strictly sequential, float64, no FMA):
//     c = p0*v0 ; for i>=1: c = fl(c + fl(p_i*v_i)); if c >= T: emit i; c = fl(c - T)
// The carried remainder makes every boundary depend on the rounding history of all earlier ticks, so a reassociated
// (scan) sum cannot reproduce the reference's decisions at near-ties.  This file restates the recurrence as
// independent TASKS whose results are provably identical to the sequential run, plus an exact integer carry chain:
//
//  * The stream is cut into chunks of CH ticks.  Task k starts right after the first (approximately located) bar
//    boundary B_k inside chunk k and replays the *exact* recurrence until it emits a boundary at an index
//    >= (k+1)*CH -- which is B_{k+1} if the guesses are consistent.
//  * Right after a boundary the true carry r is a multiple of u = ulp(T).  The task does not know r, only an
//    approximation g (from a double-double prefix sum).  Translation property: for states below 4*2^e(T), if the
//    start state is shifted by D = 4*u*z, every float add rounds identically and the whole trajectory is shifted by
//    exactly D, provided no value comes within |D| of a decision threshold (T) or a binade boundary (a power of two).
//    Ties-to-even and the coarser ulp one binade above T depend on r mod 4u only, so the task replays FOUR chains
//    (start = g0 + rho*u, rho = 0..3) and records for each: end state, and the smallest margin seen.
//  * A sequential-in-bars but trivially cheap chain (dollar_chain_step, run as a parallel scan on the device) then
//    propagates the TRUE start state: Delta = s_k - g0_k, rho = Delta mod 4, D = Delta - rho; the task is certified
//    iff D == 0 or |D|*u < margin[rho] and chain rho emitted exactly where chain 0 did; s_{k+1} = end[rho] + D.
//  * Certification failures (near-ties, giant trades, inconsistent guesses) are repaired by an exact serial replay
//    from the last certified state (dollar_serial), after which the chain resumes.  Every emitted index is therefore
//    either certified identical to the sequential run or produced by the sequential run itself.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define DC_HD __host__ __device__ __forceinline__
#else
#define DC_HD static inline
#endif

DC_HD double dc_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;  // host build uses -ffp-contract=off
#endif
}
DC_HD double dc_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
DC_HD double dc_sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, -b);
#else
    return a - b;
#endif
}
DC_HD uint64_t dc_bits(double x) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t b; memcpy(&b, &x, 8); return b;
#endif
}
DC_HD double dc_from_bits(uint64_t b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double x; memcpy(&x, &b, 8); return x;
#endif
}

// ---- double-double helpers (approximate prefix sums for the guesses only) -------------------------------------
struct dd_t { double hi, lo; };
DC_HD dd_t dd_add_d(dd_t a, double b) {   // a + b (TwoSum)
    double s = dc_add(a.hi, b);
    double bb = dc_sub(s, a.hi);
    double e = dc_add(dc_sub(a.hi, dc_sub(s, bb)), dc_sub(b, bb));
    e = dc_add(e, a.lo);
    double hi = dc_add(s, e);
    double lo = dc_sub(e, dc_sub(hi, s));
    dd_t r = {hi, lo};
    return r;
}
DC_HD dd_t dd_add(dd_t a, dd_t b) {
    dd_t r = dd_add_d(a, b.hi);
    double e = dc_add(r.lo, b.lo);
    double hi = dc_add(r.hi, e);
    double lo = dc_sub(e, dc_sub(hi, r.hi));
    dd_t q = {hi, lo};
    return q;
}

// Guess of (number of emitted boundaries, carry) before a chunk whose exclusive prefix of dollars is P, in exact
// arithmetic ignoring the one-emit-per-tick rule: K = floor(P/T), carry = P - K*T.  Only a guess.
DC_HD void dollar_guess(dd_t P, double T, int64_t *K, double *carry) {
    double q = floor(P.hi / T);
    double ph = dc_mul(q, T);
    double pl = fma(q, T, -ph);
    double rem = dc_add(dc_sub(dc_sub(P.hi, ph), pl), P.lo);
    int guard = 0;
    while (rem < 0 && guard++ < 8) { rem = dc_add(rem, T); q -= 1.0; }
    while (rem >= T && guard++ < 16) { rem = dc_sub(rem, T); q += 1.0; }
    if (!(rem >= 0)) rem = 0;
    *K = (int64_t)q;
    *carry = rem;
}

constexpr int DC_NCH = 4;

struct DollarTaskRec {
    int64_t start_idx;        // B_k (boundary the task starts after); -1 = empty task (no boundary located in chunk)
    int64_t k_start;          // ordinal of B_k in the output index array (out[k_start] == B_k)
    int64_t end_idx;          // last boundary emitted by chain 0 (>= next chunk start); -2 = ran to the end of data
    int64_t count;            // boundaries emitted by chain 0
    int64_t start_units;      // chain-0 start state / u ; task 0: unused (exact start)
    int64_t end_units[DC_NCH];
    double margin[DC_NCH];    // smallest distance of any value to a decision/binade boundary
    int32_t bad;              // bit r set: chain r left the regime where the translation property holds
    int32_t nch;              // 1 for task 0 (exact start), DC_NCH otherwise
};

struct DollarParams {
    double T, u, top_lim, sub_lim;  // u = ulp(T); top_lim = 4*2^e; sub_lim = 2*2^e
    double margin0;                 // 2^(e-35): distance guaranteed by the coarse per-tick tests
    double inv_half_u;              // 2/u
    uint32_t hi_tiny, hi_near, hi_top, hi_sub;   // high words of 2^(e-24), 2^(e-34), top_lim, sub_lim
    uint32_t eTb;                                // biased exponent of T
    int64_t n, CH, cap;
};

DC_HD uint32_t dc_hi(double x) { return (uint32_t)(dc_bits(x) >> 32); }

DC_HD bool dollar_params_init(DollarParams *P, double T, int64_t n, int64_t CH, int64_t cap) {
    P->T = T; P->n = n; P->CH = CH; P->cap = cap;
    if (!(T > 1e-250) || !(T < 1e250)) return false;
    uint64_t b = dc_bits(T);
    double pow2 = dc_from_bits(b & 0xFFF0000000000000ull);  // 2^e
    P->u = pow2 * 2.220446049250313e-16;                    // 2^(e-52)
    P->top_lim = pow2 * 4.0;
    P->sub_lim = pow2 * 2.0;
    P->margin0 = pow2 * 2.9103830456733704e-11;             // 2^-35
    P->inv_half_u = 2.0 / P->u;
    P->hi_tiny = dc_hi(pow2 * 5.9604644775390625e-08);      // 2^-24
    P->hi_near = dc_hi(pow2 * 5.820766091346741e-11);       // 2^-34
    P->hi_top = dc_hi(P->top_lim);
    P->hi_sub = dc_hi(P->sub_lim);
    P->eTb = (uint32_t)(b >> 52) & 0x7FFu;
    return true;
}

// One exact step of the reference recurrence for one chain, with margin tracking.
// Returns true when the chain emits at this tick.
//
// Margin = smallest distance of any intermediate value to a decision threshold (T) or a binade boundary (a power of
// two).  Computing it exactly every tick costs more than the recurrence itself, so the common path only runs cheap
// integer tests on the high word of r = c + d and of r - T that PROVE the distance is >= margin0 = 2^(e-35):
//   * r >= 2^(e-24) and the top 10 mantissa bits of r are neither all 0 nor all 1  =>  r is >= 2^(e-34) away from
//     every power of two;
//   * |r - T| >= 2^(e-34).
// Only when a test fails (a few 1e-3 of the ticks) is the exact distance folded into the running minimum.
DC_HD bool dollar_chain_tick(double &c, double d, const DollarParams &P, uint64_t &mbits, bool &bad) {
    const double r = dc_add(c, d);
    const double c2 = dc_sub(r, P.T);
    const uint32_t hr = dc_hi(r);
    const uint32_t hc = dc_hi(c2) & 0x7FFFFFFFu;
    const bool rare = (((hr + 0x400u) & 0xFF800u) == 0u) | (hr < P.hi_tiny) | (hr >= P.hi_top) | (hc < P.hi_near);
    if (rare) {
        const uint64_t rb = dc_bits(r);
        const double lowb = dc_from_bits(rb & 0xFFF0000000000000ull);
        const double dlo = dc_sub(r, lowb);
        const double dhi = dc_sub(lowb, dlo);
        const double am = fabs(c2);
        // non-negative doubles order like their bit patterns; negative/NaN values are caught by the range test
        uint64_t m = mbits, x;
        x = dc_bits(dlo); m = x < m ? x : m;
        x = dc_bits(dhi); m = x < m ? x : m;
        x = dc_bits(am);  m = x < m ? x : m;
        mbits = m;
        if (!(r >= 0.0) || !(r < P.top_lim)) bad = true;
    }
    if (r >= P.T) {
        if (dc_hi(c2) >= P.hi_sub) bad = true;     // carry would leave the binades where it is a multiple of u
        c = c2;
        return true;
    }
    c = r;
    return false;
}

// ---- virtual chains ------------------------------------------------------------------------------------------
// Replaying four chains costs four dependent float recurrences per tick.  But the chains differ only by small
// multiples of u, and translation by a multiple of u commutes with a float add EXCEPT at three kinds of "residue
// sensitive" adds, all of which can be recognised from chain 0 alone:
//   (b) result in T's binade (ulp u) and the exact sum is a tie (|err| == u/2): ties-to-even sends odd offsets the
//       other way  ->  o' = o + sign(err) for odd o;
//   (c) result one binade above T (ulp 2u): the result is 2u * rhe((X0 + o)/2), derived from chain 0's result R,
//       sign(err) and whether |err| == u;
//   (a) result below T's binade (ulp <= u/2): never sensitive.
// So only chain 0 is run in floating point; the other three are integer offsets o[rho] (units of u) relative to it,
// updated in a rare branch that needs the TwoSum error of the add.  A cheap integer pre-filter (is d's fraction
// w.r.t. u exactly one half?  did the binade change?  is the result above T's binade?) guards that branch.
// DC_EXPLICIT_CHAINS keeps the four explicit chains (reference implementation for the CPU cross-check).
struct DollarTask {
#ifdef DC_EXPLICIT_CHAINS
    double c[DC_NCH];
    uint64_t mb[DC_NCH];
    bool bad[DC_NCH];
#else
    double c[1];
    uint64_t mb[1];
    bool bad[1];
    int32_t o[DC_NCH];   // virtual chain offsets relative to chain 0, in units of u
#endif
    double x;            // phase 1: approximate carry
    int64_t B, K, cnt, end_idx, hi, start_units;
    int nch, phase;      // phase 1: locating the first boundary; 2: exact replay; 3: finished
    bool last_chunk;
};

#ifdef DC_EXPLICIT_CHAINS
constexpr int DC_NSIM = DC_NCH;
#else
constexpr int DC_NSIM = 1;
#endif

DC_HD int dc_ffsll(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __ffsll((long long)x);
#else
    return __builtin_ffsll((long long)x);
#endif
}

// first tick the task wants to see
DC_HD int64_t dollar_task_init(DollarTask &t, const DollarParams &P, int64_t k, double carry_in, int64_t K_in, double d0) {
    const int64_t lo = k * P.CH;
    t.hi = lo + P.CH < P.n ? lo + P.CH : P.n;
    t.last_chunk = t.hi >= P.n;
    t.B = -1; t.K = K_in; t.cnt = 0; t.end_idx = -2; t.start_units = 0; t.x = carry_in;
    for (int r = 0; r < DC_NSIM; r++) { t.mb[r] = dc_bits(P.margin0); t.bad[r] = false; t.c[r] = 0.0; }
#ifndef DC_EXPLICIT_CHAINS
    for (int r = 0; r < DC_NCH; r++) t.o[r] = r;
#endif
    if (k == 0) {
        t.B = 0; t.K = 0; t.nch = 1; t.phase = 2;
        t.c[0] = d0;                 // exact start: cum = prices[0] * volumes[0]
        return 1;
    }
    t.nch = DC_NCH; t.phase = 1;
    return lo;
}

#ifndef DC_EXPLICIT_CHAINS
// rare branch of the virtual-chain tick: update the offsets at a residue-sensitive add (c + d -> r)
DC_HD void dollar_virtual_update(DollarTask &t, const DollarParams &P, double c, double d, double r, bool in_upper) {
    // TwoSum error: err = (c + d) - r exactly
    const double bb = dc_sub(r, c);
    const double err = dc_add(dc_sub(c, dc_sub(r, bb)), dc_sub(d, bb));
    const int sd = err > 0.0 ? 1 : (err < 0.0 ? -1 : 0);
    const double ae = fabs(err);
    if (!in_upper) {
        // (b) ulp(r) == u: only an exact tie is residue sensitive
        if (ae == 0.5 * P.u) {
            for (int q = 1; q < DC_NCH; q++) if (t.o[q] & 1) t.o[q] += sd;
        }
    } else {
        // (c) ulp(r) == 2u: r = 2u*R, exact sum = 2u*(R + delta), delta = err/(2u) in [-1/2, 1/2]
        const int64_t R = (int64_t)(r / (2.0 * P.u));
        const int Rpar = (int)(R & 1);
        const bool half = (ae == P.u);
        for (int q = 1; q < DC_NCH; q++) {
            const int o = t.o[q];
            const int a = o >> 1, b = o & 1;     // o = 2a + b, b in {0,1} (arithmetic shift = floor)
            int res;
            if (b == 0) {
                if (!half) res = a;
                else res = (((Rpar + a) & 1) == 0) ? a : a + sd;          // tie at (R+a) + sd/2 -> even
            } else {
                if (half) res = a + (sd > 0 ? 1 : 0);                      // delta + 1/2 is exactly 0 or 1
                else if (sd < 0) res = a;
                else if (sd > 0) res = a + 1;
                else res = (((Rpar + a) & 1) == 0) ? a : a + 1;           // exact tie at (R+a) + 1/2 -> even
            }
            t.o[q] = 2 * res;
        }
    }
    for (int q = 1; q < DC_NCH; q++) if (t.o[q] > 32 || t.o[q] < -32) t.bad[0] = true;
}

DC_HD bool dollar_virtual_tick(DollarTask &t, double d, const DollarParams &P) {
    const double c = t.c[0];
    const double r = dc_add(c, d);
    const double c2 = dc_sub(r, P.T);
    const uint32_t hr = dc_hi(r);
    const uint32_t hcp = dc_hi(c);
    const uint32_t hc = dc_hi(c2) & 0x7FFFFFFFu;
    const bool rare_m = (((hr + 0x400u) & 0xFF800u) == 0u) | (hr < P.hi_tiny) | (hr >= P.hi_top) | (hc < P.hi_near);
    // residue-sensitivity pre-filter
    const uint32_t eb = hr >> 20;                         // sign + biased exponent of r
    const bool in_upper = eb > P.eTb;                     // r >= 2^(e+1)
    const bool in_top = eb == P.eTb;
    const bool crossing = ((hr ^ hcp) >> 20) != 0u;
    // pre-filter for "d/u has fractional part exactly one half" (the only way an add inside T's binade can be a tie):
    // t = 2d/u is then an odd integer.  For t < 2^52, t + 2^52 rounds t to an integer whose parity is the low mantissa
    // bit.  Larger d (>= T/2-ish) simply take the rare branch, which decides exactly from the TwoSum error anyway.
    const double tq = dc_mul(d, P.inv_half_u);
    const double tm = dc_add(tq, 4503599627370496.0);
    const bool dtie = ((dc_sub(tm, 4503599627370496.0) == tq) & ((uint32_t)dc_bits(tm) & 1u)) | !(tq < 4503599627370496.0);
    const bool sens = in_upper | (in_top & (dtie | crossing));
    if (rare_m | sens) {
        if (rare_m) {
            const uint64_t rb = dc_bits(r);
            const double lowb = dc_from_bits(rb & 0xFFF0000000000000ull);
            const double dlo = dc_sub(r, lowb);
            const double dhi = dc_sub(lowb, dlo);
            const double am = fabs(c2);
            uint64_t m = t.mb[0], x;
            x = dc_bits(dlo); m = x < m ? x : m;
            x = dc_bits(dhi); m = x < m ? x : m;
            x = dc_bits(am);  m = x < m ? x : m;
            t.mb[0] = m;
            if (!(r >= 0.0) || !(r < P.top_lim)) t.bad[0] = true;
        }
        if (sens && t.nch == DC_NCH && r < P.top_lim) dollar_virtual_update(t, P, c, d, r, in_upper);
    }
    if (r >= P.T) {
        if (dc_hi(c2) >= P.hi_sub) t.bad[0] = true;
        t.c[0] = c2;
        return true;
    }
    t.c[0] = r;
    return false;
}
#endif

// chain 0 emitted a boundary at tick i: record it; returns true when the task is finished
DC_HD bool dollar_task_emit(DollarTask &t, const DollarParams &P, int64_t i, int64_t *out) {
    t.cnt++;
    if (t.K + t.cnt < P.cap) out[t.K + t.cnt] = i;   // speculative writes stay in bounds; cap bounds the true count
    if (!t.last_chunk && i >= t.hi) { t.end_idx = i; t.phase = 3; return true; }
    return false;
}

// Feed tick i (d = fl(p_i * v_i)).  Returns true when the task is finished.
DC_HD bool dollar_task_consume(DollarTask &t, const DollarParams &P, int64_t i, double d, int64_t *out) {
    if (t.phase == 1) {
        t.x = dc_add(t.x, d);
        if (t.x >= P.T) {
            t.x = dc_sub(t.x, P.T);
            t.B = i;
            t.K += 1;
            double units = rint(t.x / P.u);
            if (!(units >= 0)) units = 0;
            t.start_units = (int64_t)units;
            for (int r = 0; r < DC_NSIM; r++) {
                t.c[r] = dc_mul((double)(t.start_units + r), P.u);
                // the start must be exactly representable (it is unless units ~ 2^53)
                if (t.c[r] / P.u != (double)(t.start_units + r)) t.bad[r] = true;
            }
#ifndef DC_EXPLICIT_CHAINS
            if (!(units < 4503599627370000.0)) t.bad[0] = true;   // start + 3 must stay below 2^52
#endif
            t.phase = 2;
            return false;
        }
        if (i + 1 >= t.hi) { t.phase = 3; return true; }   // no boundary located in this chunk: empty task
        return false;
    }
#ifdef DC_EXPLICIT_CHAINS
    const bool e0 = dollar_chain_tick(t.c[0], d, P, t.mb[0], t.bad[0]);
    if (t.nch == DC_NCH) {
        for (int r = 1; r < DC_NCH; r++) {
            const bool er = dollar_chain_tick(t.c[r], d, P, t.mb[r], t.bad[r]);
            if (er != e0) t.bad[r] = true;
        }
    }
#else
    const bool e0 = dollar_virtual_tick(t, d, P);
#endif
    if (e0) return dollar_task_emit(t, P, i, out);
    return false;
}

DC_HD void dollar_task_finish(const DollarTask &t, const DollarParams &P, DollarTaskRec *rec) {
    rec->start_idx = t.B; rec->k_start = t.B >= 0 ? t.K : 0; rec->end_idx = t.end_idx; rec->count = t.cnt;
    rec->start_units = t.start_units; rec->nch = t.nch;
    int32_t bm = 0;
#ifdef DC_EXPLICIT_CHAINS
    for (int r = 0; r < DC_NCH; r++) {
        rec->end_units[r] = 0;
        rec->margin[r] = dc_from_bits(t.mb[r]);
        bool bad = t.bad[r];
        if (t.end_idx >= 0 && r < t.nch) {
            const double eu = t.c[r] / P.u;          // exact when c is a multiple of u below 2^(e+1)
            rec->end_units[r] = (int64_t)eu;
            if ((double)rec->end_units[r] != eu || dc_mul(eu, P.u) != t.c[r]) bad = true;
        }
        if (bad) bm |= (1 << r);
    }
#else
    bool bad = t.bad[0];
    int64_t e0 = 0;
    if (t.end_idx >= 0) {
        const double eu = t.c[0] / P.u;              // exact when c is a multiple of u below 2^(e+1)
        e0 = (int64_t)eu;
        if ((double)e0 != eu || dc_mul(eu, P.u) != t.c[0]) bad = true;
    }
    // the virtual chains sit within 32u of chain 0: certify them against chain 0's margin minus that slack
    double m = dc_from_bits(t.mb[0]) - 64.0 * P.u;
    if (!(m > 0.0)) m = 0.0;
    for (int r = 0; r < DC_NCH; r++) {
        rec->end_units[r] = e0 + (r < t.nch ? t.o[r] : 0);
        rec->margin[r] = (r == 0) ? dc_from_bits(t.mb[0]) : m;
        if (bad) bm |= (1 << r);
    }
#endif
    rec->bad = bm;
}

// Host-style driver of the state machine (CPU emulation; the device kernel stages ticks through shared memory).
template <typename LoadP, typename LoadV>
DC_HD void dollar_task(LoadP p, LoadV v, const DollarParams &P, int64_t k, double carry_in, int64_t K_in,
                       int64_t *out, DollarTaskRec *rec) {
    DollarTask t;
    int64_t i = dollar_task_init(t, P, k, carry_in, K_in, k == 0 ? dc_mul(p(0), v(0)) : 0.0);
    if (!(k > 0 && i >= t.hi)) {
        for (; i < P.n; i++)
            if (dollar_task_consume(t, P, i, dc_mul(p(i), v(i)), out)) break;
    }
    dollar_task_finish(t, P, rec);
}

// ---- carry chain ---------------------------------------------------------------------------------------------
// Transfer function of a task on Delta = s - start_units: f(Delta) = Delta + off[Delta & 3].
struct DollarXfer { int64_t off[4]; };

DC_HD DollarXfer dollar_xfer_identity() { DollarXfer f; for (int r = 0; r < 4; r++) f.off[r] = 0; return f; }
// h = g after f
DC_HD DollarXfer dollar_xfer_compose(const DollarXfer &f, const DollarXfer &g) {
    DollarXfer h;
    for (int r = 0; r < 4; r++) {
        int64_t mid = r + f.off[r];
        h.off[r] = f.off[r] + g.off[mid & 3];
    }
    return h;
}

// Exact serial replay from a known state (repair path / degenerate thresholds).  Starts after boundary `pos` with
// carry c and `K` boundaries emitted so far; stops at the first emitted boundary i for which stop(i, K_after) is
// true, or at the end of the data.  Returns the number emitted; *c_out / *pos_out describe the stop.
template <typename LoadP, typename LoadV, typename Stop>
DC_HD int64_t dollar_serial(LoadP p, LoadV v, int64_t n, double T, int64_t pos, double c, int64_t K, int64_t *out,
                            int64_t cap, int *overflow, Stop stop, double *c_out, int64_t *pos_out) {
    int64_t cnt = 0;
    *pos_out = -2;
    for (int64_t i = pos + 1; i < n; i++) {
        c = dc_add(c, dc_mul(p(i), v(i)));
        if (c >= T) {
            c = dc_sub(c, T);
            cnt++;
            if (K + cnt < cap) out[K + cnt] = i; else *overflow = 1;
            if (stop(i, K + cnt, c)) { *pos_out = i; break; }
        }
    }
    *c_out = c;
    return cnt;
}

// ---- chain walk over task records -----------------------------------------------------------------------------
// Absolute-state transfer function of task k: F(s) = s + off[s & 3] (s = true start state in units of u).
DC_HD DollarXfer dollar_task_xfer(const DollarTaskRec &t) {
    DollarXfer f;
    for (int a = 0; a < 4; a++) {
        int rho = (int)((a - t.start_units) & 3);
        f.off[a] = t.end_units[rho] - t.start_units - rho;
    }
    return f;
}

struct DollarWalk {
    int64_t s;           // true state (units of u) right after boundary `pos`
    int64_t pos;         // last certified boundary index
    int64_t K;           // its ordinal in out[]
    int64_t fail_task;   // first task that could not be certified (-1: none)
    int64_t done;        // 1: a certified task ran to the end of the data; K_total valid
    int64_t K_total;
};

// Certify task t given the walk state w (which describes the boundary the task must start from).
// Returns 0 = certified & advanced, 1 = certified tail (done), 2 = failure (repair serially, resync after this task),
// 3 = task is stale or empty (skip), 4 = failure before the task (repair serially, resync may happen at this task).
// strict: a stale task (start behind the certified frontier) is reported as a failure instead of skipped -- the
// parallel chain composes transfer functions assuming perfect linkage, so any deviation must go to the repair path.
DC_HD int dollar_walk_step(const DollarTaskRec &t, double u, DollarWalk &w, bool strict) {
    if (t.start_idx < 0) return 3;
    if (t.start_idx < w.pos) return strict ? 4 : 3;
    if (t.start_idx > w.pos) return 4;
    if (t.k_start != w.K) return 2;
    if (t.nch != DC_NCH) return 2;            // task 0 is consumed by the caller
    const int64_t delta = w.s - t.start_units;
    const int rho = (int)(delta & 3);
    const int64_t D = delta - rho;
    if ((t.bad >> rho) & 1) return 2;
    if (D != 0 || rho != 0) {
        // rho != 0: the chain is a translate of chain 0 by a few u even when D == 0, so it needs a positive margin too
        double ad = fabs((double)D) * u;
        if (!(ad < t.margin[rho])) return 2;
    }
    if (t.end_idx == -2) { w.done = 1; w.K_total = t.k_start + t.count; return 1; }
    w.s = t.end_units[rho] + D;
    w.pos = t.end_idx;
    w.K = t.k_start + t.count;
    return 0;
}

// Sequential walk over tasks [k, k_end).  Returns 0 when the range is exhausted, 1 when done (tail certified),
// 2/4 on failure with *k_fail = failing task.
DC_HD int dollar_walk_range(const DollarTaskRec *recs, int64_t k, int64_t k_end, double u, DollarWalk &w,
                            int64_t *k_fail, int64_t *n_certified, bool strict) {
    for (; k < k_end; k++) {
        int rc = dollar_walk_step(recs[k], u, w, strict);
        if (rc == 3) continue;
        if (rc == 0) { (*n_certified)++; continue; }
        *k_fail = k;
        return rc;
    }
    return 0;
}
