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

// first non-NaN sigma in [lo, hi): the answer is almost always near the front (sigma is NaN only while its own warm-up
// window fills), so the host probes geometrically growing ranges instead of streaming the whole column
__global__ void k_cusum_first(const double *__restrict__ sigma, int64_t lo, int64_t hi, unsigned long long *first) {
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // one atomic per warp at most
    const unsigned ok = __ballot_sync(0xffffffffu, i < hi && sigma[i] == sigma[i]);
    if (ok && (threadIdx.x & 31) == (unsigned)(__ffs(ok) - 1)) atomicMin(first, (unsigned long long)i);
}

__global__ void k_count_nan(const double *__restrict__ x, int64_t lo, int64_t n, unsigned long long *count) {
    unsigned long long c = 0;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        c += x[i] != x[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

// "May close" is folded into lam: where the reference skips the close test (ts[i] == ts[i+1], logic.py:206-209) lam is NaN,
// and both tests `s+ >= lam`, `s- <= -lam` are false for a NaN -- exactly the skip.  (A NaN lam from a NaN sigma behaves the
// same way in the reference: the comparisons are false.)  One array less to stream per tick.
__global__ void k_cusum_prep(const int64_t *__restrict__ ts, const double *__restrict__ p,
                             const double *__restrict__ sigma, int64_t n, double floor_, double mult,
                             double *__restrict__ r, double *__restrict__ lam, int *__restrict__ nonfinite) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ri = i > 0 ? log(__ddiv_rn(p[i], p[i - 1])) : 0.0;
    r[i] = ri;
    if (!isfinite(ri)) *nonfinite = 1;             // zero / negative / NaN prices: the chain then takes the generic compare path
    const double l = __dmul_rn(mult, sigma[i]);
    const double lm = (floor_ > l) ? floor_ : l;   // python max(l, floor): floor only if floor > l (NaN l stays NaN)
    const bool allowed = !(i + 1 < n && ts[i] == ts[i + 1]);
    lam[i] = allowed ? __dadd_rn(lm, 0.0) : ff_nan();   // + 0.0: -0.0 becomes +0.0 (same value; lets the fast path compare bit patterns)
}

// cusum_filter inputs: r_i = log(x_i / x_{i-1}), thr_i (constant or per element); every tick may fire
__global__ void k_cusum_filter_prep(const double *__restrict__ x, const double *__restrict__ thr, int64_t nthr, int64_t n,
                                    double *__restrict__ r, double *__restrict__ lam) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    r[i] = i > 0 ? log(__ddiv_rn(x[i], x[i - 1])) : 0.0;
    lam[i] = nthr == 1 ? thr[0] : thr[i];
}

struct __align__(16) CusumState { double sp, sn; };

// one reference step; returns true when the bar closes / the event fires at this tick.
// MODE 0: _cusum_bar_indexer (bar/logic.py:203-218): s+ >= lam first, then s- <= -lam (lam NaN where the test is skipped).
// MODE 1: cusum_filter (sampling/filters.py:54-67): s- < -thr first, then s+ > thr (strict), every tick.
// MODE 2: tick-imbalance bars (own semantics, see fmk_imbalance_bar_index): theta += b; |theta| >= thr closes, theta = 0.
// MODE 3: tick-run bars: buys += (b > 0), sells += (b < 0); max(buys, sells) >= thr closes, both = 0.
//         (integers carried exactly in the doubles of the shared state; `lam` is not streamed, `thr` is the scalar)
template <int MODE>
__device__ __forceinline__ bool cusum_step(CusumState &s, double r, double lam, double thr) {
    if (MODE == 2) {
        const double th = __dadd_rn(s.sp, r);
        const bool hit = fabs(th) >= thr;
        s.sp = hit ? 0.0 : th;
        return hit;
    }
    if (MODE == 3) {
        const double nb = r > 0.0 ? s.sp + 1.0 : s.sp, ns = r < 0.0 ? s.sn + 1.0 : s.sn;
        const bool hit = (nb > ns ? nb : ns) >= thr;
        s.sp = hit ? 0.0 : nb;
        s.sn = hit ? 0.0 : ns;
        return hit;
    }
    if (MODE == 4) {
        // MODE 0 with the four double compares (~20 cycles each on the serial chain, measured) replaced by integer compares of
        // the bit patterns.  Preconditions, checked by the host: every r is finite (so a, b are never NaN) and lam is >= +0.0
        // or NaN (sigma_floor >= 0; -0.0 normalised by the prep kernel).  Then: a > 0.0  <=>  the int64 pattern is > 0;
        // b < 0.0  <=>  the pattern is < 0 (b = -0.0 cannot arise from x + r in round-to-nearest, and would only flip the sign
        // of a zero state); for non-negative doubles pattern order is value order, and every NaN pattern is above every finite
        // one, so `sp >= lam` and `-sn >= lam` are unsigned compares that are false for a NaN lam, like the double compares.
        const double a = __dadd_rn(s.sp, r), b = __dadd_rn(s.sn, r);
        const double sp = __double_as_longlong(a) > 0 ? a : 0.0;
        const double sn = __double_as_longlong(b) < 0 ? b : 0.0;
        const unsigned long long lb = (unsigned long long)__double_as_longlong(lam);
        const bool hp = (unsigned long long)__double_as_longlong(sp) >= lb;
        const bool hn = !hp && ((unsigned long long)__double_as_longlong(sn) & 0x7fffffffffffffffull) >= lb;
        s.sp = hp ? 0.0 : sp;
        s.sn = hn ? 0.0 : sn;
        return hp | hn;
    }
    const double a = __dadd_rn(s.sp, r), b = __dadd_rn(s.sn, r);
    const double sp = (a > 0.0) ? a : 0.0;      // python max(0.0, a): a only if a > 0.0 (NaN -> 0.0)
    const double sn = (b < 0.0) ? b : 0.0;
    bool hp, hn;
    if (MODE == 0) { hp = sp >= lam; hn = !hp && (sn <= -lam); }
    else { hn = sn < -lam; hp = !hn && (sp > lam); }
    s.sp = hp ? 0.0 : sp;
    s.sn = hn ? 0.0 : sn;
    return hp | hn;
}

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int CT_WARPS = 4;
constexpr int CT_R = 8;

// LANE per chunk (many chunks in flight: the speculative pass and the early repair rounds).  Replays [warm_start,
// chunk_start) silently from the zero state, then [chunk_start, chunk_end) with closes recorded in the bitmap (chunk bounds
// are multiples of 32 ticks, so words are never shared between lanes).  A warp's 32 chunks stream through a double-buffered
// shared-memory tile filled with cp.async (row = the next CT_R ticks of one lane's chunk), so the HBM latency never sits on
// the serial chain.
template <int MODE>
__global__ void __launch_bounds__(CT_WARPS * 32) k_cusum_tasks(const double *__restrict__ r, const double *__restrict__ lam,
                                                               int64_t n, int64_t first, int64_t CH, int64_t nchunks,
                                                               unsigned *__restrict__ bitmap,
                                                               CusumState *__restrict__ spec_start,
                                                               CusumState *__restrict__ spec_end,
                                                               uint8_t *__restrict__ touched,
                                                               const int64_t *__restrict__ work, int64_t nwork,
                                                               const CusumState *__restrict__ prev_end, double thr) {
    __shared__ double sr[2][CT_WARPS][32][CT_R + 1];
    __shared__ double sl[2][CT_WARPS][32][CT_R + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t tid = ((int64_t)blockIdx.x * CT_WARPS + w) * 32 + lane;
    // work == nullptr: speculative pass over every chunk; otherwise replay of the listed chunks from prev_end[k-1]
    int64_t k = tid;
    if (work) k = tid < nwork ? work[tid] : nchunks;
    if (k > nchunks) k = nchunks;
    const int64_t lo = first + 1 + k * CH;              // first tick of the chunk (ticks <= first are never tested)
    int64_t hi = lo + CH;
    if (hi > n) hi = n;
    bool active = k < nchunks && lo < n;
    int64_t pos = lo - CH;                              // warm-up over the previous chunk
    if (pos < first + 1) pos = first + 1;
    CusumState s{0.0, 0.0};
    if (work && active) { pos = lo; s = prev_end[k - 1]; }   // listed chunks have k >= 1
    unsigned word = 0;
    const int sub = lane >> 3, col = lane & 7;          // 8 lanes copy one 64-byte row; 4 rows per instruction
    // every lane advances by CT_R ticks per round, so a row's position is (start of its lane) + (round offset): the eight
    // rows this lane helps to copy are fetched once (the per-round shuffles were the top stall of the r02 ncu capture)
    const int64_t pos0 = pos, end0 = active ? hi : pos;
    int64_t rowpos[8];
    int rowlen[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int row = 4 * q + sub;
        rowpos[q] = __shfl_sync(0xffffffffu, pos0, row);
        rowlen[q] = (int)(__shfl_sync(0xffffffffu, end0, row) - rowpos[q]);
    }
    auto stage = [&](int buf, int off) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int row = 4 * q + sub;
            if (off + col < rowlen[q]) {
                cp_async8(&sr[buf][w][row][col], r + rowpos[q] + off + col);
                if (MODE < 2 || MODE == 4) cp_async8(&sl[buf][w][row][col], lam + rowpos[q] + off + col);
            }
        }
        cp_async_commit();
    };
    int buf = 0, off = 0;
    stage(0, 0);
    while (__any_sync(0xffffffffu, active)) {
        stage(buf ^ 1, off + CT_R);
        cp_async_wait<1>();
        __syncwarp();
        if (active) {
            const bool in_chunk = pos >= lo;
            if (pos + CT_R <= hi && (in_chunk || pos + CT_R <= lo)) {
                // fast path: the whole row lies inside the data and on one side of the chunk start (rows start at multiples
                // of CT_R ticks from the chunk start, and chunk lengths are multiples of 32)
                if (pos == lo) spec_start[k] = s;
                unsigned bits = 0;
#pragma unroll
                for (int tt = 0; tt < CT_R; tt++)
                    bits |= (unsigned)cusum_step<MODE>(s, sr[buf][w][lane][tt], (MODE < 2 || MODE == 4) ? sl[buf][w][lane][tt] : 0.0, thr) << tt;
                if (in_chunk) {
                    const int64_t rel = pos - (first + 1);
                    const int sh = (int)(rel & 31);
                    word |= bits << sh;
                    if (sh == 32 - CT_R || pos + CT_R == hi) { bitmap[rel >> 5] = word; word = 0; }
                }
            } else {
#pragma unroll 1
                for (int tt = 0; tt < CT_R; tt++) {
                    const int64_t i = pos + tt;
                    if (i >= hi) { active = false; break; }
                    if (i == lo) spec_start[k] = s;
                    const bool close = cusum_step<MODE>(s, sr[buf][w][lane][tt], (MODE < 2 || MODE == 4) ? sl[buf][w][lane][tt] : 0.0, thr);
                    if (i >= lo) {
                        const int64_t rel = i - (first + 1);
                        if (close) word |= 1u << (rel & 31);
                        if ((rel & 31) == 31 || i + 1 == hi) { bitmap[rel >> 5] = word; word = 0; }
                    }
                }
            }
            pos += CT_R;
            if (pos >= hi) active = false;
        }
        __syncwarp();
        buf ^= 1;
        off += CT_R;
    }
    cp_async_wait<0>();
    if (k < nchunks && lo < n) { spec_end[k] = s; if (touched) touched[k] = 1; }
}

// WARP per walker (few inconsistent chunks left: the repair tail).  A walker starts at an inconsistent chunk ("head") from
// its predecessor's end state and KEEPS WALKING through the following chunks -- the whole warp streams the ticks through a
// double-buffered shared-memory tile with coalesced cp.async and every lane steps the same chain from broadcast reads -- until
// (a) the end state of a chunk is bit-identical to the one already stored (the trajectories have coalesced: everything
// downstream is consistent with what is stored), (b) the next chunk is another walker's head in this round, or (c) the data
// ends.  Each chunk is owned by exactly one walker per round, so the writes never collide.  The round loop of the host then
// needs one round per *chain of walkers that ran into each other*, not one per chunk of the longest non-coalescing run.
constexpr int CW_WARPS = 4;
constexpr int CW_TILE = 256;
template <int MODE>
__global__ void __launch_bounds__(CW_WARPS * 32) k_cusum_walk(const double *__restrict__ r, const double *__restrict__ lam,
                                                              int64_t n, int64_t first, int64_t CH, int64_t nchunks,
                                                              unsigned *__restrict__ bitmap, CusumState *__restrict__ ss,
                                                              const CusumState *__restrict__ se,
                                                              CusumState *__restrict__ se_next, uint8_t *__restrict__ touched,
                                                              const uint8_t *__restrict__ is_head,
                                                              const int64_t *__restrict__ work, int64_t nwork, double thr,
                                                              int max_chunks) {
    __shared__ double sr[CW_WARPS][2][CW_TILE];
    __shared__ double sl[CW_WARPS][2][CW_TILE];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t wg = (int64_t)blockIdx.x * CW_WARPS + w;
    if (wg >= nwork) return;
    int64_t c = work[wg];
    CusumState s = se[c - 1];                      // heads have c >= 1
    const int64_t base0 = first + 1, m_ticks = n - base0;
    int64_t rel = c * CH;                          // relative index of the next tick (always a multiple of 32)
    int64_t chunk_end = rel + CH < m_ticks ? rel + CH : m_ticks;
    if (lane == 0) ss[c] = s;
    auto stage = [&](int b, int64_t rel0) {
#pragma unroll
        for (int q = 0; q < CW_TILE / 32; q++) {
            const int64_t idx = base0 + rel0 + q * 32 + lane;
            if (idx < n) {
                cp_async8(&sr[w][b][q * 32 + lane], r + idx);
                if (MODE < 2 || MODE == 4) cp_async8(&sl[w][b][q * 32 + lane], lam + idx);
            }
        }
        cp_async_commit();
    };
    int buf = 0;
    stage(0, rel);
    bool done = false;
    while (!done) {
        stage(buf ^ 1, rel + CW_TILE);
        cp_async_wait<1>();
        __syncwarp();
        const double *tr = sr[w][buf], *tl = sl[w][buf];
#pragma unroll 1
        for (int wd = 0; wd < CW_TILE / 32; wd++) {
            const int64_t wrel = rel + wd * 32;
            const int cnt = (int)(chunk_end - wrel < 32 ? chunk_end - wrel : 32);     // > 0: chunk_end is handled below
            unsigned word = 0;
            if (cnt == 32) {
#pragma unroll
                for (int t = 0; t < 32; t++)
                    if (cusum_step<MODE>(s, tr[wd * 32 + t], (MODE < 2 || MODE == 4) ? tl[wd * 32 + t] : 0.0, thr)) word |= 1u << t;
            } else {
                for (int t = 0; t < cnt; t++)
                    if (cusum_step<MODE>(s, tr[wd * 32 + t], (MODE < 2 || MODE == 4) ? tl[wd * 32 + t] : 0.0, thr)) word |= 1u << t;
            }
            if (lane == 0) bitmap[wrel >> 5] = word;
            if (wrel + cnt >= chunk_end) {        // end of chunk c
                const CusumState old = se[c];
                const bool changed = __double_as_longlong(old.sp) != __double_as_longlong(s.sp) ||
                                     __double_as_longlong(old.sn) != __double_as_longlong(s.sn);
                if (lane == 0) { se_next[c] = s; touched[c] = 1; }
                if (!changed || c + 1 >= nchunks || chunk_end >= m_ticks || is_head[c + 1] || --max_chunks <= 0) { done = true; break; }
                c++;
                chunk_end = chunk_end + CH < m_ticks ? chunk_end + CH : m_ticks;
                if (lane == 0) ss[c] = s;
            }
        }
        rel += CW_TILE;
        buf ^= 1;
        __syncwarp();
    }
    cp_async_wait<0>();
}

// C3: consistency check of the chunk chain; inconsistent chunks become heads and are appended to the work list
__global__ void k_cusum_check(const CusumState *__restrict__ start, const CusumState *__restrict__ end, int64_t nchunks,
                              int64_t *__restrict__ work, uint8_t *__restrict__ is_head, unsigned long long *count) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (k >= nchunks) return;
    const CusumState a = start[k], b = end[k - 1];
    const bool bad = __double_as_longlong(a.sp) != __double_as_longlong(b.sp) || __double_as_longlong(a.sn) != __double_as_longlong(b.sn);
    is_head[k] = bad;
    if (bad) work[atomicAdd(count, 1ull)] = k;
}
// end states written in this round become visible to the next check
__global__ void k_cusum_commit(int64_t nchunks, const CusumState *__restrict__ end_next, CusumState *__restrict__ end,
                               uint8_t *__restrict__ touched) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nchunks && touched[k]) { end[k] = end_next[k]; touched[k] = 0; }
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
__global__ void k_iota_i64(int64_t *out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

// Shared driver of the chunk chain: speculative pass, parallel fix-point, bitmap -> ordered index list.
// idx_out[0] is left for the caller (open marker); idx_out[1..total] are the closing / event ticks.
#define CUSUM_TASKS(grid, block, smem, ...)                                                      \
    do {                                                                                         \
        if (mode == 0) FMK_LAUNCH(ctx, k_cusum_tasks<0>, grid, block, smem, __VA_ARGS__);        \
        else if (mode == 1) FMK_LAUNCH(ctx, k_cusum_tasks<1>, grid, block, smem, __VA_ARGS__);   \
        else if (mode == 2) FMK_LAUNCH(ctx, k_cusum_tasks<2>, grid, block, smem, __VA_ARGS__);   \
        else if (mode == 3) FMK_LAUNCH(ctx, k_cusum_tasks<3>, grid, block, smem, __VA_ARGS__);   \
        else FMK_LAUNCH(ctx, k_cusum_tasks<4>, grid, block, smem, __VA_ARGS__);                  \
    } while (0)
#define CUSUM_WALK(grid, block, smem, ...)                                                       \
    do {                                                                                         \
        if (mode == 0) FMK_LAUNCH(ctx, k_cusum_walk<0>, grid, block, smem, __VA_ARGS__);         \
        else if (mode == 1) FMK_LAUNCH(ctx, k_cusum_walk<1>, grid, block, smem, __VA_ARGS__);    \
        else if (mode == 2) FMK_LAUNCH(ctx, k_cusum_walk<2>, grid, block, smem, __VA_ARGS__);    \
        else if (mode == 3) FMK_LAUNCH(ctx, k_cusum_walk<3>, grid, block, smem, __VA_ARGS__);    \
        else FMK_LAUNCH(ctx, k_cusum_walk<4>, grid, block, smem, __VA_ARGS__);                   \
    } while (0)

// exact_starts (imbalance bars): a callback that, given the chunking, fills pe[0 .. nchunks] with the TRUE state in front of
// every chunk (pe[k] = state before chunk k); the chain then needs no speculation and no repair rounds -- one replay pass.
struct ExactStarts {
    const int8_t *sides;
    int m;
};
static int imbalance_exact_starts(fmk_ctx *ctx, const ExactStarts &es, int64_t n, int64_t base0, int64_t CH, int64_t nchunks,
                                  CusumState *pe);

static int cusum_chain(fmk_ctx *ctx, const Scratch<double> &r, const Scratch<double> &lam, int64_t n, int64_t first, int mode,
                       int64_t **idx_out, int64_t *total_out, double thr = 0.0, const ExactStarts *exact = nullptr) {
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
        // walkers (warp per inconsistent chunk) take over from the lane-per-chunk kernel once few chunks are left
        int64_t walk_below = 1024;
        if (const char *e = getenv("FMK_CUSUM_WALK_BELOW")) walk_below = atoll(e);       // test hook: 0 = never, huge = always
        // chunks a walker may cover per round: the round lasts as long as its longest walk, and a walk that runs into
        // another walker's head has to wait for the next round anyway (measured at 1e9 ticks, see profiles/)
        int walk_max = 2;
        if (const char *e = getenv("FMK_CUSUM_WALK_MAX")) walk_max = atoi(e) > 0 ? atoi(e) : walk_max;
        const int64_t nchunks = cdiv(m_ticks, CH);
        const int64_t nwords = cdiv(m_ticks, 32);
        Scratch<unsigned> bitmap(ctx);
        Scratch<CusumState> ss(ctx), se(ctx), se_next(ctx);
        Scratch<int64_t> work(ctx), dtotal(ctx);
        Scratch<uint8_t> touched(ctx), is_head(ctx);
        Scratch<unsigned long long> dcount(ctx);
        FMK_TRY(bitmap.alloc(nwords)); FMK_TRY(ss.alloc(nchunks)); FMK_TRY(se.alloc(nchunks)); FMK_TRY(se_next.alloc(nchunks));
        FMK_TRY(work.alloc(nchunks)); FMK_TRY(dcount.alloc(1)); FMK_TRY(touched.alloc(nchunks)); FMK_TRY(is_head.alloc(nchunks));
        FMK_TRY(dtotal.alloc(1));
        {
            // both chain kernels want several resident blocks per SM (36 / 32 KB of static shared memory each): ask for the
            // largest shared-memory carve-out, the default one left ONE block per SM (rounds took one wave per 148 blocks)
            static bool once = false;
            if (!once) {
                once = true;
                cudaFuncSetAttribute(k_cusum_tasks<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k_cusum_tasks<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k_cusum_tasks<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k_cusum_tasks<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k_cusum_walk<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k_cusum_walk<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k_cusum_walk<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k_cusum_walk<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k_cusum_tasks<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k_cusum_walk<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                if (getenv("FMK_CUSUM_DEBUG")) {
                    int b0 = 0, b1 = 0;
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_cusum_tasks<0>, CT_WARPS * 32, 0);
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_cusum_walk<0>, CW_WARPS * 32, 0);
                    fprintf(stderr, "[cusum] occupancy: lane kernel %d blocks/SM, walker %d blocks/SM\n", b0, b1);
                }
            }
        }
        FMK_CUDA(ctx, cudaMemsetAsync(touched.p, 0, (size_t)nchunks, ctx->stream));
        FMK_CUDA(ctx, cudaMemsetAsync(is_head.p, 0, (size_t)nchunks, ctx->stream));
        int64_t hrep = 0, rounds = 0;
        if (exact) {
            Scratch<CusumState> pe(ctx);
            FMK_TRY(pe.alloc(nchunks + 1));
            FMK_TRY(imbalance_exact_starts(ctx, *exact, n, first + 1, CH, nchunks, pe.p));
            FMK_LAUNCH(ctx, k_iota_i64, (unsigned)cdiv(nchunks, 256), 256, 0, work.p, nchunks);
            // "replay the listed chunks from prev_end[k - 1]" with prev_end = pe + 1: chunk k starts from pe[k]
            CUSUM_TASKS((unsigned)cdiv(cdiv(nchunks, 32), CT_WARPS), CT_WARPS * 32, 0, (const double *)r.p, (const double *)lam.p, n,
                        first, CH, nchunks, bitmap.p, ss.p, se.p, (uint8_t *)nullptr, (const int64_t *)work.p, nchunks,
                        (const CusumState *)(pe.p + 1), thr);
        } else {
        CUSUM_TASKS((unsigned)cdiv(cdiv(nchunks, 32), CT_WARPS), CT_WARPS * 32, 0, (const double *)r.p, (const double *)lam.p, n,
                    first, CH, nchunks, bitmap.p, ss.p, se.p, (uint8_t *)nullptr, (const int64_t *)nullptr, (int64_t)0,
                    (const CusumState *)nullptr, thr);
        }
        for (; !exact;) {
            unsigned long long hcount = 0;
            FMK_CUDA(ctx, cudaMemsetAsync(dcount.p, 0, 8, ctx->stream));
            if (nchunks > 1)
                FMK_LAUNCH(ctx, k_cusum_check, (unsigned)cdiv(nchunks - 1, 256), 256, 0, (const CusumState *)ss.p,
                           (const CusumState *)se.p, nchunks, work.p, is_head.p, dcount.p);
            FMK_CUDA(ctx, cudaMemcpyAsync(&hcount, dcount.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
            FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (hcount == 0) break;
            const int64_t nw = (int64_t)hcount;
            static const bool dbg = getenv("FMK_CUSUM_DEBUG") != nullptr;
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (dbg) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, ctx->stream); }
            if (nw <= walk_below) {
                CUSUM_WALK((unsigned)cdiv(nw, CW_WARPS), CW_WARPS * 32, 0, (const double *)r.p, (const double *)lam.p, n, first, CH,
                           nchunks, bitmap.p, ss.p, (const CusumState *)se.p, se_next.p, touched.p, (const uint8_t *)is_head.p,
                           (const int64_t *)work.p, nw, thr, walk_max);
            } else {
                CUSUM_TASKS((unsigned)cdiv(cdiv(nw, 32), CT_WARPS), CT_WARPS * 32, 0, (const double *)r.p, (const double *)lam.p, n,
                            first, CH, nchunks, bitmap.p, ss.p, se_next.p, touched.p, (const int64_t *)work.p, nw,
                            (const CusumState *)se.p, thr);
            }
            FMK_LAUNCH(ctx, k_cusum_commit, (unsigned)cdiv(nchunks, 256), 256, 0, nchunks, (const CusumState *)se_next.p, se.p, touched.p);
            if (dbg) {
                float ms = 0.f;
                cudaEventRecord(e1, ctx->stream); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
                fprintf(stderr, "[cusum] round %lld: %lld heads, %s, %.3f ms (CH %lld, %lld chunks)\n", (long long)rounds, (long long)nw,
                        nw <= walk_below ? "walkers" : "lane kernel", ms, (long long)CH, (long long)nchunks);
                cudaEventDestroy(e0); cudaEventDestroy(e1);
            }
            hrep += nw; rounds++;
        }
        // count, allocate, write
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
    unsigned long long hfirst = ~0ull;
    for (int64_t lo = 0, len = 1 << 22; lo < n && hfirst == ~0ull; lo += len, len *= 8) {
        const int64_t hi = lo + len < n ? lo + len : n;
        FMK_LAUNCH(ctx, k_cusum_first, (unsigned)cdiv(hi - lo, 256), 256, 0, (const double *)sg, lo, hi, dfirst.p);
        FMK_CUDA(ctx, cudaMemcpyAsync(&hfirst, dfirst.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    const int64_t first = hfirst == ~0ull ? 0 : (int64_t)hfirst;
    ctx->cusum_filled = 0;
    if (hfirst != ~0ull) {
        // NaNs after the first valid element are what the reference's in-place forward fill changes (logic.py:186-189); when
        // there are none the fill is skipped, and a host wrapper that mirrors the mutation knows it has nothing to copy back
        Scratch<unsigned long long> dn(ctx);
        FMK_TRY(dn.alloc(1));
        FMK_CUDA(ctx, cudaMemsetAsync(dn.p, 0, 8, ctx->stream));
        FMK_LAUNCH(ctx, k_count_nan, ctx->sm_count * 16, 256, 0, (const double *)sg, first, n, dn.p);
        unsigned long long hn = 0;
        FMK_CUDA(ctx, cudaMemcpyAsync(&hn, dn.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->cusum_filled = (int64_t)hn;
        if (hn) FMK_TRY((device_inclusive_scan<FF>(ctx, FFIn{sg}, FFOut{sg, first}, n, (FF *)nullptr)));
    }

    Scratch<double> r(ctx), lam(ctx);
    Scratch<int> nonfinite(ctx);
    FMK_TRY(r.alloc(n)); FMK_TRY(lam.alloc(n)); FMK_TRY(nonfinite.alloc(1));
    FMK_CUDA(ctx, cudaMemsetAsync(nonfinite.p, 0, sizeof(int), ctx->stream));
    FMK_LAUNCH(ctx, k_cusum_prep, (unsigned)cdiv(n, 256), 256, 0, (const int64_t *)t->ts, (const double *)t->price,
               (const double *)sg, n, sigma_floor, sigma_mult, r.p, lam.p, nonfinite.p);
    int hnonfinite = 1;
    FMK_CUDA(ctx, cudaMemcpyAsync(&hnonfinite, nonfinite.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // integer-compare chain (mode 4) when its preconditions hold: finite returns, thresholds >= +0.0 or NaN
    const bool no_fast = getenv("FMK_CUSUM_NO_FAST") != nullptr;             // test hook
    const int chain_mode = (!hnonfinite && sigma_floor >= 0.0 && !no_fast) ? 4 : 0;

    int64_t total = 0;
    int64_t *idx = nullptr;
    FMK_TRY(cusum_chain(ctx, r, lam, n, first, chain_mode, &idx, &total));
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


extern "C" int64_t fmk_cusum_filled_count(fmk_ctx *ctx) { return ctx->cusum_filled; }

// cusum_filter (sampling/filters.py:6-70): host series / thresholds in, device buffer of int64 event indices out.
extern "C" int fmk_cusum_filter(fmk_ctx *ctx, const double *series, int64_t n, const double *threshold, int64_t n_thr,
                                fmk_buf **events_out, int64_t *n_events) {
    FMK_ENTER(ctx);
    *events_out = nullptr; *n_events = 0;
    if (n <= 1) return fmk_fail(ctx, FMK_ERR_ARG, "Input time series must have at least 2 elements.");
    if (n_thr != 1 && n_thr != n)
        return fmk_fail(ctx, FMK_ERR_ARG, "Threshold array must either contain 1 const. element or len(raw_time_series) elements.");
    Scratch<double> x(ctx), th(ctx), r(ctx), lam(ctx);
    FMK_TRY(x.alloc(n)); FMK_TRY(th.alloc(n_thr)); FMK_TRY(r.alloc(n)); FMK_TRY(lam.alloc(n));
    FMK_CUDA(ctx, cudaMemcpyAsync(x.p, series, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(th.p, threshold, (size_t)n_thr * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_LAUNCH(ctx, k_cusum_filter_prep, (unsigned)cdiv(n, 256), 256, 0, (const double *)x.p, (const double *)th.p, n_thr, n,
               r.p, lam.p);
    int64_t *idx = nullptr, total = 0;
    FMK_TRY(cusum_chain(ctx, r, lam, n, 0, 1, &idx, &total));
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


// ---------------------------------------------------------------------------------------------------------------
// a6: tick-imbalance / tick-run bars.  The reference only has stubs (bar/logic.py:224-261 raise NotImplementedError), so
// the semantics are OURS (pinned by our own CPU oracle only -- "parity unpinned"):
//   b_t = the trade's aggressor side when the handle has a side column and use_side != 0, else the tick rule on the prices
//         (the reference's comp_trade_side_vector, bar/utils.py:12-46); index list starts with 0 like logic.py:73-84;
//   kind 0 (imbalance, AFML 2.3.2.1 with a fixed expected imbalance): theta += b_t from the tick after the previous close,
//         the bar closes at the first t with |theta| >= threshold, then theta = 0;
//   kind 1 (runs, AFML 2.3.2.2): buys / sells counted since the previous close, close when max(buys, sells) >= threshold.
// The state is an integer that resets to exactly 0, so the chunk chain of the CUSUM bars applies unchanged (modes 2 / 3).
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_sides_to_r(const int8_t *__restrict__ side, int64_t n, double *__restrict__ r) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = side[i];
        r[i] = b > 0 ? 1.0 : (b < 0 ? -1.0 : 0.0);
    }
}

int fmk_tick_rule_device(fmk_ctx *ctx, const double *price_dev, int64_t n, int8_t *sides_dev);   // ingest.cu

// ---- exact chunk start states for tick-imbalance bars -------------------------------------------------------------------
// theta lives in (-m, m) (m = ceil(threshold)) and every chunk acts on it as a MAP G_k: start state -> end state.  Trajectories
// from different start states do not coalesce on their own (all-buy data: never), so speculation + repair rounds can take as
// many rounds as there are chunks (measured: 155 s at 1e9 ticks, threshold 200).  The map itself, however, costs O(1) per tick
// when it is built BACKWARDS: G_{t-1}(x) = G_t(x + b_t) for |x + b_t| < m, and = G_t(0) for the one x that hits +-m.  Stored
// in a circular buffer that is a pure re-indexing (offset += b_t) plus ONE element write per tick.  A lane builds the map of
// its chunk (2m - 1 entries in shared memory), the maps are composed in two levels (thread per start state inside a group of
// chunks, a short serial hop over groups), and every chunk's TRUE start state drops out -- the forward pass then replays each
// chunk exactly once.  Integer arithmetic throughout: exact by construction.
__global__ void k_imb_backmap(const int8_t *__restrict__ side, int64_t n, int64_t base0, int64_t CH, int64_t nchunks, int m,
                              int Wp, uint16_t *__restrict__ maps) {
    extern __shared__ uint16_t imb_sm[];
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nchunks) return;
    uint16_t *buf = imb_sm + (size_t)threadIdx.x * (Wp + 2);
    const int mask = Wp - 1, W = 2 * m - 1;
    for (int x = -(m - 1); x <= m - 1; x++) buf[x & mask] = (uint16_t)(x + m - 1);     // G at the chunk end: identity
    int off = 0;
    const int64_t lo = base0 + k * CH;
    const int64_t hi = lo + CH < n ? lo + CH : n;
    for (int64_t i = hi - 1; i >= lo; i--) {
        const int sd = side[i];
        if (sd == 0) continue;
        const int b = sd > 0 ? 1 : -1;
        const uint16_t v0 = buf[off & mask];            // G_t(0)
        off += b;
        buf[((b > 0 ? m - 1 : -(m - 1)) + off) & mask] = v0;
    }
    uint16_t *out = maps + k * W;
    for (int x = -(m - 1); x <= m - 1; x++) out[x + m - 1] = buf[(x + off) & mask];
}

// group composite: thread per start state walks the group's chunk maps
__global__ void k_imb_group(const uint16_t *__restrict__ maps, int64_t nchunks, int W, int gs, uint16_t *__restrict__ gmaps) {
    const int64_t g = blockIdx.x;
    const int64_t c0 = g * gs, c1 = c0 + gs < nchunks ? c0 + gs : nchunks;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        int v = x;
        for (int64_t c = c0; c < c1; c++) v = maps[c * W + v];
        gmaps[g * W + x] = (uint16_t)v;
    }
}
// one thread: true state in front of every group
__global__ void k_imb_groups_serial(const uint16_t *__restrict__ gmaps, int64_t ngroups, int W, int zero, int *__restrict__ gstart) {
    int v = zero;
    for (int64_t g = 0; g < ngroups; g++) { gstart[g] = v; v = gmaps[g * W + v]; }
}
// thread per group: true state in front of every chunk -> the chain's state records (theta carried exactly in a double)
__global__ void k_imb_starts(const uint16_t *__restrict__ maps, int64_t nchunks, int W, int gs, int m, const int *__restrict__ gstart,
                             int64_t ngroups, CusumState *__restrict__ pe) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    const int64_t c0 = g * gs, c1 = c0 + gs < nchunks ? c0 + gs : nchunks;
    int v = gstart[g];
    for (int64_t c = c0; c < c1; c++) {
        CusumState st;
        st.sp = (double)(v - (m - 1));
        st.sn = 0.0;
        pe[c] = st;
        v = maps[c * W + v];
    }
    if (c1 == nchunks) { CusumState st; st.sp = (double)(v - (m - 1)); st.sn = 0.0; pe[nchunks] = st; }
}

constexpr int IMB_MAX_M = 2048;     // 2m - 1 map entries per chunk; beyond that the generic chain is used
static int imbalance_exact_starts(fmk_ctx *ctx, const ExactStarts &es, int64_t n, int64_t base0, int64_t CH, int64_t nchunks,
                                  CusumState *pe) {
    const int m = es.m, W = 2 * m - 1;
    int Wp = 2;
    while (Wp < 2 * m) Wp <<= 1;
    // lanes per block: as many private maps as fit ~96 KB of shared memory
    int lanes = (int)((96 * 1024) / ((size_t)(Wp + 2) * sizeof(uint16_t)));
    lanes = lanes >= 128 ? 128 : (lanes >= 64 ? 64 : (lanes >= 32 ? 32 : (lanes >= 8 ? 8 : 1)));
    const size_t smem = (size_t)lanes * (Wp + 2) * sizeof(uint16_t);
    FMK_CUDA(ctx, cudaFuncSetAttribute(k_imb_backmap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int gs = 256;
    const int64_t ngroups = cdiv(nchunks, gs);
    Scratch<uint16_t> maps(ctx), gmaps(ctx);
    Scratch<int> gstart(ctx);
    FMK_TRY(maps.alloc(nchunks * W)); FMK_TRY(gmaps.alloc(ngroups * W)); FMK_TRY(gstart.alloc(ngroups));
    FMK_LAUNCH(ctx, k_imb_backmap, (unsigned)cdiv(nchunks, lanes), lanes, smem, es.sides, n, base0, CH, nchunks, m, Wp, maps.p);
    FMK_LAUNCH(ctx, k_imb_group, (unsigned)ngroups, 256, 0, (const uint16_t *)maps.p, nchunks, W, gs, gmaps.p);
    FMK_LAUNCH(ctx, k_imb_groups_serial, 1, 1, 0, (const uint16_t *)gmaps.p, ngroups, W, m - 1, gstart.p);
    FMK_LAUNCH(ctx, k_imb_starts, (unsigned)cdiv(ngroups, 64), 64, 0, (const uint16_t *)maps.p, nchunks, W, gs, m, (const int *)gstart.p,
               ngroups, pe);
    return FMK_OK;
}

extern "C" int fmk_imbalance_bar_index(fmk_ctx *ctx, const fmk_trades *t, double threshold, int use_side, int kind,
                                       fmk_index **out_ix) {
    FMK_ENTER(ctx);
    *out_ix = nullptr;
    const int64_t n = t->n;
    if (n <= 0) return fmk_fail(ctx, FMK_ERR_ARG, "empty trades");
    if (!(threshold > 0) || !(threshold < 4e15)) return fmk_fail(ctx, FMK_ERR_ARG, "threshold must be positive and finite");
    if (kind != 0 && kind != 1) return fmk_fail(ctx, FMK_ERR_ARG, "kind must be 0 (imbalance) or 1 (runs)");
    ctx->stats[0] = ctx->stats[1] = ctx->stats[2] = 0;
    Scratch<double> r(ctx), lam(ctx);
    Scratch<int8_t> sd(ctx);
    FMK_TRY(r.alloc(n));
    const int8_t *sides = t->side;
    if (!use_side || !sides) {
        FMK_TRY(sd.alloc(n));
        FMK_TRY(fmk_tick_rule_device(ctx, t->price, n, sd.p));
        sides = sd.p;
    }
    FMK_LAUNCH(ctx, k_sides_to_r, ctx->sm_count * 16, 256, 0, sides, n, r.p);
    int64_t total = 0;
    int64_t *idx = nullptr;
    // tick-imbalance bars with a moderate threshold: exact chunk start states from the backward maps, one replay pass;
    // run bars (two counters) and very large thresholds go through the speculative chain
    const double mceil = ceil(threshold);
    ExactStarts es{sides, (int)mceil};
    const bool no_exact = getenv("FMK_IMBALANCE_NO_EXACT") != nullptr;             // test hook: force the speculative chain
    const bool use_exact = kind == 0 && mceil >= 1.0 && mceil <= (double)IMB_MAX_M && !no_exact;
    FMK_TRY(cusum_chain(ctx, r, lam, n, 0, 2 + kind, &idx, &total, threshold, use_exact ? &es : nullptr));
    {
        auto launch = [&]() -> int { FMK_LAUNCH(ctx, k_set_first, 1, 1, 0, idx, (int64_t)0); return FMK_OK; };
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
