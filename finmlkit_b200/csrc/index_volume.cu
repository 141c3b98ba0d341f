// index_volume.cu -- bit-exact parallel volume-bar indexer (reference: finmlkit/bar/logic.py:87-115).
//
// Reference recurrence: cum = v0; for i>=1: cum = fl(cum + v_i); if cum >= T: emit i; cum = 0.0   (no carry-over).
// The reset to exactly 0.0 means the state after a boundary b is known exactly, so the boundary that follows b is a
// pure function next(b) = min{ j > b : seqsum(v[b+1..j]) >= T } of the data.  The boundaries are the orbit of the
// first boundary under `next`, which is a pointer chase.  It is resolved in parallel with two levels of exit tables:
//
//   V1  prefix        P = inclusive scan of v                                     (scan.cuh; 8 B/tick in, 8 B/tick out)
//   V2  k_volume_next next[i] for EVERY tick i: binary search of P for P[i]+T-guard, certain when the candidate
//                     clears T+guard as well; otherwise (exact ties, 2.5 % of bars on quantised sizes -- SURVEY H5)
//                     the reference's sequential float64 sum is replayed from i+1.  guard bounds |seqsum - (P[j]-P[i])|.
//   V3  k_volume_exit0 per 2048-tick chunk, pointer jumping in shared memory: exit0[i] = first orbit element of i
//                     beyond the chunk
//       k_volume_exit1 per 1024-chunk superchunk, backward over chunks: exit1[i] = first orbit element beyond the
//                     superchunk
//       k_volume_chase2 / chase1: serial hops over superchunks, then (parallel over superchunks) over chunks
//       k_volume_count / emit: per chunk, walk next[] from the chunk's entry and write the indices in order
//
// Requires v >= 0 and finite (P monotone); anything else falls back to the exact serial device kernel.
#include <new>
#include "common.cuh"
#include "scan.cuh"

constexpr int VC0 = 2048;            // ticks per chunk
constexpr int VC1 = 1024;            // chunks per superchunk
constexpr int64_t VSC = (int64_t)VC0 * VC1;
constexpr int VN_THREADS = 256;

struct VolIn {
    const double *v;
    __device__ double operator()(int64_t i) const { return v[i]; }
};
struct VolOut {
    double *P;
    int *bad;
    const double *v;
    __device__ void operator()(int64_t i, double cs) const {
        P[i] = cs;
        const double x = v[i];
        if (!(x >= 0.0) || !(x < 1e300)) *bad = 1;
    }
};

// first index in [lo, hi) with P[idx] >= key (hi if none)
__device__ __forceinline__ int64_t lower_bound_f64(const double *__restrict__ P, int64_t lo, int64_t hi, double key) {
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(P + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ int64_t volume_replay(const double *__restrict__ v, int64_t n, int64_t i, double T) {
    double cum = 0.0;
    for (int64_t t = i + 1; t < n; t++) {
        cum = __dadd_rn(cum, v[t]);
        if (cum >= T) return t;
    }
    return n;
}

// Bracket pre-pass: the candidates of a block's 256 ticks are monotone, so the first and the last tick's candidates
// bracket all of them.  One THREAD per bracket: a full binary search is ~30 dependent loads, but 2 n / 256 of them run
// concurrently, so occupancy hides the latency (two threads of every block doing it stalled the whole block for ~15 us).
__global__ void k_volume_brackets(const double *__restrict__ P, int64_t n, double tlo, int64_t nblk, int64_t *__restrict__ br) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= 2 * nblk) return;
    int64_t i = (q >> 1) * VN_THREADS + ((q & 1) ? VN_THREADS - 1 : 0);
    if (i > n - 1) i = n - 1;
    br[q] = lower_bound_f64(P, i + 1, n, __dadd_rn(P[i], tlo));
}

// Ambiguous starts (candidate inside the guard band) are not replayed inline -- that would stall the whole warp on one
// lane's serial sum -- but marked with -1 for k_volume_replay_seg.
constexpr int VN_STAGE = 1536;      // prefix values staged in shared memory for the per-tick searches
__global__ void __launch_bounds__(VN_THREADS) k_volume_next(const double *__restrict__ P, const double *__restrict__ v,
                                                            int64_t n, double T, double guard,
                                                            const int64_t *__restrict__ br, int32_t *__restrict__ next) {
    __shared__ double Ps[VN_STAGE];
    const int64_t i0 = (int64_t)blockIdx.x * VN_THREADS;
    const int64_t i = i0 + threadIdx.x;
    const double tlo = __dadd_rn(T, -guard), thi = __dadd_rn(T, guard);
    const int64_t b0 = br[2 * (int64_t)blockIdx.x], b1 = br[2 * (int64_t)blockIdx.x + 1];
    // stage P[b0 .. b1] (+1 for the certainty test) when it fits: the 256 searches then run on shared memory
    const int64_t span = b1 - b0 + 2;
    const bool staged = span <= VN_STAGE;
    if (staged)
        for (int64_t q = threadIdx.x; q < span; q += VN_THREADS) Ps[q] = b0 + q < n ? __ldg(P + b0 + q) : INFINITY;
    __syncthreads();
    if (i >= n) return;
    const double base = P[i];
    int64_t lo = b0, hi = b1;
    if (lo < i + 1) lo = i + 1;
    if (hi < lo) hi = lo;
    // candidate: first j > i whose approximate bar volume reaches T - guard
    const double key = __dadd_rn(base, tlo);
    int64_t j;
    double pj;
    if (staged) {
        int64_t l = lo - b0, h = hi - b0;
        while (l < h) {
            const int64_t mid = l + ((h - l) >> 1);
            if (Ps[mid] < key) l = mid + 1; else h = mid;
        }
        j = b0 + l;
        pj = Ps[l];                 // l <= b1 - b0 + ... < span
    } else {
        j = lower_bound_f64(P, lo, hi, key);
        pj = j < n ? __ldg(P + j) : 0.0;
    }
    int64_t r;
    if (j >= n) r = n;                                            // even T - guard is never reached: no boundary
    else if (pj >= __dadd_rn(base, thi)) r = j;                  // clears T + guard: certain
    else r = -1;                                                 // within the guard band: the reference's own sum decides
    next[i] = (int32_t)r;
}

// Exact replay of the ambiguous starts (exact ties: 2.6 % of all ticks on exchange-quantised sizes, i.e. ~27 per 1024
// ticks, each needing the reference's sequential sum over a whole bar).  One thread per entry reading its ~1300 sizes from
// L2 moved 270 GB at 1e9 ticks (94 ms).  Neighbouring entries sum over almost the same ticks, so a warp now owns a
// 1024-tick segment of starts: it collects the segment's entries, and for 32 of them at a time streams the union of their
// ranges through a shared-memory tile (coalesced loads, once) while every lane runs its own sequential sum from the tile.
// Occupancy: every lane runs one long dependent chain of float64 adds, so throughput comes from resident warps, and those are
// bounded by shared memory: 6 KB per warp (512-tick tile + 16-bit entry list) keeps ~36 warps per SM resident (the r01
// layout, 12 KB per warp, ran at 15 -- ncu `warps active 24 %`).
constexpr int VR_WARPS = 4;
constexpr int VR_SEG = 1024;      // starts per warp-segment
constexpr int VR_TILE = 512;      // ticks per staged tile
__global__ void __launch_bounds__(VR_WARPS * 32) k_volume_replay_seg(const double *__restrict__ v, int64_t n, double T,
                                                                     int32_t *__restrict__ next,
                                                                     unsigned long long *nreplays) {
    __shared__ double tile_s[VR_WARPS][VR_TILE];
    __shared__ uint16_t list_s[VR_WARPS][VR_SEG];     // entry offsets inside the segment
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double *tile = tile_s[w];
    uint16_t *list = list_s[w];
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nseg = (n + VR_SEG - 1) / VR_SEG;
    for (int64_t seg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; seg < nseg; seg += nwarps) {
        const int64_t base = seg * VR_SEG;
        int cnt = 0;
#pragma unroll 4
        for (int g = 0; g < VR_SEG / 32; g++) {
            const int64_t i = base + g * 32 + lane;
            const bool amb = i < n && next[i] == -1;
            const unsigned m = __ballot_sync(0xffffffffu, amb);
            if (amb) list[cnt + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(i - base);
            cnt += __popc(m);
        }
        __syncwarp();
        if (cnt == 0) continue;
        if (lane == 0) atomicAdd(nreplays, (unsigned long long)cnt);
        for (int b0 = 0; b0 < cnt; b0 += 32) {
            const bool have = b0 + lane < cnt;
            const int64_t i = have ? base + list[b0 + lane] : n;
            int64_t pos = i + 1;                    // next tick this lane adds
            double cum = 0.0;
            int64_t res = n;
            bool active = have && pos < n;
            // entries are ascending in i: lane 0 has the smallest start
            int64_t a = __shfl_sync(0xffffffffu, pos, 0);
            while (__any_sync(0xffffffffu, active) && a < n) {
                __syncwarp();
#pragma unroll 8
                for (int q = lane; q < VR_TILE; q += 32) tile[q] = a + q < n ? __ldg(v + a + q) : 0.0;
                __syncwarp();
                // a lane whose first tick lies beyond this tile (the tile is shorter than the segment of starts) just waits
                if (active && pos < a + VR_TILE) {
                    int64_t end = a + VR_TILE < n ? a + VR_TILE : n;
                    int64_t t = pos > a ? pos : a;
                    // sizes are >= 0 on this path (anything else took the serial fallback), so the running sum is
                    // monotone: add eight ticks with independent loads, test once, and only then locate the crossing
                    for (; t + 8 <= end; t += 8) {
                        const double *x = tile + (t - a);
                        const double x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3], x4 = x[4], x5 = x[5], x6 = x[6], x7 = x[7];
                        const double c0 = __dadd_rn(cum, x0), c1 = __dadd_rn(c0, x1), c2 = __dadd_rn(c1, x2), c3 = __dadd_rn(c2, x3),
                                     c4 = __dadd_rn(c3, x4), c5 = __dadd_rn(c4, x5), c6 = __dadd_rn(c5, x6), c7 = __dadd_rn(c6, x7);
                        if (c7 >= T) {
                            const int k = c0 >= T ? 0 : c1 >= T ? 1 : c2 >= T ? 2 : c3 >= T ? 3 : c4 >= T ? 4 : c5 >= T ? 5 : c6 >= T ? 6 : 7;
                            res = t + k; active = false;
                            break;
                        }
                        cum = c7;
                    }
                    if (active)
                        for (; t < end; t++) {
                            cum = __dadd_rn(cum, tile[t - a]);
                            if (cum >= T) { res = t; active = false; break; }
                        }
                    pos = end;
                    if (pos >= n) active = false;
                }
                a += VR_TILE;
            }
            if (have) next[i] = (int32_t)res;
            __syncwarp();
        }
        __syncwarp();
    }
}

// first boundary: cum = v[0]; for i >= 1: cum += v[i]; cum >= T -> i   (logic.py:106-111; tick 0 is never tested)
__global__ void k_volume_first(const double *__restrict__ v, int64_t n, double T, int64_t *first) {
    double cum = v[0];
    int64_t r = n;
    for (int64_t t = 1; t < n; t++) {
        cum = __dadd_rn(cum, v[t]);
        if (cum >= T) { r = t; break; }
    }
    *first = r;
}

// exit0[i] = first element of the orbit of i that lies beyond i's chunk (n if the orbit ends)
__global__ void __launch_bounds__(256) k_volume_exit0(const int32_t *__restrict__ next, int64_t n,
                                                      int32_t *__restrict__ exit0) {
    __shared__ int32_t tgt[VC0];
    const int64_t base = (int64_t)blockIdx.x * VC0;
    const int64_t end = base + VC0 < n ? base + VC0 : n;
    for (int k = threadIdx.x; k < VC0; k += 256) tgt[k] = base + k < n ? next[base + k] : (int32_t)n;
    __syncthreads();
    for (int round = 0; round < 12; round++) {
        int32_t nv[VC0 / 256];
        int ch = 0;
#pragma unroll
        for (int q = 0; q < VC0 / 256; q++) {
            const int k = threadIdx.x + q * 256;
            int32_t t = tgt[k];
            if ((int64_t)t < end) { t = tgt[t - base]; ch = 1; }
            nv[q] = t;
        }
        // barrier + block-wide vote in one step: every thread sees the same verdict (a shared flag that is reset by
        // thread 0 at the top of the next round would race with the threads still reading it)
        const int any = __syncthreads_or(ch);
#pragma unroll
        for (int q = 0; q < VC0 / 256; q++) tgt[threadIdx.x + q * 256] = nv[q];
        __syncthreads();
        if (!any) break;
    }
    for (int k = threadIdx.x; k < VC0; k += 256)
        if (base + k < n) exit0[base + k] = tgt[k];
}

// exit1[i] = first orbit element beyond i's superchunk; chunks are processed back to front so that the exit1 of
// every later chunk in the superchunk is final when it is gathered.
__global__ void __launch_bounds__(256) k_volume_exit1(const int32_t *__restrict__ exit0, int64_t n,
                                                      int32_t *exit1) {
    const int64_t sbase = (int64_t)blockIdx.x * VSC;
    const int64_t send = sbase + VSC < n ? sbase + VSC : n;
    const int64_t nchunks = (send - sbase + VC0 - 1) / VC0;
    for (int64_t c = nchunks - 1; c >= 0; c--) {
        const int64_t base = sbase + c * VC0;
        for (int k = threadIdx.x; k < VC0; k += 256) {
            const int64_t i = base + k;
            if (i < send) {
                int32_t e = exit0[i];
                if ((int64_t)e < send) e = exit1[e];   // written in an earlier iteration (later chunk)
                exit1[i] = e;
            }
        }
        __threadfence_block();
        __syncthreads();
    }
}

// serial hops over superchunks: entryS[S] = first boundary inside superchunk S (-1 if none)
__global__ void k_volume_chase2(const int32_t *__restrict__ exit1, const int64_t *first, int64_t n,
                                int64_t *entryS) {
    int64_t e = *first;
    while (e < n) {
        entryS[e / VSC] = e;
        const int64_t nx = exit1[e];
        if (nx <= e) break;          // next() always moves forward; never spin on a corrupt table
        e = nx;
    }
}

// per superchunk: entryC[c] = first boundary inside chunk c (-1 if none)
__global__ void k_volume_chase1(const int32_t *__restrict__ exit0, const int64_t *__restrict__ entryS, int64_t nS,
                                int64_t n, int64_t *entryC) {
    const int64_t S = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (S >= nS) return;
    int64_t e = entryS[S];
    if (e < 0) return;
    const int64_t send = (S + 1) * VSC < n ? (S + 1) * VSC : n;
    while (e < send) {
        entryC[e / VC0] = e;
        const int64_t nx = exit0[e];
        if (nx <= e) break;
        e = nx;
    }
}

__global__ void k_volume_count(const int32_t *__restrict__ next, const int64_t *__restrict__ entryC, int64_t nC,
                               int64_t n, int64_t *counts) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    int64_t e = entryC[c], cnt = 0;
    const int64_t cend = (c + 1) * VC0 < n ? (c + 1) * VC0 : n;
    while (e >= 0 && e < cend) { cnt++; const int64_t nx = next[e]; if (nx <= e) break; e = nx; }
    counts[c] = cnt;
}

struct CntIn {
    const int64_t *c;
    __device__ int64_t operator()(int64_t i) const { return c[i]; }
};
struct CntOut {
    int64_t *off;
    __device__ void operator()(int64_t i, int64_t cs) const { off[i + 1] = cs; }
};

__global__ void k_volume_emit(const int32_t *__restrict__ next, const int64_t *__restrict__ entryC,
                              const int64_t *__restrict__ off, int64_t nC, int64_t n, int64_t *out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    int64_t e = entryC[c];
    int64_t w = 1 + off[c];            // out[0] = 0 is the open marker
    const int64_t cend = (c + 1) * VC0 < n ? (c + 1) * VC0 : n;
    while (e >= 0 && e < cend) { out[w++] = e; const int64_t nx = next[e]; if (nx <= e) break; e = nx; }
    if (c == 0) out[0] = 0;
}

// exact serial fallback (negative / non-finite volumes, NaN threshold ...): two passes, count then write
__global__ void k_volume_serial(const double *__restrict__ v, int64_t n, double T, int64_t *out, int64_t cap,
                                int64_t *count) {
    int64_t m = 0;
    if (out && m < cap) out[m] = 0;
    m++;
    double cum = v[0];
    for (int64_t i = 1; i < n; i++) {
        cum = __dadd_rn(cum, v[i]);
        if (cum >= T) {
            if (out && m < cap) out[m] = i;
            m++;
            cum = 0.0;
        }
    }
    *count = m;
}

template <typename T>
static int fetch(fmk_ctx *ctx, T *host, const T *dev) {
    FMK_CUDA(ctx, cudaMemcpyAsync(host, dev, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

static int finish_index(fmk_ctx *ctx, const fmk_trades *t, int64_t *idx, int64_t m, fmk_index **out_ix) {
    fmk_index *ix = new (std::nothrow) fmk_index();
    if (!ix) { fmk_dfree(ctx, idx); return FMK_ERR_ALLOC; }
    memset(ix, 0, sizeof(*ix));
    ix->m = m;
    ix->n_ticks = t->n;
    ix->sorted = 1;
    ix->close_idx = idx;
    int rc = fmk_gather_close_ts(ctx, t, ix);
    if (rc) { fmk_index_free(ctx, ix); return rc; }
    *out_ix = ix;
    return FMK_OK;
}

static int volume_serial(fmk_ctx *ctx, const fmk_trades *t, double T, fmk_index **out_ix) {
    Scratch<int64_t> cnt(ctx);
    FMK_TRY(cnt.alloc(1));
    FMK_LAUNCH(ctx, k_volume_serial, 1, 1, 0, t->amount, t->n, T, (int64_t *)nullptr, (int64_t)0, cnt.p);
    int64_t m = 0;
    FMK_TRY(fetch(ctx, &m, cnt.p));
    int64_t *idx = nullptr;
    FMK_TRY(fmk_dalloc(ctx, &idx, m));
    FMK_LAUNCH(ctx, k_volume_serial, 1, 1, 0, t->amount, t->n, T, idx, m, cnt.p);
    ctx->stats[1]++;
    return finish_index(ctx, t, idx, m, out_ix);
}

int fmk_volume_index_impl(fmk_ctx *ctx, const fmk_trades *t, double T, fmk_index **out_ix) {
    FMK_ENTER(ctx);
    *out_ix = nullptr;
    const int64_t n = t->n;
    if (n <= 0) return fmk_fail(ctx, FMK_ERR_ARG, "empty trades");
    ctx->stats[0] = ctx->stats[1] = ctx->stats[2] = 0;
    if (!(T > 0) || !(T < 1e300) || n >= 2147483000ll) return volume_serial(ctx, t, T, out_ix);

    Scratch<double> P(ctx), tot(ctx);
    Scratch<int> bad(ctx);
    FMK_TRY(P.alloc(n)); FMK_TRY(tot.alloc(1)); FMK_TRY(bad.alloc(1));
    FMK_CUDA(ctx, cudaMemsetAsync(bad.p, 0, sizeof(int), ctx->stream));
    FMK_TRY(device_inclusive_scan<double>(ctx, VolIn{t->amount}, VolOut{P.p, bad.p, t->amount}, n, tot.p));
    double total = 0;
    int hbad = 0;
    FMK_CUDA(ctx, cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FMK_TRY(fetch(ctx, &total, tot.p));
    if (hbad || !(total >= 0) || !(total < 1e300)) return volume_serial(ctx, t, T, out_ix);
    // guard >= |seqsum(v[i+1..j]) - (P[j]-P[i])|: scan rounding (<= ~50 eps P_total, taken 4096x) + the sequential
    // sum's own rounding (<= (j-i) eps T, taken with j-i = n) + the two key additions
    const double eps = 1.1102230246251565e-16;
    const double guard = eps * (4096.0 * total + 4.0 * (double)n * T + 16.0 * T);

    Scratch<int32_t> next(ctx), exit0(ctx), exit1(ctx);
    Scratch<int64_t> first(ctx), entryS(ctx), entryC(ctx), counts(ctx), off(ctx), replays(ctx);
    const int64_t nC = cdiv(n, VC0), nS = cdiv(n, VSC);
    FMK_TRY(next.alloc(n)); FMK_TRY(exit0.alloc(n)); FMK_TRY(exit1.alloc(n));
    FMK_TRY(first.alloc(1)); FMK_TRY(entryS.alloc(nS)); FMK_TRY(entryC.alloc(nC)); FMK_TRY(counts.alloc(nC));
    FMK_TRY(off.alloc(nC + 1)); FMK_TRY(replays.alloc(1));
    FMK_CUDA(ctx, cudaMemsetAsync(replays.p, 0, 8, ctx->stream));
    FMK_CUDA(ctx, cudaMemsetAsync(entryS.p, 0xff, (size_t)nS * 8, ctx->stream));
    FMK_CUDA(ctx, cudaMemsetAsync(entryC.p, 0xff, (size_t)nC * 8, ctx->stream));
    FMK_CUDA(ctx, cudaMemsetAsync(off.p, 0, 8, ctx->stream));
    {
        const int64_t nblk = cdiv(n, VN_THREADS);
        Scratch<int64_t> br(ctx);
        FMK_TRY(br.alloc(2 * nblk));
        FMK_LAUNCH(ctx, k_volume_brackets, (unsigned)cdiv(2 * nblk, 256), 256, 0, (const double *)P.p, n,
                   T - guard, nblk, br.p);
        FMK_LAUNCH(ctx, k_volume_next, (unsigned)nblk, VN_THREADS, 0, (const double *)P.p, t->amount, n, T, guard,
                   (const int64_t *)br.p, next.p);
    }
    {
        int64_t blocks = cdiv(cdiv(n, VR_SEG), VR_WARPS);
        const int64_t maxb = (int64_t)ctx->sm_count * 32;
        if (blocks > maxb) blocks = maxb;
        FMK_LAUNCH(ctx, k_volume_replay_seg, (unsigned)blocks, VR_WARPS * 32, 0, t->amount, n, T, next.p,
                   (unsigned long long *)replays.p);
    }
    FMK_LAUNCH(ctx, k_volume_first, 1, 1, 0, t->amount, n, T, first.p);
    FMK_LAUNCH(ctx, k_volume_exit0, (unsigned)nC, 256, 0, (const int32_t *)next.p, n, exit0.p);
    FMK_LAUNCH(ctx, k_volume_exit1, (unsigned)nS, 256, 0, (const int32_t *)exit0.p, n, exit1.p);
    FMK_LAUNCH(ctx, k_volume_chase2, 1, 1, 0, (const int32_t *)exit1.p, (const int64_t *)first.p, n, entryS.p);
    FMK_LAUNCH(ctx, k_volume_chase1, (unsigned)cdiv(nS, 64), 64, 0, (const int32_t *)exit0.p, (const int64_t *)entryS.p, nS, n, entryC.p);
    FMK_LAUNCH(ctx, k_volume_count, (unsigned)cdiv(nC, 128), 128, 0, (const int32_t *)next.p, (const int64_t *)entryC.p, nC, n, counts.p);
    FMK_TRY(device_inclusive_scan<int64_t>(ctx, CntIn{counts.p}, CntOut{off.p}, nC, (int64_t *)nullptr));
    int64_t nb = 0, hrep = 0;
    FMK_CUDA(ctx, cudaMemcpyAsync(&hrep, replays.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_TRY(fetch(ctx, &nb, off.p + nC));
    ctx->stats[0] = nC; ctx->stats[1] = hrep; ctx->stats[2] = 1;
    int64_t *idx = nullptr;
    FMK_TRY(fmk_dalloc(ctx, &idx, nb + 1));
    auto emit = [&]() -> int {
        FMK_LAUNCH(ctx, k_volume_emit, (unsigned)cdiv(nC, 128), 128, 0, (const int32_t *)next.p, (const int64_t *)entryC.p,
                   (const int64_t *)off.p, nC, n, idx);
        return FMK_OK;
    };
    const int erc = emit();
    if (erc) { fmk_dfree(ctx, idx); return erc; }
    return finish_index(ctx, t, idx, nb + 1, out_ix);
}

