// index_dollar.cu -- device pipeline of the bit-exact dollar-bar indexer (see dollar_core.h for the algorithm and
// its exactness argument; reference: finmlkit/bar/logic.py:118-149).
//
//   A1  k_dollar_chunk_sums : double-double sum of fl(p*v) per chunk of CH ticks        (reads 16 B/tick, coalesced)
//   A2  k_dollar_prefix     : exclusive double-double scan over chunks -> (K_in, carry) guess per chunk
//   B   k_dollar_tasks      : one thread per chunk: locate first boundary, replay 4 exact chains (reads 16 B/tick)
//   C   k_dollar_chain      : single-block parallel composition of the per-task integer transfer functions,
//                             certification of every task against the true carried state
//   S   k_dollar_serial     : exact serial replay from the last certified state (repair path only)
#include <stdlib.h>
#include <new>
#include "common.cuh"
#include "dollar_core.h"
#include "scan.cuh"

constexpr int DOLLAR_CH = 4096;          // default ticks per task; large inputs pick a wave-filling size (dollar_pick_chunk)
constexpr int DOLLAR_A_THREADS = 256;

struct LdG {
    const double *a;
    __device__ __forceinline__ double operator()(int64_t i) const { return __ldg(a + i); }
};

__global__ void __launch_bounds__(DOLLAR_A_THREADS) k_dollar_chunk_sums(const double *__restrict__ p,
                                                                        const double *__restrict__ v, int64_t n,
                                                                        int CH, dd_t *__restrict__ sums) {
    __shared__ dd_t sm[DOLLAR_A_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * CH;
    dd_t s = {0.0, 0.0};
#pragma unroll 8
    for (int k = 0; k < CH / DOLLAR_A_THREADS; k++) {
        const int64_t i = base + threadIdx.x + (int64_t)k * DOLLAR_A_THREADS;
        if (i < n) s = dd_add_d(s, __dmul_rn(__ldg(p + i), __ldg(v + i)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dd_t y;
        y.hi = __shfl_xor_sync(0xffffffffu, s.hi, o);
        y.lo = __shfl_xor_sync(0xffffffffu, s.lo, o);
        s = dd_add(s, y);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm[w] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        dd_t t = sm[0];
        for (int k = 1; k < DOLLAR_A_THREADS / 32; k++) t = dd_add(t, sm[k]);
        sums[blockIdx.x] = t;
    }
}

// Exclusive double-double scan over the chunk sums (scan.cuh, 3 small grid-wide kernels) -> the (K_in, carry) guess of
// every chunk and the grand total.
struct DD {
    double hi, lo;
    __device__ DD() {}
    __device__ explicit DD(int) { hi = 0.0; lo = 0.0; }
};
__device__ __forceinline__ DD operator+(const DD &a, const DD &b) {
    dd_t x = {a.hi, a.lo}, y = {b.hi, b.lo};
    dd_t r = dd_add(x, y);
    DD o;
    o.hi = r.hi; o.lo = r.lo;
    return o;
}
__device__ __forceinline__ DD __shfl_up_sync(unsigned m, const DD &x, int o) {
    DD r;
    r.hi = ::__shfl_up_sync(m, x.hi, o);
    r.lo = ::__shfl_up_sync(m, x.lo, o);
    return r;
}
struct SumIn {
    const dd_t *s;
    __device__ DD operator()(int64_t k) const { DD d; d.hi = s[k].hi; d.lo = s[k].lo; return d; }
};
struct GuessOut {
    int64_t *K_in;
    double *carry;
    int64_t nt;
    double T;
    __device__ void operator()(int64_t k, const DD &incl) const {
        // inclusive prefix through chunk k = exclusive prefix of chunk k+1
        if (k == 0) { K_in[0] = 0; carry[0] = 0.0; }
        if (k + 1 < nt) {
            dd_t P = {incl.hi, incl.lo};
            int64_t K;
            double c;
            dollar_guess(P, T, &K, &c);
            K_in[k + 1] = K;
            carry[k + 1] = c;
        }
    }
};

// One lane per task; a warp's 32 tasks stream their ticks through a shared-memory tile that the warp fills with
// coalesced asynchronous copies (cp.async: row r = the next DT_R ticks of lane r's task).  The tile is double
// buffered: the copies of round k+1 are in flight while round k is replayed, so the HBM latency never sits on the
// serial float chain.
constexpr int DT_WARPS = 4;
constexpr int DT_R = 8;

// PF = L2 prefetch size of the copy (0: none): a task row advances 64 bytes per round, so without it DRAM serves ~85 000
// interleaved 64-byte streams; with .L2::128B / .L2::256B the row's next rounds are already in L2.
template <int PF = 0>
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    if (PF == 256) asm volatile("cp.async.ca.shared.global.L2::256B [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
    else if (PF == 128) asm volatile("cp.async.ca.shared.global.L2::128B [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int PF>
__global__ void __launch_bounds__(DT_WARPS * 32) k_dollar_tasks_t(const double *__restrict__ p, const double *__restrict__ v,
                                                                DollarParams P, int64_t nt,
                                                                const int64_t *__restrict__ K_in,
                                                                const double *__restrict__ carry,
                                                                int64_t *__restrict__ out,
                                                                DollarTaskRec *__restrict__ recs) {
    __shared__ double sp[2][DT_WARPS][32][DT_R + 1];
    __shared__ double sv[2][DT_WARPS][32][DT_R + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t k = ((int64_t)blockIdx.x * DT_WARPS + w) * 32 + lane;
    const int64_t n = P.n;
    DollarTask t;
    int64_t pos = 0;
    bool active = k < nt;
    if (active) pos = dollar_task_init(t, P, k, k > 0 ? carry[k] : 0.0, k > 0 ? K_in[k] : 0, k == 0 ? __dmul_rn(p[0], v[0]) : 0.0);
    else { t.B = -1; t.K = 0; t.cnt = 0; t.end_idx = -2; t.nch = DC_NCH; t.start_units = 0; t.phase = 3; }
    // Every task of the warp advances DT_R ticks per round, so row r of the tile (task kw + r) is always at tick
    // (kw + r) * CH + adv: the copy addresses need no per-row hand-over (the shuffle version spent 170 of the ~600 warp
    // instructions of a round on it), only the ballot of the lanes that are still replaying.
    const int sub = lane >> 3, col = lane & 7;     // 8 lanes copy one 64-byte row; 4 rows per instruction
    const int64_t kw = ((int64_t)blockIdx.x * DT_WARPS + w) * 32;
    const int64_t row0 = (kw + sub) * P.CH + col, rowstep = 4 * P.CH;
    int skip = 0;
    if (k == 0) { pos = 0; skip = 1; }             // task 0 starts at tick 1 with c = p0 * v0: its row still starts at tick 0
    auto stage = [&](int buf, int64_t adv, unsigned act) {
        const bool inside = (kw + 32) * P.CH + adv + DT_R <= n;      // warp-uniform: no row of this round reaches past n
        if (inside) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int row = 4 * q + sub;
                const int64_t idx = row0 + q * rowstep + adv;
                if ((act >> row) & 1u) {
                    cp_async8<PF>(&sp[buf][w][row][col], p + idx);
                    cp_async8<PF>(&sv[buf][w][row][col], v + idx);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int row = 4 * q + sub;
                const int64_t idx = row0 + q * rowstep + adv;
                if (((act >> row) & 1u) && idx < n) {
                    cp_async8<PF>(&sp[buf][w][row][col], p + idx);
                    cp_async8<PF>(&sv[buf][w][row][col], v + idx);
                }
            }
        }
        cp_async_commit();
    };
    int buf = 0;
    int64_t adv = 0;
    stage(0, 0, __ballot_sync(0xffffffffu, active));
    while (__any_sync(0xffffffffu, active)) {
        stage(buf ^ 1, adv + DT_R, __ballot_sync(0xffffffffu, active));          // prefetch the next round
        adv += DT_R;
        cp_async_wait<1>();                           // the current round's copies have landed
        __syncwarp();
        if (active) {
            int m = DT_R;
            if (pos + DT_R > n) m = (int)(n - pos);
            int tt = skip;
            skip = 0;
            if (t.phase == 1) {   // approximate walk to the first boundary of the chunk
#pragma unroll 1
                for (; tt < m && t.phase == 1; tt++) {
                    const double d = __dmul_rn(sp[buf][w][lane][tt], sv[buf][w][lane][tt]);
                    if (dollar_task_consume(t, P, pos + tt, d, out)) active = false;
                }
            }
#ifndef DC_EXPLICIT_CHAINS
            if (t.phase == 2) {   // exact replay: the tight loop
                // the next tick's product is formed while the current one walks the dependent chain (the shared-memory
                // load -> DMUL latency was the largest single stall of the loop); column DT_R is the row's padding
                double d = __dmul_rn(sp[buf][w][lane][tt], sv[buf][w][lane][tt]);
#pragma unroll 1
                for (; tt < m; tt++) {
                    const double dn = __dmul_rn(sp[buf][w][lane][tt + 1], sv[buf][w][lane][tt + 1]);
                    if (dollar_virtual_tick(t, d, P)) {
                        if (dollar_task_emit(t, P, pos + tt, out)) { active = false; break; }
                    }
                    d = dn;
                }
            }
#else
            for (; tt < m && active; tt++) {
                const double d = __dmul_rn(sp[buf][w][lane][tt], sv[buf][w][lane][tt]);
                if (dollar_task_consume(t, P, pos + tt, d, out)) active = false;
            }
#endif
            pos += DT_R;
            if (pos >= n) active = false;
        }
        __syncwarp();
        buf ^= 1;
    }
    cp_async_wait<0>();
    if (k < nt) {
        DollarTaskRec rec;
        dollar_task_finish(t, P, &rec);
        recs[k] = rec;
    }
}

// The shipped kernel copies with .L2::256B (measured at 1e9 ticks: 3.61 ms without, 3.61 ms with 128B, 3.32 ms with 256B);
// FMK_DOLLAR_L2PF = 0 / 128 launches the other two (profiling A/B).
static constexpr auto k_dollar_tasks = &k_dollar_tasks_t<256>;
static constexpr auto k_dollar_tasks_pf128 = &k_dollar_tasks_t<128>;
static constexpr auto k_dollar_tasks_pf0 = &k_dollar_tasks_t<0>;

// status block shared between chain / serial kernels and the host
struct DollarStatus {
    int64_t event;      // 0 = range exhausted, 1 = done, 2 = failure (resync after k_ev), 4 = failure (resync at k_ev)
    int64_t k_ev;
    int64_t s, pos, K;  // last certified boundary state (s in units of u)
    int64_t K_total;
    double c;           // serial start state as a double (used when the state is not a multiple of u)
    int64_t use_c;
    unsigned long long ev_min;   // scratch: smallest task index that raised an event in the current chain pass
};

// ---- carry chain as a 3-kernel scan over tasks [k0, nt) --------------------------------------------------------
constexpr int DC_THREADS = 256;
struct ChainElem { DollarXfer f; long long last; };

__device__ __forceinline__ ChainElem chain_elem(const DollarTaskRec &t, int64_t k) {
    ChainElem e;
    e.f = dollar_xfer_identity();
    e.last = -1;
    if (t.start_idx >= 0) {
        if (t.nch == DC_NCH && t.end_idx != -2) e.f = dollar_task_xfer(t);
        e.last = k;
    }
    return e;
}
// a then b
__device__ __forceinline__ ChainElem chain_combine(const ChainElem &a, const ChainElem &b) {
    ChainElem r;
    r.f = dollar_xfer_compose(a.f, b.f);
    r.last = b.last >= 0 ? b.last : a.last;
    return r;
}
// block-wide inclusive scan (Hillis-Steele in shared memory); returns the inclusive value of this thread
__device__ ChainElem chain_block_scan(ChainElem x, ChainElem *buf /* 2 * DC_THREADS */) {
    int cur = 0;
    buf[threadIdx.x] = x;
    __syncthreads();
#pragma unroll 1
    for (int o = 1; o < DC_THREADS; o <<= 1) {
        ChainElem y = buf[cur * DC_THREADS + threadIdx.x];
        if ((int)threadIdx.x >= o) y = chain_combine(buf[cur * DC_THREADS + threadIdx.x - o], y);
        buf[(cur ^ 1) * DC_THREADS + threadIdx.x] = y;
        cur ^= 1;
        __syncthreads();
    }
    ChainElem r = buf[cur * DC_THREADS + threadIdx.x];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(DC_THREADS) k_dollar_chain_reduce(const DollarTaskRec *__restrict__ recs, int64_t nt,
                                                                    int64_t k0, ChainElem *agg) {
    __shared__ ChainElem buf[2 * DC_THREADS];
    const int64_t k = k0 + (int64_t)blockIdx.x * DC_THREADS + threadIdx.x;
    ChainElem e;
    e.f = dollar_xfer_identity(); e.last = -1;
    if (k < nt) e = chain_elem(recs[k], k);
    ChainElem inc = chain_block_scan(e, buf);
    if (threadIdx.x == DC_THREADS - 1) agg[blockIdx.x] = inc;
}

// single block: in-place exclusive scan of the per-block aggregates
__global__ void __launch_bounds__(DC_THREADS) k_dollar_chain_top(ChainElem *agg, int64_t nblocks) {
    __shared__ ChainElem buf[2 * DC_THREADS];
    __shared__ ChainElem carry;
    if (threadIdx.x == 0) { carry.f = dollar_xfer_identity(); carry.last = -1; }
    __syncthreads();
    for (int64_t b = 0; b < nblocks; b += DC_THREADS) {
        const int64_t j = b + threadIdx.x;
        ChainElem e;
        e.f = dollar_xfer_identity(); e.last = -1;
        if (j < nblocks) e = agg[j];
        ChainElem inc = chain_block_scan(e, buf);
        const ChainElem c = carry;
        // exclusive = carry + (inclusive of previous thread)
        buf[threadIdx.x] = inc;
        __syncthreads();
        ChainElem ex = c;
        if (threadIdx.x > 0) ex = chain_combine(c, buf[threadIdx.x - 1]);
        if (j < nblocks) agg[j] = ex;
        __syncthreads();
        if (threadIdx.x == DC_THREADS - 1) carry = chain_combine(c, inc);
        __syncthreads();
    }
}

// mode 0: find the first task that raises an event (atomicMin into st->ev_min)
// mode 1: the thread owning that task (or the last task when there is no event) writes the walk state
__global__ void __launch_bounds__(DC_THREADS) k_dollar_chain_apply(const DollarTaskRec *__restrict__ recs, int64_t nt,
                                                                   int64_t k0, int64_t s0, int64_t pos0, int64_t K0,
                                                                   double u, const ChainElem *__restrict__ agg, int mode,
                                                                   DollarStatus *st) {
    __shared__ ChainElem buf[2 * DC_THREADS];
    const int64_t k = k0 + (int64_t)blockIdx.x * DC_THREADS + threadIdx.x;
    ChainElem e;
    e.f = dollar_xfer_identity(); e.last = -1;
    DollarTaskRec t;
    t.start_idx = -1;
    if (k < nt) { t = recs[k]; e = chain_elem(t, k); }
    ChainElem inc = chain_block_scan(e, buf);
    buf[threadIdx.x] = inc;
    __syncthreads();
    ChainElem ex = agg[blockIdx.x];
    if (threadIdx.x > 0) ex = chain_combine(ex, buf[threadIdx.x - 1]);
    if (k >= nt) return;
    DollarWalk w;
    w.fail_task = -1; w.done = 0; w.K_total = 0;
    w.s = s0 + ex.f.off[s0 & 3];
    if (ex.last >= 0) {
        const DollarTaskRec &pt = recs[ex.last];
        w.pos = pt.end_idx;
        w.K = pt.k_start + pt.count;
    } else { w.pos = pos0; w.K = K0; }
    const DollarWalk w_in = w;
    const int rc = dollar_walk_step(t, u, w, true);
    const bool event = (rc == 1 || rc == 2 || rc == 4);
    if (mode == 0) {
        if (event) atomicMin(&st->ev_min, (unsigned long long)k);
        return;
    }
    const unsigned long long evk = st->ev_min;
    if (evk == ~0ull) {
        if (k == nt - 1) {   // no event: range exhausted; state after the last task
            st->event = 0; st->k_ev = nt; st->s = w.s; st->pos = w.pos; st->K = w.K; st->use_c = 0;
        }
    } else if ((unsigned long long)k == evk) {
        st->event = rc; st->k_ev = k; st->s = w_in.s; st->pos = w_in.pos; st->K = w_in.K; st->K_total = w.K_total; st->use_c = 0;
    }
}

// task 0 (exact start) is consumed on its own: it decides how the chain is entered
__global__ void k_dollar_chain_task0(const DollarTaskRec *__restrict__ recs, const double *p, const double *v,
                                     DollarStatus *st) {
    const DollarTaskRec t0 = recs[0];
    st->ev_min = ~0ull;
    if (t0.end_idx == -2) { st->event = 1; st->k_ev = 0; st->K_total = t0.count; }
    else if (t0.bad & 1) {
        st->event = 2; st->k_ev = 0; st->s = 0; st->pos = 0; st->K = 0; st->c = __dmul_rn(p[0], v[0]); st->use_c = 1;
    } else {
        st->event = 7;   // proceed with the chain from task 1
        st->k_ev = 1; st->s = t0.end_units[0]; st->pos = t0.end_idx; st->K = t0.count; st->use_c = 0;
    }
}

__global__ void k_dollar_reset_ev(DollarStatus *st) { st->ev_min = ~0ull; }

// exact serial repair from (pos, c, K); resync allowed at tasks >= kmin.  One warp: the lanes fetch 32 ticks at a time
// (coalesced) and every lane replays the same recurrence on the shuffled products, so the chain never waits on a
// dependent global load.  Semantics are those of dollar_serial() in dollar_core.h.
__global__ void __launch_bounds__(32) k_dollar_serial(const double *__restrict__ p, const double *__restrict__ v, int64_t n,
                                                      double T, double u, double sub_lim, int64_t CH, int64_t nt,
                                                      const DollarTaskRec *__restrict__ recs, int64_t kmin, int64_t pos,
                                                      double c, int64_t K, int c_from_first, int64_t *out, int64_t cap,
                                                      DollarStatus *st) {
    const int lane = threadIdx.x;
    if (c_from_first) c = __dmul_rn(p[0], v[0]);   // logic.py:142 cum_dollar = prices[0] * volumes[0]
    int64_t cnt = 0, pos_out = -2;
    bool overflow = false;
    auto tile = [&](int64_t base) {                    // this lane's product of the 32-tick tile starting at `base`
        const int64_t i = base + lane;
        return i < n ? __dmul_rn(__ldg(p + i), __ldg(v + i)) : 0.0;
    };
    double d_next = tile(pos + 1);
    for (int64_t base = pos + 1; base < n && pos_out == -2; base += 32) {
        const double d = d_next;
        d_next = tile(base + 32);                      // in flight while the chain walks the current tile (the loads were half of the loop's time)
        const int m = (n - base) < 32 ? (int)(n - base) : 32;
        for (int q = 0; q < m; q++) {
            c = __dadd_rn(c, __shfl_sync(0xffffffffu, d, q));
            if (c >= T) {
                c = __dadd_rn(c, -T);
                cnt++;
                const int64_t at = base + q;
                if (K + cnt < cap) { if (lane == 0) out[K + cnt] = at; } else overflow = true;
                const int64_t kk = at / CH;
                if (kk >= kmin && kk < nt) {
                    const DollarTaskRec &t = recs[kk];
                    if (t.start_idx == at && t.k_start == K + cnt) {
                        const double eu = c / u;
                        if ((double)(int64_t)eu == eu && c < sub_lim) { pos_out = at; break; }
                    }
                }
            }
        }
    }
    if (lane != 0) return;
    if (overflow) { st->event = -1; return; }
    if (pos_out == -2) { st->event = 1; st->K_total = K + cnt; return; }
    st->event = 5;   // resynchronised
    st->k_ev = pos_out / CH;
    st->s = (int64_t)(c / u);
    st->pos = pos_out;
    st->K = K + cnt;
    st->use_c = 0;
}

__global__ void k_set_i64(int64_t *p, int64_t v) { *p = v; }

// Every task (one lane) replays its chunk plus about one bar, and all tasks take the same time, so the task kernel's
// duration is (number of waves) x (chunk + overlap).  With the default 4096-tick chunks 1e9 ticks are 2.6 waves of the
// resident lanes, i.e. three rounds with the last one mostly idle.  For inputs of at least one full wave the chunk is
// sized so that the tasks fill an integer number of waves (multiple of 256 ticks, 4096..16384).
static int64_t dollar_pick_chunk(fmk_ctx *ctx, int64_t n) {
    int bps = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_dollar_tasks, DT_WARPS * 32, 0) != cudaSuccess || bps <= 0)
        return DOLLAR_CH;
    const int64_t resident = (int64_t)bps * (ctx->sm_count - ctx->reserved_sms) * DT_WARPS * 32;
    if (n < resident * DOLLAR_CH) return DOLLAR_CH;
    int64_t best = DOLLAR_CH;
    double best_cost = 1e300;
    for (int64_t W = 1; W <= 64; W++) {
        int64_t ch = cdiv(cdiv(n, W * resident), DOLLAR_A_THREADS) * DOLLAR_A_THREADS;
        if (ch > 16384) continue;
        if (ch < DOLLAR_CH) break;
        const int64_t waves = cdiv(cdiv(n, ch), resident);
        const double cost = (double)waves * (double)(ch + 768);     // 768: a typical bar of overlap
        if (cost < best_cost) { best_cost = cost; best = ch; }
    }
    return best;
}

int fmk_dollar_index_impl(fmk_ctx *ctx, const fmk_trades *t, double T, fmk_index **out_ix) {
    FMK_ENTER(ctx);
    *out_ix = nullptr;
    const int64_t n = t->n;
    if (n <= 0) return fmk_fail(ctx, FMK_ERR_ARG, "empty trades");
    if (T != T) return fmk_fail(ctx, FMK_ERR_ARG, "threshold is NaN");
    static const char *ch_env = getenv("FMK_DOLLAR_CH");            // profiling override (multiple of 256)
    const int64_t CH = ch_env ? (int64_t)atoll(ch_env) : dollar_pick_chunk(ctx, n);
    const int64_t nt = cdiv(n, CH);
    ctx->stats[0] = nt; ctx->stats[1] = 0; ctx->stats[2] = 0;

    Scratch<dd_t> sums(ctx);
    Scratch<int64_t> K_in(ctx);
    Scratch<double> carry(ctx);
    Scratch<DollarTaskRec> recs(ctx);
    Scratch<DollarStatus> st(ctx);
    FMK_TRY(sums.alloc(nt + 1));
    FMK_TRY(K_in.alloc(nt));
    FMK_TRY(carry.alloc(nt));
    FMK_TRY(recs.alloc(nt));
    FMK_TRY(st.alloc(1));

    DollarParams P;
    const bool fast = dollar_params_init(&P, T, n, CH, 0);
    int64_t cap = n + 1;
    if (fast) {
        FMK_LAUNCH(ctx, k_dollar_chunk_sums, (unsigned)nt, DOLLAR_A_THREADS, 0, t->price, t->amount, n, (int)CH, sums.p);
        Scratch<DD> dtot(ctx);
        FMK_TRY(dtot.alloc(1));
        FMK_TRY((device_inclusive_scan<DD>(ctx, SumIn{sums.p}, GuessOut{K_in.p, carry.p, nt, T}, nt, dtot.p)));
        dd_t total;
        FMK_CUDA(ctx, cudaMemcpyAsync(&total, dtot.p, sizeof(dd_t), cudaMemcpyDeviceToHost, ctx->stream));
        FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        // every emission removes T from a non-negative running sum, so count <= total/T (+ slack for rounding)
        double bound = total.hi / T;
        if (bound >= 0 && bound < (double)n) cap = (int64_t)bound + 4;
        if (cap > n + 1) cap = n + 1;
    }
    P.cap = cap;
    int64_t *idx = nullptr;
    FMK_TRY(fmk_dalloc(ctx, &idx, cap));
    struct Guard {  // frees idx on early error returns
        fmk_ctx *c; int64_t **p;
        ~Guard() { if (*p) fmk_dfree(c, *p); }
    } guard{ctx, &idx};
    FMK_LAUNCH(ctx, k_set_i64, 1, 1, 0, idx, (int64_t)0);
    FMK_CUDA(ctx, cudaMemsetAsync(st.p, 0, sizeof(DollarStatus), ctx->stream));

    DollarStatus hs;
    memset(&hs, 0, sizeof(hs));
    int64_t K_total = -1;
    if (!fast) {
        // degenerate threshold (<= 0, inf, denormal): the recurrence is replayed serially on the device
        ctx->stats[1]++;
        FMK_LAUNCH(ctx, k_dollar_serial, 1, 32, 0, t->price, t->amount, n, T, 1.0, 0.0, CH, (int64_t)0,
                   (const DollarTaskRec *)recs.p, (int64_t)0, (int64_t)0, 0.0, (int64_t)0, 1, idx, cap, st.p);
        FMK_CUDA(ctx, cudaMemcpyAsync(&hs, st.p, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
        FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (hs.event != 1) return fmk_fail(ctx, FMK_ERR_INTERNAL, "dollar serial replay failed");
        K_total = hs.K_total;
    }
    if (fast) {
    static const int l2pf = getenv("FMK_DOLLAR_L2PF") ? atoi(getenv("FMK_DOLLAR_L2PF")) : 256;
    if (l2pf == 0)
        FMK_LAUNCH(ctx, k_dollar_tasks_pf0, (unsigned)cdiv(nt, DT_WARPS * 32), DT_WARPS * 32, 0, t->price, t->amount, P, nt,
                   (const int64_t *)K_in.p, (const double *)carry.p, idx, recs.p);
    else if (l2pf == 128)
        FMK_LAUNCH(ctx, k_dollar_tasks_pf128, (unsigned)cdiv(nt, DT_WARPS * 32), DT_WARPS * 32, 0, t->price, t->amount, P, nt,
                   (const int64_t *)K_in.p, (const double *)carry.p, idx, recs.p);
    else
        FMK_LAUNCH(ctx, k_dollar_tasks, (unsigned)cdiv(nt, DT_WARPS * 32), DT_WARPS * 32, 0, t->price, t->amount, P, nt,
                   (const int64_t *)K_in.p, (const double *)carry.p, idx, recs.p);
    Scratch<ChainElem> agg(ctx);
    FMK_TRY(agg.alloc(cdiv(nt, DC_THREADS) + 1));
    // task 0 runs from the exact initial state and decides how the chain is entered
    FMK_LAUNCH(ctx, k_dollar_chain_task0, 1, 1, 0, (const DollarTaskRec *)recs.p, (const double *)t->price,
               (const double *)t->amount, st.p);
    FMK_CUDA(ctx, cudaMemcpyAsync(&hs, st.p, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t k0 = 1, s0 = hs.s, pos0 = hs.pos, K0 = hs.K;
    bool need_chain = hs.event == 7;
    if (hs.event == 1) K_total = hs.K_total;
    while (K_total < 0) {
        if (need_chain) {
            ctx->stats[2]++;
            if (k0 >= nt) {
                hs.event = 0; hs.k_ev = nt; hs.s = s0; hs.pos = pos0; hs.K = K0; hs.use_c = 0;
            } else {
                const int64_t nblk = cdiv(nt - k0, DC_THREADS);
                FMK_LAUNCH(ctx, k_dollar_reset_ev, 1, 1, 0, st.p);
                FMK_LAUNCH(ctx, k_dollar_chain_reduce, (unsigned)nblk, DC_THREADS, 0, (const DollarTaskRec *)recs.p, nt, k0, agg.p);
                FMK_LAUNCH(ctx, k_dollar_chain_top, 1, DC_THREADS, 0, agg.p, nblk);
                FMK_LAUNCH(ctx, k_dollar_chain_apply, (unsigned)nblk, DC_THREADS, 0, (const DollarTaskRec *)recs.p, nt, k0, s0, pos0,
                           K0, P.u, (const ChainElem *)agg.p, 0, st.p);
                FMK_LAUNCH(ctx, k_dollar_chain_apply, (unsigned)nblk, DC_THREADS, 0, (const DollarTaskRec *)recs.p, nt, k0, s0, pos0,
                           K0, P.u, (const ChainElem *)agg.p, 1, st.p);
                FMK_CUDA(ctx, cudaMemcpyAsync(&hs, st.p, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
                FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            }
            if (hs.event == 1) { K_total = hs.K_total; break; }
        }
        // failure (2/4) or range exhausted (0): exact serial repair from the last certified state
        const int64_t kmin = hs.event == 4 ? hs.k_ev : (hs.event == 0 ? nt : hs.k_ev + 1);
        const double c = hs.use_c ? hs.c : (double)hs.s * P.u;
        ctx->stats[1]++;
        FMK_LAUNCH(ctx, k_dollar_serial, 1, 32, 0, t->price, t->amount, n, T, P.u, P.sub_lim, CH, nt,
                   (const DollarTaskRec *)recs.p, kmin, hs.pos, c, hs.K, 0, idx, cap, st.p);
        FMK_CUDA(ctx, cudaMemcpyAsync(&hs, st.p, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
        FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (hs.event == -1) return fmk_fail(ctx, FMK_ERR_INTERNAL, "dollar index overflow");
        if (hs.event == 1) { K_total = hs.K_total; break; }
        k0 = hs.k_ev; s0 = hs.s; pos0 = hs.pos; K0 = hs.K;
        need_chain = true;
        if (k0 <= 0) return fmk_fail(ctx, FMK_ERR_INTERNAL, "dollar chain resync at task 0");
    }
    }
    fmk_index *ix = new (std::nothrow) fmk_index();
    if (!ix) return FMK_ERR_ALLOC;
    memset(ix, 0, sizeof(*ix));
    ix->m = K_total + 1;
    ix->n_ticks = n;
    ix->sorted = 1;
    ix->close_idx = idx;
    idx = nullptr;  // ownership moved
    int rc = fmk_gather_close_ts(ctx, t, ix);
    if (rc) { fmk_index_free(ctx, ix); return rc; }
    *out_ix = ix;
    return FMK_OK;
}
