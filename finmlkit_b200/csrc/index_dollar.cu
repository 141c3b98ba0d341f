// index_dollar.cu -- device pipeline of the bit-exact dollar-bar indexer (see dollar_core.h for the algorithm and
// its exactness argument; reference: finmlkit/bar/logic.py:118-149).
//
//   A1  k_dollar_chunk_sums : double-double sum of fl(p*v) per chunk of CH ticks        (reads 16 B/tick, coalesced)
//   A2  k_dollar_prefix     : exclusive double-double scan over chunks -> (K_in, carry) guess per chunk
//   B   k_dollar_tasks      : one thread per chunk: locate first boundary, replay 4 exact chains (reads 16 B/tick)
//   C   k_dollar_chain      : single-block parallel composition of the per-task integer transfer functions,
//                             certification of every task against the true carried state
//   S   k_dollar_serial     : exact serial replay from the last certified state (repair path only)
#include <new>
#include "common.cuh"
#include "dollar_core.h"

constexpr int DOLLAR_CH = 2048;
constexpr int DOLLAR_A_THREADS = 256;

struct LdG {
    const double *a;
    __device__ __forceinline__ double operator()(int64_t i) const { return __ldg(a + i); }
};

__global__ void __launch_bounds__(DOLLAR_A_THREADS) k_dollar_chunk_sums(const double *__restrict__ p,
                                                                        const double *__restrict__ v, int64_t n,
                                                                        dd_t *__restrict__ sums) {
    __shared__ dd_t sm[DOLLAR_A_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * DOLLAR_CH;
    dd_t s = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < DOLLAR_CH / DOLLAR_A_THREADS; k++) {
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

// single block: exclusive dd scan over chunk sums; emits the (K_in, carry) guess per chunk and the grand total
constexpr int DOLLAR_P_THREADS = 1024;
__global__ void __launch_bounds__(DOLLAR_P_THREADS) k_dollar_prefix(const dd_t *__restrict__ sums, int64_t nt, double T,
                                                                    int64_t *__restrict__ K_in,
                                                                    double *__restrict__ carry, dd_t *total) {
    __shared__ dd_t seg[DOLLAR_P_THREADS];
    const int64_t per = (nt + DOLLAR_P_THREADS - 1) / DOLLAR_P_THREADS;
    const int64_t a = (int64_t)threadIdx.x * per;
    int64_t b = a + per;
    if (b > nt) b = nt;
    dd_t s = {0.0, 0.0};
    for (int64_t k = a; k < b; k++) s = dd_add(s, sums[k]);
    seg[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        dd_t run = {0.0, 0.0};
        for (int t = 0; t < DOLLAR_P_THREADS; t++) {
            dd_t x = seg[t];
            seg[t] = run;
            run = dd_add(run, x);
        }
        *total = run;
    }
    __syncthreads();
    dd_t run = seg[threadIdx.x];
    for (int64_t k = a; k < b; k++) {
        int64_t K;
        double c;
        dollar_guess(run, T, &K, &c);
        K_in[k] = K;
        carry[k] = c;
        run = dd_add(run, sums[k]);
    }
}

__global__ void __launch_bounds__(128) k_dollar_tasks(const double *__restrict__ p, const double *__restrict__ v,
                                                      DollarParams P, int64_t nt, const int64_t *__restrict__ K_in,
                                                      const double *__restrict__ carry, int64_t *__restrict__ out,
                                                      DollarTaskRec *__restrict__ recs) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nt) return;
    DollarTaskRec rec;
    dollar_task(LdG{p}, LdG{v}, P, k, k > 0 ? carry[k] : 0.0, k > 0 ? K_in[k] : 0, out, &rec);
    recs[k] = rec;
}

// status block shared between chain / serial kernels and the host
struct DollarStatus {
    int64_t event;      // 0 = range exhausted, 1 = done, 2 = failure (resync after k_ev), 4 = failure (resync at k_ev)
    int64_t k_ev;
    int64_t s, pos, K;  // last certified boundary state (s in units of u)
    int64_t K_total;
    double c;           // serial start state as a double (used when the state is not a multiple of u)
    int64_t use_c;
    int64_t n_certified;
};

constexpr int DOLLAR_C_THREADS = 1024;

// Chain over tasks [k0, nt) entering with state (s0,pos0,K0).  k0 == 0 additionally consumes task 0 (exact start).
__global__ void __launch_bounds__(DOLLAR_C_THREADS) k_dollar_chain(const DollarTaskRec *__restrict__ recs, int64_t nt,
                                                                   int64_t k0, int64_t s0, int64_t pos0, int64_t K0,
                                                                   double u, const double *p, const double *v,
                                                                   DollarStatus *st) {
    __shared__ DollarXfer pre[DOLLAR_C_THREADS];
    __shared__ long long lastne[DOLLAR_C_THREADS];
    __shared__ unsigned long long ev_min;
    __shared__ int64_t sh_s0, sh_pos0, sh_K0, sh_k0;
    __shared__ int early;
    if (threadIdx.x == 0) {
        ev_min = ~0ull;
        early = 0;
        sh_s0 = s0; sh_pos0 = pos0; sh_K0 = K0; sh_k0 = k0;
        st->n_certified = 0;
        if (k0 == 0) {
            const DollarTaskRec t0 = recs[0];
            if (t0.end_idx == -2) {
                st->event = 1; st->k_ev = 0; st->K_total = t0.count; early = 1;
            } else if (t0.bad & 1) {
                st->event = 2; st->k_ev = 0; st->s = 0; st->pos = 0; st->K = 0;
                st->c = __dmul_rn(p[0], v[0]); st->use_c = 1; early = 1;
            } else {
                sh_s0 = t0.end_units[0]; sh_pos0 = t0.end_idx; sh_K0 = t0.count; sh_k0 = 1;
            }
        }
    }
    __syncthreads();
    if (early) return;
    const int64_t kb = sh_k0;
    if (kb >= nt) {
        if (threadIdx.x == 0) { st->event = 0; st->k_ev = nt; st->s = sh_s0; st->pos = sh_pos0; st->K = sh_K0; st->use_c = 0; }
        return;
    }
    const int64_t per = (nt - kb + DOLLAR_C_THREADS - 1) / DOLLAR_C_THREADS;
    const int64_t a = kb + (int64_t)threadIdx.x * per;
    int64_t b = a + per;
    if (b > nt) b = nt;
    // phase 1: segment composite under the assumption that every task is valid
    DollarXfer f = dollar_xfer_identity();
    long long last = -1;
    for (int64_t k = a; k < b; k++) {
        const DollarTaskRec &t = recs[k];
        if (t.start_idx < 0) continue;
        if (t.nch == DC_NCH && t.end_idx != -2) f = dollar_xfer_compose(f, dollar_task_xfer(t));
        last = k;
    }
    pre[threadIdx.x] = f;
    lastne[threadIdx.x] = last;
    __syncthreads();
    // phase 2: exclusive scan over segments (serial over 1024 small elements)
    if (threadIdx.x == 0) {
        DollarXfer run = dollar_xfer_identity();
        long long lr = -1;
        for (int t = 0; t < DOLLAR_C_THREADS; t++) {
            DollarXfer x = pre[t];
            long long l = lastne[t];
            pre[t] = run;
            lastne[t] = lr;
            run = dollar_xfer_compose(run, x);
            if (l >= 0) lr = l;
        }
    }
    __syncthreads();
    // phase 3: walk the segment with the true entering state
    DollarWalk w;
    w.fail_task = -1; w.done = 0; w.K_total = 0;
    w.s = sh_s0 + pre[threadIdx.x].off[sh_s0 & 3];
    if (lastne[threadIdx.x] >= 0) {
        const DollarTaskRec &t = recs[lastne[threadIdx.x]];
        w.pos = t.end_idx;
        w.K = t.k_start + t.count;
    } else {
        w.pos = sh_pos0; w.K = sh_K0;
    }
    int64_t kf = -1, ncert = 0;
    int rc = 0;
    if (a < b) rc = dollar_walk_range(recs, a, b, u, w, &kf, &ncert, true);
    if (rc != 0) atomicMin(&ev_min, (unsigned long long)kf);
    __syncthreads();
    const unsigned long long evk = ev_min;
    if (evk == ~0ull) {
        // no event anywhere: the thread whose segment reaches nt holds the final state
        atomicAdd((unsigned long long *)&st->n_certified, (unsigned long long)ncert);
        if (b == nt && a < b) {
            st->event = 0; st->k_ev = nt; st->s = w.s; st->pos = w.pos; st->K = w.K; st->use_c = 0;
        }
        return;
    }
    if (rc != 0 && (unsigned long long)kf == evk) {
        st->event = rc; st->k_ev = kf; st->s = w.s; st->pos = w.pos; st->K = w.K; st->K_total = w.K_total; st->use_c = 0;
    }
}

// single-thread exact repair from (pos, c, K); resync allowed at tasks >= kmin
__global__ void k_dollar_serial(const double *__restrict__ p, const double *__restrict__ v, int64_t n, double T,
                                double u, double sub_lim, int64_t CH, int64_t nt,
                                const DollarTaskRec *__restrict__ recs, int64_t kmin, int64_t pos, double c, int64_t K,
                                int c_from_first, int64_t *out, int64_t cap, DollarStatus *st) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (c_from_first) c = __dmul_rn(p[0], v[0]);   // logic.py:142 cum_dollar = prices[0] * volumes[0]
    int overflow = 0;
    double c_out;
    int64_t pos_out;
    auto stop = [&](int64_t i, int64_t Kafter, double cc) {
        const int64_t kk = i / CH;
        if (kk < kmin || kk >= nt) return false;
        const DollarTaskRec &t = recs[kk];
        if (t.start_idx != i || t.k_start != Kafter) return false;
        const double eu = cc / u;
        return (double)(int64_t)eu == eu && cc < sub_lim;
    };
    const int64_t cnt = dollar_serial(LdG{p}, LdG{v}, n, T, pos, c, K, out, cap, &overflow, stop, &c_out, &pos_out);
    if (overflow) { st->event = -1; return; }
    if (pos_out == -2) { st->event = 1; st->K_total = K + cnt; return; }
    st->event = 5;   // resynchronised
    st->k_ev = pos_out / CH;
    st->s = (int64_t)(c_out / u);
    st->pos = pos_out;
    st->K = K + cnt;
    st->use_c = 0;
}

__global__ void k_set_i64(int64_t *p, int64_t v) { *p = v; }

int fmk_dollar_index_impl(fmk_ctx *ctx, const fmk_trades *t, double T, fmk_index **out_ix) {
    *out_ix = nullptr;
    const int64_t n = t->n;
    if (n <= 0) return fmk_fail(ctx, FMK_ERR_ARG, "empty trades");
    if (T != T) return fmk_fail(ctx, FMK_ERR_ARG, "threshold is NaN");
    const int64_t CH = DOLLAR_CH;
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
        FMK_LAUNCH(ctx, k_dollar_chunk_sums, (unsigned)nt, DOLLAR_A_THREADS, 0, t->price, t->amount, n, sums.p);
        FMK_LAUNCH(ctx, k_dollar_prefix, 1, DOLLAR_P_THREADS, 0, sums.p, nt, T, K_in.p, carry.p, sums.p + nt);
        dd_t total;
        FMK_CUDA(ctx, cudaMemcpyAsync(&total, sums.p + nt, sizeof(dd_t), cudaMemcpyDeviceToHost, ctx->stream));
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
        FMK_LAUNCH(ctx, k_dollar_serial, 1, 1, 0, t->price, t->amount, n, T, 1.0, 0.0, CH, (int64_t)0,
                   (const DollarTaskRec *)recs.p, (int64_t)0, (int64_t)0, 0.0, (int64_t)0, 1, idx, cap, st.p);
        FMK_CUDA(ctx, cudaMemcpyAsync(&hs, st.p, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
        FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (hs.event != 1) return fmk_fail(ctx, FMK_ERR_INTERNAL, "dollar serial replay failed");
        K_total = hs.K_total;
    }
    if (fast) {
    FMK_LAUNCH(ctx, k_dollar_tasks, (unsigned)cdiv(nt, 128), 128, 0, t->price, t->amount, P, nt,
               (const int64_t *)K_in.p, (const double *)carry.p, idx, recs.p);
    int64_t k0 = 0, s0 = 0, pos0 = 0, K0 = 0;
    for (int iter = 0;; iter++) {
        ctx->stats[2]++;
        FMK_LAUNCH(ctx, k_dollar_chain, 1, DOLLAR_C_THREADS, 0, (const DollarTaskRec *)recs.p, nt, k0, s0, pos0, K0, P.u,
                   (const double *)t->price, (const double *)t->amount, st.p);
        FMK_CUDA(ctx, cudaMemcpyAsync(&hs, st.p, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
        FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (hs.event == 1) { K_total = hs.K_total; break; }
        // failure (2/4) or range exhausted (0): exact serial repair from the last certified state
        const int64_t kmin = hs.event == 4 ? hs.k_ev : (hs.event == 0 ? nt : hs.k_ev + 1);
        const double c = hs.use_c ? hs.c : (double)hs.s * P.u;
        ctx->stats[1]++;
        FMK_LAUNCH(ctx, k_dollar_serial, 1, 1, 0, t->price, t->amount, n, T, P.u, P.sub_lim, CH, nt,
                   (const DollarTaskRec *)recs.p, kmin, hs.pos, c, hs.K, 0, idx, cap, st.p);
        FMK_CUDA(ctx, cudaMemcpyAsync(&hs, st.p, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
        FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (hs.event == -1) return fmk_fail(ctx, FMK_ERR_INTERNAL, "dollar index overflow");
        if (hs.event == 1) { K_total = hs.K_total; break; }
        k0 = hs.k_ev; s0 = hs.s; pos0 = hs.pos; K0 = hs.K;
        if (k0 <= 0) return fmk_fail(ctx, FMK_ERR_INTERNAL, "dollar chain resync at task 0");
    }
    }
    fmk_index *ix = new (std::nothrow) fmk_index();
    if (!ix) return FMK_ERR_ALLOC;
    memset(ix, 0, sizeof(*ix));
    ix->m = K_total + 1;
    ix->n_ticks = n;
    ix->close_idx = idx;
    idx = nullptr;  // ownership moved
    int rc = fmk_gather_close_ts(ctx, t, ix);
    if (rc) { fmk_index_free(ctx, ix); return rc; }
    *out_ix = ix;
    return FMK_OK;
}
