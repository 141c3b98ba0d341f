// CPU emulation of the device dollar-bar pipeline (same dollar_core.h the kernels use): pass A (chunk sums, guesses),
// pass B (tasks), pass C (chain + serial repairs).  Built and run by tests/test_dollar_core_cpu.py.
//   extern "C" int64_t dollar_emulate(p, v, n, T, CH, out, cap, stats[4])
#include <vector>
#include <stdio.h>
#include "../../finmlkit_b200/csrc/dollar_core.h"

struct Ld { const double *a; double operator()(int64_t i) const { return a[i]; } };

// per-task records only (pass A + B), for the explicit-vs-virtual chain cross-check:
// rec_out[k*10 + ...] = start_idx, end_idx, count, start_units, end_units[0..3], bad, nch
extern "C" int64_t dollar_task_records(const double *p, const double *v, int64_t n, double T, int64_t CH, int64_t *rec_out,
                                       double *margin_out) {
    DollarParams P;
    if (!dollar_params_init(&P, T, n, CH, n + 2)) return -1;
    Ld lp{p}, lv{v};
    std::vector<int64_t> out(n + 2);
    const int64_t nt = (n + CH - 1) / CH;
    dd_t run = {0.0, 0.0};
    for (int64_t k = 0; k < nt; k++) {
        int64_t K_in = 0; double carry = 0;
        if (k > 0) dollar_guess(run, T, &K_in, &carry);
        DollarTaskRec r;
        dollar_task(lp, lv, P, k, carry, K_in, out.data(), &r);
        int64_t *o = rec_out + k * 10;
        o[0] = r.start_idx; o[1] = r.end_idx; o[2] = r.count; o[3] = r.start_units;
        for (int q = 0; q < 4; q++) { o[4 + q] = r.end_units[q]; margin_out[k * 4 + q] = r.margin[q]; }
        o[8] = r.bad; o[9] = r.nch;
        int64_t hi = (k + 1) * CH < n ? (k + 1) * CH : n;
        dd_t s = {0.0, 0.0};
        for (int64_t i = k * CH; i < hi; i++) s = dd_add_d(s, p[i] * v[i]);
        run = dd_add(run, s);
    }
    return nt;
}

extern "C" int64_t dollar_emulate(const double *p, const double *v, int64_t n, double T, int64_t CH, int64_t *out,
                                  int64_t cap, int64_t *stats) {
    DollarParams P;
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    Ld lp{p}, lv{v};
    int overflow = 0;
    out[0] = 0;
    if (!dollar_params_init(&P, T, n, CH, cap)) {
        // degenerate threshold: fully serial
        double c; int64_t pos;
        int64_t cnt = dollar_serial(lp, lv, n, T, 0, p[0] * v[0], 0, out, cap, &overflow,
                                    [](int64_t, int64_t, double) { return false; }, &c, &pos);
        stats[1] = 1;
        return overflow ? -1 : cnt + 1;
    }
    const int64_t nt = (n + CH - 1) / CH;
    // pass A: chunk sums (double-double) and exclusive prefix
    std::vector<dd_t> pre(nt + 1);
    dd_t run = {0.0, 0.0};
    for (int64_t k = 0; k < nt; k++) {
        pre[k] = run;
        int64_t hi = (k + 1) * CH < n ? (k + 1) * CH : n;
        dd_t s = {0.0, 0.0};
        for (int64_t i = k * CH; i < hi; i++) s = dd_add_d(s, p[i] * v[i]);
        run = dd_add(run, s);
    }
    // pass B
    std::vector<DollarTaskRec> recs(nt);
    for (int64_t k = 0; k < nt; k++) {
        int64_t K_in = 0; double carry = 0;
        if (k > 0) dollar_guess(pre[k], T, &K_in, &carry);
        dollar_task(lp, lv, P, k, carry, K_in, out, &recs[k]);
    }
    stats[0] = nt;
    // pass C
    DollarWalk w;
    w.fail_task = -1; w.done = 0; w.K_total = 0;
    int64_t k = 1;
    bool need_serial = false;
    double sc = 0; int64_t spos = 0, sK = 0;   // serial start
    {   // task 0: exact start, always the true trajectory
        const DollarTaskRec &t0 = recs[0];
        if (t0.end_idx == -2) return overflow ? -1 : t0.count + 1;
        if (t0.bad & 1) { need_serial = true; sc = p[0] * v[0]; spos = 0; sK = 0; }
        else { w.s = t0.end_units[0]; w.pos = t0.end_idx; w.K = t0.count; }
    }
    for (;;) {
        if (need_serial) {
            stats[1]++;
            double c; int64_t pos;
            int64_t kmin = k;
            auto stop = [&](int64_t i, int64_t Kafter, double cc) {
                int64_t kk = i / CH;
                if (kk < kmin || kk >= nt) return false;
                const DollarTaskRec &t = recs[kk];
                if (t.start_idx != i || t.k_start != Kafter) return false;
                double eu = cc / P.u;
                return (double)(int64_t)eu == eu && cc < P.sub_lim;
            };
            int64_t cnt = dollar_serial(lp, lv, n, T, spos, sc, sK, out, cap, &overflow, stop, &c, &pos);
            if (pos == -2) return overflow ? -1 : sK + cnt + 1;
            w.s = (int64_t)(c / P.u); w.pos = pos; w.K = sK + cnt;
            k = pos / CH;
            need_serial = false;
        }
        int64_t kf = -1;
        int rc = dollar_walk_range(recs.data(), k, nt, P.u, w, &kf, &stats[2], false);
        if (rc == 1) return overflow ? -1 : w.K_total + 1;
        if (rc == 0) {
            // ran out of tasks without a certified tail task (trailing chunks located no boundary): finish serially
            need_serial = true; sc = (double)w.s * P.u; spos = w.pos; sK = w.K; k = nt;
            stats[3]++;
            continue;
        }
        k = (rc == 4) ? kf : kf + 1;   // where a resync may happen
        need_serial = true; sc = (double)w.s * P.u; spos = w.pos; sK = w.K;
    }
}
