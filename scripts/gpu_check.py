"""Exploration script (run under gpurun): device-generated stream -> indexers/reductions vs the CPU oracle, with timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from finmlkit_b200 import core
import oracle

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
T = float(sys.argv[2]) if len(sys.argv) > 2 else 1e6
ctx = core.default_context(0)
t0 = time.time(); tr = core.DeviceTrades.synth(N, seed=42, ctx=ctx); print(f"synth {N} ticks: {time.time()-t0:.3f}s")
ts, px, qty, side = tr.download()
print("ts[:3]", ts[:3], "px[:3]", px[:3], "qty[:3]", qty[:3], "side[:8]", side[:8], "dup ts frac", np.mean(np.diff(ts) == 0))
print("mean notional", float(np.mean(px[:1000000] * qty[:1000000])))

def timed(name, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        ctx.flush_l2(); ctx.sync()
        ctx.timer_start(); r = fn(); ms = ctx.timer_stop(); best = min(best, ms)
    print(f"{name:34s} {best:9.3f} ms  {N/best/1e6:10.2f} Gticks/s... ({N/(best*1e-3)/1e9:.3f}e9 ticks/s)")
    return r

ix = timed("dollar_bar_index", lambda: core.dollar_bar_index(tr, T))
print("  stats", ctx.index_stats(), "bars", ix.m - 1)
t0 = time.time(); ref = oracle.dollar_bar_indexer(px, qty, T); dt = time.time() - t0
print(f"  oracle dollar: {dt:.3f}s  {N/dt/1e6:.1f} Mticks/s")
cts, cidx = ix.download()
print("  dollar idx bit-exact:", np.array_equal(cidx, ref), len(cidx), len(ref))
assert np.array_equal(cts, ts[ref])
timed("ohlcv (device-resident, +median)", lambda: ctx.check(ctx._L.fmk_bar_ohlcv_device(ctx.h, tr.h, ix.h, 1)))
timed("ohlcv (device-resident, no median)", lambda: ctx.check(ctx._L.fmk_bar_ohlcv_device(ctx.h, tr.h, ix.h, 0)))
got = core.bar_ohlcv(tr, ix)
t0 = time.time(); o = oracle.comp_bar_ohlcv(px, qty, ref); dt = time.time() - t0
print(f"  oracle ohlcv: {dt:.3f}s ({oracle.num_threads()} threads)")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from helpers import check_ohlcv, check_directional, check_trade_size, check_footprint_csr
check_ohlcv(got, o, "ohlcv"); print("  ohlcv parity OK")
tix = timed("time_bar_index 60s", lambda: core.time_bar_index(tr, 60.0))
clock, tidx = tix.download(); rc, ri = oracle.time_bar_indexer(ts, 60.0)
print("  time idx exact:", np.array_equal(tidx, ri) and np.array_equal(clock, rc), tix.m - 1)
timed("ohlcv time bars (+median)", lambda: ctx.check(ctx._L.fmk_bar_ohlcv_device(ctx.h, tr.h, tix.h, 1)))
timed("ohlcv time bars (no median)", lambda: ctx.check(ctx._L.fmk_bar_ohlcv_device(ctx.h, tr.h, tix.h, 0)))
check_ohlcv(core.bar_ohlcv(tr, tix), oracle.comp_bar_ohlcv(px, qty, ri), "time ohlcv"); print("  time ohlcv parity OK")
d = timed("directional (host out)", lambda: core.bar_directional(tr, ix), reps=2)
check_directional(d, oracle.comp_bar_directional_features(px, qty, ref, side), "dir"); print("  directional parity OK")
th = np.full(ix.m - 1, 0.05)
s = timed("trade size (host out)", lambda: core.bar_trade_size(tr, ix, th, 5.0), reps=2)
check_trade_size(s, oracle.comp_bar_trade_size_features(qty, th, ref, 5.0), "ts"); print("  trade-size parity OK")
f = timed("footprints (host out)", lambda: core.bar_footprints_csr(tr, ix, 0.1, o[2], o[1], 3.0), reps=2)
t0 = time.time(); fr = oracle.comp_bar_footprints_csr(px, qty, ref, side, 0.1, o[2], o[1], 3.0); print(f"  oracle footprints {time.time()-t0:.3f}s")
check_footprint_csr(f, fr[0], list(fr[1:]), float(np.max(np.abs(fr[1]))), "fp"); print("  footprint parity OK")
M = min(N, 20_000_000)
r = timed("lagged returns 1h (host io, 20M)", lambda: core.lagged_returns(ts[:M], px[:M], 3600.0, True, ctx=ctx), reps=2)
t0 = time.time(); rr = oracle.comp_lagged_returns(ts[:M], px[:M], 3600.0, True); print(f"  oracle lagret {time.time()-t0:.3f}s")
from helpers import assert_f64, assert_exact
assert_f64(r, rr, "lagret", atol=1e-15); print("  lagret parity OK")
sg = timed("ewmst 1h (host io, 20M)", lambda: core.ewmst_series(ts[:M], rr, 3600.0, ctx=ctx), reps=2)
t0 = time.time(); sr = oracle.ewmst(ts[:M], rr, 3600.0); print(f"  oracle ewmst {time.time()-t0:.3f}s")
assert_f64(sg, sr, "ewmst", atol=1e-18); print("  ewmst parity OK")
ev = ref[1:]; ev = ev[ev < M]; ev = ev[np.isfinite(sr[ev])]; ev = ev[ts[ev] + 3600 * 10**9 <= ts[M-1]]
tg = sr[ev] * 2
tr2 = core.DeviceTrades.upload(ts[:M], px[:M], qty[:M], side[:M], ctx=ctx)
lab = timed(f"triple barrier {len(ev)} events", lambda: core.triple_barrier_dev(tr2, ev, tg, (2.0, 2.0), 3600.0, 1.0, None, 0.0), reps=2)
t0 = time.time(); lr = oracle.triple_barrier(ts[:M], px[:M], ev, tg, (2.0, 2.0), 3600.0, 1.0, None, 0.0); print(f"  oracle tbm {time.time()-t0:.3f}s")
print("  tbm labels equal:", np.array_equal(lab[0], lr[0]), "touch equal:", np.array_equal(lab[1], lr[1]), "mean path", float(np.mean(lr[1]-ev)))
assert_f64(lab[2], lr[2], "tbm rets", atol=1e-15); assert_f64(lab[3], lr[3], "tbm ratios", atol=1e-15); print("  tbm parity OK")
print("launches", ctx.launch_count())
vix = timed("volume_bar_index T=5", lambda: core.volume_bar_index(tr, 5.0))
print("  stats", ctx.index_stats(), "bars", vix.m - 1)
t0 = time.time(); vref = oracle.volume_bar_indexer(qty, 5.0); dt = time.time() - t0
print(f"  oracle volume: {dt:.3f}s  {N/dt/1e6:.1f} Mticks/s; exact:", np.array_equal(vix.download()[1], vref))
print("prof", ctx.prof_report())
