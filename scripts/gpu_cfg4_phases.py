"""config 4 phase timing on one GPU (host wall clock with a sync after every phase) + per-round CUSUM debug on stderr.
   FMK_CUSUM_DEBUG=1 python scripts/gpu_cfg4_phases.py [N=1e9]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from finmlkit_b200 import core
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
ctx = core.default_context(0)
tr = core.DeviceTrades.synth(N, seed=42, ctx=ctx)
last_ts = int(core.DeviceIndex.from_host(tr, np.array([N - 1, N - 1], np.int64)).download()[0][0])
def phase(name, fn, t):
    ctx.sync(); t0 = time.perf_counter(); out = fn(); ctx.sync(); t[name] = t.get(name, 0) + (time.perf_counter() - t0) * 1e3; return out
for rep in range(3):
    t = {}
    r = phase("lagged", lambda: core.lagged_returns_dev(tr, 3600.0, True), t)
    sig = phase("ewmst", lambda: core.ewmst_dev(tr, r, 3600.0), t)
    del r
    cix = phase("cusum", lambda: core.cusum_bar_index(tr, sig, 5e-4, 2.0), t)
    cts, cidx = phase("download", lambda: cix.download(), t)
    tg = phase("gather", lambda: sig.gather(cidx[1:]), t)
    ev = cidx[1:]
    keep = np.isfinite(tg) & (cts[1:] + 3600 * 10**9 <= last_ts)
    ev, tg = ev[keep], tg[keep]
    lab = phase("tbm", lambda: core.triple_barrier_dev(tr, ev, tg, (2.0, 2.0), 3600.0, 1.0, None, 0.0), t)
    phase("weights", lambda: core.sample_weights_dev(tr, ev, lab[1]), t)
    print(rep, {k: round(v, 2) for k, v in t.items()}, "total", round(sum(t.values()), 1), ctx.index_stats() , flush=True)
