"""One device-resident step of BASELINE configs 3, 4 (+ imbalance bars) and 5 at N ticks, for ncu captures.
   python scripts/gpu_cfg_steps.py [N=2e8] [reps=1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from finmlkit_b200 import core
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = core.default_context(0)
tr = core.DeviceTrades.synth(N, seed=42, ctx=ctx)
last_ts = int(core.DeviceIndex.from_host(tr, np.array([N - 1, N - 1], np.int64)).download()[0][0])
for rep in range(REPS):
    vix = core.volume_bar_index(tr, 50.0)
    fr = core.bar_features_device(tr, vix, core.F_OHLCV | core.F_MEDIAN | core.F_DIRECTIONAL | core.F_FOOTPRINT, price_tick_size=0.1)
    print("config3", vix.m - 1, fr.n_levels, flush=True)
    del fr, vix
    r = core.lagged_returns_dev(tr, 3600.0, True)
    sig = core.ewmst_dev(tr, r, 3600.0)
    del r
    cix = core.cusum_bar_index(tr, sig, 5e-4, 2.0)
    cts, cidx = cix.download()
    ev, tg = cidx[1:], sig.gather(cidx[1:])
    keep = np.isfinite(tg) & (cts[1:] + 3600 * 10**9 <= last_ts)
    ev, tg = ev[keep], tg[keep]
    lab = core.triple_barrier_dev(tr, ev, tg, (2.0, 2.0), 3600.0, 1.0, None, 0.0)
    core.sample_weights_dev(tr, ev, lab[1])
    print("config4", cix.m - 1, len(ev), ctx.index_stats(), flush=True)
    iix = core.imbalance_bar_index(tr, 200.0)
    print("imbalance", iix.m - 1, flush=True)
    dix = core.dollar_bar_index(tr, 1e6)
    fr = core.bar_features_device(tr, dix, core.F_ALL, price_tick_size=0.1)
    print("config5", dix.m - 1, fr.n_levels, flush=True)
    del fr, dix, iix, cix, sig
ctx.sync()
