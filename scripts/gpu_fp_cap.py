"""A/B of the footprint kernel's shared-memory capacity (FMK_FP_CAP) on configs 3 and 5.  python scripts/gpu_fp_cap.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from finmlkit_b200 import core
ctx = core.default_context(0)
tr = core.DeviceTrades.synth(1_000_000_000, seed=42, ctx=ctx)
vix = core.volume_bar_index(tr, 50.0)
dix = core.dollar_bar_index(tr, 1e6)
F = core.F_OHLCV | core.F_FOOTPRINT
for cap in (128, 192, 256, 320, 448, 576, 832):
    os.environ["FMK_FP_CAP"] = str(cap)
    out = []
    for ix in (vix, dix):
        for rep in range(2):
            ctx.sync(); ctx.prof_enable(True); fr = core.bar_features_device(tr, ix, F, price_tick_size=0.1); ctx.sync(); ctx.prof_enable(False)
            pr = ctx.prof_report()
        out.append(round(pr["k_bar_footprint"][1], 2))
        del fr
    print("cap", cap, "k_bar_footprint ms: volume bars", out[0], " dollar bars", out[1], flush=True)
