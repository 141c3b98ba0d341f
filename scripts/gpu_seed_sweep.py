"""Headline step (dollar bars $1e6 + OHLCV incl. median, 1e9 ticks) on the symbol streams the N-GPU bench gives ranks 0..7
(seed 42 + rank), one after the other on ONE GPU: per-stream ms/step, index statistics and per-kernel ms.  Separates
data-dependent step time (exact serial repairs of near-ties) from communication effects in the weak-scaling numbers.
   python scripts/gpu_seed_sweep.py [N=1e9] [first_seed=42] [count=8]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finmlkit_b200 import core
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
S0 = int(sys.argv[2]) if len(sys.argv) > 2 else 42
CNT = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ctx = core.default_context(0)
for seed in range(S0, S0 + CNT):
    tr = core.DeviceTrades.synth(N, seed=seed, ctx=ctx)
    def step():
        ix = core.dollar_bar_index(tr, 1e6)
        fr = core.bar_features_device(tr, ix, core.F_OHLCV | core.F_MEDIAN)
        return ix.m - 1
    for _ in range(3):
        nb = step()
    ctx.sync()
    ctx.prof_enable(True)
    ctx.timer_start()
    for _ in range(5):
        step()
    ms = ctx.timer_stop() / 5
    ctx.prof_enable(False)
    prof = {k: round(v[1] / 5, 3) for k, v in sorted(ctx.prof_report().items(), key=lambda kv: -kv[1][1])[:6]}
    print(f"seed {seed}: {ms:.3f} ms/step, {nb} bars, {ctx.index_stats()}, {prof}", flush=True)
    del tr
    ctx.trim() if hasattr(ctx, "trim") else None
