import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from finmlkit_b200 import core
import oracle
N = int(float(sys.argv[1])); T = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
ctx = core.default_context(0)
tr = core.DeviceTrades.synth(N, seed=42, ctx=ctx)
ts, px, qty, side = tr.download()
ctx.prof_enable(True)
t0 = time.time(); ix = core.volume_bar_index(tr, T); ctx.sync(); dt = time.time() - t0
print(f"N={N} volume index wall {dt*1e3:.2f} ms stats {ctx.index_stats()} bars {ix.m-1}", flush=True)
for k, v in sorted(ctx.prof_report().items(), key=lambda kv: -kv[1][1]): print(f"   {k:40s} {v[0]:3d} {v[1]:10.3f} ms")
ref = oracle.volume_bar_indexer(qty, T)
print("  exact:", np.array_equal(ix.download()[1], ref), flush=True)
