#!/bin/bash
# next-bar L2 prefetch in k_bar_ohlcv_median (A/B), .L2::256B copies in k_dollar_tasks / k_cusum_tasks: parity, then times
O=gpurun_out; mkdir -p $O
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], {k: round(v,3) for k,v in list(d["roofline"]["all_kernels_ms_per_step"].items())[:3]}, d.get("time_bars_1min",{}).get("ms_per_step"))'
timeout 900 python -m pytest tests -q -m gpu -x -k "dollar or golden or headline or ohlcv or cusum or fullsize or time_bar" 2>&1 | tail -2
for pf in 1 0; do
  echo "MEDIAN_PREFETCH=$pf"
  FMK_MEDIAN_PREFETCH=$pf timeout 300 python bench.py --steps 10 --warmup 3 --no-sub --no-e2e 2>/dev/null | python -c "$show"
done
timeout 300 python scripts/gpu_cfg4_phases.py 1e9 2>/dev/null | tail -2
