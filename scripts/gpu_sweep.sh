#!/bin/bash
# usage: bash scripts/gpu_sweep.sh "VAR1=a VAR2=b" "VAR1=c" ...   -- one short bench per environment setting
for envs in "$@"; do
  out=$(env $envs timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e 2>/tmp/sweep.err)
  echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); k=d['roofline']['all_kernels_ms_per_step']
    print('$envs', '| step', round(d['ms_per_step'],3), {x:round(k[x],3) for x in list(k)[:3]}, '| time bars+median', round(d['time_bars_1min']['ohlcv+median']['ms_per_step'],3))
except Exception as e:
    print('$envs', 'FAILED', e); print(open('/tmp/sweep.err').read()[-500:])
"
done
