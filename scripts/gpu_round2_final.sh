#!/bin/bash
# end-of-round evidence on one B200: the whole GPU suite, the default bench line, the slowest rank stream alone, then the ncu
# launch list + --set full captures (scripts/gpu_round2_profile.sh)
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/r02_pytest_gpu.log; cat $O/r02_pytest_gpu.log
timeout 900 python bench.py > $O/r02_bench.json 2> $O/r02_bench.err; echo "bench rc=$?"; tail -3 $O/r02_bench.err
timeout 120 python scripts/gpu_seed_sweep.py 1e9 49 1 2>&1 | tail -1
bash scripts/gpu_round2_profile.sh 2>&1 | tail -3
python - <<'P'
import json
d=[json.loads(l) for l in open('gpurun_out/r02_bench.json') if l.startswith('{')][-1]
print('headline ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline'].get('step_level',{}).get('frac'))
for k in ('time_bars_1min','config3','config4','config5','config1','e2e_wrapper'):
    c=d.get(k)
    if c: print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in c.items() if a in ('ms_per_step','ticks_per_s','value','cold_ms','warm_ms')})
P
