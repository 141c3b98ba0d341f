#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/d_pytest_all.log
tail -n 15 gpurun_out/d_pytest_all.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "bench rc=$?"; tail -n 5 gpurun_out/d_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/d_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
for k in ('config3','config4','config5'):
    c=d[k]; print(k, 'ms', round(c['ms_per_step'],2), 'tps', c['ticks_per_s'], 'e2e', c.get('e2e',{}).get('value'), 'cpu', c.get('cpu_baseline',{}).get('value'))
    print('   ', {a:round(b,2) for a,b in list((c.get('kernels_ms_per_step') or c.get('kernels_ms_per_step_rank0')).items())[:10]})
print('config1', d['config1']['ours'], d['config1']['cpu_baseline']['value'])
print('wrapper', d['e2e_wrapper']['ours'], d['e2e_wrapper'].get('cpu_baseline',{}).get('value'))
print('imb', d['config4'].get('imbalance_bars'))
P
