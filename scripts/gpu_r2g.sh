#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 > gpurun_out/g_pytest_all.log
tail -n 6 gpurun_out/g_pytest_all.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench rc=$?"; tail -n 4 gpurun_out/g_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/g_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'roof', d['roofline']['frac'], d['roofline']['kernel'])
print('  ', {a:round(b,2) for a,b in list(d['roofline']['all_kernels_ms_per_step'].items())[:4]})
tb=d['time_bars_1min']; print('time bars', tb['ohlcv+median']['ms_per_step'], tb['ohlcv+median']['roofline']['frac'], tb['ohlcv']['ms_per_step'])
for k in ('config3','config4','config5'):
    c=d[k]; print(k, 'ms', round(c['ms_per_step'],2), 'tps', c['ticks_per_s'], 'e2e', c.get('e2e',{}).get('value'), 'cpu', c.get('cpu_baseline',{}).get('value'))
    print('   ', {a:round(b,2) for a,b in list((c.get('kernels_ms_per_step') or c.get('kernels_ms_per_step_rank0')).items())[:12]})
print('config1', d['config1']['ours'], d['config1']['cpu_baseline']['value'])
print('wrapper', d['e2e_wrapper']['ours'], d['e2e_wrapper'].get('cpu_baseline',{}).get('value'))
i=d['config4'].get('imbalance_bars'); print('imb', i['ms_per_step'], i['bars'])
P
