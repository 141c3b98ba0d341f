#!/bin/bash
# 2-GPU re-validation with tight timeouts: headline only, then headline + config5
mkdir -p gpurun_out
run() { FMK_BENCH_WATCHDOG_S=90 timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 10 --warmup 3 "${@:3}"; }
run 150 29541 --no-sub --no-e2e > gpurun_out/i_n2_headline.json 2> gpurun_out/i_n2_headline.err; echo "headline rc=$?"
grep "bench\]" gpurun_out/i_n2_headline.err | tail -n 4
run 240 29542 --no-e2e > gpurun_out/i_n2_sub.json 2> gpurun_out/i_n2_sub.err; echo "sub rc=$?"
grep "bench\]" gpurun_out/i_n2_sub.err | tail -n 6
python - <<'P'
import json
for f in ('gpurun_out/i_n2_headline.json','gpurun_out/i_n2_sub.json'):
    try:
        d=json.load(open(f)); print(f, 'value', d['value'], 'ms', d['ms_per_step'], (d.get('config5') or {}).get('ms_per_step'), (d.get('config5') or {}).get('gather'))
    except Exception as e: print(f, 'no json', e)
P
