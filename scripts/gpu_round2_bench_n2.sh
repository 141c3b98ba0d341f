#!/bin/bash
# N = 2, default bench line (identical stream per rank, peer-to-peer payload path): headline + sub-records without the e2e legs
mkdir -p gpurun_out
FMK_BENCH_WATCHDOG_S=120 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "rc=$?"
grep "bench\]" gpurun_out/r02_bench_n2.err | tail -n 4
python - <<'P'
import json
d=[json.loads(l) for l in open('gpurun_out/r02_bench_n2.json') if l.startswith('{')][-1]
print('ms', d['ms_per_step'], 'value', d['value'], d['config'].get('rank_ms_per_step'), d['config'].get('streams')[:40])
c=d.get('config5'); print('config5', c['ms_per_step'], c.get('rank_ms_per_step'), c.get('gather'))
P
