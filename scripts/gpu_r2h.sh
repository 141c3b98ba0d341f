#!/bin/bash
# 2-GPU pass: communicator tests with the lag-2 pipeline, bench N=2 (headline + config5)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "comm" 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/h_bench_n2.json 2> gpurun_out/h_bench_n2.err
echo "bench n2 rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/h_bench_n2.err | tail -n 4
python - <<'P'
import json
d=json.load(open('gpurun_out/h_bench_n2.json'))
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'gather bytes', d['config']['gather_bytes_per_step'])
c=d['config5']; print('config5 ms', c['ms_per_step'], 'tps', c['ticks_per_s'], c.get('gather'))
P
