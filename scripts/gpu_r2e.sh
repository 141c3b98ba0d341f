#!/bin/bash
# 2-GPU pass: communicator test (bytes received == bytes packed), then the bench at N=2 through torchrun
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/e_smi.txt
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "comm" 2>&1 | tail -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/e_bench_n2.json 2> gpurun_out/e_bench_n2.err
echo "bench n2 rc=$?"; tail -n 6 gpurun_out/e_bench_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/e_bench_n2.json'))
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'gather bytes', d['config']['gather_bytes_per_step'])
c=d['config5']; print('config5 ms', c['ms_per_step'], 'tps', c['ticks_per_s'], c.get('gather'))
print({k:round(v,2) for k,v in list(d['roofline']['all_kernels_ms_per_step'].items())[:6]})
P
