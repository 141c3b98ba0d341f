#!/bin/bash
# 8-GPU pass: headline-only A/B of the NCCL CTA budget, then the full bench line at N=8
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 10 --warmup 3 "${@:2}"; }
for c in 4 16 32; do
  FMK_NCCL_MAX_CTAS=$c run 2951$c --no-sub --no-e2e > gpurun_out/f_n8_headline_ctas$c.json 2> gpurun_out/f_n8_headline_ctas$c.err
  python -c "
import json; d=json.load(open('gpurun_out/f_n8_headline_ctas$c.json')); print('ctas $c', 'value', d['value'], 'ms', d['ms_per_step'], {k:round(v,2) for k,v in list(d['roofline']['all_kernels_ms_per_step'].items())[:3]})"
done
for c in 16 48; do
  FMK_NCCL_MAX_CTAS=$c run 2952$c --no-e2e > gpurun_out/f_n8_sub_ctas$c.json 2> gpurun_out/f_n8_sub_ctas$c.err
  python -c "
import json; d=json.load(open('gpurun_out/f_n8_sub_ctas$c.json')); c5=d['config5']; print('ctas $c', 'config5 ms', c5['ms_per_step'], 'tps', c5['ticks_per_s'], c5.get('gather'))"
done
run 29530 > gpurun_out/f_bench_n8.json 2> gpurun_out/f_bench_n8.err
echo "full n8 rc=$?"; tail -n 4 gpurun_out/f_bench_n8.err
python -c "
import json; d=json.load(open('gpurun_out/f_bench_n8.json')); print('N8 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
