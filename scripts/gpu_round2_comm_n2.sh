#!/bin/bash
# 2-GPU validation of the peer-to-peer payload path: the two-rank byte-for-byte test (p2p, p2p + reset, send/recv fallback),
# then the bench at N = 2 with p2p and with the fallback (headline, then headline + config5).  Tight timeouts everywhere.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q -x -k "comm" 2>&1 | tail -15
run() { FMK_BENCH_WATCHDOG_S=90 timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 10 --warmup 3 "${@:3}"; }
run 150 29541 --no-sub --no-e2e > gpurun_out/m_n2_headline_p2p.json 2> gpurun_out/m_n2_headline_p2p.err; echo "headline p2p rc=$?"
FMK_COMM_P2P=0 run 150 29543 --no-sub --no-e2e > gpurun_out/m_n2_headline_nccl.json 2> gpurun_out/m_n2_headline_nccl.err; echo "headline nccl rc=$?"
run 240 29542 --no-e2e > gpurun_out/m_n2_sub.json 2> gpurun_out/m_n2_sub.err; echo "sub rc=$?"
grep "bench\]" gpurun_out/m_n2_sub.err | tail -n 4
python - <<'P'
import json
for f in ('gpurun_out/m_n2_headline_p2p.json','gpurun_out/m_n2_headline_nccl.json','gpurun_out/m_n2_sub.json'):
    try:
        d=json.load(open(f)); print(f, 'ms', d['ms_per_step'], {k: round(v,2) for k,v in list(d['roofline']['all_kernels_ms_per_step'].items())[:3]}, (d.get('config5') or {}).get('ms_per_step'), (d.get('config5') or {}).get('gather'))
    except Exception as e: print(f, 'no json', e); print(open(f.replace('.json','.err')).read()[-1500:])
P
