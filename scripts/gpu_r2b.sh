#!/bin/bash
# round-2 GPU pass B: CUSUM walkers -- golden tiny-chunk tests in both kernels, 1e7 vs oracle, full-size config 4, bench config 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_vectors.py tests/test_gpu_kits.py -q -m gpu 2>&1 | tail -15 > gpurun_out/b_pytest_golden.log
FMK_CUSUM_WALK_BELOW=1000000000 timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -k "cusum" 2>&1 | tail -15 > gpurun_out/b_pytest_cusum_walk_always.log
FMK_CUSUM_WALK_BELOW=0 timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -k "cusum" 2>&1 | tail -15 > gpurun_out/b_pytest_cusum_walk_never.log
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_ref_suite.py tests/test_gpu_round2.py tests/test_gpu_edge_cases.py -q -m gpu 2>&1 | tail -25 > gpurun_out/b_pytest_full.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-config5 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
for f in gpurun_out/b_pytest_*.log; do echo "== $f"; tail -n 4 $f; done
python - <<'P'
import json
d=json.load(open('gpurun_out/b_bench.json'))
c=d['config4']; print('config4 ms', c['ms_per_step'], c['cusum_stats']); print({k:round(v,2) for k,v in list(c['kernels_ms_per_step'].items())[:12]})
P
