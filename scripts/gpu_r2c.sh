#!/bin/bash
mkdir -p gpurun_out
for G in 1 2 4; do
  FMK_CUSUM_WALK_MAX=$G FMK_CUSUM_DEBUG=1 python scripts/gpu_cfg4_phases.py 1e9 > gpurun_out/c_phases_G$G.log 2> gpurun_out/c_phases_G$G.err
  echo "== G=$G"; tail -2 gpurun_out/c_phases_G$G.log
  grep cusum gpurun_out/c_phases_G$G.err | tail -n 40 | awk '{s+=$(NF-5)} END {print "sum of last-step round ms:", s, NR}'
done
grep cusum gpurun_out/c_phases_G2.err | tail -n 40
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_round2.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | tail -5
