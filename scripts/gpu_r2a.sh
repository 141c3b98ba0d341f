#!/bin/bash
# round-2 GPU pass A: new tests first (fast feedback), then the bench lines, then the full suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/smi.txt 2>&1
free -g >> gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_round2.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
echo "bench rc=$?" >> gpurun_out/bench_ours.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 1500 python -m pytest tests/test_gpu_ref_suite.py -q -m gpu 2>&1 | tail -60 > gpurun_out/pytest_refsuite.log
timeout 2400 python -m pytest tests -q -m gpu --deselect tests/test_gpu_ref_suite.py --deselect tests/test_gpu_round2.py 2>&1 | tail -60 > gpurun_out/pytest_all.log
tail -5 gpurun_out/pytest_round2.log gpurun_out/pytest_refsuite.log gpurun_out/pytest_all.log
tail -3 gpurun_out/bench_ours.err
