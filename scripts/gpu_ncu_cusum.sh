#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_cusum_tasks -c 3 -o gpurun_out/r02_cusum_tasks python scripts/gpu_cfg4_phases.py 4e8 > gpurun_out/ncu_cusum_tasks.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cusum_walk -c 2 -o gpurun_out/r02_cusum_walk python scripts/gpu_cfg4_phases.py 4e8 > gpurun_out/ncu_cusum_walk.log 2>&1
ls -la gpurun_out/*.ncu-rep
