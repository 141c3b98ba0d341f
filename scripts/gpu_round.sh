#!/bin/bash
# One GPU-box call: parity tests, bench (both arms), ncu launch list + full capture of the dominant kernels.
# usage: gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh TAG'
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref exit $?"
cut -c1-400 $O/${TAG}_bench_ref.json
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
cat $O/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e > $O/${TAG}_ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_dollar_tasks|k_bar_ohlcv_median|k_dollar_chunk_sums' -c 4 \
    -f -o $O/${TAG}_full python bench.py --steps 1 --warmup 0 --no-e2e > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 900 ncu --set full --clock-control none -k 'regex:k_volume_next|k_volume_replay_seg|k_bar_footprint|k_bar_directional|k_lagged_returns|k_ewm_apply|k_triple_barrier|k_w_tile_sums' -c 9 \
    -f -o $O/${TAG}_cfg python scripts/gpu_configs.py 2e8 > $O/${TAG}_ncu_cfg.log 2>&1; echo "ncu cfg exit $?"
timeout 600 python scripts/gpu_configs.py 1e9 > $O/${TAG}_configs.log 2>&1; echo "configs exit $?"
grep config $O/${TAG}_configs.log | cut -c1-300
ls -la $O | tail -20
