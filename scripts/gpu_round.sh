#!/bin/bash
# One GPU-box call: parity tests, bench (both arms), ncu launch list + full capture of the dominant kernels.
# usage: gpurun --timeout 1800 -- 'bash scripts/gpu_round.sh TAG'
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
cat $O/${TAG}_bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref exit $?"
cat $O/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e > $O/${TAG}_ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_dollar_tasks|k_bar_ohlcv_median|k_dollar_chunk_sums|k_bar_order_stats' -c 4 \
    -f -o $O/${TAG}_full python bench.py --steps 1 --warmup 0 --no-e2e > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $O
