#!/bin/bash
# round-2 evidence: launch list of the bench command + ncu --set full of the headline kernels (1e9) and of the config 3/4/5 kernels
O=gpurun_out
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench_1e9.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-sub > $O/r02_ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_dollar_tasks|k_bar_ohlcv_median|k_dollar_chunk_sums' -c 3 \
    -f -o $O/r02_full_headline python bench.py --steps 1 --warmup 0 --no-e2e --no-sub > $O/r02_ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 1200 ncu --set full --clock-control none -k 'regex:k_volume_next|k_volume_replay_seg|k_volume_exit1|k_bar_footprint|k_footprint_features|k_bar_directional|k_lagged_returns|k_ewm_apply|k_ewm_reduce|k_triple_barrier|k_w_tile_sums|k_cusum_prep|k_imb_backmap|k_bar_order_stats|k_bar_trade_size' -c 18 \
    -f -o $O/r02_full_cfg python scripts/gpu_cfg_steps.py 2e8 > $O/r02_ncu_cfg.log 2>&1; echo "ncu cfg exit $?"
# gpurun merges at most 64 MiB back: export the raw / source pages here and keep only the (small) headline report
ncu -i $O/r02_full_headline.ncu-rep --page raw --csv > $O/r02_full_headline_raw.csv 2>/dev/null
ncu -i $O/r02_full_headline.ncu-rep --page source --csv > $O/r02_full_headline_source.csv 2>/dev/null
ncu -i $O/r02_full_cfg.ncu-rep --page raw --csv > $O/r02_full_cfg_raw.csv 2>/dev/null
rm -f $O/r02_full_cfg.ncu-rep
gzip -f $O/r02_full_headline_source.csv
ls -la $O/
