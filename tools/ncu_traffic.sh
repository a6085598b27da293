#!/bin/bash
# DRAM traffic and duration of EVERY launch of the mergeUntil kernels over one full cfg3 run (1 GB corpus, 32 000 merges),
# plus K1: application replay, three metrics (a launch rewrites GBs of state, kernel replay cannot restore it).
cd "$(dirname "$0")/.."
export REPS=1
python tools/time_cfg3.py > gpurun_out/traffic_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --replay-mode application --clock-control none \
    -k regex:"k_merge_rounds|k_merge_loop|k_hist|k_scatter|k_alloc_lists|k_ingest_ids|k_count_bins|k_build_hot|k_rehash" --csv --log-file gpurun_out/traffic_cfg3.csv \
    python tools/time_cfg3.py > gpurun_out/traffic_ncu.log 2>&1
tail -3 gpurun_out/traffic_ncu.log
