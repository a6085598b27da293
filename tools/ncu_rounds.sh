#!/bin/bash
# One k_merge_rounds launch of the full cfg3 run (1 GB corpus) under ncu, application replay (a launch rewrites GBs of
# state: kernel replay would have to save and restore all of it).  $1 = launches of the kernel to skip (default 9: a launch
# in the latency-bound regime of small merges), $2 = output name.
cd "$(dirname "$0")/.."
SKIP=${1:-9}; OUT=${2:-prof_rounds}
export REPS=1
python tools/time_cfg3.py > gpurun_out/${OUT}_plain.log 2>&1 &&
ncu --set full --replay-mode application --clock-control none --import-source on -k regex:k_merge_rounds -s $SKIP -c 1 -f -o gpurun_out/$OUT \
    python tools/time_cfg3.py > gpurun_out/${OUT}_ncu.log 2>&1
tail -5 gpurun_out/${OUT}_ncu.log
