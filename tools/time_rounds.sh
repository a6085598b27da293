#!/bin/bash
# A/B of the mergeUntil kernels on the full cfg3 workload: rounds (K = 16, 8) vs one merge per iteration
cd "$(dirname "$0")/.."
REPS=4 ENVSWEEP="BPE_LOOP_ROUNDS=0 BPE_LOOP_ROUNDS=1,BPE_LOOP_K=16 BPE_LOOP_ROUNDS=1,BPE_LOOP_K=8 BPE_LOOP_ROUNDS=1,BPE_LOOP_K=16" BPE_TRACE=1 python tools/time_cfg3.py 2>&1 | grep -v "k_merge_loop:" 
