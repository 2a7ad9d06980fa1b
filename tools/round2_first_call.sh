#!/bin/bash
# First GPU call of the next round: what round 1 could not run any more (GPU budget spent).
#   1. the experimental warp-specialised tcgen05 exact pass A (tuning exact_tcgen05 = 2): correctness, then timing
#   2. launch list of the three exact variants at the C2 size
# Usage (from the repo root):  gpurun --timeout 200 -- 'bash tools/round2_first_call.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 python tools/exact_tcgen05_check.py 2 > gpurun_out/tc5ws_check.log 2>&1; echo "ws check rc=$?"; tail -2 gpurun_out/tc5ws_check.log | cut -c1-600
timeout 60 python tools/exact_perf.py 1000000 10000 2 > gpurun_out/exact_perf_r2.log 2>&1; echo "perf rc=$?"; tail -1 gpurun_out/exact_perf_r2.log
