#!/bin/bash
# round-end check on one B200: GPU tests, smoke, default bench, launch list of the exact top-k kernels
cd "$(dirname "$0")/.."
timeout 110 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log
timeout 75 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_final.json
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"exact|merge_keys" --csv --log-file gpurun_out/exact_launches2.csv python tools/exact_perf.py 1000000 10000 1 > gpurun_out/exact_ncu2.log 2>&1; echo "ncu rc=$?"
