timeout 300 python tools/sweep.py 1000000 48 > gpurun_out/sweep7.log 2>&1; cat gpurun_out/sweep7.log
timeout 900 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench7.json; grep -o '"kernel_ms_avg": [0-9.]*' gpurun_out/bench7.json
