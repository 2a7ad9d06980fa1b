rm -f gpurun_out/sweep9.log
timeout 300 python tools/sweep.py 1000000 48 >> gpurun_out/sweep9.log 2>&1
timeout 300 python tools/sweep.py 1000000 64 >> gpurun_out/sweep9.log 2>&1
cat gpurun_out/sweep9.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest16.log
tail -5 gpurun_out/pytest16.log
