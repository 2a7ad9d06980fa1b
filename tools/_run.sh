rm -f gpurun_out/sweep13.log gpurun_out/build9.log
timeout 300 python tools/sweep.py 1000000 48 >> gpurun_out/sweep13.log 2>&1
timeout 300 python tools/sweep.py 1000000 48 >> gpurun_out/sweep13.log 2>&1
cat gpurun_out/sweep13.log
timeout 300 python tools/build_only.py 1000000 >> gpurun_out/build9.log 2>&1
cut -c1-200 gpurun_out/build9.log
