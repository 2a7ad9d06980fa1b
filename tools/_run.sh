timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest17.log
tail -40 gpurun_out/pytest17.log
