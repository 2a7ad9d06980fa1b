rm -f gpurun_out/build6.log
for cfg in "vis_global_build=16384 build_batch_max=16384 build_batch_ratio=32" "vis_global_build=16384 build_batch_max=16384 build_batch_ratio=16" "vis_global_build=16384 build_batch_max=8192 build_batch_ratio=8" "vis_global_build=16384 build_batch_max=12288 build_batch_ratio=16" "vis_global_build=0 build_batch_max=8192 build_batch_ratio=16"; do
  echo "== $cfg" >> gpurun_out/build6.log
  timeout 300 python tools/build_only.py 1000000 $cfg >> gpurun_out/build6.log 2>&1
done
cat gpurun_out/build6.log
