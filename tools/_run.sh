N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --mode partitioned --steps 20 --warmup 5 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "c4 rc=$?"; cut -c1-1400 gpurun_out/bench_c4_n$N.json
