TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --mode partitioned --steps 20 --warmup 3 > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err; echo "c4 rc=$?"; cat gpurun_out/bench_c4_n2.json | cut -c1-1500
timeout 300 python tools/partition_check.py 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/partition_check.py 2>&1 | tail -3
