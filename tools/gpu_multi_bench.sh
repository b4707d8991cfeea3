#!/bin/bash
# bench only: gpurun --gpus N --timeout 600 -- 'bash tools/gpu_multi_bench.sh N'
N=${1:-2}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?" >> $O/bench_n$N.err
