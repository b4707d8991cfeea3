#!/bin/bash
# multi-GPU check + bench: gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh N'
N=${1:-2}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/multi_gpu_check.py > $O/multi_check_$N.log 2>&1; echo "check rc=$?" >> $O/multi_check_$N.log
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?" >> $O/bench_n$N.err
tail -n 3 $O/multi_check_$N.log
