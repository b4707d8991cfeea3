#!/bin/bash
# exactness check only: gpurun --gpus N --timeout 600 -- 'bash tools/gpu_multi_check.sh N'
N=${1:-2}
O=gpurun_out; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $O/multi_check_$N.log 2>&1
echo "check rc=$?" >> $O/multi_check_$N.log
tail -n 3 $O/multi_check_$N.log
