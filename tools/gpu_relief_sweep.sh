#!/bin/bash
# Root relief (samples fewer on rank 0) of frames shared by samples: gpurun --gpus N -- 'bash tools/gpu_relief_sweep.sh N "0 1 2"'
N=${1:-8}
for relief in ${2:-0 1 2}; do
VT_ROOT_RELIEF_SPP=$relief timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-configs --cpu-spp 1 2>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('N=$N root relief $relief spp', d['run']['spp_per_rank'], 'step_ms %.4f e2e_ms %.4f kernel_max %.4f root_kernel %.4f closeup %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms_max_over_ranks'], d['roofline']['kernel_ms'], d['secondary']['ms_per_step']))"
done
