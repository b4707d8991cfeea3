#!/bin/bash
# Sweep of the root's relief and the item order for frames shared by tile rows (N GPUs):
#   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_rows_sweep.sh N "0/8 1/16 1/8" "0 1"'
N=${1:-2}
for relief in ${2:-0/8 1/16 1/8}; do for order in ${3:-0 1}; do
VT_ITEM_ORDER=$order VT_ROOT_RELIEF_ROWS=$relief timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-configs --cpu-spp 1 2>gpurun_out/sweep.err | python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('N=$N relief $relief order $order step_ms %.4f e2e_ms %.4f kernel_max %.4f root_kernel %.4f closeup %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms_max_over_ranks'], d['roofline']['kernel_ms'], d['secondary']['ms_per_step']))
except Exception as e:
    print('failed', e); print(open('gpurun_out/sweep.err').read()[-1500:])"
done; done
if [ -n "$4" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-configs --cpu-spp 1 --partition samples 2>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('N=$N samples step_ms %.4f e2e_ms %.4f kernel_max %.4f root_kernel %.4f closeup %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms_max_over_ranks'], d['roofline']['kernel_ms'], d['secondary']['ms_per_step']))"
fi
