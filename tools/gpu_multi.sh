#!/bin/bash
# Multi-GPU exactness check + bench lines for both ways of sharing a frame (run on N GPUs of one box):
#   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh N'
N=${1:-2}; O=gpurun_out; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29511 tests/multi_gpu_check.py > $O/multi_gpu_check_n$N.log 2>&1; echo "check rc=$?"; tail -n 2 $O/multi_gpu_check_n$N.log
timeout 600 $RUN --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-configs > $O/bench_n${N}_rows.json 2> $O/bench_n${N}_rows.err; echo "rows rc=$?"
timeout 600 $RUN --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --no-configs --partition samples > $O/bench_n${N}_samples.json 2> $O/bench_n${N}_samples.err; echo "samples rc=$?"
for f in rows samples; do python - $O/bench_n${N}_$f.json $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'n', d['n_gpus'], 'step_ms %.4f' % d['ms_per_step'], 'e2e_ms %.4f' % d['e2e']['ms_per_step'], 'kernel_ms %.4f' % d['roofline']['kernel_ms_max_over_ranks'], 'exact', d['parity']['exact'], d['parity']['frame_fnv'], 'secondary %.4f' % d['secondary']['ms_per_step'])
except Exception as e:
    print(sys.argv[2], 'failed', e)
PY
done
