#!/bin/bash
# A/B timing of side-by-side builds (variants/*/librender.so, see vtrace_b200/build.py --variant) and tuning knobs:
#   gpurun --timeout 900 -- 'bash tools/gpu_ab.sh "base w8s96c2 ..." "6 12" "16"'
#     $1 = variants (base = vtrace_b200/librender.so), $2 = VT_REFILL_BATCH values, $3 = VT_REFILL values
O=gpurun_out; mkdir -p $O; : > $O/ab.txt
for v in $1; do for b in ${2:-6}; do for r in ${3:-16}; do
  lib=variants/$v/librender.so; [ "$v" = base ] && lib=vtrace_b200/librender.so
  VT_LIBRENDER=$PWD/$lib VT_REFILL_BATCH=$b VT_REFILL=$r timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-configs 2>>$O/ab.err | \
    python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$v batch=$b refill=$r', 'kernel_ms=%.4f step_ms=%.4f closeup_ms=%.4f fnv=%s' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['secondary']['ms_per_step'], d['parity']['frame_fnv']))" >> $O/ab.txt
done; done; done
cat $O/ab.txt
