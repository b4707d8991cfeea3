#!/bin/bash
# device timings of the brick-volume configurations for each listed variant (base = vtrace_b200/librender.so):
#   gpurun --timeout 900 -- 'bash tools/gpu_configs.sh "base nopf"'
for v in ${1:-base}; do
  lib=variants/$v/librender.so; [ "$v" = base ] && lib=vtrace_b200/librender.so
  for c in "heightmap_4k" "sparse_rays --frames 5"; do
    VT_LIBRENDER=$PWD/$lib python tools/run_config.py --config $c | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$v', d['config'], 'trace_ms', round(d['trace_ms_median'],4), 'min', round(d['trace_ms_min'],4), 'giters/s', round(d['giters_per_s'],1), 'rays', d['rays'], 'iters', d['iterations'])"
  done
done
