#!/bin/bash
# GPU tests + device timings of the configurations that run the general path kernel (trace_paths_kernel):
#   gpurun --timeout 900 -- 'bash tools/gpu_general_paths.sh'
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in "world_paths --frames 5" "temple_paths --grid --spp 8 --frames 5" "heightmap_paths --frames 5" "temple_primary --grid"; do
    timeout 120 python tools/run_config.py --config $c | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['config'], 'grid' if d['grid'] else '', 'trace_ms', round(d['trace_ms_median'],4), 'rays', d['rays'], 'iters', d['iterations'])"
done
