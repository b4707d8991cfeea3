#!/bin/bash
# One `ncu --set full` capture of the bench kernel (trace_paths_wave_kernel) plus a quick bench line:
#   gpurun --timeout 900 -- 'bash tools/gpu_prof.sh [extra bench.py args]'
O=gpurun_out; mkdir -p $O
cp vtrace_b200/librender.so $O/librender_profiled.so
timeout 600 python bench.py --steps 10 --warmup 3 --no-configs "$@" > $O/bench_quick.json 2> $O/prof.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -c 1 -f -k regex:trace_paths_wave_kernel -s 4 -o $O/prof_paths \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-configs "$@" > $O/ncu_paths.log 2>&1; echo "ncu rc=$?"
tail -n 3 $O/ncu_paths.log
