#!/bin/bash
# ncu --set full captures of the two brick-volume kernels (configs[3]: trace_primary_kernel brick + shadow variant,
# configs[4]: trace_rays_kernel):  gpurun --timeout 900 -- 'bash tools/gpu_prof_bricks.sh'
O=gpurun_out; mkdir -p $O
cp vtrace_b200/librender.so $O/librender_profiled.so
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 600 $NCU -k regex:trace_primary_kernel -s 3 -o $O/prof_heightmap python tools/run_config.py --config heightmap_4k --frames 3 --warmup 1 > $O/ncu_heightmap.log 2>&1; echo "rc=$?"
timeout 600 $NCU -k regex:trace_rays_kernel -s 2 -o $O/prof_rays python tools/run_config.py --config sparse_rays --frames 2 --warmup 1 > $O/ncu_rays.log 2>&1; echo "rc=$?"
