#!/bin/bash
O=gpurun_out; mkdir -p $O
cp vtrace_b200/librender.so $O/librender_spp1.so
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 600 $NCU -k regex:trace_paths_wave_kernel -s 4 -o $O/prof_spp1 python tools/run_config.py --config temple_paths --spp 1 --frames 4 > $O/ncu_spp1.log 2>&1
timeout 600 $NCU -k regex:trace_paths_wave_kernel -s 4 -o $O/prof_spp8 python tools/run_config.py --config temple_paths --spp 8 --frames 4 > $O/ncu_spp8.log 2>&1
