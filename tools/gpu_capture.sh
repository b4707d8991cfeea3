#!/bin/bash
# Round-end capture on ONE B200 (run through gpurun): GPU tests, the bench lines, the ncu launch list of the
# bench command, one `ncu --set full` capture per trace kernel, device timings of every configuration.
# Everything lands in gpurun_out/; tools/ncu_summary.py and tools/ncu_lines.py condense it into profiles/.
#   gpurun --timeout 1800 -- 'bash tools/gpu_capture.sh'
set -x
O=gpurun_out
mkdir -p $O
cp vtrace_b200/librender.so $O/librender_profiled.so   # tools/ncu_lines.py --so needs the exact binary
timeout 900 python -m pytest tests -m gpu -x -q > $O/t.log 2>&1; echo "pytest rc=$?" >> $O/t.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2>> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > $O/bench_under_ncu.log 2>&1
: > $O/configs.jsonl
for c in "treasure_primary" "temple_primary" "temple_primary --closeup" "temple_paths" "temple_paths --closeup" \
         "treasure_paths --closeup" "temple_primary --grid" "temple_paths --grid --spp 8 --frames 5" \
         "world_primary" "world_paths --frames 5" "heightmap_4k" "sparse_rays --frames 5" "heightmap_paths --frames 5"; do
    timeout 300 python tools/run_config.py --config $c >> $O/configs.jsonl 2>> $O/bench.err
done
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 600 $NCU -k regex:trace_paths_wave_kernel -s 4 -o $O/prof_paths python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > $O/ncu_paths.log 2>&1
timeout 600 $NCU -k regex:trace_primary_kernel -s 5 -o $O/prof_primary python tools/run_config.py --config temple_primary --closeup --frames 4 > $O/ncu_primary.log 2>&1
timeout 600 $NCU -k regex:trace_primary_kernel -s 3 -o $O/prof_heightmap python tools/run_config.py --config heightmap_4k --frames 3 --warmup 1 > $O/ncu_heightmap.log 2>&1
timeout 600 $NCU -k regex:trace_rays_kernel -s 2 -o $O/prof_rays python tools/run_config.py --config sparse_rays --frames 2 --warmup 1 > $O/ncu_rays.log 2>&1
timeout 600 $NCU -k regex:trace_paths_kernel -s 2 -o $O/prof_world python tools/run_config.py --config world_paths --frames 3 --warmup 1 > $O/ncu_world.log 2>&1
timeout 600 $NCU -k regex:trace_paths_kernel -s 2 -o $O/prof_grid python tools/run_config.py --config temple_paths --grid --spp 8 --frames 3 --warmup 1 > $O/ncu_grid.log 2>&1
