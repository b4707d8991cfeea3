#!/bin/bash
# Condenses what tools/gpu_capture.sh brought back in gpurun_out/ into the tracked files under profiles/ (run here, on the CPU):
#   bash tools/make_profiles.sh r02
R=${1:-r02}; O=gpurun_out; P=profiles; SO=$O/librender_profiled.so
for f in bench_n1 bench_reference_arm; do tail -n 1 $O/$f.json > $P/${R}_$f.json; done
cp $O/configs.jsonl $P/${R}_configs.jsonl
cp $O/launches_bench.csv $P/${R}_launches_bench.csv
python tools/ncu_summary.py \
  "trace_paths_wave_kernel (bench workload, configs[2])=$O/prof_paths.ncu-rep" \
  "trace_primary_kernel (configs[1], close-up camera)=$O/prof_primary.ncu-rep" \
  "trace_primary_kernel, brick + shadow variant (configs[3], 1024^3 heightmap at 4K)=$O/prof_heightmap.ncu-rep" \
  "trace_rays_kernel (configs[4], 2^26 incoherent rays through 4096^3 sparse bricks)=$O/prof_rays.ncu-rep" \
  "trace_paths_kernel (the reference's default world, 343 instances, 1000x1000, 8 spp, masks in global memory)=$O/prof_world.ncu-rep" \
  "trace_paths_kernel (11x11 entity grid, 1080p, 8 spp, masks in shared memory)=$O/prof_grid.ncu-rep" > $P/${R}_ncu_summary.json
python tools/ncu_lines.py $O/prof_paths.ncu-rep trace_paths_wave_kernelILb1E --so $SO --buckets > $P/${R}_trace_paths_regions.txt
python tools/ncu_lines.py $O/prof_paths.ncu-rep trace_paths_wave_kernelILb1E --so $SO --top 60 > $P/${R}_trace_paths_lines.txt
python tools/ncu_lines.py $O/prof_heightmap.ncu-rep trace_primary_kernelILb0ELb1ELb1E --so $SO --buckets > $P/${R}_trace_heightmap_regions.txt
python tools/ncu_lines.py $O/prof_rays.ncu-rep trace_rays_kernel --so $SO --buckets > $P/${R}_trace_rays_regions.txt
python tools/ncu_lines.py $O/prof_world.ncu-rep trace_paths_kernelILb0ELb0E --so $SO --buckets > $P/${R}_trace_world_regions.txt
python tools/ncu_lines.py $O/prof_grid.ncu-rep trace_paths_kernelILb1ELb0E --so $SO --buckets > $P/${R}_trace_grid_regions.txt
python - <<PY
import json
d = json.load(open("$P/${R}_ncu_summary.json"))
e = d["trace_paths_wave_kernel (bench workload, configs[2])"]
def nbytes(v):
    x, u = v.split()
    return float(x) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
rd, wr = nbytes(e["dram__bytes_read.sum"]), nbytes(e["dram__bytes_write.sum"])
json.dump({"trace_paths_dram_bytes_per_launch": rd + wr,
           "source": "profiles/${R}_ncu_summary.json (ncu --set full, bench.py workload, N=1, trace_paths_wave_kernel; a lean frame only touches the accumulators of the instance's screen rectangle)",
           "dram_bytes_read": rd, "dram_bytes_write": wr}, open("$P/traffic.json", "w"), indent=1)
PY
head -n 2 $P/${R}_trace_*_regions.txt | cut -c1-160
