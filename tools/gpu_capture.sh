set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t.log
: > gpurun_out/configs_a.jsonl
for c in "heightmap_4k" "sparse_rays --frames 5"; do
  timeout 300 python tools/run_config.py --config $c >> gpurun_out/configs_a.jsonl 2>&1
done
cp vtrace_b200/librender.so gpurun_out/librender_a.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_rays_kernel -s 2 -c 1 -o gpurun_out/rays_a -f python tools/run_config.py --config sparse_rays --frames 2 --warmup 1 > gpurun_out/ncu_rays_a.log 2>&1
