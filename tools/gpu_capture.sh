timeout 900 python -m pytest tests -m gpu -x -q -k "config3 or config4 or brick or shadow or rays" > gpurun_out/t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t.log
: > gpurun_out/configs_a.jsonl
for c in "heightmap_4k" "sparse_rays --frames 5"; do
  timeout 300 python tools/run_config.py --config $c >> gpurun_out/configs_a.jsonl 2>&1
done
