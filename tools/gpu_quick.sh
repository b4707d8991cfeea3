#!/bin/bash
# scratch runner: selected GPU tests (argument: pytest -k expression) + timings of the bench workload at several spp / knobs
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "$1" > $O/t.log 2>&1; echo "pytest rc=$?" >> $O/t.log
: > $O/configs_a.jsonl
for ipw in 1 2 3 6 12; do
  for spp in 64 8; do
    echo "items_per_warp $ipw" >> $O/configs_a.jsonl
    VT_ITEMS_PER_WARP=$ipw timeout 300 python tools/run_config.py --config temple_paths --spp $spp >> $O/configs_a.jsonl 2>&1
  done
done
for spp in 64 8 1; do
echo "no sky" >> $O/configs_a.jsonl
VT_DEBUG_NO_SKY=1 timeout 300 python tools/run_config.py --config temple_paths --spp $spp >> $O/configs_a.jsonl 2>&1
done
echo "spp 1" >> $O/configs_a.jsonl
timeout 300 python tools/run_config.py --config temple_paths --spp 1 >> $O/configs_a.jsonl 2>&1
