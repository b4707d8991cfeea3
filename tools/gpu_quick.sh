#!/bin/bash
# scratch runner: selected GPU tests (argument: pytest -k expression) + timings of the bench workload at several spp / knobs
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "$1" > $O/t.log 2>&1; echo "pytest rc=$?" >> $O/t.log
: > $O/configs_a.jsonl
for item in 24 32 64; do
  for spp in 64 8; do
    echo "item_spp $item" >> $O/configs_a.jsonl
    VT_ITEM_SPP=$item timeout 300 python tools/run_config.py --config temple_paths --spp $spp >> $O/configs_a.jsonl 2>&1
  done
done
