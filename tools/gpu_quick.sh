#!/bin/bash
# scratch runner: selected GPU tests + a few config timings (arguments: pytest -k expression)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "$1" > $O/t.log 2>&1; echo "pytest rc=$?" >> $O/t.log
: > $O/configs_a.jsonl
timeout 300 python tools/run_config.py --config temple_paths --grid --spp 8 --frames 5 >> $O/configs_a.jsonl 2>&1
timeout 300 python tools/run_config.py --config heightmap_paths --frames 5 >> $O/configs_a.jsonl 2>&1
