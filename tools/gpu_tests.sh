#!/bin/bash
# GPU parity tests only (argument: optional pytest -k expression), then smoke() and a short bench:
#   gpurun --timeout 1200 -- 'bash tools/gpu_tests.sh [expr]'
O=gpurun_out; mkdir -p $O
if [ -n "$1" ]; then timeout 1200 python -m pytest tests -m gpu -x -q -k "$1" > $O/t.log 2>&1; else timeout 1200 python -m pytest tests -m gpu -x -q > $O/t.log 2>&1; fi
echo "pytest rc=$?" >> $O/t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $O/t.log 2>&1; echo "smoke rc=$?" >> $O/t.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_quick.json 2>> $O/t.log; echo "bench rc=$?" >> $O/t.log
tail -n 14 $O/t.log
