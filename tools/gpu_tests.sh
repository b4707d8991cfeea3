#!/bin/bash
# GPU parity tests only (argument: optional pytest -k expression): gpurun --timeout 900 -- 'bash tools/gpu_tests.sh [expr]'
O=gpurun_out; mkdir -p $O
if [ -n "$1" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$1" > $O/t.log 2>&1; else timeout 900 python -m pytest tests -m gpu -x -q > $O/t.log 2>&1; fi
echo "pytest rc=$?" >> $O/t.log
tail -n 25 $O/t.log
