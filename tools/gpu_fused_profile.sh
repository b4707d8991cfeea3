#!/bin/bash
# Launch list of the fused multi-GPU data path run with ONE rank (bench.py --force-fused): durations of the
# push / flag / resolve kernels that sit next to the trace kernel at N > 1.
O=gpurun_out; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_fused.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --force-fused > $O/bench_fused_ncu.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --force-fused > $O/bench_fused_single.json 2>> $O/bench_fused_ncu.log
