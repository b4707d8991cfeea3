#!/usr/bin/env python
"""Condenses `ncu --set full` reports into the counters profiles/ keeps.

    python tools/ncu_summary.py label=report.ncu-rep [label=report.ncu-rep ...] > profiles/rNN_ncu_summary.json

One entry per report (its first captured launch).  Runs here, on the CPU container, on reports
brought back from the GPU box in gpurun_out/.
"""
from __future__ import annotations

import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum",  # x 32 B = L2 traffic
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__grid_size",
    "launch__block_size",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
]


def summarise(path: str) -> dict:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units, vals = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(names)}
    d = {"kernel": vals[col["Kernel Name"]]}
    for k in KEEP:
        if k in col:
            u = units[col[k]]
            d[k] = f"{vals[col[k]]} {u}".strip()
    return d


def main():
    res = {}
    for arg in sys.argv[1:]:
        label, _, path = arg.partition("=")
        res[label] = summarise(path)
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
