#!/usr/bin/env python
"""Joins an ncu report's per-SASS-instruction counters with nvdisasm line info and prints where
a kernel's issue slots and stall samples go, aggregated per CUDA source line.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep trace_paths_kernelILb1 [--top 40]

Needs the librender.so that was profiled (the cubin is extracted from it with cuobjdump).
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(so_path: str, kernel_substr: str):
    """[(offset, text, file, line)] for the first function whose mangled name contains kernel_substr."""
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", so_path], cwd=td, check=True, capture_output=True)
        out = []
        for cubin in sorted(os.listdir(td)):
            if not cubin.endswith(".cubin"):
                continue
            txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
            cur_fn, cur_file, cur_line, active = None, None, None, False
            for ln in txt.splitlines():
                m = re.match(r"\s*\.text\.(\S+):", ln)
                if m:
                    cur_fn = m.group(1)
                    active = kernel_substr in cur_fn and not out
                    continue
                if ln.strip().startswith(".section") or re.match(r"\s*\.text\.", ln):
                    if out:
                        active = False
                m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
                if m:
                    cur_file, cur_line = m.group(1), int(m.group(2))
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if m and active:
                    out.append((int(m.group(1), 16), m.group(2).strip(), cur_file, cur_line))
            if out:
                return out
    raise SystemExit(f"kernel {kernel_substr} not found in {so_path}")


def ncu_sass(rep: str, kernel_regex: str | None):
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
    if kernel_regex:
        cmd += ["-k", f"regex:{kernel_regex}"]
    txt = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    return hdr, body


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel", help="substring of the mangled kernel name, e.g. trace_paths_kernelILb1")
    ap.add_argument("--so", default=os.path.join(ROOT, "vtrace_b200", "librender.so"))
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--regex", default=None, help="ncu -k regex when the report holds several kernels")
    ap.add_argument("--sass", default=None, help="FILE:LINE — list every SASS instruction attributed to that source line with its counters")
    ap.add_argument("--buckets", action="store_true",
                    help="aggregate per code region: a region starts at a '// ----' banner comment or a function definition")
    args = ap.parse_args()

    sass = sass_lines(os.path.abspath(args.so), args.kernel)
    hdr, body = ncu_sass(args.report, args.regex)
    col = {h: i for i, h in enumerate(hdr)}
    if len(body) != len(sass):
        print(f"warning: ncu has {len(body)} instructions, cubin has {len(sass)} (profile from another build?)", file=sys.stderr)
    n = min(len(body), len(sass))
    src = {}
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])  # inst, thread inst, samples, not-issued samples
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    stalls = collections.Counter()
    tot_inst = tot_thr = tot_samp = 0
    for k in range(n):
        r = body[k]
        inst = int(r[col["Instructions Executed"]] or 0)
        thr = int(r[col["Thread Instructions Executed"]] or 0)
        samp = int(r[col["# Samples"]] or 0)
        _, text, f, line = sass[k]
        key = (os.path.basename(f or "?"), line)
        a = agg[key]
        a[0] += inst; a[1] += thr; a[2] += samp
        tot_inst += inst; tot_thr += thr; tot_samp += samp
        for h in stall_cols:
            stalls[h] += int(r[col[h]] or 0)
    print(f"kernel instructions: {n}; warp-inst executed {tot_inst:,}; thread-inst {tot_thr:,}; "
          f"SIMT efficiency {tot_thr / max(tot_inst, 1) / 32:.3f}; samples {tot_samp:,}")
    print("stall reasons (all samples):", ", ".join(f"{k[6:]}={v}" for k, v in stalls.most_common(8)))
    if args.sass:
        want_f, _, want_l = args.sass.partition(":")
        print(f"{'offset':>8}{'inst%':>7}{'samp%':>7}{'lanes':>6}  top stall        sass")
        for k in range(n):
            off, text, f, line = sass[k]
            if os.path.basename(f or "?") != want_f or (want_l and line != int(want_l)):
                continue
            r = body[k]
            inst = int(r[col["Instructions Executed"]] or 0)
            thr = int(r[col["Thread Instructions Executed"]] or 0)
            samp = int(r[col["# Samples"]] or 0)
            st = max(stall_cols, key=lambda h: int(r[col[h]] or 0)) if samp else ""
            print(f"{off:8x}{100 * inst / max(tot_inst, 1):7.2f}{100 * samp / max(tot_samp, 1):7.2f}{thr / max(inst, 1):6.1f}  {st[6:]:<16} {text[:80]}")
        return
    lines_cache = {}

    def text_of(fname, line):
        path = os.path.join(ROOT, "vtrace_b200", "csrc", fname)
        if path not in lines_cache:
            try:
                lines_cache[path] = open(path).read().splitlines()
            except OSError:
                lines_cache[path] = []
        ls = lines_cache[path]
        return ls[line - 1].strip()[:90] if line and 0 < line <= len(ls) else ""

    if args.buckets:
        import bisect
        bounds = {}

        def region(fname, line):
            path = os.path.join(ROOT, "vtrace_b200", "csrc", fname)
            if fname not in bounds:
                try:
                    ls = open(path).read().splitlines()
                except OSError:
                    ls = []
                marks = [(i + 1, l.strip()) for i, l in enumerate(ls)
                         if re.match(r"\s*// ----", l) or re.match(r"(template|__device__|__global__|static|inline|auto \w+ = \[)", l.strip())
                         or re.match(r"\s*auto \w+ = \[&\]", l)]
                bounds[fname] = marks
            marks = bounds[fname]
            k = bisect.bisect_right([m[0] for m in marks], line or 0) - 1
            if k < 0:
                return fname
            # a template line is followed by the real signature: prefer that
            name = marks[k][1]
            if name.startswith("template") and k + 1 < len(marks) and marks[k + 1][0] <= (line or 0) + 0:
                name = marks[k + 1][1]
            return f"{fname}:{marks[k][0]} {name[:70]}"
        bagg = collections.defaultdict(lambda: [0, 0, 0])
        for (fname, line), a in agg.items():
            b = bagg[region(fname, line)]
            b[0] += a[0]; b[1] += a[1]; b[2] += a[2]
        print(f"{'inst%':>7}{'samp%':>7}{'lanes':>7}  region")
        for name, a in sorted(bagg.items(), key=lambda kv: -kv[1][0])[: args.top]:
            print(f"{100 * a[0] / max(tot_inst, 1):7.2f}{100 * a[2] / max(tot_samp, 1):7.2f}{a[1] / max(a[0], 1):7.1f}  {name}")
        return
    print(f"{'file:line':<22}{'inst%':>7}{'samp%':>7}{'thr/inst':>9}  source")
    for (fname, line), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[: args.top]:
        print(f"{fname + ':' + str(line):<22}{100 * a[0] / max(tot_inst, 1):7.2f}{100 * a[2] / max(tot_samp, 1):7.2f}"
              f"{a[1] / max(a[0], 1):9.1f}  {text_of(fname, line)}")


if __name__ == "__main__":
    main()
