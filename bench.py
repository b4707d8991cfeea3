#!/usr/bin/env python
"""bench.py — headline benchmark of the vtrace voxel-tracing hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm (oracle port)

Workload (BASELINE.json configs[2]): AncientTemple.vox, 1920x1080, path tracing,
4 bounces, 64 spp per frame; the 64 samples are sharded over the N ranks (rank g renders samples
s = g mod N), the fixed-point accumulation buffers are summed with one NCCL all-reduce and
resolved — so the total work is fixed as N grows ("scaling": "strong").  One step = one frame.

Metric: Mrays/s = ray segments traced (primary + bounce) by all ranks / device time, max over
ranks.  `value` is measured with everything resident in HBM; `e2e` goes through the
reference-facing C ABI with HOST buffers every step (instance matrices written into the pinned
staging returned by start_update_instances, projection/camera passed by host pointer to the
frame call, the finished RGBA8 frame read back to host memory), copies inside the timed region.

The reference (Vulkan + GLSL + Rust, needs a window) cannot run on the box; its CPU arm here is
the oracle = C transcription of trace.frag/trace.vert ("kind": "port"), see DESIGN.md §2.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "Mrays/s (primary+bounce) at 1080p"
UNIT = "Mrays/s"
WIDTH, HEIGHT, SPP, BOUNCES, SEED = 1920, 1080, 64, 4, 0x5EED
WORKLOAD = "AncientTemple.vox 1920x1080 path tracing, 4 bounces, 64 spp (configs[2]), camera eye=(1.6,-0.9,1.2) fov 80deg"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the trace kernel from the committed ncu capture, or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("trace_paths_dram_bytes_per_launch")
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def scene_inputs():
    from tools import scenes
    chunk = scenes.load_asset("AncientTemple")
    P, V = scenes.camera(WIDTH, HEIGHT)
    return chunk, P, V, scenes.single_instance(0)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host's cores

def host_threads() -> int:
    """All host threads this process may use (torchrun pins OMP_NUM_THREADS=1; the CPU arm ignores that)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample(threads: int, spp: int):
    """Times `spp` samples/pixel of the full 1080p workload on the oracle; returns (Mrays/s, s, rays)."""
    import oracle_lib
    chunk, P, V, inst = scene_inputs()
    sc = oracle_lib.OracleScene()
    sc.add_texture(chunk.get_raw(), *chunk.dims())
    sc.set_instances(inst)
    t0 = time.perf_counter()
    _, rays, iters = sc.render_paths(P, V, WIDTH, HEIGHT, spp=spp, bounces=BOUNCES, seed=SEED, threads=threads)
    dt = time.perf_counter() - t0
    return rays / dt / 1e6, dt, rays, iters


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # the CPU arm runs once, on rank 0; other ranks exit without work
    threads = host_threads()
    spp = args.ref_spp
    for _ in range(args.warmup):
        cpu_sample(threads, 1)
    times, rays_total = [], 0
    for _ in range(args.steps):
        _, dt, rays, _ = cpu_sample(threads, spp)
        times.append(dt)
        rays_total += rays
    total = sum(times)
    value = rays_total / total / 1e6
    sample = f"{spp} of the {SPP} spp of the 1080p frame per step (samples 0..{spp - 1}), {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference = C transcription of trace.frag/trace.vert on host cores; "
                                                  "Vulkan/lavapipe/rustc unavailable on the box"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# CUDA arm

def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from vtrace_b200 import abi
    from vtrace_b200.distributed import reduce_accum, setup_fused_reduce, shard_samples, stream_barrier
    from vtrace_b200.renderer import Renderer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    if SPP % world:
        raise SystemExit(f"{SPP} spp do not shard over {world} ranks")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    chunk, P, V, inst = scene_inputs()
    r = Renderer()  # entry(): picks LOCAL_RANK's device
    r.add_texture(chunk)
    r.update_instances_raw(inst)
    first, stride, count = shard_samples(SPP, rank, world)
    r.configure(width=WIDTH, height=HEIGHT, mode=abi.MODE_PATHS, flags=abi.FLAG_NO_HIT_RECORDS, spp=count,
                bounces=BOUNCES, seed=SEED, sample_first=first, sample_stride=stride, total_spp=SPP, max_frames=0)
    # a non-default torch stream becomes the current stream; the library enqueues on it too, so torch
    # CUDA events bracket the library's kernels and NCCL is ordered with them
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    r.set_stream(stream.cuda_stream)
    # N > 1: by default every rank pushes its sums of the covered rectangle straight into rank 0's memory over
    # NVLink (vt_fused_reduce_*); --reduce allreduce keeps a per-rank buffer and sums them with NCCL
    fused = (world > 1 or args.force_fused) and args.reduce == "fused"
    fused_flags = os.environ.get("VT_FUSED_SYNC", "1") != "0"  # the library's own flag synchronisation (default)
    accum = None
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    if fused and world == 1:
        r.fused_reduce_export(1)  # (--force-fused: the multi-GPU data path with a single rank, to profile its kernels)
    elif fused and not setup_fused_reduce(r, rank, world, dev):
        fused = False  # no peer access between the GPUs of this box: sum the accumulators with NCCL instead
    if not fused:
        accum = torch.zeros((HEIGHT, WIDTH, 3), dtype=torch.int64, device=dev)  # 2^-24 fixed-point radiance sums
        r.set_accum_buffer(accum.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    frame_pinned = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.uint8, pin_memory=True)  # the host-side frame buffer
    frame_host = frame_pinned.numpy()
    lib = abi.load()

    def trace_and_reduce():
        if fused:
            r.fused_reduce_next_frame()
            r.render_async(P, V)   # trace, push the partial sums into rank 0's memory, raise this rank's flag
            if not fused_flags:
                stream_barrier(flag)   # VT_FUSED_SYNC=0: order the ranks with a 4-byte NCCL all-reduce instead
            if rank == 0:
                r.resolve()        # waits for every rank's flag, sums the slots, encodes the frame
        else:
            accum.zero_()
            r.render_async(P, V)
            reduce_accum(accum)
            r.resolve()

    def step_resident():
        """One frame, everything device-resident, no host copies."""
        trace_and_reduce()

    def step_e2e():
        """One frame through the reference-facing ABI with host buffers."""
        r.update_instances_raw(inst)          # host matrices -> pinned staging -> device
        if world == 1 and not fused:
            assert r.render_tick_raw(P, V)    # projection/camera by host pointer; clear + trace + resolve
        else:
            trace_and_reduce()
        if rank == 0 or not fused:
            n = lib.vt_read_color(frame_host.ctypes.data, frame_host.nbytes)  # finished frame -> host
            assert n == frame_host.nbytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps."""
        ms, trace_ms, rays, iters = [], [], 0, 0
        for _ in range(steps):
            flush.fill_(1)  # untimed: evict the previous frame from L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_fn()
            e1.record(stream)
            e1.synchronize()
            r.synchronize()
            ms.append(e0.elapsed_time(e1))
            st = r.stats()
            trace_ms.append(st.last_trace_ms)
            rays += st.rays
            iters += st.iterations
        return ms, trace_ms, rays, iters

    def timed_resident(steps):
        """K device-resident steps enqueued back to back (the host never waits inside the timed region, as
        a renderer that pipelines its frames would); each step is bracketed by its own CUDA events on the
        launching stream, with the L2 flushed in between; one synchronisation at the end."""
        r.stats()  # folds everything so far; the kernel-time sums restart here
        evs = []
        for _ in range(steps):
            flush.fill_(1)  # untimed: evict the previous frame from L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_resident()
            e1.record(stream)
            evs.append((e0, e1))
        torch.cuda.synchronize()
        st = r.stats()
        assert st.trace_frames == steps, (st.trace_frames, steps)
        ms = [a.elapsed_time(b) for a, b in evs]
        per_frame_trace = st.trace_ms_sum / steps
        return ms, [per_frame_trace] * steps, st.rays * steps, st.iterations * steps  # every step renders the same frame

    for _ in range(max(args.warmup, 3)):
        step_resident()
        step_e2e()
    barrier()
    launches0 = r.stats().launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    ms, trace_ms, rays, iters = timed_resident(args.steps)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = r.stats().launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    ms_e, _, rays_e, _ = timed(step_e2e, args.steps)
    barrier()

    # max over ranks of the device time; sum over ranks of the work
    t = torch.tensor([sum(ms), sum(ms_e), sum(trace_ms)], dtype=torch.float64, device=dev)
    w = torch.tensor([rays, rays_e, iters, launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    t_res, t_e2e, t_trace = (float(x) for x in t.tolist())
    rays_all, rays_e_all, iters_all, launches_all = (int(x) for x in w.tolist())

    if rank == 0:
        steps = args.steps
        value = rays_all / (t_res * 1e-3) / 1e6
        e2e_value = rays_e_all / (t_e2e * 1e-3) / 1e6
        # roofline of the dominant kernel (trace_paths_wave_kernel), per launch on THIS rank:
        # algorithmic bytes = 4 B per DDA iteration (one RGBA8 voxel record, trace.frag:76) +
        # 16 B per pixel of accumulator read-modify-write
        st = r.stats()
        peak, peak_src = measured_peaks()
        alg_bytes = 4.0 * (iters / steps) + 16.0 * WIDTH * HEIGHT
        kernel_s = (sum(trace_ms) / steps) * 1e-3
        achieved = alg_bytes / kernel_s / 1e9
        cpu = None
        if world == 1 and not args.no_cpu:
            threads = host_threads()
            v, dt, _, _ = cpu_sample(threads, args.cpu_spp)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_spp} of the {SPP} spp of the same 1080p frame (samples 0..{args.cpu_spp - 1}), "
                             f"{dt:.2f} s wall on {threads} threads"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": t_res / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "spp_per_rank": SPP // world, "partition": (f"spp sharded over {world} rank(s), " + (
                           ("partial sums pushed into rank 0's memory over NVLink peer stores, ordered by " +
                            ("flags in peer memory" if fused_flags else "a 4-byte NCCL stream barrier")) if fused
                           else "NCCL all-reduce of 3*w*h int64")) if world > 1 else "single rank",
                       "l2": "flushed between steps (256 MiB fill, untimed); scene itself is 256 KB and lives in shared memory/L2 by design",
                       "masks_in_smem": bool(st.masks_in_smem)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(inst.nbytes + 128),
                    "d2h_bytes_per_step": int(frame_host.nbytes + 16), "ms_per_step": t_e2e / steps},
            "gpu_launches": launches_all,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(), "kernel": "trace_paths_wave_kernel", "kernel_ms": kernel_s * 1e3,
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "kernel_ms_max_over_ranks": t_trace / steps,
                         "dda_iterations_per_launch": iters / steps,
                         "dda_iterations_per_s": (iters / steps) / kernel_s},
            "clocks": clocks,
            "wall_s": t_wall,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if not fused:
        r.set_accum_buffer(None)
    r.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--cpu-spp", type=int, default=64, help="spp of the bounded cpu_baseline sample (64 = the whole frame)")
    ap.add_argument("--ref-spp", type=int, default=64, help="spp per step of the --impl reference arm (64 = the whole frame)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--force-fused", action="store_true", help="N=1 only: run the fused multi-GPU data path with one rank (profiling aid)")
    ap.add_argument("--reduce", default="fused", choices=["fused", "allreduce"], help="cross-GPU accumulation for N > 1")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
