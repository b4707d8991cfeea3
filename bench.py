#!/usr/bin/env python
"""bench.py — headline benchmark of the vtrace voxel-tracing hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm (oracle port)

Workload (BASELINE.json configs[2]): AncientTemple.vox, 1920x1080, path tracing,
4 bounces, 64 spp per frame; the 64 samples are sharded over the N ranks (rank g renders samples
s = g mod N) and the ranks' fixed-point radiance sums are combined over NVLink — so the total work
is fixed as N grows ("scaling": "strong").  One step = one frame.

Metric: Mrays/s = ray segments (primary + bounce) of the frame, all ranks / device time, max over
ranks.  `value` is measured with everything resident in HBM; `e2e` goes through the
reference-facing C ABI with HOST buffers every step (instance matrices written into the pinned
staging returned by start_update_instances, projection/camera passed by host pointer to the
frame call, the finished RGBA8 frame read back to host memory), copies inside the timed region.
`traced_rays` / `traced_mrays_per_s` exclude the camera samples of pixels outside the instance's
screen rectangle, which the kernel resolves analytically (spp x sky) without marching anything.

After the timed regions the line also carries: `parity` (the frame all ranks produced, compared bit for
bit with the CPU oracle's), `secondary` (the same scene from a frame-filling camera) and `configs` (the other
BASELINE.json configurations, each timed a few frames; configs[4] on every rank, ray ids sharded).

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
CLOSEUP_EYE = (0.8, -0.45, 0.6)
WORKLOAD = "AncientTemple.vox 1920x1080 path tracing, 4 bounces, 64 spp (configs[2]), camera eye=(1.6,-0.9,1.2) fov 80deg"
DATA = "reference asset AncientTemple.vox (committed fixture) + fixed synthetic camera"
# identical in both arms (the driver compares the dicts); per-run details live in "run"
CONFIG = {"workload": WORKLOAD,
          "l2": "GPU arm: L2 flushed between steps (256 MiB fill, untimed); the scene itself is 256 KB and lives in "
                "shared memory by design.  CPU arm: n/a"}
SM_CLOCK_GHZ, SMSP_PER_SM, STEP_INSTRUCTIONS = 1.965, 4, 17  # issue roof: one warp instruction per SMSP per clock


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


def fnv1a64(data: bytes) -> str:
    """64-bit digest of a buffer: FNV-1a 64 (what host/vtrace_headless prints) folded over the buffer's SHA-256.
    FNV itself is sequential — 8 MB in pure Python takes seconds — so the heavy pass runs at C speed."""
    import hashlib
    h = 1469598103934665603
    for b in hashlib.sha256(data).digest():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


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


def cpu_sample(threads: int, spp: int, eye=None):
    """Times `spp` samples/pixel of the full 1080p workload on the oracle; returns (Mrays/s, s, rays, iters, accum)."""
    import oracle_lib
    from tools import scenes
    chunk, P, V, inst = scene_inputs()
    if eye is not None:
        P, V = scenes.camera(WIDTH, HEIGHT, eye=eye)
    sc = oracle_lib.OracleScene()
    sc.add_texture(chunk.get_raw(), *chunk.dims())
    sc.set_instances(inst)
    t0 = time.perf_counter()
    accum, rays, iters = sc.render_paths(P, V, WIDTH, HEIGHT, spp=spp, bounces=BOUNCES, seed=SEED, threads=threads)
    dt = time.perf_counter() - t0
    return rays / dt / 1e6, dt, rays, iters, accum


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
        _, dt, rays, _, _ = cpu_sample(threads, spp)
        times.append(dt)
        rays_total += rays
    total = sum(times)
    value = rays_total / total / 1e6
    sample = f"{spp} of the {SPP} spp of the 1080p frame per step (samples 0..{spp - 1}), {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": DATA,
        "config": CONFIG,
        "run": {"note": "reference = C transcription of trace.frag/trace.vert on host cores; Vulkan/lavapipe/rustc "
                        "unavailable on the box"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# CUDA arm

def run_cuda(args):
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    from tools import scenes
    from vtrace_b200 import abi
    from vtrace_b200.distributed import reduce_accum, setup_fused_reduce, shard_samples, shard_samples_weighted, stream_barrier
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
    if not fused and world > 1:
        accum = torch.zeros((HEIGHT, WIDTH, 3), dtype=torch.int64, device=dev)  # 2^-24 fixed-point radiance sums
        r.set_accum_buffer(accum.data_ptr())
    shares = [shard_samples(SPP, k, world) for k in range(world)]
    by_rows = fused and args.partition == "rows" and (world > 1 or bool(os.environ.get("VT_FUSED_ROWS_AS_WORLD")))  # (env: profiling aid)
    if by_rows:
        # the frame is shared out by rows of 8x4 tiles: rank k traces all 64 samples of the tile rows ty = k (mod world), so
        # per-pixel work (camera set-up, accumulator and NVLink traffic, the root's pass over the slots) is divided too
        shares = [(0, 1, SPP) for _ in range(world)]
        r.configure(width=WIDTH, height=HEIGHT, mode=abi.MODE_PATHS, flags=abi.FLAG_NO_HIT_RECORDS, spp=SPP,
                    bounces=BOUNCES, seed=SEED, sample_first=0, sample_stride=1, total_spp=SPP, max_frames=0)
        # ... and the root, which also sums the slots and encodes the frame, gets fewer rows (eighths of a full share)
        # (about 45 us of extra work against a kernel of 0.54 / 0.30 / 0.19 ms at 2 / 4 / 8 ranks)
        relief_default = {2: "1/16", 4: "1/8", 8: "2/8"}.get(world, "1/8")
        row_relief, row_relief_den = (int(v) for v in os.environ.get("VT_ROOT_RELIEF_ROWS", relief_default).split("/"))
        r.fused_reduce_partition(True, row_relief, row_relief_den)
    elif fused and world > 1:
        # the root also sums the partial sums and encodes the frame (about two samples' worth of time): it traces fewer
        relief = float(os.environ.get("VT_ROOT_RELIEF_SPP", "2"))
        shares = [shard_samples_weighted(SPP, k, world, relief) for k in range(world)]
        first, stride, count = shares[rank]
        r.configure(width=WIDTH, height=HEIGHT, mode=abi.MODE_PATHS, flags=abi.FLAG_NO_HIT_RECORDS, spp=count,
                    bounces=BOUNCES, seed=SEED, sample_first=first, sample_stride=stride, total_spp=SPP, max_frames=0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    frames_pinned = [torch.empty((HEIGHT, WIDTH, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)]  # host-side frame buffers
    frame_host = frames_pinned[0].numpy()
    frames_host = [f.numpy() for f in frames_pinned]
    lib = abi.load()
    cam = {"P": P, "V": V}

    def trace_and_reduce():
        if fused:
            r.fused_reduce_next_frame()
            r.render_async(cam["P"], cam["V"])   # trace, push the partial sums into rank 0's memory, raise this rank's flag
            if not fused_flags:
                stream_barrier(flag)   # VT_FUSED_SYNC=0: order the ranks with a 4-byte NCCL all-reduce instead
            if rank == 0:
                r.resolve()        # waits for every rank's flag, sums the slots, encodes the frame
        elif world > 1:
            accum.zero_()
            r.render_async(cam["P"], cam["V"])
            reduce_accum(accum)
            r.resolve()
        else:
            r.render_frame_async(cam["P"], cam["V"])   # render_tick without the wait: clear + trace + resolve

    def step_resident():
        """One frame, everything device-resident, no host copies."""
        trace_and_reduce()

    def step_e2e():
        """One frame through the reference-facing ABI with host buffers, host waiting for it (render_tick + vt_read_color)."""
        r.update_instances_raw(inst)          # host matrices -> pinned staging -> device
        if world == 1 and not fused:
            assert r.render_tick_raw(cam["P"], cam["V"])    # projection/camera by host pointer; clear + trace + resolve
        else:
            trace_and_reduce()
        if rank == 0 or not fused:
            n = lib.vt_read_color(frame_host.ctypes.data, frame_host.nbytes)  # finished frame -> host
            assert n == frame_host.nbytes

    def timed_e2e_pipelined(steps):
        """The same per-step traffic (instance matrices host -> device, finished RGBA8 frame device -> pinned host), but the
        host pipelines its frames the way a renderer with two frames in flight does: frame k+1 is traced while the copy
        engine moves frame k over PCIe (vt_render_frame_async + vt_read_color_async, two colour buffers, two host buffers).
        Step k's interval [e0, e1] on the launching stream covers its upload, its trace / reduce / resolve and the rest of
        frame k-1's read-back (the stream waits for it before e1); the last read-back is timed on its own and added.  The L2
        flush between steps stays outside the intervals."""
        r.stats()
        reads = rank == 0 or not fused
        evs = []
        for k in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            r.update_instances_raw(inst)
            if world == 1 and not fused:
                r.render_frame_async(cam["P"], cam["V"])
            else:
                trace_and_reduce()
            if reads:
                r.read_color_fence()          # (device-side) frame k-1 has landed in host memory
            e1.record(stream)
            if reads:
                r.read_color_async(frames_host[k & 1])
            evs.append((e0, e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if reads:
            r.read_color_fence()
        e1.record(stream)
        evs.append((e0, e1))
        if reads:
            r.read_color_wait()
        torch.cuda.synchronize()
        st = r.stats()
        return [a.elapsed_time(b) for a, b in evs], st.rays_sum

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps."""
        ms, rays = [], 0
        r.stats()
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
        rays = st.rays_sum
        return ms, rays

    def timed_resident(steps):
        """K device-resident steps enqueued back to back (the host never waits inside the timed region, as
        a renderer that pipelines its frames would); each step is bracketed by its own CUDA events on the
        launching stream, with the L2 flushed in between; one synchronisation at the end.  The work counters
        are summed per frame by the library (vt_stats.*_sum), not extrapolated from one frame."""
        r.stats()  # folds everything so far; the sums restart here
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
        return ms, st.trace_ms_sum, st.rays_sum, st.iterations_sum, st.analytic_rays_sum

    def reduce_over_ranks(times, work):
        """max over ranks of the device times; sum over ranks of the work counters"""
        t = torch.tensor(times, dtype=torch.float64, device=dev)
        w = torch.tensor(work, dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(w, op=dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()], [int(x) for x in w.tolist()]

    steps = args.steps
    for _ in range(max(args.warmup, 3)):
        step_resident()
        step_e2e()
    timed_e2e_pipelined(max(args.warmup, 3))  # (untimed: the second colour buffer and the copy stream are created on first use)
    barrier()
    launches0 = r.stats().launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    ms, trace_ms_sum, rays, iters, analytic = timed_resident(steps)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = r.stats().launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    ms_e, rays_e = timed(step_e2e, steps)
    barrier()
    ms_p, rays_p = timed_e2e_pipelined(steps)
    if rank == 0 or not fused:
        assert np.array_equal(frames_host[0], frames_host[1]) and np.array_equal(frames_host[0], frame_host)  # same camera, same frame
    barrier()
    (t_res, t_e2e, t_e2e_pipe, t_trace), (rays_all, rays_e_all, rays_p_all, iters_all, analytic_all, launches_all) = reduce_over_ranks(
        [sum(ms), sum(ms_e), sum(ms_p), trace_ms_sum], [rays, rays_e, rays_p, iters, analytic, launches])

    # ---- parity of the frame the timed configuration produces (all ranks' samples), against the CPU oracle ----
    trace_and_reduce()
    got_accum = got_color = None
    if rank == 0:
        if accum is not None:
            r.synchronize()
            got_accum = accum.cpu().numpy().view(np.uint64)
        else:
            got_accum = r.read_accum()
        got_color = r.read_color()
    barrier()

    # ---- secondary: the same scene from a frame-filling camera (nearly every sample is a marched ray) ----
    cam["P"], cam["V"] = scenes.camera(WIDTH, HEIGHT, eye=CLOSEUP_EYE)
    for _ in range(3):
        step_resident()
    barrier()
    sec_steps = max(3, min(steps, 10))
    ms2, trace2_sum, rays2, iters2, analytic2 = timed_resident(sec_steps)
    barrier()
    (t2, t2_trace), (rays2_all, iters2_all, analytic2_all) = reduce_over_ranks([sum(ms2), trace2_sum], [rays2, iters2, analytic2])
    cam["P"], cam["V"] = P, V

    masks_in_smem = bool(r.stats().masks_in_smem)

    # ---- measured cache rooflines (plain streaming kernels of the library, untimed region) ----
    peaks = {}
    if rank == 0:
        for kind, name in ((0, "l2_read_gbs"), (1, "smem_read_gbs"), (2, "hbm_read_gbs")):
            v = C.c_double(0.0)
            if lib.vt_measure_peak(kind, C.byref(v)) == 0:
                peaks[name] = v.value

    # ---- the other BASELINE.json configurations ----
    if fused:
        r.fused_reduce_disable()
    if accum is not None:
        r.set_accum_buffer(None)
    hbm_peak, peak_src = measured_peaks()
    configs = [] if args.no_configs else run_other_configs(r, rank, world, stream, flush, reduce_over_ranks, barrier, hbm_peak, peaks, args)

    if rank == 0:
        value = rays_all / (t_res * 1e-3) / 1e6
        e2e_serial = rays_e_all / (t_e2e * 1e-3) / 1e6
        e2e_value = rays_p_all / (t_e2e_pipe * 1e-3) / 1e6
        traced = rays_all - analytic_all
        # roofline of the dominant kernel (trace_paths_wave_kernel), per launch on THIS rank:
        # algorithmic bytes = 4 B per DDA iteration (one RGBA8 voxel record, trace.frag:76) +
        # 16 B per pixel of accumulator read-modify-write
        alg_bytes = 4.0 * (iters / steps) + 16.0 * WIDTH * HEIGHT
        kernel_s = (trace_ms_sum / steps) * 1e-3
        achieved = alg_bytes / kernel_s / 1e9
        iters_per_s = (iters / steps) / kernel_s
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        issue_roof = sm_count * SMSP_PER_SM * SM_CLOCK_GHZ * 1e9 * 32 / STEP_INSTRUCTIONS
        cpu = parity = None
        if not args.no_cpu:
            threads = host_threads()
            v, dt, _, _, want = cpu_sample(threads, args.cpu_spp)
            if world == 1:
                cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                       "sample": f"{args.cpu_spp} of the {SPP} spp of the same 1080p frame (samples 0..{args.cpu_spp - 1}), "
                                 f"{dt:.2f} s wall on {threads} threads"}
            if args.cpu_spp == SPP:
                import oracle_lib
                want_color = oracle_lib.resolve(want, SPP)
                parity = {"exact": bool(np.array_equal(got_accum, want) and np.array_equal(got_color, want_color)),
                          "frame_fnv": fnv1a64(got_color.tobytes()), "oracle_frame_fnv": fnv1a64(want_color.tobytes()),
                          "accum_fnv": fnv1a64(got_accum.tobytes()), "oracle_accum_fnv": fnv1a64(want.tobytes()),
                          "what": f"full-size frame of all {world} rank(s): 3 x u64 radiance sums per pixel and the RGBA8 "
                                  "frame vs the CPU oracle, bit for bit (fnv = FNV-1a 64 of the buffer's SHA-256)"}
        if parity is None:
            parity = {"exact": None, "frame_fnv": fnv1a64(got_color.tobytes()), "accum_fnv": fnv1a64(got_accum.tobytes()),
                      "what": "oracle leg skipped (--no-cpu or --cpu-spp != 64): digests only"}
        kernel2_s = (trace2_sum / sec_steps) * 1e-3
        secondary = {
            "workload": f"same scene and settings, frame-filling camera eye={CLOSEUP_EYE}", "steps": sec_steps,
            "value": rays2_all / (t2 * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": t2 / sec_steps,
            "traced_rays_per_step": (rays2_all - analytic2_all) / sec_steps,
            "traced_mrays_per_s": (rays2_all - analytic2_all) / (t2 * 1e-3) / 1e6,
            "dda_iterations_per_s": (iters2 / sec_steps) / kernel2_s,
            "roofline_frac_hbm": (4.0 * (iters2 / sec_steps) + 16.0 * WIDTH * HEIGHT) / kernel2_s / 1e9 / hbm_peak,
            "issue_frac": (iters2 / sec_steps) / kernel2_s / issue_roof,
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": t_res / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": DATA,
            "config": CONFIG,
            "run": {"spp_per_rank": [c for _, _, c in shares], "partition": ((
                        f"rows of 8x4-pixel tiles dealt round-robin to {world} rank(s) (the root {row_relief_den - row_relief} for every {row_relief_den} of another rank), "
                        f"every rank tracing all {SPP} samples of its rows, "
                        if by_rows else f"spp sharded over {world} rank(s), ") + (
                        ("partial sums pushed into rank 0's memory over NVLink peer stores, ordered by " +
                         ("flags in peer memory" if fused_flags else "a 4-byte NCCL stream barrier")) if fused
                        else "NCCL all-reduce of 3*w*h int64")) if world > 1 else "single rank",
                    "masks_in_smem": masks_in_smem, "ms_per_step_min": min(ms), "ms_per_step_median": statistics.median(ms)},
            "traced_rays_per_step": traced / steps, "traced_mrays_per_s": traced / (t_res * 1e-3) / 1e6,
            "analytic_sky_samples_per_step": analytic_all / steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(inst.nbytes + 128),
                    "d2h_bytes_per_step": int(frame_host.nbytes + 32), "ms_per_step": t_e2e_pipe / steps,
                    "traced_mrays_per_s": (rays_p_all - analytic_all) / (t_e2e_pipe * 1e-3) / 1e6,
                    "how": "every step: instance matrices from host memory -> device, trace (+ cross-GPU accumulation), resolve, "
                           "RGBA8 frame -> pinned host memory; two frames in flight (vt_render_frame_async + vt_read_color_async: "
                           "frame k+1 is traced while the copy engine moves frame k)",
                    "host_waits_every_frame": {"value": e2e_serial, "ms_per_step": t_e2e / steps,
                                               "how": "render_tick + vt_read_color, the host blocking on each"}},
            "gpu_launches": launches_all,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": recorded_traffic(), "kernel": "trace_paths_wave_kernel", "kernel_ms": kernel_s * 1e3,
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "kernel_ms_max_over_ranks": t_trace / steps,
                         "dda_iterations_per_launch": iters / steps,
                         "dda_iterations_per_s": iters_per_s,
                         # the scene is shared-memory resident: the same algorithmic bytes against the measured cache peaks,
                         # and the voxel steps against the instruction-issue roof (the limit this kernel actually runs into)
                         "vs_l2_peak": {"peak": peaks.get("l2_read_gbs"), "frac": achieved / peaks["l2_read_gbs"] if peaks.get("l2_read_gbs") else None},
                         "vs_smem_peak": {"peak": peaks.get("smem_read_gbs"), "frac": achieved / peaks["smem_read_gbs"] if peaks.get("smem_read_gbs") else None},
                         "vs_issue_roof": {"peak_steps_per_s": issue_roof, "frac": iters_per_s / issue_roof,
                                           "how": f"{sm_count} SMs x {SMSP_PER_SM} issue slots x {SM_CLOCK_GHZ} GHz x 32 lanes / {STEP_INSTRUCTIONS} "
                                                  "SASS instructions per voxel step"},
                         "hbm_read_gbs_measured_here": peaks.get("hbm_read_gbs")},
            "parity": parity,
            "secondary": secondary,
            "configs": configs,
            "clocks": clocks,
            "wall_s": t_wall,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_other_configs(r, rank, world, stream, flush, reduce_over_ranks, barrier, hbm_peak, peaks, args):
    """configs[0], [1], [3] on rank 0 and configs[4] on every rank (ray ids sharded, no collective): a few frames each,
    timed on the device by the library's own CUDA events around the trace kernel."""
    import numpy as np

    import oracle_lib
    from tools import scenes
    from vtrace_b200 import abi

    out = []
    frames = max(3, min(args.steps, 5))

    def fresh():
        r.reset()
        r.set_stream(stream.cuda_stream)  # (the L2 flush is enqueued on this stream too: ordered with the frames)

    def timed_frames(P, V):
        ms = []
        for i in range(2 + frames):
            flush.fill_(1)
            assert r.render_tick_raw(P, V)
            st = r.stats()
            if i >= 2:
                ms.append(st.last_trace_ms)
        return float(np.median(ms)), st

    def entry(name, st, ms, n_rays_bytes, bytes_per_ray, resident, extra=None):
        t = ms * 1e-3
        alg = 4.0 * st.iterations + bytes_per_ray * n_rays_bytes
        e = {"workload": name, "trace_kernel_ms": ms, "rays": int(st.rays), "mrays_per_s": st.rays / t / 1e6,
             "giters_per_s": st.iterations / t / 1e9, "algorithmic_gb_per_s": alg / t / 1e9,
             "roofline_frac_hbm": alg / t / 1e9 / hbm_peak}
        if resident and peaks.get("l2_read_gbs"):
            e["roofline_frac_l2"] = alg / t / 1e9 / peaks["l2_read_gbs"]
        if extra:
            e.update(extra)
        return e

    # configs[0], [1]: the reference's own pass on the two assets, checked against the oracle on the spot
    if rank == 0:
        for name, asset, w, h in (("configs[0] Treasure.vox 640x480 primary", "Treasure", 640, 480),
                                  ("configs[1] AncientTemple.vox 1920x1080 primary", "AncientTemple", 1920, 1080)):
            fresh()
            chunk = scenes.load_asset(asset)
            r.add_texture(chunk)
            r.update_instances_raw(scenes.single_instance(0))
            P, V = scenes.camera(w, h)
            r.configure(width=w, height=h, mode=abi.MODE_PRIMARY, flags=0, spp=1, sample_first=0, sample_stride=1, total_spp=1, max_frames=0)
            ms, st = timed_frames(P, V)
            exact = None
            if not args.no_cpu:
                sc = oracle_lib.OracleScene()
                sc.add_texture(chunk.get_raw(), *chunk.dims())
                sc.set_instances(scenes.single_instance(0))
                want, wcolor, _, witers = sc.render_primary(P, V, w, h)
                got = r.read_hits()
                exact = bool(all(np.array_equal(got[f], want[f]) for f in ("hit_voxel", "packed", "instance", "iters"))
                             and np.array_equal(r.read_color(), wcolor) and st.iterations == witers)
            out.append(entry(name, st, ms, st.rays, 12.0, True, {"parity_exact": exact}))
        # configs[3]: 1024^3 heightmap, 4K, primary + shadow rays (parity: tests/test_gpu_parity.py::test_config3_*)
        fresh()
        r.add_volume_procedural(abi.VOLUME_HEIGHTMAP, 1024, 1024, 1024, 1)
        r.update_instances_raw(scenes.single_instance(0))
        P, V = scenes.camera(3840, 2160, eye=(0.9, -0.8, 0.9))
        r.configure(width=3840, height=2160, mode=abi.MODE_PRIMARY, flags=abi.FLAG_SHADOW_RAYS, spp=1, sample_first=0, sample_stride=1,
                    total_spp=1, max_frames=0)
        ms, st = timed_frames(P, V)
        out.append(entry("configs[3] heightmap 1024^3, 3840x2160 primary + shadow rays", st, ms, st.rays, 12.0, False))
    barrier()
    # configs[4]: 4096^3 sparse bricks, 2^26 incoherent rays PER GPU; rank g traces ray ids [g * 2^26, (g + 1) * 2^26)
    fresh()
    r.add_volume_procedural(abi.VOLUME_SPARSE_BRICKS, 4096, 4096, 4096, 2)
    r.update_instances_raw(scenes.single_instance(0))
    P, V = scenes.camera(8192, 8192)
    r.configure(width=8192, height=8192, mode=abi.MODE_RAYS, flags=0, spp=1, seed=3, sample_first=rank, sample_stride=1, total_spp=1, max_frames=0)
    barrier()
    ms, st = timed_frames(P, V)
    (t4,), (rays4, iters4) = reduce_over_ranks([ms], [st.rays, st.iterations])
    if rank == 0:
        t = t4 * 1e-3
        alg = 4.0 * iters4 + 16.0 * rays4
        out.append({"workload": f"configs[4] sparse 4096^3 bricks, 2^26 incoherent rays per GPU x {world} GPU(s) (ray ids sharded, no collective)",
                    "trace_kernel_ms": t4, "rays": rays4, "mrays_per_s": rays4 / t / 1e6, "giters_per_s": iters4 / t / 1e9,
                    "algorithmic_gb_per_s": alg / t / 1e9, "roofline_frac_hbm": alg / t / 1e9 / (hbm_peak * world), "n_gpus": world,
                    "scaling": "weak"})
    barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--cpu-spp", type=int, default=64, help="spp of the bounded cpu_baseline sample (64 = the whole frame; also enables the parity check)")
    ap.add_argument("--ref-spp", type=int, default=64, help="spp per step of the --impl reference arm (64 = the whole frame)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity legs")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE.json configurations")
    ap.add_argument("--force-fused", action="store_true", help="N=1 only: run the fused multi-GPU data path with one rank (profiling aid)")
    ap.add_argument("--reduce", default="fused", choices=["fused", "allreduce"], help="cross-GPU accumulation for N > 1")
    ap.add_argument("--partition", default="samples", choices=["rows", "samples"],
                    help="N > 1, fused: share a frame by samples (default; measured faster on this workload at 2, 4 and 8 GPUs) or by rows "
                         "of tiles (each rank traces all samples of its rows: faster on the close-up view)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
