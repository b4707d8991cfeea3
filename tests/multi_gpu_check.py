#!/usr/bin/env python
"""Multi-GPU exactness check, run under torchrun on N GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank traces its shard of the samples.  The frame must be bit-identical (a) to rank 0 tracing all
samples alone, (b) with the NCCL all-reduce of the accumulators, (c) with the fused accumulation into the
root's buffer over NVLink peer stores ordered by flags (vt_fused_reduce_*), over several frames (double
buffering), (d) the same ordered by an external stream barrier.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools import scenes  # noqa: E402
from vtrace_b200 import abi  # noqa: E402
from vtrace_b200.distributed import reduce_accum, setup_fused_reduce, shard_samples, stream_barrier  # noqa: E402
from vtrace_b200.renderer import Renderer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    w, h, spp = 640, 360, 16
    chunk = scenes.load_asset("AncientTemple")
    P, V = scenes.camera(w, h, eye=(0.8, -0.45, 0.6))
    r = Renderer()
    r.add_texture(chunk)
    r.update_instances_raw(scenes.single_instance(0))
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    r.set_stream(stream.cuda_stream)

    # (a) the whole frame on one rank
    whole = None
    if rank == 0:
        r.configure(width=w, height=h, mode=abi.MODE_PATHS, spp=spp, bounces=4, seed=0x5EED, sample_first=0, sample_stride=1, total_spp=spp)
        assert r.render_tick_raw(P, V)
        whole, whole_color = r.read_accum(), r.read_color()

    # (b) sharded + NCCL all-reduce
    first, stride, count = shard_samples(spp, rank, world)
    r.configure(width=w, height=h, mode=abi.MODE_PATHS, spp=count, bounces=4, seed=0x5EED, sample_first=first,
                sample_stride=stride, total_spp=spp)
    accum = torch.zeros((h, w, 3), dtype=torch.int64, device=dev)
    r.set_accum_buffer(accum.data_ptr())
    r.render_async(P, V)
    reduce_accum(accum)
    r.resolve()
    r.synchronize()
    got = accum.cpu().numpy().view(np.uint64)
    if rank == 0:
        assert np.array_equal(got, whole), "all-reduce path differs from the single-rank frame"
        assert np.array_equal(r.read_color(), whole_color)
    r.set_accum_buffer(None)

    # (c) fused accumulation: every rank pushes its partial sums into rank 0's memory over NVLink and raises a
    # flag there; the root waits for the flags.  Six frames, no host synchronisation and no collective in
    # between: both halves of the double buffer get reused, so the "consumed" flags are exercised too.
    os.environ["VT_FUSED_SYNC"] = "1"
    assert setup_fused_reduce(r, rank, world, dev), "fused accumulation could not be set up (no peer access?)"
    for frame in range(6):
        r.fused_reduce_next_frame()
        r.render_async(P, V)
        if rank == 0:
            fused = r.read_accum()
            assert np.array_equal(fused, whole), f"fused accumulation differs from the single-rank frame (frame {frame})"
            r.resolve()
            assert np.array_equal(r.read_color(), whole_color)
    r.synchronize()
    dist.barrier()
    # (d) the same with the ranks ordered by an external stream barrier instead of the flags
    r.fused_reduce_disable()
    os.environ["VT_FUSED_SYNC"] = "0"
    assert setup_fused_reduce(r, rank, world, dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    for frame in range(3):
        r.fused_reduce_next_frame()
        r.render_async(P, V)
        stream_barrier(flag)
        if rank == 0:
            fused = r.read_accum()
            assert np.array_equal(fused, whole), f"fused accumulation (external barrier) differs (frame {frame})"
        r.synchronize()
    dist.barrier()
    # (e) more than 255 samples: the partial sums no longer fit 3 x u32 and travel in the wide layout
    r.fused_reduce_disable()
    os.environ["VT_FUSED_SYNC"] = "1"
    w2, h2, spp2 = 320, 180, 264
    P2, V2 = scenes.camera(w2, h2, eye=(0.8, -0.45, 0.6))
    if rank == 0:
        r.configure(width=w2, height=h2, mode=abi.MODE_PATHS, spp=spp2, bounces=4, seed=0x5EED, sample_first=0, sample_stride=1, total_spp=spp2)
        assert r.render_tick_raw(P2, V2)
        whole2 = r.read_accum()
    first, stride, count = shard_samples(spp2, rank, world)
    r.configure(width=w2, height=h2, mode=abi.MODE_PATHS, spp=count, bounces=4, seed=0x5EED, sample_first=first,
                sample_stride=stride, total_spp=spp2)
    assert setup_fused_reduce(r, rank, world, dev)
    for frame in range(3):
        r.fused_reduce_next_frame()
        r.render_async(P2, V2)
        if rank == 0:
            assert np.array_equal(r.read_accum(), whole2), f"fused accumulation, wide layout, differs (frame {frame})"
    r.synchronize()
    dist.barrier()
    # (f) the frame shared out by rows of tiles instead of by samples (vt_fused_reduce_partition): every rank traces all
    # samples of the tile rows ty = rank (mod world); wide layout first (264 samples), then the compact one with the
    # camera moving from frame to frame (the screen rectangle, and with it the first owned row, changes)
    r.fused_reduce_disable()
    r.configure(width=w2, height=h2, mode=abi.MODE_PATHS, spp=spp2, bounces=4, seed=0x5EED, sample_first=0, sample_stride=1, total_spp=spp2)
    assert setup_fused_reduce(r, rank, world, dev)
    r.fused_reduce_partition(True, 1, 16)
    for frame in range(3):
        r.fused_reduce_next_frame()
        r.render_async(P2, V2)
        if rank == 0:
            assert np.array_equal(r.read_accum(), whole2), f"fused accumulation by tile rows, wide layout, differs (frame {frame})"
    r.synchronize()
    dist.barrier()
    r.fused_reduce_disable()
    eyes = [(0.8, -0.45, 0.6), (1.6, -0.9, 1.2), (2.4, -1.3, 1.9), (0.5, -0.2, 0.35), (1.1, 0.4, -0.9)]
    wholes = []
    if rank == 0:
        r.configure(width=w, height=h, mode=abi.MODE_PATHS, spp=spp, bounces=4, seed=0x5EED, sample_first=0, sample_stride=1, total_spp=spp)
        for eye in eyes:
            Pm, Vm = scenes.camera(w, h, eye=eye)
            assert r.render_tick_raw(Pm, Vm)
            wholes.append((r.read_accum(), r.read_color()))
    r.configure(width=w, height=h, mode=abi.MODE_PATHS, spp=spp, bounces=4, seed=0x5EED, sample_first=0, sample_stride=1, total_spp=spp)
    assert setup_fused_reduce(r, rank, world, dev)
    r.fused_reduce_partition(True, 3, 8)   # the root is dealt 5 tile rows for every 8 of another rank
    rays_rows = 0
    for frame, eye in enumerate(eyes + eyes):
        Pm, Vm = scenes.camera(w, h, eye=eye)
        r.fused_reduce_next_frame()
        r.render_async(Pm, Vm)
        if rank == 0:
            r.resolve()
            for name, got, want in (("colour", r.read_color(), wholes[frame % len(eyes)][1]), ("sums", r.read_accum(), wholes[frame % len(eyes)][0])):
                if not np.array_equal(got, want):
                    bad = np.argwhere((got != want).reshape(h, w, -1).any(axis=2))
                    raise AssertionError(f"fused accumulation by tile rows: {name} differ (frame {frame}) at {len(bad)} pixels, rows "
                                         f"{bad[:, 0].min()}..{bad[:, 0].max()}, columns {bad[:, 1].min()}..{bad[:, 1].max()}; first {bad[0]}: "
                                         f"{got.reshape(h, w, -1)[bad[0][0], bad[0][1]]} != {want.reshape(h, w, -1)[bad[0][0], bad[0][1]]}")
        r.synchronize()
        if frame == 0:
            rays_rows = r.stats().rays
    # every path of the frame is traced by exactly one rank: the ray counters add up to the single-rank frame's
    t = torch.tensor([rays_rows], dtype=torch.int64, device=dev)
    dist.all_reduce(t)
    if rank == 0:
        r.fused_reduce_disable()
        r.configure(width=w, height=h, mode=abi.MODE_PATHS, spp=spp, bounces=4, seed=0x5EED, sample_first=0, sample_stride=1, total_spp=spp)
        Pm, Vm = scenes.camera(w, h, eye=eyes[0])
        assert r.render_tick_raw(Pm, Vm)
        assert r.stats().rays == int(t.item()), f"rays over the ranks {int(t.item())} != single-rank frame {r.stats().rays}"
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU check ok: {world} ranks, all-reduce and fused NVLink accumulation (flags, external barrier, by samples and by tile rows) bit-identical to one rank")
    r.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
