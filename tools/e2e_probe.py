#!/usr/bin/env python
"""Where the end-to-end frame time goes (one GPU): the bench's pipelined e2e loop with its parts switched off one at a time.
    python tools/e2e_probe.py            # variants: full, no upload, no read-back, no fence, no flush
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import scenes  # noqa: E402
from vtrace_b200 import abi  # noqa: E402
from vtrace_b200.renderer import Renderer  # noqa: E402

W, H, SPP = 1920, 1080, 64
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
r = Renderer()
r.add_texture(scenes.load_asset("AncientTemple"))
inst = scenes.single_instance(0)
r.update_instances_raw(inst)
P, V = scenes.camera(W, H)
r.configure(width=W, height=H, mode=abi.MODE_PATHS, flags=abi.FLAG_NO_HIT_RECORDS, spp=SPP, bounces=4, seed=0x5EED,
            sample_first=0, sample_stride=1, total_spp=SPP, max_frames=0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
r.set_stream(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
frames = [torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]


def run(steps, upload=True, read=True, fence=True, do_flush=True):
    evs = []
    for k in range(steps):
        if do_flush:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if upload:
            r.update_instances_raw(inst)
        r.render_frame_async(P, V)
        if read and fence:
            r.read_color_fence()
        e1.record(stream)
        if read:
            r.read_color_async(frames[k & 1])
        evs.append((e0, e1))
    if read:
        r.read_color_wait()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    return float(np.median(ms)), float(np.mean(ms))


for name, kw in (("full", {}), ("no upload", {"upload": False}), ("no read-back", {"read": False}), ("no fence", {"fence": False}),
                 ("no flush", {"do_flush": False}), ("no upload, no read-back", {"upload": False, "read": False})):
    run(5, **kw)
    med, mean = run(20, **kw)
    st = r.stats()
    print(f"{name:26s} median {med:.4f} ms  mean {mean:.4f} ms  trace kernel {st.last_trace_ms:.4f} ms  frame (library events) {st.last_frame_ms:.4f} ms")
r.close()
