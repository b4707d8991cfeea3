#!/usr/bin/env python
"""Runs one named configuration on the GPU a few times and prints device timings as JSON.
Used for profiling (ncu wraps this) and for the per-config tables in DESIGN.md; not the
headline benchmark (that is bench.py).

    python tools/run_config.py --config temple_primary [--closeup] [--frames 20]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from tools import scenes  # noqa: E402
from vtrace_b200 import abi  # noqa: E402
from vtrace_b200.renderer import Renderer  # noqa: E402

CLOSEUP_EYE = (0.8, -0.45, 0.6)

CONFIGS = {
    # name: (asset, width, height, mode, spp)
    "treasure_primary": ("Treasure", 640, 480, abi.MODE_PRIMARY, 0),          # configs[0]
    "temple_primary": ("AncientTemple", 1920, 1080, abi.MODE_PRIMARY, 0),     # configs[1]
    "temple_paths": ("AncientTemple", 1920, 1080, abi.MODE_PATHS, 64),        # configs[2]
    "treasure_paths": ("Treasure", 1920, 1080, abi.MODE_PATHS, 64),
    # extension scenes (procedural brick volumes)
    "heightmap_4k": ("heightmap1024", 3840, 2160, abi.MODE_PRIMARY, 0),        # configs[3]: primary + shadow rays
    "sparse_rays": ("sparse4096", 8192, 8192, abi.MODE_RAYS, 0),               # configs[4]: 2^26 incoherent rays
    "heightmap_paths": ("heightmap1024", 1920, 1080, abi.MODE_PATHS, 8),       # the configs[3] volume, path traced
    # the reference's default world (src/world.rs:143-198): 121 entities + the paged terrain chunks, 1000x1000 window
    "world_primary": ("world", 1000, 1000, abi.MODE_PRIMARY, 0),
    "world_paths": ("world", 1000, 1000, abi.MODE_PATHS, 8),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="temple_primary", choices=sorted(CONFIGS))
    ap.add_argument("--closeup", action="store_true")
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--spp", type=int, default=None)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--grid", action="store_true", help="11x11 entity grid of src/world.rs instead of one instance")
    args = ap.parse_args()
    asset, w, h, mode, spp = CONFIGS[args.config]
    if args.spp is not None:
        spp = args.spp
    with Renderer() as r:
        if args.grid:
            a = r.add_texture(scenes.load_asset("Treasure"))
            b = r.add_texture(scenes.load_asset("AncientTemple"))
            r.update_instances_raw(scenes.entity_grid(a, b))
            from vtrace_b200 import glm
            P = glm.perspective(glm.REFERENCE_FOV, np.float32(w) / np.float32(h), glm.REFERENCE_NEAR, glm.REFERENCE_FAR)
            V = glm.look_at((0.0, -2.0, 0.0), (3.0, -5.0, 2.0), (0.0, 1.0, 0.0))
        elif asset == "world":
            textures, inst = scenes.default_world()
            ids = [r.add_texture(c) for c in textures]
            from vtrace_b200 import glm
            r.update_instances_raw(np.stack([glm.with_texture_id(m, ids[t]).reshape(16) for m, t in inst]))
            P, V = scenes.camera(w, h, eye=(9.0, -9.0, 7.0), center=(0.0, -2.0, 0.0))
        elif asset == "heightmap1024":
            r.add_volume_procedural(abi.VOLUME_HEIGHTMAP, 1024, 1024, 1024, 1)
            r.update_instances_raw(scenes.single_instance(0))
            P, V = scenes.camera(w, h, eye=(0.9, -0.8, 0.9))
            if mode == abi.MODE_PRIMARY:
                args.flags |= abi.FLAG_SHADOW_RAYS
        elif asset == "sparse4096":
            r.add_volume_procedural(abi.VOLUME_SPARSE_BRICKS, 4096, 4096, 4096, 2)
            r.update_instances_raw(scenes.single_instance(0))
            P, V = scenes.camera(w, h)
        else:
            r.add_texture(scenes.load_asset(asset))
            r.update_instances_raw(scenes.single_instance(0))
            P, V = scenes.camera(w, h, eye=CLOSEUP_EYE if args.closeup else scenes.EYE)
        r.configure(width=w, height=h, mode=mode, flags=args.flags, spp=max(spp, 1), bounces=4, seed=3 if mode == abi.MODE_RAYS else 0x5EED,
                    sample_first=0, sample_stride=1, total_spp=max(spp, 1), max_frames=0)
        trace_ms, frame_ms = [], []
        for i in range(args.warmup + args.frames):
            assert r.render_tick_raw(P, V)
            st = r.stats()
            if i >= args.warmup:
                trace_ms.append(st.last_trace_ms)
                frame_ms.append(st.last_frame_ms)
        st = r.stats()
        t = float(np.median(trace_ms)) * 1e-3
        out = {
            "config": args.config, "closeup": args.closeup, "grid": args.grid, "width": w, "height": h, "spp": spp,
            "flags": args.flags, "masks_in_smem": int(st.masks_in_smem),
            "rays": int(st.rays), "iterations": int(st.iterations),
            "trace_ms_median": t * 1e3, "trace_ms_min": float(np.min(trace_ms)), "frame_ms_median": float(np.median(frame_ms)),
            "mrays_per_s": st.rays / t / 1e6, "giters_per_s": st.iterations / t / 1e9,
            "alg_gb_per_s": (4.0 * st.iterations + (16.0 if mode == abi.MODE_PATHS else 12.0) * (w * h if mode == abi.MODE_PATHS else st.rays)) / t / 1e9,
        }
        print(json.dumps(out))


if __name__ == "__main__":
    main()
