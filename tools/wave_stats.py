#!/usr/bin/env python
"""Per-phase counters of the wavefront kernel on the bench workload (needs a -DVT_WAVE_STATS variant build):
    python -m vtrace_b200.build --variant stats -DVT_WAVE_STATS
    VT_LIBRENDER=$PWD/variants/stats/librender.so python tools/wave_stats.py [--closeup]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from tools import scenes  # noqa: E402
from vtrace_b200 import abi  # noqa: E402
from vtrace_b200.renderer import Renderer  # noqa: E402

w, h, spp = 1920, 1080, int(os.environ.get("SPP", "64"))
eye = (0.8, -0.45, 0.6) if "--closeup" in sys.argv else (1.6, -0.9, 1.2)
with Renderer() as r:
    r.add_texture(scenes.load_asset("AncientTemple"))
    r.update_instances_raw(scenes.single_instance(0))
    P, V = scenes.camera(w, h, eye=eye)
    r.configure(width=w, height=h, mode=abi.MODE_PATHS, flags=0, spp=spp, bounces=4, seed=0x5EED, sample_first=0, sample_stride=1,
                total_spp=spp, max_frames=0)
    lib = abi.load()
    fn = lib.vt_debug_wave_stats
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p]
    out = np.zeros(32, dtype=np.uint64)
    ft = lib.vt_debug_wave_times
    ft.restype, ft.argtypes = C.c_int, [C.c_void_p]
    times = np.zeros(192, dtype=np.uint32)
    rows_as = os.environ.get("VT_FUSED_ROWS_AS_WORLD")  # one rank of an N-rank job that shares frames by tile rows

    def frame():
        if rows_as:
            r.fused_reduce_next_frame()
            r.render_async(P, V)
            r.synchronize()
        else:
            assert r.render_tick_raw(P, V)

    if rows_as:
        r.fused_reduce_export(1)
        r.fused_reduce_partition(True)
    frame()
    fn(out.ctypes.data)  # (drop the first frame)
    ft(times.ctypes.data)
    frame()
    assert fn(out.ctypes.data) == 0
    s = [int(x) for x in out]
    st = r.stats()
    print(f"rays {st.rays - st.analytic_rays}  iterations {st.iterations}")
    for name, i in (("primary", 0), ("bounce", 2), ("sky", 4)):
        print(f"{name:8s} phases {s[i]:9d}  lanes/phase {s[i + 1] / max(s[i], 1):6.2f}")
    print(f"march    phases {s[6]:9d}  ready rays/phase {s[7] / max(s[6], 1):6.2f}  walk calls {s[8]:9d} ({s[8] / max(s[6], 1):.2f}/phase)"
          f"  stopped lanes/call {s[9] / max(s[8], 1):6.2f}  lanes with a ray/call {s[11] / max(s[8], 1):6.2f}  parked/phase {s[10] / max(s[6], 1):6.2f}")
    tot = sum(s[12:21])
    print("ray length histogram (fast-path rays that ended in a march):", ", ".join(
        f"{n}: {100 * s[12 + k] / max(tot, 1):.1f}%" for k, n in enumerate(["0", "1", "2", "3", "4-7", "8-15", "16-31", "32-63", "64+"])))
    assert ft(times.ctypes.data) == 0
    for name, h in (("first item claimed", times[128:192]), ("work exhausted", times[0:64]), ("warp finished", times[64:128])):
        nz = np.nonzero(h)[0]
        print(f"{name:18s} (5 us buckets since the warp started, warps):", " ".join(f"{5 * b}us:{h[b]}" for b in nz))
    print("kernel ms", r.stats().last_trace_ms)
