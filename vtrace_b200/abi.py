"""ctypes binding of librender.so — the same symbols src/render.rs:110-128 declares in its
`extern "C"` block, plus the vt_* headless extensions (include/vtrace_abi.h).

There is no fallback: if the CUDA library is missing or cannot initialise, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

VT_MISS = 0xFFFFFFFF
MODE_PRIMARY, MODE_PATHS, MODE_RAYS = 0, 1, 2
VOLUME_HEIGHTMAP, VOLUME_SPARSE_BRICKS = 1, 2
FLAG_VIEWPORT_H_IS_W, FLAG_NO_HIT_RECORDS, FLAG_FORCE_GLOBAL_MASKS, FLAG_PER_PIXEL_PATHS, FLAG_NO_BINNING, FLAG_SHADOW_RAYS = 1, 2, 4, 16, 32, 64

HIT_DTYPE = np.dtype([("hit_voxel", "<u4"), ("packed", "<u4"), ("instance", "<u4"), ("iters", "<u4")])


class RenderTickInfo(C.Structure):  # src/render.rs:177-181, lib/common.h:89-92
    _fields_ = [("perspective", C.c_void_p), ("camera", C.c_void_p)]


class UserInput(C.Structure):  # src/render.rs:37-51, lib/common.h:145-151
    _fields_ = [("keys", C.c_uint8 * 6), ("mouse_x", C.c_double), ("mouse_y", C.c_double),
                ("last_mouse_x", C.c_double), ("last_mouse_y", C.c_double)]


class VtConfig(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("mode", C.c_uint32), ("flags", C.c_uint32),
                ("spp", C.c_uint32), ("bounces", C.c_uint32), ("seed", C.c_uint32), ("sample_first", C.c_uint32),
                ("sample_stride", C.c_uint32), ("total_spp", C.c_uint32), ("max_frames", C.c_int32),
                ("device", C.c_int32)]


class VtStats(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("rays", C.c_uint64), ("iterations", C.c_uint64), ("launches", C.c_uint64),
                ("last_trace_ms", C.c_float), ("last_frame_ms", C.c_float), ("masks_in_smem", C.c_uint32),
                ("trace_frames", C.c_uint32), ("trace_ms_sum", C.c_float), ("bin_list_grown", C.c_uint32),
                ("analytic_rays", C.c_uint64), ("rays_sum", C.c_uint64), ("iterations_sum", C.c_uint64),
                ("analytic_rays_sum", C.c_uint64)]


# every symbol include/vtrace_abi.h declares: name -> (restype, argtypes)
_vp, _u32, _i32, _i64, _sz = C.c_void_p, C.c_uint32, C.c_int32, C.c_int64, C.c_size_t
SYMBOLS = {
    # Part 1 — the reference's FFI
    "entry": (C.c_uint64, []),
    "render_tick": (_i32, [C.POINTER(_i32), C.POINTER(_i32), C.POINTER(RenderTickInfo)]),
    "get_input_data_pointer": (C.POINTER(UserInput), []),
    "add_texture": (_i32, [_vp, _u32, _u32, _u32]),
    "vt_add_volume_procedural": (_i32, [_u32, _u32, _u32, _u32, _u32]),
    "vt_add_volume_bricks": (_i32, [_vp, _vp, _vp, _sz, _u32, _u32, _u32]),
    "start_update_instances": (_vp, [_u32]),
    "end_update_instances": (_i32, [_u32]),
    "cleanup": (None, []),
    # Part 2 — extensions
    "vt_get_config": (_i32, [C.POINTER(VtConfig)]),
    "vt_configure": (_i32, [C.POINTER(VtConfig)]),
    "vt_render_async": (_i32, [_vp, _vp]),
    "vt_render_frame_async": (_i32, [_vp, _vp]),
    "vt_synchronize": (_i32, []),
    "vt_read_hits": (_i64, [_vp, _sz]),
    "vt_read_color": (_i64, [_vp, _sz]),
    "vt_read_color_async": (_i64, [_vp, _sz]),
    "vt_read_color_wait": (_i32, []),
    "vt_read_color_fence": (_i32, []),
    "vt_read_color_bgra": (_i64, [_vp, _sz]),
    "vt_write_ppm": (_i32, [C.c_char_p]),
    "vt_read_depth": (_i64, [_vp, _sz]),
    "vt_read_accum": (_i64, [_vp, _sz]),
    "vt_accum_device_ptr": (_vp, []),
    "vt_set_accum_buffer": (_i32, [_vp]),
    "vt_clear_accum": (_i32, []),
    "vt_resolve": (_i32, []),
    "vt_fused_reduce_export": (_i32, [_vp, _u32]),
    "vt_fused_reduce_import": (_i32, [_vp, _u32, _u32]),
    "vt_fused_reduce_next_frame": (_i32, []),
    "vt_fused_reduce_disable": (_i32, []),
    "vt_fused_reduce_partition": (_i32, [_u32, _u32, _u32]),
    "vt_set_stream": (_i32, [_vp]),
    "vt_get_stats": (_i32, [C.POINTER(VtStats)]),
    "vt_measure_peak": (_i32, [_u32, C.POINTER(C.c_double)]),
    "vt_set_user_input": (_i32, [C.POINTER(UserInput)]),
    "vt_last_error": (C.c_char_p, []),
}

_lib = None


def library_path() -> str:
    # VT_LIBRENDER: a side-by-side build of the same sources (python -m vtrace_b200.build --variant ...), for A/B timing
    return os.environ.get("VT_LIBRENDER") or _build.LIB


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen librender.so and type every symbol.  Does not touch the GPU."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            if not build_if_missing:
                raise FileNotFoundError(path)
            _build.build()
        lib = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def last_error() -> str:
    return (load().vt_last_error() or b"").decode()
