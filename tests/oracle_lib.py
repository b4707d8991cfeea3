"""ctypes binding of the CPU oracle (oracle/vtrace_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (vtrace_b200/) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libvtrace_oracle.so")

VO_MISS = 0xFFFFFFFF
FLAG_VIEWPORT_H_IS_W = 1
FLAG_SHADOW_RAYS = 64
FLAG_SUBSET_8 = 0x10000
VOLUME_HEIGHTMAP, VOLUME_SPARSE_BRICKS = 1, 2

HIT_DTYPE = np.dtype([("hit_voxel", "<u4"), ("packed", "<u4"), ("instance", "<u4"), ("iters", "<u4")])


def build_oracle(force: bool = False) -> str:
    src = [os.path.join(ORACLE_DIR, f) for f in ("vtrace_oracle.c", "vtrace_oracle.h", "Makefile")]
    stale = force or not os.path.exists(ORACLE_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in src
    )
    if stale:
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return ORACLE_SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        vp, u32, i32, u64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64
        L.vo_scene_create.restype = vp
        L.vo_scene_destroy.argtypes = [vp]
        L.vo_add_texture.argtypes = [vp, vp, u32, u32, u32]
        L.vo_add_texture.restype = i32
        L.vo_add_volume_procedural.argtypes = [vp, u32, u32, u32, u32, u32]
        L.vo_add_volume_procedural.restype = i32
        L.vo_add_volume_bricks.argtypes = [vp, vp, vp, vp, u64, u32, u32, u32]
        L.vo_add_volume_bricks.restype = i32
        L.vo_last_shadow_rays.restype = u64
        L.vo_render_rays.argtypes = [vp, u64, u64, u32, vp, vp, C.c_int]
        L.vo_render_rays.restype = u64
        L.vo_set_instances.argtypes = [vp, vp, u32]
        L.vo_render_primary.argtypes = [vp, vp, vp, C.c_int, C.c_int, u32, vp, vp, vp, C.c_int]
        L.vo_render_primary.restype = u64
        L.vo_render_paths.argtypes = [vp, vp, vp, C.c_int, C.c_int, u32, u32, u32, u32, u32, u32, vp, vp, C.c_int]
        L.vo_resolve.argtypes = [vp, C.c_int, C.c_int, u32, vp]
        L.vo_frag_main.argtypes = [vp, vp, vp, vp, vp, vp, u32, u32, u32, vp, vp, vp]
        L.vo_mat4_inverse.argtypes = [vp, vp]
        L.vo_mat4_mul.argtypes = [vp, vp, vp]
        L.vo_load_vox.argtypes = [C.c_char_p, vp, vp, u64]
        L.vo_load_vox.restype = C.c_int
        L.vo_load_vox_model.argtypes = [C.c_char_p, C.c_uint32, vp, vp, u64]
        L.vo_load_vox_model.restype = C.c_int
        L.vo_srgb_decode.argtypes = [C.c_uint8]
        L.vo_srgb_decode.restype = C.c_float
        L.vo_srgb_encode.argtypes = [C.c_float]
        L.vo_srgb_encode.restype = C.c_uint8
        L.vo_read_texels.argtypes = [vp, u32, u32, u32, u32, u32, u32, u32, vp]
        L.vo_read_texels.restype = C.c_int
        L.vo_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _m(a) -> np.ndarray:
    """mat4 -> 16 contiguous float32, column-major (as glm-rs stores it)."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(16))


class OracleScene:
    """Holds what crosses the reference's C ABI: textures + instance matrices."""

    def __init__(self):
        self._h = lib().vo_scene_create()

    def close(self):
        if self._h:
            lib().vo_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_texture(self, rgba: np.ndarray, w: int, h: int, d: int) -> int:
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        assert rgba.size == 4 * w * h * d
        return lib().vo_add_texture(self._h, _p(rgba), w, h, d)

    def add_volume_procedural(self, kind: int, w: int, h: int, d: int, seed: int) -> int:
        return lib().vo_add_volume_procedural(self._h, kind, w, h, d, seed)

    def add_volume_bricks(self, coords, masks, colors, w: int, h: int, d: int) -> int:
        coords = np.ascontiguousarray(coords, dtype=np.uint32).reshape(-1, 3)
        masks = np.ascontiguousarray(masks, dtype=np.uint32).reshape(-1, 16)
        colors = np.ascontiguousarray(colors, dtype=np.uint8).reshape(-1, 4)
        return lib().vo_add_volume_bricks(self._h, _p(coords), _p(masks), _p(colors), len(coords), w, h, d)

    def render_rays(self, n: int, seed: int, first: int = 0, threads=0, want_color=True):
        rec = np.empty(n, dtype=HIT_DTYPE)
        rgba = np.empty((n, 4), dtype=np.uint8) if want_color else None
        iters = lib().vo_render_rays(self._h, n, first, seed, _p(rec), _p(rgba), threads)
        return rec, rgba, int(iters)

    def read_texels(self, tex: int, box=None, dims=None) -> np.ndarray:
        """RGBA8 texels of volume `tex` (any kind), x fastest; box = (x0, y0, z0, nx, ny, nz) or the whole volume (dims)."""
        x0, y0, z0, nx, ny, nz = box if box is not None else (0, 0, 0, *dims)
        out = np.empty(4 * nx * ny * nz, dtype=np.uint8)
        if lib().vo_read_texels(self._h, tex, x0, y0, z0, nx, ny, nz, _p(out)) != 0:
            raise ValueError("vo_read_texels: bad texture id or box")
        return out

    def last_shadow_rays(self) -> int:
        return int(lib().vo_last_shadow_rays())

    def set_instances(self, mats: np.ndarray):
        mats = np.ascontiguousarray(mats, dtype=np.float32).reshape(-1, 16)
        lib().vo_set_instances(self._h, _p(mats) if len(mats) else None, len(mats))

    def render_primary(self, P, V, width, height, flags=0, threads=0, want_color=True, want_depth=False):
        rec = np.zeros(width * height, dtype=HIT_DTYPE)
        rgba = np.zeros((height, width, 4), dtype=np.uint8) if want_color else None
        depth = np.zeros((height, width), dtype=np.float32) if want_depth else None
        P, V = _m(P), _m(V)
        iters = lib().vo_render_primary(self._h, _p(P), _p(V), width, height, flags, _p(rec), _p(rgba), _p(depth), threads)
        return rec.reshape(height, width), rgba, depth, int(iters)

    def render_paths(self, P, V, width, height, spp, bounces=4, seed=0x5EED, flags=0, sample_first=0,
                     sample_stride=1, accum=None, threads=0):
        if accum is None:
            accum = np.zeros((height, width, 3), dtype=np.uint64)
        stats = np.zeros(2, dtype=np.uint64)
        P, V = _m(P), _m(V)
        lib().vo_render_paths(self._h, _p(P), _p(V), width, height, flags, bounces, seed, sample_first,
                              sample_stride, spp, _p(accum), _p(stats), threads)
        return accum, int(stats[0]), int(stats[1])


def resolve(accum: np.ndarray, total_spp: int) -> np.ndarray:
    h, w, _ = accum.shape
    out = np.empty((h, w, 4), dtype=np.uint8)
    lib().vo_resolve(_p(np.ascontiguousarray(accum)), w, h, total_spp, _p(out))
    return out


def frag_main(P, V, M, sp, mp, rgba, w, h, d):
    out = np.zeros(6, dtype=np.int32)
    color = np.zeros(4, dtype=np.float32)
    depth = np.zeros(1, dtype=np.float32)
    P, V, M = _m(P), _m(V), _m(M)
    sp = np.ascontiguousarray(sp, dtype=np.float32)
    mp = np.ascontiguousarray(mp, dtype=np.float32)
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    lib().vo_frag_main(_p(P), _p(V), _p(M), _p(sp), _p(mp), _p(rgba), w, h, d, _p(out), _p(color), _p(depth))
    return out, color, float(depth[0])


def load_vox_model(path: str, model: int):
    """(raw RGBA bytes as n x 4, dims) of one model of a .vox file; None when the file has fewer models."""
    dims = np.zeros(3, dtype=np.uint32)
    rc = lib().vo_load_vox_model(path.encode(), model, _p(dims), None, 0)
    if rc == -6:
        return None
    if rc != 0:
        raise RuntimeError(f"vo_load_vox_model({path}, {model}) failed: {rc}")
    out = np.zeros(4 * int(dims.prod()), dtype=np.uint8)
    rc = lib().vo_load_vox_model(path.encode(), model, _p(dims), _p(out), out.size)
    if rc != 0:
        raise RuntimeError(f"vo_load_vox_model({path}, {model}) failed: {rc}")
    return out.reshape(-1, 4), tuple(int(d) for d in dims)


def load_vox(path: str):
    dims = np.zeros(3, dtype=np.uint32)
    rc = lib().vo_load_vox(path.encode(), _p(dims), None, 0)
    if rc != 0:
        raise RuntimeError(f"vo_load_vox({path}) failed: {rc}")
    out = np.empty(4 * int(dims[0]) * int(dims[1]) * int(dims[2]), dtype=np.uint8)
    rc = lib().vo_load_vox(path.encode(), _p(dims), _p(out), out.size)
    if rc != 0:
        raise RuntimeError(f"vo_load_vox({path}) failed: {rc}")
    return out, tuple(int(x) for x in dims)


def mat4_inverse(m) -> np.ndarray:
    out = np.empty(16, dtype=np.float32)
    m = _m(m)
    lib().vo_mat4_inverse(_p(m), _p(out))
    return out.reshape(4, 4)


def max_threads() -> int:
    return int(lib().vo_max_threads())
