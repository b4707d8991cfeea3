"""Host-side mirror of the reference's render facade (src/render.rs) over librender.so.

Same names and call order as the Rust engine: `Renderer()` calls `entry()` and raises where
Rust panics; `update_instances(scene)` is start/end_update_instances; `render_tick(pos, dir,
queue)` pops at most ONE queued texture, uploads it with add_texture, renders a frame, then
rebuilds perspective/camera for the next frame (so frame k uses frame k-1's pose,
src/render.rs:293); dropping the renderer calls `cleanup()`.

The headless extras (configure / read_* / stats) wrap the vt_* extension symbols.
"""
from __future__ import annotations

import ctypes as C
from collections import deque

import numpy as np

from . import abi, glm


class TextureUploadQueue:
    """src/render.rs:150-175: FIFO of (chunk, handle); handles count up from 0."""

    def __init__(self):
        self.queue = deque()
        self.num_textures_added = 0

    def add_texture(self, texture) -> int:
        handle = self.num_textures_added
        self.queue.append((texture, handle))
        self.num_textures_added += 1
        return handle

    def pop(self):
        return self.queue.popleft() if self.queue else None


class Renderer:
    def __init__(self):
        self._lib = abi.load()
        code = self._lib.entry()  # src/render.rs:184-188
        if code != 0:
            raise RuntimeError(f"ERROR: renderer initialization failed (code {code:#x}): {abi.last_error()}")
        self._alive = True
        self.window_width = 1
        self.window_height = 1
        self.prev_window_width = 1
        self.prev_window_height = 1
        self.fov = glm.REFERENCE_FOV
        self.frame_num = 0
        self.perspective = self.create_perspective(self.fov, 1.0)                 # src/render.rs:199
        self.camera = self.create_camera((0.0, 0.0, 0.0), (1.0, 0.0, 0.0))        # src/render.rs:200
        self.texture_handle_lookup: dict[int, int] = {}

    # -- src/render.rs:208-214 ---------------------------------------------------------------
    @staticmethod
    def create_perspective(fov, aspect):
        return glm.perspective(fov, aspect, glm.REFERENCE_NEAR, glm.REFERENCE_FAR)

    @staticmethod
    def create_camera(position, direction):
        position = np.asarray(position, dtype=np.float32)
        return glm.look_at(position, position + np.asarray(direction, dtype=np.float32), (0.0, 1.0, 0.0))

    def get_input_data_pointer(self):
        return self._lib.get_input_data_pointer()

    # -- src/render.rs:220-238 ---------------------------------------------------------------
    def update_instances(self, scene):
        """scene: iterable of (model mat4, texture_handle), e.g. a flattened SceneGraph."""
        scene = list(scene)
        ptr = self._lib.start_update_instances(len(scene))
        if not ptr:
            raise RuntimeError("ERROR: Updating instances failed: " + abi.last_error())
        n_upper = max(len(scene), 1)
        dst = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n_upper, 16))
        true_count = 0
        for model, handle in scene:
            tex = self.texture_handle_lookup.get(handle)
            if tex is None:
                continue  # texture not resident yet: the instance is skipped (src/render.rs:227-233)
            dst[true_count] = glm.with_texture_id(model, tex).reshape(16)
            true_count += 1
        if self._lib.end_update_instances(true_count) != 0:
            raise RuntimeError("ERROR: Updating instances failed: " + abi.last_error())

    def update_instances_raw(self, mats: np.ndarray):
        """Instances already packed as GPUInstance (id in [3][3]); n x 16 float32."""
        mats = np.ascontiguousarray(mats, dtype=np.float32).reshape(-1, 16)
        ptr = self._lib.start_update_instances(len(mats))
        if not ptr:
            raise RuntimeError("ERROR: Updating instances failed: " + abi.last_error())
        if len(mats):
            C.memmove(ptr, mats.ctypes.data, mats.nbytes)
        if self._lib.end_update_instances(len(mats)) != 0:
            raise RuntimeError("ERROR: Updating instances failed: " + abi.last_error())

    # -- src/render.rs:240-313 ---------------------------------------------------------------
    def render_tick(self, pos, direction, texture_upload_queue: TextureUploadQueue | None = None):
        if texture_upload_queue is not None:
            item = texture_upload_queue.pop()
            if item is not None:
                texture, handle = item
                w, h, d = texture.dims()
                assert w > 0 and h > 0 and d > 0
                raw = np.ascontiguousarray(texture.get_raw(), dtype=np.uint8)
                tex_id = self._lib.add_texture(raw.ctypes.data, w, h, d)
                if tex_id < 0:
                    raise RuntimeError("ERROR: Adding texture failed: " + abi.last_error())
                self.texture_handle_lookup[handle] = tex_id
        ok = self.render_tick_raw(self.perspective, self.camera)
        if self.window_width != self.prev_window_width or self.window_height != self.prev_window_height:
            self.perspective = self.create_perspective(self.fov, np.float32(self.window_width) / np.float32(self.window_height))
            self.prev_window_width, self.prev_window_height = self.window_width, self.window_height
        self.camera = self.create_camera(pos, direction)
        self.frame_num += 1
        return ok

    def render_tick_raw(self, perspective, camera) -> bool:
        """The bare FFI call with caller-supplied matrices (host pointers)."""
        P = np.ascontiguousarray(perspective, dtype=np.float32).reshape(16)
        V = np.ascontiguousarray(camera, dtype=np.float32).reshape(16)
        info = abi.RenderTickInfo(P.ctypes.data, V.ctypes.data)
        w, h = C.c_int32(self.window_width), C.c_int32(self.window_height)
        code = self._lib.render_tick(C.byref(w), C.byref(h), C.byref(info))
        self.window_width, self.window_height = w.value, h.value
        return code == 0

    def add_texture(self, chunk) -> int:
        w, h, d = chunk.dims()
        raw = np.ascontiguousarray(chunk.get_raw(), dtype=np.uint8)
        tex_id = self._lib.add_texture(raw.ctypes.data, w, h, d)
        if tex_id < 0:
            raise RuntimeError("ERROR: Adding texture failed: " + abi.last_error())
        return tex_id

    def add_volume_procedural(self, kind: int, w: int, h: int, d: int, seed: int) -> int:
        tex_id = self._lib.vt_add_volume_procedural(kind, w, h, d, seed)
        if tex_id < 0:
            raise RuntimeError("ERROR: Adding procedural volume failed: " + abi.last_error())
        return tex_id

    def add_volume_bricks(self, coords, masks, colors, w: int, h: int, d: int) -> int:
        coords = np.ascontiguousarray(coords, dtype=np.uint32).reshape(-1, 3)
        masks = np.ascontiguousarray(masks, dtype=np.uint32).reshape(-1, 16)
        colors = np.ascontiguousarray(colors, dtype=np.uint8).reshape(-1, 4)
        assert len(coords) == len(masks) == len(colors)
        tex_id = self._lib.vt_add_volume_bricks(coords.ctypes.data, masks.ctypes.data, colors.ctypes.data, len(coords), w, h, d)
        if tex_id < 0:
            raise RuntimeError("ERROR: Adding brick volume failed: " + abi.last_error())
        return tex_id

    # -- headless extensions -------------------------------------------------------------------
    def get_config(self) -> abi.VtConfig:
        cfg = abi.VtConfig()
        if self._lib.vt_get_config(C.byref(cfg)) != 0:
            raise RuntimeError("vt_get_config failed")
        return cfg

    def configure(self, **kw) -> abi.VtConfig:
        cfg = self.get_config()
        for k, v in kw.items():
            if not hasattr(cfg, k):
                raise AttributeError(k)
            setattr(cfg, k, v)
        if self._lib.vt_configure(C.byref(cfg)) != 0:
            raise RuntimeError("vt_configure failed: " + abi.last_error())
        return cfg

    def _size(self):
        cfg = self.get_config()
        return cfg.width, cfg.height

    def _read(self, fn, arr):
        n = fn(arr.ctypes.data, arr.nbytes)
        if n != arr.nbytes:
            raise RuntimeError(f"read-back failed ({n}): " + abi.last_error())
        return arr

    def read_hits(self) -> np.ndarray:
        w, h = self._size()
        return self._read(self._lib.vt_read_hits, np.empty((h, w), dtype=abi.HIT_DTYPE))

    def read_color(self) -> np.ndarray:
        w, h = self._size()
        return self._read(self._lib.vt_read_color, np.empty((h, w, 4), dtype=np.uint8))

    def read_color_bgra(self) -> np.ndarray:
        w, h = self._size()
        return self._read(self._lib.vt_read_color_bgra, np.empty((h, w, 4), dtype=np.uint8))

    def write_ppm(self, path: str):
        if self._lib.vt_write_ppm(path.encode()) != 0:
            raise RuntimeError("vt_write_ppm failed: " + abi.last_error())

    def read_depth(self) -> np.ndarray:
        w, h = self._size()
        return self._read(self._lib.vt_read_depth, np.empty((h, w), dtype=np.float32))

    def read_accum(self) -> np.ndarray:
        w, h = self._size()
        return self._read(self._lib.vt_read_accum, np.empty((h, w, 3), dtype=np.uint64))

    def render_async(self, perspective, camera):
        P = np.ascontiguousarray(perspective, dtype=np.float32).reshape(16)
        V = np.ascontiguousarray(camera, dtype=np.float32).reshape(16)
        if self._lib.vt_render_async(P.ctypes.data, V.ctypes.data) != 0:
            raise RuntimeError("vt_render_async failed: " + abi.last_error())

    def render_frame_async(self, perspective, camera):
        """render_tick without the wait (clear + trace + resolve enqueued on the library's stream)."""
        P = np.ascontiguousarray(perspective, dtype=np.float32).reshape(16)
        V = np.ascontiguousarray(camera, dtype=np.float32).reshape(16)
        if self._lib.vt_render_frame_async(P.ctypes.data, V.ctypes.data) != 0:
            raise RuntimeError("vt_render_frame_async failed: " + abi.last_error())

    def read_color_async(self, pinned: np.ndarray):
        """Pipelined read-back of the frame enqueued last into page-locked memory (valid after read_color_wait)."""
        if self._lib.vt_read_color_async(pinned.ctypes.data, pinned.nbytes) != pinned.nbytes:
            raise RuntimeError("vt_read_color_async failed: " + abi.last_error())

    def read_color_wait(self):
        if self._lib.vt_read_color_wait() != 0:
            raise RuntimeError("vt_read_color_wait failed: " + abi.last_error())

    def read_color_fence(self):
        if self._lib.vt_read_color_fence() != 0:
            raise RuntimeError("vt_read_color_fence failed: " + abi.last_error())

    def synchronize(self):
        if self._lib.vt_synchronize() != 0:
            raise RuntimeError("vt_synchronize failed: " + abi.last_error())

    def clear_accum(self):
        if self._lib.vt_clear_accum() != 0:
            raise RuntimeError("vt_clear_accum failed: " + abi.last_error())

    def resolve(self):
        if self._lib.vt_resolve() != 0:
            raise RuntimeError("vt_resolve failed: " + abi.last_error())

    def set_stream(self, cuda_stream_handle: int | None):
        if self._lib.vt_set_stream(C.c_void_p(cuda_stream_handle or 0)) != 0:
            raise RuntimeError("vt_set_stream failed: " + abi.last_error())

    def fused_reduce_export(self, world: int) -> bytes:
        buf = (C.c_uint8 * 64)()
        if self._lib.vt_fused_reduce_export(buf, world) != 0:
            raise RuntimeError("vt_fused_reduce_export failed: " + abi.last_error())
        return bytes(buf)

    def fused_reduce_import(self, handle: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        if self._lib.vt_fused_reduce_import(buf, rank, world) != 0:
            raise RuntimeError("vt_fused_reduce_import failed: " + abi.last_error())

    def fused_reduce_next_frame(self):
        if self._lib.vt_fused_reduce_next_frame() != 0:
            raise RuntimeError("vt_fused_reduce_next_frame failed: " + abi.last_error())

    def fused_reduce_disable(self):
        if self._lib.vt_fused_reduce_disable() != 0:
            raise RuntimeError("vt_fused_reduce_disable failed: " + abi.last_error())

    def fused_reduce_partition(self, by_tile_rows: bool, relief_num: int = 0, relief_den: int = 8):
        """False: ranks share a frame by samples; True: by rows of 8x4 tiles, every rank tracing all samples of its rows (the
        root relief_den - relief_num rows for every relief_den of another rank)."""
        if self._lib.vt_fused_reduce_partition(1 if by_tile_rows else 0, relief_num, relief_den) != 0:
            raise RuntimeError("vt_fused_reduce_partition failed: " + abi.last_error())

    def set_accum_buffer(self, device_ptr: int | None):
        if self._lib.vt_set_accum_buffer(C.c_void_p(device_ptr or 0)) != 0:
            raise RuntimeError("vt_set_accum_buffer failed: " + abi.last_error())

    def accum_device_ptr(self) -> int:
        return int(self._lib.vt_accum_device_ptr() or 0)

    def stats(self) -> abi.VtStats:
        st = abi.VtStats()
        if self._lib.vt_get_stats(C.byref(st)) != 0:
            raise RuntimeError("vt_get_stats failed")
        return st

    def reset(self):
        """cleanup() + entry(): drops every texture / instance, back to the initial state."""
        self._lib.cleanup()
        self._alive = False
        code = self._lib.entry()
        if code != 0:
            raise RuntimeError(f"ERROR: renderer initialization failed (code {code:#x}): {abi.last_error()}")
        self._alive = True
        self.texture_handle_lookup.clear()
        self.window_width = self.window_height = self.prev_window_width = self.prev_window_height = 1
        self.frame_num = 0

    # -- Drop (src/render.rs:316-320) ----------------------------------------------------------
    def close(self):
        if getattr(self, "_alive", False):
            self._lib.cleanup()
            self._alive = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
