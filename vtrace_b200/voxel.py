"""Host-side mirror of the reference's voxel content path that feeds add_texture.

  * load_magica_voxel  <- src/voxel/magica_voxel.rs:18-44 (over dot_vox 4.1.0, Cargo.lock:127-130)
  * RawDynamicChunk    <- src/voxel/rawchunk.rs:254-320 (linear index z + dz*(y + dy*x), :292)
  * Color              <- src/voxel/common.rs:65-87 (r,g,b,a bytes; from_uint little-endian)

The Rust engine keeps doing this itself in the real drop-in; this file exists so the Python
harness (tests, bench) can hand the C ABI the same bytes the engine would.
"""
from __future__ import annotations

import struct

import numpy as np


class RawDynamicChunk:
    """Dense x-major/z-fastest voxel container of RGBA `Color`s (src/voxel/rawchunk.rs:254-320)."""

    def __init__(self, dim_x: int, dim_y: int, dim_z: int):
        self.dim_x, self.dim_y, self.dim_z = dim_x, dim_y, dim_z
        # data[x][y][z][rgba]  ==  flat index z + dim_z*(y + dim_y*x)
        self.data = np.zeros((dim_x, dim_y, dim_z, 4), dtype=np.uint8)

    def at_mut(self, x, y, z):
        if not (0 <= x < self.dim_x and 0 <= y < self.dim_y and 0 <= z < self.dim_z):
            return None
        return self.data[x, y, z]

    def get_raw(self) -> np.ndarray:
        """The byte buffer add_texture receives (src/render.rs:254-261)."""
        return np.ascontiguousarray(self.data).reshape(-1)

    def dims(self):
        """(width, height, depth) exactly as src/render.rs:255-260 passes them."""
        return self.dim_x, self.dim_y, self.dim_z


def _chunks(buf: bytes, off: int, end: int):
    while off + 12 <= end:
        cid = buf[off:off + 4]
        n, m = struct.unpack_from("<II", buf, off + 4)
        yield cid, buf[off + 12:off + 12 + n]
        off += 12 + n + m


def default_palette() -> np.ndarray:
    """MagicaVoxel's default palette (published with the .vox format: index 0 unused, a 6x6x6 colour cube without black,
    then ramps of red, green, blue, grey), laid out like an RGBA chunk: row k = colour of file index k + 1.  dot_vox 4.1.0
    supplies its copy when a file has no RGBA chunk; that copy is not available offline — unpinned restatement."""
    lv = [0xFF, 0xCC, 0x99, 0x66, 0x33, 0x00]
    ramp = [0xEE, 0xDD, 0xBB, 0xAA, 0x88, 0x77, 0x55, 0x44, 0x22, 0x11]
    t = [0] * 257
    for k in range(215):  # 0xAABBGGRR: blue runs fastest, red slowest
        t[k + 1] = 0xFF000000 | lv[k % 6] << 16 | lv[(k // 6) % 6] << 8 | lv[k // 36]
    for j in range(10):
        t[216 + j] = 0xFF000000 | ramp[j]
        t[226 + j] = 0xFF000000 | ramp[j] << 8
        t[236 + j] = 0xFF000000 | ramp[j] << 16
        t[246 + j] = 0xFF000000 | ramp[j] * 0x010101
    return np.array(t[1:], dtype="<u4").view(np.uint8).reshape(256, 4)


def load_magica_voxel(path: str) -> list[RawDynamicChunk]:
    """One RawDynamicChunk per model, axes remapped as magica_voxel.rs:31-37 does."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:4] != b"VOX " or buf[8:12] != b"MAIN":
        raise ValueError("not a MagicaVoxel file")
    sizes, xyzis, palette, pending = [], [], None, None
    for cid, body in _chunks(buf, 20, len(buf)):
        if cid == b"SIZE":
            pending = struct.unpack_from("<III", body)
        elif cid == b"XYZI" and pending is not None:  # one model per SIZE / XYZI pair, in file order
            n = struct.unpack_from("<I", body)[0]
            sizes.append(pending)
            xyzis.append(np.frombuffer(body, dtype=np.uint8, count=4 * n, offset=4).reshape(n, 4))
            pending = None
        elif cid == b"RGBA" and palette is None:
            palette = np.frombuffer(body, dtype=np.uint8, count=1024).reshape(256, 4)
    if palette is None:
        palette = default_palette()
    out = []
    for (sx, sy, sz), vox in zip(sizes, xyzis):
        chunk = RawDynamicChunk(sx, sy, sz)
        x = vox[:, 0].astype(np.int64)
        y = sy - vox[:, 2].astype(np.int64) - 1
        z = vox[:, 1].astype(np.int64)
        if (y < 0).any() or (z >= sz).any() or (x >= sx).any():
            raise ValueError("voxel outside chunk (the reference would panic on unwrap)")
        idx = np.maximum(vox[:, 3].astype(np.int64) - 1, 0)  # dot_vox: i = file index - 1
        # later voxels overwrite earlier ones, as the sequential Rust loop does
        for k in range(len(vox)):
            chunk.data[x[k], y[k], z[k]] = palette[idx[k]]
        out.append(chunk)
    return out
