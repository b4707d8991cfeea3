"""Named benchmark/test scenes (BASELINE.json configs).  Pure host-side data:
camera matrices and instance matrices exactly as the Rust engine would pass them."""
from __future__ import annotations

import os

import numpy as np

from vtrace_b200 import glm, voxel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "tests", "golden", "assets")

# fixed forever
EYE = (1.6, -0.9, 1.2)
CENTER = (0.0, 0.0, 0.0)
UP = (0.0, 1.0, 0.0)


def load_asset(name: str) -> voxel.RawDynamicChunk:
    return voxel.load_magica_voxel(os.path.join(ASSETS, f"{name}.vox"))[0]


def camera(width: int, height: int, eye=EYE, center=CENTER):
    P = glm.perspective(glm.REFERENCE_FOV, np.float32(width) / np.float32(height), glm.REFERENCE_NEAR, glm.REFERENCE_FAR)
    V = glm.look_at(eye, center, UP)
    return P, V


def single_instance(texture_id: int = 0) -> np.ndarray:
    return glm.with_texture_id(glm.identity(), texture_id).reshape(1, 16)


def entity_grid(tex_treasure: int, tex_temple: int, half: int = 5) -> np.ndarray:
    """The 11x11 alternating entity grid of src/world.rs:143-161."""
    mats = []
    for x in range(-half, half + 1):
        for z in range(-half, half + 1):
            model = glm.translate(glm.identity(), (x * 1.5, -5.0, z * 1.5))
            tid = tex_treasure if (x + z + 10) % 2 == 0 else tex_temple
            mats.append(glm.with_texture_id(model, tid).reshape(16))
    return np.stack(mats)


# ---- the reference's default world (src/world.rs:143-198): the 11x11 entity grid plus the paged terrain ---------------

CHUNK_VOXEL_SIZE = 16      # src/gen/pager.rs:24
CHUNK_WORLD_SIZE = 2.0     # src/gen/pager.rs:25


def _hash3(x, y, z, seed):
    """lowbias32 over integer lattice points (vectorised); stands in for the `noise` crate (un-vendored, unpinned)."""
    def mix(v):
        v = v.copy()
        v ^= v >> np.uint32(16); v *= np.uint32(0x7FEB352D); v ^= v >> np.uint32(15); v *= np.uint32(0x846CA68B); v ^= v >> np.uint32(16)
        return v
    with np.errstate(over="ignore"):
        h = mix(x.astype(np.uint32) * np.uint32(0x9E3779B1) + np.uint32(seed))
        h = mix(h ^ (y.astype(np.uint32) * np.uint32(0x85EBCA77)))
        h = mix(h ^ (z.astype(np.uint32) * np.uint32(0xC2B2AE3D)))
    return h


def _value_noise(px, py, pz, seed):
    """trilinear value noise in [-1, 1] at float64 positions"""
    x0, y0, z0 = np.floor(px).astype(np.int64), np.floor(py).astype(np.int64), np.floor(pz).astype(np.int64)
    fx, fy, fz = px - x0, py - y0, pz - z0
    ux, uy, uz = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy), fz * fz * (3 - 2 * fz)
    acc = 0.0
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                v = _hash3(x0 + dx, y0 + dy, z0 + dz, seed).astype(np.float64) / 4294967295.0 * 2.0 - 1.0
                acc = acc + v * (ux if dx else 1 - ux) * (uy if dy else 1 - uy) * (uz if dz else 1 - uz)
    return acc


def terrain_chunk(cx: int, cy: int, cz: int, seed: int = 0):
    """One 16^3 chunk of the paged terrain, shaped like src/gen/terrain.rs:37-94: voxels inside a sphere of radius 50
    voxels around the origin, solid where a 3-D noise at 0.1 x position is positive, dirt colour within 4 voxels of the
    surface (else stone), shaded by a second noise.  The two noises are value noise over an integer hash — the reference's
    OpenSimplex / Billow come from the un-vendored `noise` 0.7.0 crate and cannot be restated offline (parity unpinned);
    chunk count, chunk size, instance transforms and texture layout are the reference's.  Returns None for an empty chunk
    (src/gen/terrain.rs:89-93), else a RawDynamicChunk written as chunk.at_mut(z, y, x) (:84)."""
    n = CHUNK_VOXEL_SIZE
    ax = np.arange(n, dtype=np.int64)
    x, y, z = np.meshgrid(ax + cx * n, ax + cy * n, ax + cz * n, indexing="ij")
    wx, wy, wz = x.astype(np.float64), y.astype(np.float64), z.astype(np.float64)
    inside = wx * wx + wy * wy + wz * wz <= 50.0 * 50.0
    if not inside.any():
        return None
    surface = wx * wx + (wy - 4.0) * (wy - 4.0) + wz * wz > 50.0 * 50.0
    solid = inside & (_value_noise(wx * 0.1, wy * 0.1, wz * 0.1, seed) > 0.0)
    if not solid.any():
        return None
    tone = 0.5 * np.abs(_value_noise(wx * 0.1 + 17.0, wy * 0.1 - 5.0, wz * 0.1 + 3.0, seed + 1)) + 0.5
    col = np.where(surface[..., None], np.array([255.0, 200.0, 100.0]), np.array([150.0, 150.0, 150.0])) * tone[..., None]
    rgba = np.zeros((n, n, n, 4), dtype=np.uint8)
    rgba[..., :3] = col.astype(np.uint8)
    rgba[..., 3] = 255
    rgba[~solid] = 0
    chunk = voxel.RawDynamicChunk(n, n, n)
    chunk.data[:] = rgba.transpose(2, 1, 0, 3)  # at_mut(z, y, x) = voxel (x, y, z)
    return chunk


def default_world(camera_position=(0.0, 0.0, 0.0), load_dist: int = 10, seed: int = 0):
    """(chunks to upload, instances) of the reference's default world around `camera_position`: textures 0 / 1 are
    AncientTemple / Treasure (upload order of src/world.rs:55-56), then one texture per non-empty terrain chunk in paging
    order; instances = 121 entities followed by the terrain chunks (src/world.rs:143-198).  `instances` is a list of
    (model matrix, texture index)."""
    textures = [load_asset("AncientTemple"), load_asset("Treasure")]
    inst = []
    for x in range(-5, 6):
        for z in range(-5, 6):
            model = glm.translate(glm.identity(), (x * 1.5, -5.0, z * 1.5))
            inst.append((model, 1 if (x + z + 10) % 2 == 0 else 0))
    cp = tuple(int(np.floor(np.float32(c) / np.float32(CHUNK_WORLD_SIZE))) for c in camera_position)
    r = range(-load_dist, load_dist + 1)
    for x in r:
        for y in r:
            for z in r:
                c = (cp[0] + x, cp[1] + y, cp[2] + z)
                if max(abs(c[0]), abs(c[1]), abs(c[2])) > 4:  # 50 voxels = 3.125 chunks: farther chunks are empty (quick reject)
                    continue
                chunk = terrain_chunk(*c, seed=seed)
                if chunk is None:
                    continue
                t = glm.translate(glm.identity(), tuple(np.float32(v) * np.float32(CHUNK_WORLD_SIZE) for v in c))
                model = glm.scale(t, (CHUNK_WORLD_SIZE, CHUNK_WORLD_SIZE, CHUNK_WORLD_SIZE))
                inst.append((model, len(textures)))
                textures.append(chunk)
    return textures, inst
