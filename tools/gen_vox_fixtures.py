#!/usr/bin/env python
"""Writes the synthetic .vox fixtures under tests/golden/assets/ that exercise what the two reference assets do not:
several models in one file (PACK + one SIZE / XYZI pair each) and a file without RGBA chunk (default palette).

    python tools/gen_vox_fixtures.py

The voxel lists are a function of a fixed seed; `models()` is imported by the tests to rebuild the expected bytes
independently of the three loaders.
"""
from __future__ import annotations

import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "assets")


def models():
    """[(size xyz, voxels n x 4 (x, y, z, palette index 1..255))]: size.y == size.z, as in both reference assets."""
    rng = np.random.default_rng(0x70C5)
    out = []
    for sx, sy, sz, n in ((6, 5, 5, 60), (3, 9, 9, 90), (8, 8, 8, 200)):
        cells = rng.choice(sx * sy * sz, size=n, replace=False)
        v = np.stack([cells % sx, (cells // sx) % sy, cells // (sx * sy), rng.integers(1, 256, size=n)], axis=1).astype(np.uint8)
        v[0, 3], v[1, 3] = 1, 255  # both ends of the palette
        out.append(((sx, sy, sz), v))
    return out


def palette():
    rng = np.random.default_rng(0xA11E7)
    p = rng.integers(0, 256, size=(256, 4), dtype=np.uint8)
    p[:, 3] = 255
    p[255] = 0
    return p


def chunk(cid: bytes, body: bytes, children: bytes = b"") -> bytes:
    return cid + struct.pack("<II", len(body), len(children)) + body + children


def write(path: str, with_rgba: bool):
    ms = models()
    kids = chunk(b"PACK", struct.pack("<I", len(ms)))
    for (sx, sy, sz), v in ms:
        kids += chunk(b"SIZE", struct.pack("<III", sx, sy, sz))
        kids += chunk(b"XYZI", struct.pack("<I", len(v)) + v.tobytes())
    if with_rgba:
        kids += chunk(b"RGBA", palette().tobytes())
    with open(path, "wb") as f:
        f.write(b"VOX " + struct.pack("<I", 150) + chunk(b"MAIN", b"", kids))


if __name__ == "__main__":
    write(os.path.join(OUT, "three_models_rgba.vox"), True)
    write(os.path.join(OUT, "three_models_default_palette.vox"), False)
    print("wrote", OUT)
