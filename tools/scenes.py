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
