import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def assets():
    from tools import scenes
    return {name: scenes.load_asset(name) for name in ("Treasure", "AncientTemple")}


@pytest.fixture(scope="session")
def renderer():
    """One librender instance per test session (the ABI is a process-wide singleton)."""
    from vtrace_b200.renderer import Renderer
    r = Renderer()
    yield r
    r.close()


def make_volume(rng: np.random.Generator, w, h, d, fill=0.3, alpha_choices=(255,)):
    """Random RGBA8 volume in the x-fastest byte order add_texture consumes."""
    vol = np.zeros((d, h, w, 4), dtype=np.uint8)
    filled = rng.random((d, h, w)) < fill
    vol[..., :3] = rng.integers(0, 256, size=(d, h, w, 3), dtype=np.uint8)
    vol[..., 3] = np.where(filled, rng.choice(np.array(alpha_choices, dtype=np.uint8), size=(d, h, w)), 0)
    vol[~filled] = 0
    return vol.reshape(-1)


class RawVolume:
    """Duck-types RawDynamicChunk for Renderer.add_texture."""

    def __init__(self, raw, w, h, d):
        self.raw, self.w, self.h, self.d = raw, w, h, d

    def get_raw(self):
        return self.raw

    def dims(self):
        return self.w, self.h, self.d
