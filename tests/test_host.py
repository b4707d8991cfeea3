"""CPU tests of the host-side mirrors of the reference's Rust facade (glm builders, .vox loader,
texture upload queue, scene helpers)."""
import os

import numpy as np
import pytest

from tools import scenes
from vtrace_b200 import glm, voxel
from vtrace_b200.distributed import shard_samples, shard_samples_weighted
from vtrace_b200.renderer import TextureUploadQueue


def test_perspective_matches_the_gl_formula():
    P = glm.perspective(glm.REFERENCE_FOV, 16 / 9, 0.01, 10000.0)
    ys = 1.0 / np.tan(float(glm.REFERENCE_FOV) / 2.0)
    assert np.isclose(P[1][1], ys, rtol=1e-6) and np.isclose(P[0][0], ys / (16 / 9), rtol=1e-6)
    assert P[2][3] == -1.0 and P[3][3] == 0.0
    assert np.isclose(P[2][2], -(10000.0 + 0.01) / (10000.0 - 0.01), rtol=1e-6)
    assert np.isclose(P[3][2], -2 * 10000.0 * 0.01 / (10000.0 - 0.01), rtol=1e-6)


def test_look_at_is_a_rigid_transform():
    V = glm.look_at(scenes.EYE, scenes.CENTER, scenes.UP)
    R = V[:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6)
    eye_h = np.append(np.asarray(scenes.EYE, dtype=np.float32), 1.0)
    assert np.allclose(V.T @ eye_h, [0, 0, 0, 1], atol=1e-6)           # the eye maps to the origin
    c = V.T @ np.array([0, 0, 0, 1], dtype=np.float32)
    assert c[2] < 0 and abs(c[0]) < 1e-6 and abs(c[1]) < 1e-6          # looks down -Z at the centre


def test_texture_id_is_bit_cast_into_m33():
    m = glm.with_texture_id(glm.translate(glm.identity(), (1, 2, 3)), 0xBEEF)
    assert m.reshape(16).view(np.uint32)[15] == 0xBEEF and tuple(m[3][:3]) == (1, 2, 3)


def test_python_loader_matches_oracle_loader(oracle):
    for name in ("Treasure", "AncientTemple"):
        path = os.path.join(scenes.ASSETS, f"{name}.vox")
        raw, dims = oracle.load_vox(path)
        chunk = voxel.load_magica_voxel(path)[0]
        assert chunk.dims() == dims and np.array_equal(chunk.get_raw(), raw)


def test_chunk_layout_is_z_fastest():
    c = voxel.RawDynamicChunk(2, 3, 4)
    c.at_mut(1, 2, 3)[:] = (9, 8, 7, 6)
    flat = c.get_raw().reshape(-1, 4)
    assert tuple(flat[3 + 4 * (2 + 3 * 1)]) == (9, 8, 7, 6)   # z + dz*(y + dy*x), src/voxel/rawchunk.rs:292
    assert c.at_mut(2, 0, 0) is None


def test_texture_upload_queue_is_fifo_with_dense_handles():
    q = TextureUploadQueue()
    assert [q.add_texture(x) for x in "abc"] == [0, 1, 2]
    assert q.pop() == ("a", 0) and q.pop() == ("b", 1) and q.pop() == ("c", 2) and q.pop() is None


def test_entity_grid_is_the_reference_world():
    g = scenes.entity_grid(3, 7)
    assert g.shape == (121, 16)                                            # 11 x 11, src/world.rs:143-161
    ids = g.view(np.uint32)[:, 15]
    assert set(ids.tolist()) == {3, 7} and (ids == 3).sum() == 61
    assert np.allclose(g[:, 13], -5.0) and np.isclose(g[:, 12].min(), -7.5) and np.isclose(g[:, 14].max(), 7.5)


def test_shard_samples_partitions_every_sample_once():
    for total in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            seen = []
            for rank in range(world):
                first, stride, count = shard_samples(total, rank, world)
                seen += [first + k * stride for k in range(count)]
            assert sorted(seen) == list(range(total))


def test_weighted_sharding_partitions_every_sample_once_and_relieves_the_root():
    for total in (0, 1, 7, 64, 65, 256):
        for world in (1, 2, 3, 4, 8):
            for extra in (0.0, 2.0, 5.5, 1000.0):
                seen, counts = [], []
                for rank in range(world):
                    first, stride, count = shard_samples_weighted(total, rank, world, extra)
                    seen += [first + k * stride for k in range(count)]
                    counts.append(count)
                assert sorted(seen) == list(range(total))
                if world > 1:
                    assert counts[0] <= min(counts[1:]) + (1 if extra == 0.0 else 0)
                    assert max(counts[1:]) - min(counts[1:]) <= 1
    assert [shard_samples_weighted(64, r, 8)[2] for r in range(8)] == [6, 9, 9, 8, 8, 8, 8, 8]
    assert [shard_samples_weighted(64, r, 2)[2] for r in range(2)] == [31, 33]


def _fnv1a64(b: bytes) -> int:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.parametrize("name,with_rgba", [("three_models_rgba", True), ("three_models_default_palette", False)])
def test_multi_model_vox_and_default_palette_across_the_three_loaders(oracle, tmp_path, name, with_rgba):
    """dot_vox semantics the two reference assets do not exercise: one chunk per SIZE / XYZI pair, and
    the default palette when the file has no RGBA chunk.  The C oracle loader, the Python loader and the C++ host's loader
    must produce the same bytes — and those bytes must be what an independent construction from the fixture's voxel list
    gives.  (The default palette itself is restated from the published .vox format, not from the crate: unpinned.)"""
    import subprocess

    from tools import gen_vox_fixtures as gen
    path = os.path.join(scenes.ASSETS, f"{name}.vox")
    pal = gen.palette() if with_rgba else voxel.default_palette()
    if not with_rgba:  # spot values of the published table: index 1 white, 2 = 0xffccffff, 216 = darkest-but-one... red ramp
        assert tuple(pal[0]) == (255, 255, 255, 255) and tuple(pal[1]) == (255, 255, 204, 255)
        assert tuple(pal[215]) == (238, 0, 0, 255) and tuple(pal[254]) == (17, 17, 17, 255) and tuple(pal[255]) == (0, 0, 0, 0)
        assert len({tuple(c) for c in pal[:255]}) == 255
    chunks = voxel.load_magica_voxel(path)
    ms = gen.models()
    assert len(chunks) == len(ms) == 3
    exe = tmp_path / "vox_dump"
    host = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "host")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(host, "..", "include"), "-o", str(exe), os.path.join(host, "vox_dump.cpp")],
                   check=True)
    dumped = subprocess.run([str(exe), path], check=True, capture_output=True, text=True).stdout.splitlines()
    assert len(dumped) == 3
    for m, ((sx, sy, sz), vox) in enumerate(ms):
        want = np.zeros((sx, sy, sz, 4), dtype=np.uint8)  # chunk (x, size.y - z - 1, y), magica_voxel.rs:31-37
        for x, y, z, i in vox:
            want[x, sy - int(z) - 1, y] = pal[int(i) - 1]
        raw, dims = oracle.load_vox_model(path, m)
        assert dims == (sx, sy, sz) == chunks[m].dims()
        assert np.array_equal(raw.reshape(-1), want.reshape(-1))
        assert np.array_equal(chunks[m].get_raw(), want.reshape(-1))
        assert dumped[m] == f"model {m} dims {sx} {sy} {sz} fnv1a {_fnv1a64(want.tobytes()):016x}"
    assert oracle.load_vox_model(path, 3) is None


def test_tile_row_ownership_of_the_fused_reduction():
    """vt_fused_reduce_partition shares a frame by rows of tiles; the host-side mirror of the device rule: every row has one
    owner, the root's share shrinks by eighths, the other ranks' shares stay equal."""
    from vtrace_b200.distributed import row_owner, rows_per_rank
    assert [row_owner(t, 1) for t in range(5)] == [0] * 5
    assert [row_owner(t, 4) for t in range(8)] == [0, 1, 2, 3, 0, 1, 2, 3]
    for world in (2, 4, 8):
        for relief in range(8):
            cycle = relief * (world - 1) + (8 - relief) * world
            counts = rows_per_rank(cycle * 3, world, relief)
            assert counts[0] == 3 * (8 - relief) and all(c == 24 for c in counts[1:]), (world, relief, counts)
            assert sum(counts) == cycle * 3
    # 270 tile rows of a 1080p frame over 8 ranks, the root relieved by 2/8
    counts = rows_per_rank(270, 8, 2)
    assert sum(counts) == 270 and counts[0] < min(counts[1:]) and max(counts[1:]) - min(counts[1:]) <= 2
