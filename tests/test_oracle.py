"""CPU tests of the oracle: pinned against (1) vectors produced by executing the reference's own
compiled shaders, (2) the asset checksums of SURVEY.md §A.4, (3) committed checksums of its own
hit buffers at BASELINE.json configs[0]/[1] (regression pins)."""
import hashlib
import json
import os

import numpy as np
import pytest

from tools import scenes
from vtrace_b200 import glm

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# SURVEY.md §A.4
ASSET_SHA = {
    "Treasure": ("26b84b8b448245a54b59ec391f5e4cc584dcd80d6a0774f47d636acd340868c4", (50, 50, 50), 37069,
                 "6ed8da0358aa55c7bf0b4759ffe70b3dd791ff550aba9f760b3477d6dad612d2"),
    "AncientTemple": ("734120b480c2cb5a75baead8754a2862a53538a242e7eae11bfd1bc71e700897", (40, 40, 40), 12225,
                      "bfa4eb2201fdb18ddf0ffce643141b960e74731256ede07521054bffef395d4e"),
}


@pytest.mark.parametrize("name", sorted(ASSET_SHA))
def test_vox_loader_matches_survey_pins(oracle, name):
    file_sha, dims, filled, raw_sha = ASSET_SHA[name]
    path = os.path.join(scenes.ASSETS, f"{name}.vox")
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == file_sha
    raw, got_dims = oracle.load_vox(path)
    assert got_dims == dims
    assert int((raw.reshape(-1, 4)[:, 3] > 0).sum()) == filled
    assert hashlib.sha256(raw.tobytes()).hexdigest() == raw_sha


def _textures(vec):
    out = []
    for i, name in enumerate(vec["tex_names"]):
        dims = tuple(int(x) for x in vec["tex_dims"][i])
        if f"tex_raw_{i}" in vec:
            out.append((vec[f"tex_raw_{i}"], dims))
        else:
            ch = scenes.load_asset(str(name))
            assert ch.dims() == dims
            out.append((ch.get_raw(), dims))
    return out


def test_frag_main_matches_reference_spirv(oracle):
    """oracle/vtrace_oracle.c::vo_frag_main vs shaders/trace.frag.spv executed by tools/spirv_interp.py."""
    vec = np.load(os.path.join(GOLDEN, "frag_spirv_vectors.npz"))
    tex = _textures(vec)
    n = len(vec["tex"])
    assert n > 500
    bad = []
    for k in range(n):
        raw, (w, h, d) = tex[int(vec["tex"][k])]
        out, color, depth = oracle.frag_main(vec["P"][k], vec["V"][k], vec["M"][k], vec["sp"][k], vec["mp"][k][:3], raw, w, h, d)
        want_mask = int(vec["mask"][k][0]) | (int(vec["mask"][k][1]) << 1) | (int(vec["mask"][k][2]) << 2)
        ok = (int(out[0]) == 1 - int(vec["discard"][k]) and tuple(out[1:4]) == tuple(vec["voxel"][k]) and
              int(out[4]) == int(vec["steps"][k]) and int(out[5]) == want_mask and
              np.array_equal(color.view(np.uint32), vec["color"][k].view(np.uint32)) and
              np.float32(depth).view(np.uint32) == vec["depth"][k].view(np.uint32))
        if not ok:
            bad.append(k)
    assert not bad, f"{len(bad)} of {n} fragments differ from the shader binary, first: {bad[:5]}"
    assert int((vec["discard"] == 0).sum()) > 200 and int(vec["steps"].max()) >= 100


def test_vert_matches_reference_spirv(oracle):
    """trace.vert:32-47: texture id recovered from model[3][3], screen_position = (P*V)*(M*pos)."""
    vec = np.load(os.path.join(GOLDEN, "vert_spirv_vectors.npz"))
    lib = oracle.lib()
    for k in range(len(vec["texture_id"])):
        inst = vec["inst"][k].copy()
        tid = int(inst.view(np.uint32)[15])
        assert tid == int(vec["texture_id"][k])
        inst[15] = np.float32(1.0)
        assert np.array_equal(vec["model_matrix"][k].view(np.uint32), inst.view(np.uint32))
        PV = np.empty(16, dtype=np.float32)
        lib.vo_mat4_mul(oracle._p(vec["P"][k].copy()), oracle._p(vec["V"][k].copy()), oracle._p(PV))
        M, pos = inst.reshape(4, 4), np.append(vec["position"][k], np.float32(1.0)).astype(np.float32)

        def mv(m, v):
            r = np.zeros(4, dtype=np.float32)
            for i in range(4):
                acc = np.float32(m[0][i] * v[0])
                for c in range(1, 4):
                    acc = np.float32(acc + np.float32(m[c][i] * v[c]))
                r[i] = acc
            return r
        sp = mv(PV.reshape(4, 4), mv(M, pos))
        assert np.array_equal(sp.view(np.uint32), vec["screen_position"][k].view(np.uint32))
        assert np.array_equal(vec["model_position"][k], pos)


def _hash_records(rec):
    return hashlib.sha256(np.ascontiguousarray(rec).tobytes()).hexdigest()


CONFIG_PINS = os.path.join(GOLDEN, "oracle_config_pins.json")


def _render_config(oracle, assets, name, w, h, flags=0):
    sc = oracle.OracleScene()
    sc.add_texture(assets[name].get_raw(), *assets[name].dims())
    sc.set_instances(scenes.single_instance(0))
    P, V = scenes.camera(w, h)
    return sc.render_primary(P, V, w, h, flags=flags, want_depth=True)


@pytest.mark.parametrize("key,name,w,h,flags", [
    ("config0_treasure_640x480", "Treasure", 640, 480, 0),
    ("config1_temple_1080p", "AncientTemple", 1920, 1080, 0),
    ("config1_temple_1080p_viewport_h_is_w", "AncientTemple", 1920, 1080, 1),
])
def test_config_hit_buffers_are_pinned(oracle, assets, key, name, w, h, flags):
    rec, rgba, depth, iters = _render_config(oracle, assets, name, w, h, flags)
    got = {"records_sha256": _hash_records(rec), "color_sha256": hashlib.sha256(rgba.tobytes()).hexdigest(),
           "iterations": iters, "hits": int((rec["hit_voxel"] != oracle.VO_MISS).sum())}
    if os.environ.get("VT_UPDATE_PINS"):
        pins = json.load(open(CONFIG_PINS)) if os.path.exists(CONFIG_PINS) else {}
        pins[key] = got
        json.dump(pins, open(CONFIG_PINS, "w"), indent=1, sort_keys=True)
    pins = json.load(open(CONFIG_PINS))
    assert got == pins[key]


def test_primary_is_thread_count_independent(oracle, assets):
    a = _render_config(oracle, assets, "Treasure", 320, 240)
    sc = oracle.OracleScene()
    sc.add_texture(assets["Treasure"].get_raw(), *assets["Treasure"].dims())
    sc.set_instances(scenes.single_instance(0))
    P, V = scenes.camera(320, 240)
    b = sc.render_primary(P, V, 320, 240, threads=1, want_depth=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[3] == b[3]


def test_hit_record_semantics(oracle, assets):
    """voxel index / steps / face bits are consistent with each other and the volume."""
    rec, rgba, depth, iters = _render_config(oracle, assets, "Treasure", 320, 240)
    raw = assets["Treasure"].get_raw().reshape(-1, 4)
    hit = rec["hit_voxel"] != oracle.VO_MISS
    assert hit.any() and (~hit).any()
    assert (raw[rec["hit_voxel"][hit], 3] > 0).all()                      # the hit voxel is filled
    assert np.array_equal(rgba[hit][:, :3], raw[rec["hit_voxel"][hit], :3])  # opaque: colour = texel bytes
    steps = rec["packed"] & 0xFFFF
    face = (rec["packed"] >> 16) & 7
    assert (face[hit] != 0).all() and (rec["packed"][~hit] == 0).all()
    assert (steps[hit] <= rec["iters"][hit]).all() and int(rec["iters"].sum()) == iters
    assert (rec["instance"][hit] == 0).all() and (rec["instance"][~hit] == oracle.VO_MISS).all()
    assert (depth[~hit] == 1.0).all() and (depth[hit] < 1.0).all()
    clear = [oracle.lib().vo_srgb_encode(np.float32(c) / np.float32(100.0)) for c in (53.0, 81.0, 92.0)]
    assert (rgba[~hit] == np.array(clear + [255], dtype=np.uint8)).all()


def test_srgb_tables_round_trip(oracle):
    lib = oracle.lib()
    for c in range(256):
        assert lib.vo_srgb_encode(lib.vo_srgb_decode(c)) == c
    assert lib.vo_srgb_encode(np.float32(0.0)) == 0 and lib.vo_srgb_encode(np.float32(1.0)) == 255
    assert lib.vo_srgb_encode(np.float32(np.nan)) == 0 and lib.vo_srgb_encode(np.float32(7.0)) == 255


def test_mat4_inverse_is_an_inverse(oracle):
    rng = np.random.default_rng(0)
    for _ in range(50):
        m = glm.rotate(glm.translate(glm.identity(), rng.uniform(-3, 3, 3)), rng.uniform(-3, 3), rng.uniform(-1, 1, 3) + 1e-3)
        m = glm.scale(m, rng.uniform(0.3, 3, 3))
        inv = oracle.mat4_inverse(m)
        prod = (m.astype(np.float64).T @ inv.astype(np.float64).T)
        assert np.allclose(prod, np.eye(4), atol=1e-4)


def test_empty_scene_draws_clear_colour(oracle):
    sc = oracle.OracleScene()
    sc.set_instances(np.zeros((0, 16), dtype=np.float32))  # coerced to one stale instance (lib/memory.c:236,251)
    P, V = scenes.camera(64, 48)
    rec, rgba, _, iters = sc.render_primary(P, V, 64, 48)
    assert iters == 0 and (rec["hit_voxel"] == oracle.VO_MISS).all()


def test_paths_sample_sharding_sums_exactly(oracle, assets):
    """spp sharded over ranks: fixed-point accumulation => 1 rank == sum of N ranks."""
    from vtrace_b200.distributed import shard_samples
    sc = oracle.OracleScene()
    sc.add_texture(assets["AncientTemple"].get_raw(), *assets["AncientTemple"].dims())
    sc.set_instances(scenes.single_instance(0))
    w, h, spp = 96, 54, 7
    P, V = scenes.camera(w, h, eye=(0.8, -0.45, 0.6))
    whole, rays, iters = sc.render_paths(P, V, w, h, spp=spp)
    for world in (2, 3, 8):
        total = np.zeros_like(whole)
        r_sum = i_sum = 0
        for rank in range(world):
            first, stride, count = shard_samples(spp, rank, world)
            acc, r, i = sc.render_paths(P, V, w, h, spp=count, sample_first=first, sample_stride=stride)
            total += acc
            r_sum += r
            i_sum += i
        assert np.array_equal(total, whole) and (r_sum, i_sum) == (rays, iters)
    img = oracle.resolve(whole, spp)
    assert img.shape == (h, w, 4) and (img[..., 3] == 255).all()
    # energy bound: albedo <= 1, so nothing is brighter than the sky
    sky = np.array([int(np.float32(c) / np.float32(100.0) * np.float32(16777216.0)) for c in (53.0, 81.0, 92.0)], dtype=np.uint64)
    assert (whole <= sky * np.uint64(spp)).all()


def test_brick_volume_walk_equals_dense_walk(oracle):
    """The oracle's brick-volume occupancy test against its dense one: the same voxels given once as an RGBA8
    texture and once as uploaded 8^3 bricks must produce identical traversals (hit voxel, steps, face, iterations),
    for primary rays, shadow rays and path-traced radiance (colours equalised: one colour per brick)."""
    from conftest import make_volume

    rng = np.random.default_rng(21)
    w, h, d = 32, 16, 24
    occ = rng.random((d, h, w)) < 0.18
    occ[:, :, :8] &= rng.random((d, h, 8)) < 0.3           # a sparser region: some bricks end up empty
    occ[8:16, 8:16, 16:24] = False                          # an empty brick for sure
    brick_color = rng.integers(1, 256, size=(d // 8, h // 8, w // 8, 3), dtype=np.uint8)
    dense = np.zeros((d, h, w, 4), dtype=np.uint8)
    dense[..., :3] = np.repeat(np.repeat(np.repeat(brick_color, 8, axis=0), 8, axis=1), 8, axis=2)
    dense[..., 3] = np.where(occ, 255, 0)
    dense[~occ] = 0
    coords, masks, colors = [], [], []
    for bz in range(d // 8):
        for by in range(h // 8):
            for bx in range(w // 8):
                blk = occ[bz * 8:bz * 8 + 8, by * 8:by * 8 + 8, bx * 8:bx * 8 + 8]
                if not blk.any():
                    continue
                words = np.zeros(16, dtype=np.uint32)
                zz, yy, xx = np.nonzero(blk)
                for z, y, x in zip(zz, yy, xx):  # bit (x | (y&3) << 3) of word ((z&7) << 1 | (y&7) >> 2), include/vtrace_abi.h
                    words[(z << 1) | (y >> 2)] |= np.uint32(1) << np.uint32(x | ((y & 3) << 3))
                coords.append((bx, by, bz)); masks.append(words); colors.append((*brick_color[bz, by, bx], 255))
    assert 0 < len(coords) < (w // 8) * (h // 8) * (d // 8)

    a, b = oracle.OracleScene(), oracle.OracleScene()
    ta = a.add_texture(dense.reshape(-1), w, h, d)
    tb = b.add_volume_bricks(np.array(coords), np.array(masks), np.array(colors, dtype=np.uint8), w, h, d)
    m = glm.rotate(glm.identity(), 0.4, (0.2, 1.0, 0.1))
    a.set_instances(np.stack([glm.with_texture_id(m, ta).reshape(16)]))
    b.set_instances(np.stack([glm.with_texture_id(m, tb).reshape(16)]))
    for eye in [(1.4, -0.8, 1.1), (-1.0, 0.6, 0.9), (0.0, 0.0, 1.8)]:
        P, V = scenes.camera(160, 120, eye=eye)
        ra, ca, _, ia = a.render_primary(P, V, 160, 120, flags=oracle.FLAG_SHADOW_RAYS)
        rb, cb, _, ib = b.render_primary(P, V, 160, 120, flags=oracle.FLAG_SHADOW_RAYS)
        assert ia == ib and np.array_equal(ra, rb) and np.array_equal(ca, cb)
        assert (ra["hit_voxel"] != oracle.VO_MISS).sum() > 500
        pa, na, ja = a.render_paths(P, V, 160, 120, spp=2, bounces=3)
        pb, nb, jb = b.render_paths(P, V, 160, 120, spp=2, bounces=3)
        assert (na, ja) == (nb, jb) and np.array_equal(pa, pb)


@pytest.mark.parametrize("kind,dims,seed", [(1, (64, 32, 48), 4), (2, (64, 64, 64), 2)])
def test_procedural_volume_equals_its_dense_texture(oracle, kind, dims, seed):
    """A procedural volume (heightmap / sparse bricks) and the dense RGBA8 texture holding the same texels traverse
    identically: the large-scene kinds add no traversal rule of their own, only another occupancy source."""
    w, h, d = dims
    a, b = oracle.OracleScene(), oracle.OracleScene()
    ta = a.add_volume_procedural(kind, w, h, d, seed)
    texels = a.read_texels(ta, dims=dims)
    filled = int((texels.reshape(-1, 4)[:, 3] > 0).sum())
    assert 0 < filled < w * h * d
    tb = b.add_texture(texels, w, h, d)
    a.set_instances(np.stack([glm.with_texture_id(glm.identity(), ta).reshape(16)]))
    b.set_instances(np.stack([glm.with_texture_id(glm.identity(), tb).reshape(16)]))
    P, V = scenes.camera(200, 120, eye=(0.9, -0.8, 0.9))
    ra, ca, _, ia = a.render_primary(P, V, 200, 120, flags=oracle.FLAG_SHADOW_RAYS)
    rb, cb, _, ib = b.render_primary(P, V, 200, 120, flags=oracle.FLAG_SHADOW_RAYS)
    assert ia == ib and np.array_equal(ra, rb) and np.array_equal(ca, cb)
    assert (ra["hit_voxel"] != oracle.VO_MISS).sum() > 200
    n = 4096
    qa, _, ja = a.render_rays(n, seed=7)
    qb, _, jb = b.render_rays(n, seed=7)
    assert ja == jb and np.array_equal(qa, qb)
    with pytest.raises(ValueError):
        a.read_texels(ta, box=(0, 0, 0, w + 1, 1, 1))
