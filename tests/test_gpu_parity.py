"""GPU parity tests proper: the CUDA path, called through the C ABI (librender.so), against the
CPU oracle on the same inputs.  Integer outputs (hit voxel, face, steps, instance, iteration
counts, colour bytes, fixed-point radiance sums) must be BIT-EXACT; depth is compared bit-exact
too because both sides execute the same IEEE operations.

Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import numpy as np
import pytest

from conftest import RawVolume, make_volume
from tools import scenes
from vtrace_b200 import abi, glm

pytestmark = pytest.mark.gpu


def _assert_records_equal(got, want, what=""):
    for field in ("hit_voxel", "packed", "instance", "iters"):
        g, w = got[field], want[field]
        if not np.array_equal(g, w):
            bad = np.argwhere(g != w)
            y, x = bad[0]
            raise AssertionError(
                f"{what}: {field} differs at {len(bad)} pixels; first (x={x}, y={y}): cuda={g[y, x]:#x} oracle={w[y, x]:#x}")


class Scene:
    """Feeds the same textures / instances to librender and to the oracle."""

    _next_tex = 0  # librender's texture table only grows (as the reference's does)

    def __init__(self, renderer, oracle):
        self.r = renderer
        self.o = oracle.OracleScene()
        self.tex_map = {}

    def add(self, chunk):
        rid = self.r.add_texture(chunk)
        oid = self.o.add_texture(chunk.get_raw(), *chunk.dims())
        self.tex_map[oid] = rid
        return oid

    def set_instances(self, models_and_tex):
        """models_and_tex: list of (mat4, oracle-side texture id)"""
        if len(models_and_tex):
            o_m = np.stack([glm.with_texture_id(m, t).reshape(16) for m, t in models_and_tex])
            r_m = np.stack([glm.with_texture_id(m, self.tex_map.get(t, 60000 + t)).reshape(16) for m, t in models_and_tex])
        else:
            o_m = r_m = np.zeros((0, 16), dtype=np.float32)
        self.o.set_instances(o_m)
        self.r.update_instances_raw(r_m)

    def check_primary(self, P, V, w, h, flags=0, what="", subset=False):
        self.r.configure(width=w, height=h, mode=abi.MODE_PRIMARY, flags=flags, max_frames=0)
        assert self.r.render_tick_raw(P, V)
        assert (self.r.window_width, self.r.window_height) == (w, h)
        got, color, depth = self.r.read_hits(), self.r.read_color(), self.r.read_depth()
        oflags = (flags & (1 | abi.FLAG_SHADOW_RAYS)) | (0x10000 if subset else 0)
        want, wcolor, wdepth, iters = self.o.render_primary(P, V, w, h, flags=oflags, want_depth=True)
        st = self.r.stats()
        if subset:  # the oracle computed every 8th pixel in x and y only (full-size configurations)
            got, color, depth = got[::8, ::8], color[::8, ::8], depth[::8, ::8]
            want, wcolor, wdepth = want[::8, ::8], wcolor[::8, ::8], wdepth[::8, ::8]
        _assert_records_equal(got, want, what)
        assert np.array_equal(color, wcolor), f"{what}: colour bytes differ"
        assert np.array_equal(depth.view(np.uint32), wdepth.view(np.uint32)), f"{what}: depth bits differ"
        if not subset:
            assert st.iterations == iters, f"{what}: iteration counter {st.iterations} != {iters}"
            shadow = self.o.last_shadow_rays() if flags & abi.FLAG_SHADOW_RAYS else 0
            assert st.rays == w * h + shadow, f"{what}: ray counter {st.rays} != {w * h} + {shadow}"
        return got, st

    def check_paths(self, P, V, w, h, spp, bounces=4, seed=0x5EED, flags=0, what=""):
        self.r.configure(width=w, height=h, mode=abi.MODE_PATHS, flags=flags, spp=spp, bounces=bounces, seed=seed,
                         sample_first=0, sample_stride=1, total_spp=spp, max_frames=0)
        assert self.r.render_tick_raw(P, V)
        got = self.r.read_accum()
        color = self.r.read_color()
        want, rays, iters = self.o.render_paths(P, V, w, h, spp=spp, bounces=bounces, seed=seed, flags=flags & 1)
        if not np.array_equal(got, want):
            bad = np.argwhere((got != want).any(axis=2))
            y, x = bad[0]
            raise AssertionError(f"{what}: radiance sums differ at {len(bad)} pixels; first (x={x}, y={y}): "
                                 f"cuda={got[y, x]} oracle={want[y, x]}")
        st = self.r.stats()
        assert st.rays == rays and st.iterations == iters, f"{what}: counters {st.rays},{st.iterations} != {rays},{iters}"
        import oracle_lib
        assert np.array_equal(color, oracle_lib.resolve(want, spp)), f"{what}: resolved colour differs"
        # the tolerance the north star states for radiance, on top of the exact check above
        scale = 1.0 / (spp * 2.0**24)
        diff = (got.astype(np.float64) - want.astype(np.float64)) * scale
        assert np.sqrt((diff**2).mean()) <= 1e-4 and np.abs(diff).max() <= 1e-3
        return got, st


@pytest.fixture()
def scene(renderer, oracle):
    renderer.reset()  # fresh texture table / mask arena for every test
    return Scene(renderer, oracle)


# ---- BASELINE.json configs[0] and configs[1] ------------------------------------------------

def test_config0_treasure_640x480_primary(scene, assets):
    t = scene.add(assets["Treasure"])
    scene.set_instances([(glm.identity(), t)])
    P, V = scenes.camera(640, 480)
    got, st = scene.check_primary(P, V, 640, 480, what="Treasure 640x480")
    assert (got["hit_voxel"] != abi.VT_MISS).sum() > 10000
    assert st.masks_in_smem == 1


@pytest.mark.parametrize("flags", [0, abi.FLAG_VIEWPORT_H_IS_W])
def test_config1_temple_1080p_primary(scene, assets, flags):
    t = scene.add(assets["AncientTemple"])
    scene.set_instances([(glm.identity(), t)])
    P, V = scenes.camera(1920, 1080)
    got, _ = scene.check_primary(P, V, 1920, 1080, flags=flags, what=f"Temple 1080p flags={flags}")
    assert (got["hit_voxel"] != abi.VT_MISS).sum() > 20000


def test_bgra_and_ppm_readback(renderer, scene, assets, tmp_path):
    t = scene.add(assets["Treasure"])
    scene.set_instances([(glm.identity(), t)])
    P, V = scenes.camera(160, 120, eye=(0.9, -0.5, 0.7))
    scene.check_primary(P, V, 160, 120, what="readback formats")
    rgba, bgra = renderer.read_color(), renderer.read_color_bgra()
    assert np.array_equal(rgba[..., [2, 1, 0, 3]], bgra)  # VK_FORMAT_B8G8R8A8_SRGB byte order, lib/swapchain.c:88
    path = str(tmp_path / "f.ppm")
    renderer.write_ppm(path)
    data = open(path, "rb").read()
    head = b"P6\n160 120\n255\n"
    assert data.startswith(head) and np.array_equal(np.frombuffer(data[len(head):], dtype=np.uint8).reshape(120, 160, 3), rgba[..., :3])


def test_closeup_cameras(scene, assets):
    t = scene.add(assets["AncientTemple"])
    scene.set_instances([(glm.identity(), t)])
    for eye in [(0.8, -0.45, 0.6), (-0.7, 0.2, 0.9), (0.0, -1.2, 0.05), (0.55, 0.55, 0.55)]:
        P, V = scenes.camera(512, 288, eye=eye)
        scene.check_primary(P, V, 512, 288, what=f"closeup {eye}")


def test_global_mask_path_matches(scene, assets):
    t = scene.add(assets["Treasure"])
    scene.set_instances([(glm.identity(), t)])
    P, V = scenes.camera(640, 480, eye=(0.9, -0.5, 0.7))
    _, st = scene.check_primary(P, V, 640, 480, flags=abi.FLAG_FORCE_GLOBAL_MASKS, what="global masks")
    assert st.masks_in_smem == 0


# ---- edge cases -----------------------------------------------------------------------------

def test_camera_inside_volume_draws_nothing(scene, assets):
    t = scene.add(assets["Treasure"])
    scene.set_instances([(glm.identity(), t)])
    P, V = scenes.camera(320, 240, eye=(0.1, 0.2, -0.1), center=(1.0, 0.3, 0.2))
    got, _ = scene.check_primary(P, V, 320, 240, what="camera inside")
    assert (got["hit_voxel"] == abi.VT_MISS).all()  # back faces only -> culled (lib/pipeline.c:120-121)


def test_axis_aligned_rays(scene, assets):
    """Zero direction components: delta = inf, 0*inf = NaN inside the loop (SURVEY §7 hard parts)."""
    t = scene.add(assets["AncientTemple"])
    scene.set_instances([(glm.identity(), t)])
    for eye, center in [((0.0, 0.0, 2.0), (0.0, 0.0, 0.0)), ((2.0, 0.0, 0.0), (0.0, 0.0, 0.0)),
                        ((0.0, 0.0, -1.5), (0.0, 0.0, 0.0))]:
        # odd sizes put a pixel centre exactly on the optical axis
        P, V = scenes.camera(321, 241, eye=eye, center=center)
        scene.check_primary(P, V, 321, 241, what=f"axis aligned {eye}")


def test_empty_scene_and_bad_texture_id(scene, assets):
    P, V = scenes.camera(160, 120)
    scene.set_instances([])  # coerced to one stale instance (lib/memory.c:236,251)
    got, _ = scene.check_primary(P, V, 160, 120, what="empty scene")
    assert (got["hit_voxel"] == abi.VT_MISS).all()
    scene.set_instances([(glm.identity(), 40000)])  # texture id out of range
    scene.check_primary(P, V, 160, 120, what="bad texture id")


def test_ragged_and_noncubic_volumes(scene):
    """Sizes where floor(fl(v/s)*s) != v (texel remap) and W != H != D (reference quirk: dir not scaled)."""
    rng = np.random.default_rng(7)
    for (w, h, d) in [(22, 23, 29), (1, 1, 1), (3, 50, 7), (41, 41, 41), (64, 16, 33)]:
        t = scene.add(RawVolume(make_volume(rng, w, h, d, fill=0.15), w, h, d))
        scene.set_instances([(glm.identity(), t)])
        P, V = scenes.camera(400, 300, eye=(0.9, -0.6, 0.8))
        scene.check_primary(P, V, 400, 300, what=f"volume {w}x{h}x{d}")


def test_transparent_voxels_blend(scene):
    rng = np.random.default_rng(11)
    w = h = d = 24
    t = scene.add(RawVolume(make_volume(rng, w, h, d, fill=0.2, alpha_choices=(255, 128, 1, 77)), w, h, d))
    models = [(glm.translate(glm.identity(), (0.0, 0.0, -0.8 * k)), t) for k in range(3)]
    scene.set_instances(models)
    P, V = scenes.camera(480, 360, eye=(0.3, -0.4, 1.6))
    scene.check_primary(P, V, 480, 360, what="alpha blending across instances")


def test_multi_instance_grid(scene, assets):
    """The 11x11 entity grid of src/world.rs:143-161 (rank 1 of SURVEY §8f)."""
    a = scene.add(assets["Treasure"])
    b = scene.add(assets["AncientTemple"])
    grid = scenes.entity_grid(a, b)
    models = []
    for m in grid:
        m = m.reshape(4, 4).copy()
        tid = int(m.reshape(16).view(np.uint32)[15])
        m[3][3] = 1.0
        models.append((m, tid))
    scene.set_instances(models)
    P = glm.perspective(glm.REFERENCE_FOV, 640 / 360, glm.REFERENCE_NEAR, glm.REFERENCE_FAR)
    V = glm.look_at((0.0, -2.0, 0.0), (3.0, -5.0, 2.0), (0.0, 1.0, 0.0))
    got, _ = scene.check_primary(P, V, 640, 360, what="entity grid")
    assert len(np.unique(got["instance"])) > 10


def test_many_instances_binned(renderer, scene, monkeypatch):
    """1000 chunk instances (the shape of the reference's paged terrain, src/world.rs:163-198): the
    screen-space bins must give exactly what visiting every instance gives — and what the oracle gives."""
    rng = np.random.default_rng(3)
    t = scene.add(RawVolume(make_volume(rng, 16, 16, 16, fill=0.35), 16, 16, 16))
    models = [(glm.scale(glm.translate(glm.identity(), (2.2 * x, 2.2 * y - 9.0, 2.2 * z)), (2.0, 2.0, 2.0)), t)
              for x in range(-5, 5) for y in range(-5, 5) for z in range(-5, 5)]
    scene.set_instances(models)
    P = glm.perspective(glm.REFERENCE_FOV, 640 / 360, glm.REFERENCE_NEAR, glm.REFERENCE_FAR)
    V = glm.look_at((14.0, 4.0, 17.0), (0.0, -9.0, 0.0), (0.0, 1.0, 0.0))
    got, st = scene.check_primary(P, V, 640, 360, what="1000 instances, binned")
    assert len(np.unique(got["instance"])) > 100
    unbinned, _ = scene.check_primary(P, V, 640, 360, flags=abi.FLAG_NO_BINNING, what="1000 instances, unbinned")
    assert np.array_equal(got, unbinned)


def test_bin_list_overflow_grows_and_retries(renderer, oracle, assets, monkeypatch):
    monkeypatch.setenv("VT_BIN_CAP", "1000")  # far too small: the first attempt overflows
    renderer.reset()
    sc = Scene(renderer, oracle)
    a = sc.add(assets["Treasure"])
    b = sc.add(assets["AncientTemple"])
    grid = scenes.entity_grid(a, b)
    models = []
    for m in grid:
        m = m.reshape(4, 4).copy()
        tid = int(m.reshape(16).view(np.uint32)[15])
        m[3][3] = 1.0
        models.append((m, tid))
    sc.set_instances(models)
    P = glm.perspective(glm.REFERENCE_FOV, 640 / 360, glm.REFERENCE_NEAR, glm.REFERENCE_FAR)
    V = glm.look_at((0.0, -2.0, 0.0), (3.0, -5.0, 2.0), (0.0, 1.0, 0.0))
    sc.check_primary(P, V, 640, 360, what="bin overflow retry")
    sc.check_paths(P, V, 160, 90, spp=2, bounces=2, what="bin overflow retry, paths")


def test_rotated_scaled_instances(scene, assets):
    t = scene.add(assets["Treasure"])
    m1 = glm.scale(glm.rotate(glm.translate(glm.identity(), (0.4, 0.1, -0.3)), 0.7, (0.3, 1.0, 0.2)), (1.3, 0.7, 1.9))
    m2 = glm.rotate(glm.translate(glm.identity(), (-0.9, 0.0, 0.5)), -1.1, (1.0, 0.2, 0.0))
    scene.set_instances([(m1, t), (m2, t)])
    P, V = scenes.camera(512, 384, eye=(1.8, -1.1, 1.5))
    scene.check_primary(P, V, 512, 384, what="rotated/scaled")


def test_instances_straddling_the_eye_plane(scene, assets):
    """Screen rectangles of proxy cubes that reach behind the eye plane (instance_setup_kernel cuts them at w = eps instead
    of giving up on the whole screen): cameras in the middle of the entity grid, next to a cube face (whole-screen
    fallback), below the grid looking up, and a rotated / stretched instance that passes the camera on one side.  Primary
    records, depth and colour as well as path-traced radiance sums must still be the oracle's, bit for bit."""
    a = scene.add(assets["Treasure"])
    b = scene.add(assets["AncientTemple"])
    models = _grid_models(a, b)
    long_box = glm.scale(glm.rotate(glm.translate(glm.identity(), (0.6, -3.2, 0.4)), 0.5, (0.2, 1.0, 0.3)), (9.0, 0.8, 0.6))
    models.append((long_box, b))
    scene.set_instances(models)
    w, h = 320, 200
    P = glm.perspective(glm.REFERENCE_FOV, w / h, glm.REFERENCE_NEAR, glm.REFERENCE_FAR)
    cams = [((0.0, -4.2, 0.0), (4.0, -5.2, 3.0)),      # just above the cubes, grazing view across the grid
            ((0.75, -5.0, 0.75), (6.0, -5.0, 0.8)),    # between four cubes, at their height
            ((0.0, -5.0, 0.52), (0.0, -5.0, 3.0)),     # 0.02 in front of a cube face
            ((2.0, -8.0, 1.0), (0.0, -3.0, 0.0)),      # on the other side of the grid, looking back through it
            ((0.6, -3.2, 1.2), (0.7, -3.1, -4.0))]     # the long box passes on both sides of the camera
    for i, (eye, center) in enumerate(cams):
        V = glm.look_at(eye, center, (0.0, 1.0, 0.0))
        scene.check_primary(P, V, w, h, what=f"straddling cubes, camera {i}, primary")
        scene.check_paths(P, V, w, h, spp=2, bounces=3, what=f"straddling cubes, camera {i}, paths")


# ---- large-scene extension: shadow rays, procedural brick volumes, incoherent rays --------------

def test_shadow_rays_on_dense_volumes(scene, assets):
    t = scene.add(assets["AncientTemple"])
    scene.set_instances([(glm.identity(), t)])
    P, V = scenes.camera(640, 360, eye=(0.8, -0.45, 0.6))
    got, st = scene.check_primary(P, V, 640, 360, flags=abi.FLAG_SHADOW_RAYS, what="shadow rays, temple")
    traced = (got["packed"] >> 22) & 1
    lit = (got["packed"] >> 23) & 1
    assert traced.sum() > 1000 and 0 < lit.sum() < traced.sum()
    a = scene.add(assets["Treasure"])
    scene.set_instances([(glm.translate(glm.identity(), (0.7 * i - 0.7, 0.0, -0.9 * i)), a if i % 2 else t) for i in range(3)])
    scene.check_primary(P, V, 640, 360, flags=abi.FLAG_SHADOW_RAYS, what="shadow rays, three instances")


class ProcScene(Scene):
    def add_procedural(self, kind, w, h, d, seed):
        rid = self.r.add_volume_procedural(kind, w, h, d, seed)
        oid = self.o.add_volume_procedural(kind, w, h, d, seed)
        self.tex_map[oid] = rid
        return oid

    def check_rays(self, w, h, seed, n_check, what=""):
        self.r.configure(width=w, height=h, mode=abi.MODE_RAYS, flags=0, seed=seed, sample_first=0, max_frames=0)
        P, V = scenes.camera(w, h)
        assert self.r.render_tick_raw(P, V)
        got = self.r.read_hits().reshape(-1)[:n_check]
        color = self.r.read_color().reshape(-1, 4)[:n_check]
        want, wcolor, iters = self.o.render_rays(n_check, seed)
        for field in ("hit_voxel", "packed", "instance", "iters"):
            bad = np.flatnonzero(got[field] != want[field])
            assert len(bad) == 0, f"{what}: {field} differs for {len(bad)} rays, first ray {bad[:1]}"
        assert np.array_equal(color, wcolor), f"{what}: colours differ"
        st = self.r.stats()
        if n_check == w * h:
            assert st.iterations == iters
        return got, st


@pytest.fixture()
def pscene(renderer, oracle):
    renderer.reset()
    return ProcScene(renderer, oracle)


def test_heightmap_brick_volume_small(pscene):
    t = pscene.add_procedural(abi.VOLUME_HEIGHTMAP, 256, 128, 192, 1)
    pscene.set_instances([(glm.identity(), t)])
    for eye in [(0.9, -0.8, 0.9), (-0.6, -0.3, 1.1)]:
        P, V = scenes.camera(480, 270, eye=eye)
        got, _ = pscene.check_primary(P, V, 480, 270, flags=abi.FLAG_SHADOW_RAYS, what=f"heightmap 256x128x192 {eye}")
        assert (got["hit_voxel"] != abi.VT_MISS).sum() > 5000


def test_brick_and_dense_volumes_in_one_scene(pscene, assets):
    b = pscene.add_procedural(abi.VOLUME_SPARSE_BRICKS, 64, 64, 64, 2)
    d = pscene.add(assets["Treasure"])
    h = pscene.add_procedural(abi.VOLUME_HEIGHTMAP, 128, 128, 128, 9)
    pscene.set_instances([(glm.translate(glm.identity(), (-1.1, 0.0, 0.0)), b), (glm.identity(), d),
                          (glm.translate(glm.identity(), (1.1, 0.0, 0.2)), h)])
    P, V = scenes.camera(640, 360, eye=(0.4, -1.0, 2.4))
    got, _ = pscene.check_primary(P, V, 640, 360, flags=abi.FLAG_SHADOW_RAYS, what="bricks + dense")
    assert set(np.unique(got["instance"])) >= {0, 1, 2}


def test_uploaded_brick_volume(pscene):
    """vt_add_volume_bricks: caller-supplied sparse 8^3 bricks (coordinates, 512-bit masks, one colour each)."""
    rng = np.random.default_rng(5)
    w, h, d = 128, 64, 96
    all_coords = np.array([(x, y, z) for z in range(d // 8) for y in range(h // 8) for x in range(w // 8)], dtype=np.uint32)
    pick = rng.random(len(all_coords)) < 0.12
    coords = all_coords[pick]
    masks = rng.integers(0, 2**32, size=(len(coords), 16), dtype=np.uint64).astype(np.uint32) & \
        rng.integers(0, 2**32, size=(len(coords), 16), dtype=np.uint64).astype(np.uint32)   # ~25 % of the voxels
    colors = rng.integers(0, 256, size=(len(coords), 4), dtype=np.uint8)
    rid = pscene.r.add_volume_bricks(coords, masks, colors, w, h, d)
    oid = pscene.o.add_volume_bricks(coords, masks, colors, w, h, d)
    pscene.tex_map[oid] = rid
    pscene.set_instances([(glm.identity(), oid)])
    P, V = scenes.camera(480, 270, eye=(0.9, -0.7, 0.9))
    got, _ = pscene.check_primary(P, V, 480, 270, flags=abi.FLAG_SHADOW_RAYS, what="uploaded bricks")
    assert (got["hit_voxel"] != abi.VT_MISS).sum() > 3000
    pscene.check_rays(128, 64, seed=11, n_check=128 * 64, what="uploaded bricks, rays")
    with pytest.raises(RuntimeError):
        pscene.r.add_volume_bricks(np.array([[99, 0, 0]]), masks[:1], colors[:1], w, h, d)


def test_incoherent_rays_small(pscene, assets):
    t = pscene.add_procedural(abi.VOLUME_SPARSE_BRICKS, 512, 512, 512, 2)
    pscene.set_instances([(glm.identity(), t)])
    got, st = pscene.check_rays(256, 128, seed=3, n_check=256 * 128, what="sparse 512^3 rays")
    assert 0.2 < (got["hit_voxel"] != abi.VT_MISS).mean() < 0.8
    # the same mode over a dense volume (stop mask in global memory)
    pscene.r.reset()
    ps = ProcScene(pscene.r, __import__("oracle_lib"))
    d = ps.add(assets["AncientTemple"])
    ps.set_instances([(glm.identity(), d)])
    ps.check_rays(128, 128, seed=7, n_check=128 * 128, what="dense volume rays")


def test_config3_heightmap_1024_4k_subset(pscene):
    """BASELINE configs[3]: 1024^3 heightmap terrain, 3840x2160 primary + shadow rays; oracle parity on a
    1/64 pixel subset."""
    t = pscene.add_procedural(abi.VOLUME_HEIGHTMAP, 1024, 1024, 1024, 1)
    pscene.set_instances([(glm.identity(), t)])
    P, V = scenes.camera(3840, 2160, eye=(0.9, -0.8, 0.9))
    got, st = pscene.check_primary(P, V, 3840, 2160, flags=abi.FLAG_SHADOW_RAYS, what="config 3", subset=True)
    assert (got["hit_voxel"] != abi.VT_MISS).sum() > 2000 and st.rays > 3840 * 2160


def test_config4_sparse_4096_first_rays(pscene):
    """BASELINE configs[4]: sparse 4096^3 brick volume, incoherent rays; oracle parity on the first 2^16
    rays of a 2^22-ray launch (the full 2^26-ray launch is a benchmark, tools/run_config.py)."""
    t = pscene.add_procedural(abi.VOLUME_SPARSE_BRICKS, 4096, 4096, 4096, 2)
    pscene.set_instances([(glm.identity(), t)])
    got, st = pscene.check_rays(2048, 2048, seed=3, n_check=1 << 16, what="config 4")
    assert st.rays == 2048 * 2048 and st.iterations > (1 << 22) * 50


# ---- path-tracing extension -----------------------------------------------------------------

def test_paths_single_instance_exact(scene, assets):
    t = scene.add(assets["AncientTemple"])
    scene.set_instances([(glm.identity(), t)])
    P, V = scenes.camera(320, 180, eye=(0.8, -0.45, 0.6))
    got, st = scene.check_paths(P, V, 320, 180, spp=4, what="paths temple")
    assert st.rays > 320 * 180 * 4


def test_paths_generic_kernel_on_single_instance(scene, assets):
    """The general (multi-instance) kernel and the wavefront single-instance kernel are two
    schedules of the same paths: both must reproduce the oracle bit for bit."""
    t = scene.add(assets["Treasure"])
    scene.set_instances([(glm.identity(), t)])
    P, V = scenes.camera(256, 160, eye=(0.9, -0.5, 0.7))
    a, _ = scene.check_paths(P, V, 256, 160, spp=5, flags=abi.FLAG_PER_PIXEL_PATHS, what="general per-pixel kernel")
    w, _ = scene.check_paths(P, V, 256, 160, spp=5, flags=0, what="wavefront kernel")
    g, _ = scene.check_paths(P, V, 256, 160, spp=5, flags=abi.FLAG_FORCE_GLOBAL_MASKS, what="wavefront kernel, global masks")
    assert np.array_equal(a, w) and np.array_equal(a, g)
    b, _ = scene.check_paths(P, V, 256, 160, spp=5, flags=abi.FLAG_PER_PIXEL_PATHS | abi.FLAG_FORCE_GLOBAL_MASKS,
                             what="general per-pixel kernel, global masks")
    assert np.array_equal(a, b)


def test_paths_many_samples_chunking(scene, assets):
    """spp > 128: an item's samples are chunked so that its 32-bit per-pixel sums cannot wrap."""
    t = scene.add(assets["AncientTemple"])
    scene.set_instances([(glm.identity(), t)])
    P, V = scenes.camera(96, 64, eye=(0.8, -0.45, 0.6))
    scene.check_paths(P, V, 96, 64, spp=150, bounces=2, flags=abi.FLAG_PER_PIXEL_PATHS, what="150 spp, per-pixel kernel")
    scene.check_paths(P, V, 96, 64, spp=150, bounces=2, what="150 spp")


def test_paths_axis_aligned_and_inside(scene, assets):
    t = scene.add(assets["AncientTemple"])
    scene.set_instances([(glm.identity(), t)])
    P, V = scenes.camera(161, 121, eye=(0.0, 0.0, 2.0))
    scene.check_paths(P, V, 161, 121, spp=2, what="paths axis aligned")
    scene.check_paths(P, V, 161, 121, spp=2, flags=abi.FLAG_PER_PIXEL_PATHS, what="paths axis aligned, per-pixel kernel")
    P, V = scenes.camera(160, 120, eye=(0.1, 0.2, -0.1), center=(1.0, 0.3, 0.2))
    scene.check_paths(P, V, 160, 120, spp=2, what="paths camera inside")


def test_paths_degenerate_scenes(scene, assets):
    """No instance, an out-of-range texture id, zero bounces, one sample: every path kernel must agree with the oracle."""
    P, V = scenes.camera(96, 64)
    scene.set_instances([])
    scene.check_paths(P, V, 96, 64, spp=2, what="paths, empty scene")
    scene.set_instances([(glm.identity(), 40000)])
    scene.check_paths(P, V, 96, 64, spp=2, what="paths, bad texture id")
    t = scene.add(assets["Treasure"])
    scene.set_instances([(glm.identity(), t)])
    P, V = scenes.camera(96, 64, eye=(0.9, -0.5, 0.7))
    for flags in (0, abi.FLAG_PER_PIXEL_PATHS):
        scene.check_paths(P, V, 96, 64, spp=1, bounces=0, flags=flags, what=f"paths, 0 bounces, flags {flags}")
        scene.check_paths(P, V, 97, 63, spp=3, bounces=1, flags=flags, what=f"paths, odd size, flags {flags}")


def test_paths_multi_instance_exact(scene, assets):
    a = scene.add(assets["Treasure"])
    b = scene.add(assets["AncientTemple"])
    models = [(glm.translate(glm.identity(), (1.2 * i - 1.2, 0.0, 1.1 * j - 0.5)), a if (i + j) % 2 else b)
              for i in range(3) for j in range(2)]
    scene.set_instances(models)
    P, V = scenes.camera(240, 160, eye=(2.2, -1.4, 2.4))
    scene.check_paths(P, V, 240, 160, spp=3, bounces=3, what="paths multi")


def _grid_models(a, b):
    models = []
    for m in scenes.entity_grid(a, b):
        m = m.reshape(4, 4).copy()
        tid = int(m.reshape(16).view(np.uint32)[15])
        m[3][3] = 1.0
        models.append((m, tid))
    return models


def test_default_world_primary_and_paths(scene):
    """The reference's default world (src/world.rs:143-198): the 11x11 entity grid plus every non-empty 16^3 chunk of the
    paged terrain (tools/scenes.default_world: 343 instances, 224 textures, masks read from global memory) — primary
    frame and path-traced frame against the oracle, from outside and from the reference's start position inside chunk 0."""
    textures, inst = scenes.default_world()
    ids = [scene.add(c) for c in textures]
    scene.set_instances([(m, ids[t]) for m, t in inst])
    for eye, center in (((9.0, -9.0, 7.0), (0.0, -2.0, 0.0)), ((0.0, 0.0, 0.0), (1.0, 0.0, 0.0))):
        P, V = scenes.camera(400, 400, eye=eye, center=center)
        got, st = scene.check_primary(P, V, 400, 400, what=f"default world from {eye}")
        assert st.masks_in_smem == 0
        if eye[0] > 1.0:
            assert len(np.unique(got["instance"][got["hit_voxel"] != abi.VT_MISS])) > 150
    P, V = scenes.camera(200, 200, eye=(9.0, -9.0, 7.0), center=(0.0, -2.0, 0.0))
    scene.check_paths(P, V, 200, 200, spp=2, bounces=3, what="default world paths")


def test_paths_world_grid_entity_grid(scene, assets):
    """Bounce rays over the 11x11 entity grid go through the world-space instance grid (world_grid.cuh): the
    result must be what the oracle's loop over every instance gives, and what the same kernel gives with the
    grid (and the screen bins) switched off."""
    a = scene.add(assets["Treasure"])
    b = scene.add(assets["AncientTemple"])
    scene.set_instances(_grid_models(a, b))
    P = glm.perspective(glm.REFERENCE_FOV, 240 / 136, glm.REFERENCE_NEAR, glm.REFERENCE_FAR)
    V = glm.look_at((0.0, -2.0, 0.0), (3.0, -5.0, 2.0), (0.0, 1.0, 0.0))
    got, st = scene.check_paths(P, V, 240, 136, spp=3, bounces=4, what="entity grid paths, world grid")
    plain, _ = scene.check_paths(P, V, 240, 136, spp=3, bounces=4, flags=abi.FLAG_NO_BINNING, what="entity grid paths, every instance")
    assert np.array_equal(got, plain)
    assert st.rays > 240 * 136 * 3


def test_paths_world_grid_mixed_sizes(scene, assets):
    """Instances of very different sizes, rotated and scaled, touching and overlapping: a large slab spans many
    cells, small volumes sit on it and inside each other's boxes."""
    rng = np.random.default_rng(11)
    t = scene.add(assets["Treasure"])
    u = scene.add(RawVolume(make_volume(rng, 12, 9, 14, fill=0.45), 12, 9, 14))
    models = [(glm.scale(glm.translate(glm.identity(), (0.0, 0.9, 0.0)), (9.0, 0.6, 9.0)), u)]  # the slab (+Y is down)
    for i in range(14):
        pos = (float(rng.uniform(-3.5, 3.5)), float(rng.uniform(-0.6, 0.3)), float(rng.uniform(-3.5, 3.5)))
        m = glm.translate(glm.identity(), pos)
        m = glm.rotate(m, float(rng.uniform(-1.5, 1.5)), (float(rng.uniform(-1, 1)), 1.0, float(rng.uniform(-1, 1))))
        m = glm.scale(m, (float(rng.uniform(0.4, 1.6)), float(rng.uniform(0.4, 1.6)), float(rng.uniform(0.4, 1.6))))
        models.append((m, t if i % 2 else u))
    models.append((glm.translate(glm.identity(), (0.25, -0.2, 0.1)), t))   # overlaps its neighbours' boxes
    models.append((glm.translate(glm.identity(), (0.25, -0.2, 0.1)), u))   # coincident boxes: ties go to the lower index
    scene.set_instances(models)
    P, V = scenes.camera(200, 150, eye=(4.5, -3.0, 5.0))
    got, _ = scene.check_paths(P, V, 200, 150, spp=3, bounces=4, what="mixed sizes, world grid")
    plain, _ = scene.check_paths(P, V, 200, 150, spp=3, bounces=4, flags=abi.FLAG_NO_BINNING, what="mixed sizes, every instance")
    assert np.array_equal(got, plain)


def test_paths_world_grid_overflow_falls_back(renderer, oracle, assets, monkeypatch):
    """A list that is too small makes the frame loop over every instance (still exact); the next frame has room."""
    monkeypatch.setenv("VT_WORLD_CAP", "50")
    renderer.reset()
    sc = Scene(renderer, oracle)
    a = sc.add(assets["Treasure"])
    b = sc.add(assets["AncientTemple"])
    sc.set_instances(_grid_models(a, b))
    P = glm.perspective(glm.REFERENCE_FOV, 160 / 90, glm.REFERENCE_NEAR, glm.REFERENCE_FAR)
    V = glm.look_at((0.0, -2.0, 0.0), (3.0, -5.0, 2.0), (0.0, 1.0, 0.0))
    first, _ = sc.check_paths(P, V, 160, 90, spp=2, bounces=3, what="world grid overflow, first frame")
    second, _ = sc.check_paths(P, V, 160, 90, spp=2, bounces=3, what="world grid overflow, second frame")
    assert np.array_equal(first, second)


def test_paths_over_brick_volumes(pscene, assets):
    """Path tracing through procedural / uploaded brick volumes, alone and next to a dense volume (the general
    path kernel marches every kind of volume through march_instance())."""
    h = pscene.add_procedural(abi.VOLUME_HEIGHTMAP, 128, 64, 128, 4)
    pscene.set_instances([(glm.identity(), h)])
    P, V = scenes.camera(240, 136, eye=(0.9, -0.8, 0.9))
    got, st = pscene.check_paths(P, V, 240, 136, spp=3, bounces=4, what="paths, heightmap bricks")
    assert st.iterations > 240 * 136 * 3
    b = pscene.add_procedural(abi.VOLUME_SPARSE_BRICKS, 64, 64, 64, 2)
    d = pscene.add(assets["Treasure"])
    pscene.set_instances([(glm.translate(glm.identity(), (-1.1, 0.0, 0.0)), b), (glm.identity(), d),
                          (glm.translate(glm.identity(), (1.1, 0.0, 0.2)), h)])
    P, V = scenes.camera(240, 136, eye=(0.4, -1.0, 2.4))
    pscene.check_paths(P, V, 240, 136, spp=3, bounces=3, what="paths, bricks + dense + heightmap")
    rng = np.random.default_rng(8)
    coords = np.array([(x, y, z) for z in range(4) for y in range(4) for x in range(4)], dtype=np.uint32)[rng.random(64) < 0.4]
    masks = rng.integers(0, 2**32, size=(len(coords), 16), dtype=np.uint64).astype(np.uint32)
    colors = rng.integers(0, 256, size=(len(coords), 4), dtype=np.uint8)
    rid = pscene.r.add_volume_bricks(coords, masks, colors, 32, 32, 32)
    oid = pscene.o.add_volume_bricks(coords, masks, colors, 32, 32, 32)
    pscene.tex_map[oid] = rid
    pscene.set_instances([(glm.identity(), oid)])
    P, V = scenes.camera(160, 120, eye=(0.9, -0.7, 0.9))
    pscene.check_paths(P, V, 160, 120, spp=4, bounces=4, what="paths, uploaded bricks")


@pytest.mark.parametrize("seed", [101, 202, 303, 404, 505, 606, 707, 808])
def test_random_scenes_primary_and_paths(scene, assets, seed):
    """Randomised scenes: 1-12 instances with random rotations, non-uniform scales and overlaps, random
    (non-cubic) volumes next to the assets, random cameras incl. ones inside a volume; primary records,
    shadow rays and radiance sums must all equal the oracle's."""
    rng = np.random.default_rng(seed)
    texs = [scene.add(assets["Treasure"]), scene.add(assets["AncientTemple"])]
    for _ in range(2):
        w, h, d = (int(rng.integers(3, 40)) for _ in range(3))
        texs.append(scene.add(RawVolume(make_volume(rng, w, h, d, fill=float(rng.uniform(0.05, 0.6))), w, h, d)))
    for trial in range(3):
        n = int(rng.integers(1, 13))
        models = []
        for _ in range(n):
            m = glm.translate(glm.identity(), tuple(float(x) for x in rng.uniform(-1.6, 1.6, 3)))
            m = glm.rotate(m, float(rng.uniform(-3.1, 3.1)), tuple(float(x) for x in rng.uniform(-1, 1, 3) + np.array([0, 1.5, 0])))
            m = glm.scale(m, tuple(float(x) for x in rng.uniform(0.3, 1.8, 3)))
            models.append((m, texs[int(rng.integers(0, len(texs)))]))
        scene.set_instances(models)
        eye = tuple(float(x) for x in rng.uniform(-3.0, 3.0, 3))
        center = tuple(float(x) for x in rng.uniform(-0.5, 0.5, 3))
        w, h = int(rng.integers(60, 260)), int(rng.integers(40, 180))
        P, V = scenes.camera(w, h, eye=eye, center=center)
        what = f"random scene seed {seed} trial {trial} ({n} instances, {w}x{h})"
        scene.check_primary(P, V, w, h, flags=abi.FLAG_SHADOW_RAYS if trial == 1 else 0, what=what)
        scene.check_paths(P, V, w, h, spp=int(rng.integers(1, 5)), bounces=int(rng.integers(0, 5)), what=what + " paths")


def test_paths_sample_sharding_is_exact(renderer, scene, assets):
    """spp split over ranks: integer accumulation makes 1-rank == sum of N ranks."""
    t = scene.add(assets["AncientTemple"])
    scene.set_instances([(glm.identity(), t)])
    w, h, spp, n = 200, 120, 8, 4
    P, V = scenes.camera(w, h, eye=(0.8, -0.45, 0.6))
    whole, _ = scene.check_paths(P, V, w, h, spp=spp, what="whole")
    total = np.zeros_like(whole)
    for rank in range(n):
        renderer.configure(width=w, height=h, mode=abi.MODE_PATHS, spp=spp // n, bounces=4, seed=0x5EED,
                           sample_first=rank, sample_stride=n, total_spp=spp)
        renderer.clear_accum()
        renderer.render_async(P, V)
        total += renderer.read_accum()
    assert np.array_equal(total, whole)


def test_paths_lean_frames_moving_camera(renderer, scene, assets):
    """render_tick on a single-instance scene only clears / resolves the instance's screen rectangle; the sums outside it
    are written when somebody asks (vt_read_accum), also after the rectangle moved and after other kinds of frames."""
    t = scene.add(assets["AncientTemple"])
    scene.set_instances([(glm.identity(), t)])
    w, h, spp = 320, 200, 3
    cams = [scenes.camera(w, h, eye=e, center=c) for e, c in (((1.6, -0.9, 1.2), (0.0, 0.0, 0.0)), ((2.2, -0.4, 0.3), (0.6, 0.2, 0.0)),
                                                              ((0.9, -1.4, -1.1), (-0.3, 0.0, 0.2)))]
    renderer.configure(width=w, height=h, mode=abi.MODE_PATHS, spp=spp, bounces=4, seed=0x5EED, sample_first=0, sample_stride=1,
                       total_spp=spp, max_frames=0)
    for P, V in cams:  # three frames, nothing read in between: stale sums of the earlier rectangles are lying around
        assert renderer.render_tick_raw(P, V)
    color = renderer.read_color()
    want, _, _ = scene.o.render_paths(*cams[-1], w, h, spp=spp, bounces=4, seed=0x5EED, flags=0)
    import oracle_lib
    assert np.array_equal(color, oracle_lib.resolve(want, spp))
    assert np.array_equal(renderer.read_accum(), want)
    # a primary frame replaces the instance uniforms; the sums must have been completed before
    assert renderer.render_tick_raw(*cams[0])
    renderer.configure(width=w, height=h, mode=abi.MODE_PRIMARY, max_frames=0)
    assert renderer.render_tick_raw(*cams[1])
    want0, _, _ = scene.o.render_paths(*cams[0], w, h, spp=spp, bounces=4, seed=0x5EED, flags=0)
    assert np.array_equal(renderer.read_accum(), want0)
    # accumulating frames (the multi-GPU building block) on top of a lean frame
    renderer.configure(width=w, height=h, mode=abi.MODE_PATHS, spp=spp, bounces=4, seed=0x5EED, sample_first=0, sample_stride=1,
                       total_spp=spp, max_frames=0)
    assert renderer.render_tick_raw(*cams[2])
    renderer.render_async(*cams[2])
    want2, _, _ = scene.o.render_paths(*cams[2], w, h, spp=spp, bounces=4, seed=0x5EED, flags=0)
    assert np.array_equal(renderer.read_accum(), want2 * np.uint64(2))


def test_paths_tiles_with_a_single_covered_pixel(scene, assets):
    """A screen rectangle whose corner is the last pixel of an 8x4 tile leaves that tile with ONE covered pixel: the
    wavefront kernel's job -> (sample, pixel) split divides by the number of covered pixels with a 32-bit magic number,
    and the magic number of 1 does not fit (found by tests/multi_gpu_check.py; every sample became sample 0).  The first
    camera puts the rectangle's corner at pixel (199, 23) of a 640x360 frame; the sweep covers other corner alignments."""
    t = scene.add(assets["AncientTemple"])
    scene.set_instances([(glm.identity(), t)])
    P, V = scenes.camera(640, 360, eye=(1.1, 0.4, -0.9))
    scene.check_paths(P, V, 640, 360, spp=5, bounces=4, what="corner tile with one covered pixel")
    for i in range(24):
        a = 0.26 * i
        eye = (1.9 * np.cos(a) + 0.013 * i, -0.7 + 0.05 * i, 1.9 * np.sin(a))
        P, V = scenes.camera(168, 100, eye=eye)
        scene.check_paths(P, V, 168, 100, spp=3, bounces=3, what=f"corner alignment sweep, camera {i}")


def test_fused_accumulation_data_path_with_one_rank(renderer, scene, assets):
    """The multi-GPU data path (vt_fused_reduce_*) with a world of one rank: the root's sums never travel — its summation
    kernel takes them from the local accumulators, parks them in its slot for later calls on the same frame and clears
    them — for both ways of sharing a frame, compact and wide layout, with the camera moving between frames.  (Several
    ranks: tests/multi_gpu_check.py under torchrun.)"""
    import oracle_lib
    t = scene.add(assets["AncientTemple"])
    scene.set_instances([(glm.identity(), t)])
    w, h = 320, 200
    cams = [scenes.camera(w, h, eye=e) for e in ((1.6, -0.9, 1.2), (0.8, -0.45, 0.6), (1.1, 0.4, -0.9))]
    for spp in (3, 260):
        renderer.configure(width=w, height=h, mode=abi.MODE_PATHS, spp=spp, bounces=3, seed=0x5EED, sample_first=0, sample_stride=1,
                           total_spp=spp, max_frames=0)
        wants = [scene.o.render_paths(P, V, w, h, spp=spp, bounces=3, seed=0x5EED, flags=0) for P, V in (cams if spp == 3 else cams[:1])]
        renderer.fused_reduce_export(1)
        try:
            for by_rows in (False, True):
                renderer.fused_reduce_partition(by_rows, 2 if by_rows else 0, 8)
                for (P, V), (want, rays, iters) in zip(cams, wants):
                    renderer.fused_reduce_next_frame()
                    renderer.render_async(P, V)
                    renderer.resolve()
                    assert np.array_equal(renderer.read_color(), oracle_lib.resolve(want, spp))
                    assert np.array_equal(renderer.read_accum(), want)   # second summation of the frame: from the slot
                    renderer.resolve()
                    assert np.array_equal(renderer.read_color(), oracle_lib.resolve(want, spp))
                    st = renderer.stats()
                    assert (st.rays, st.iterations) == (rays, iters)
            # the protocol: the root sums every frame before it starts the next
            renderer.fused_reduce_next_frame()
            renderer.render_async(*cams[0])
            renderer.fused_reduce_next_frame()
            with pytest.raises(RuntimeError, match="for every frame"):
                renderer.render_async(*cams[0])
            renderer.resolve()
            # ... and every frame has its own sequence number
            renderer.render_async(*cams[0])      # (uses the number the refused frame did not consume)
            renderer.resolve()
            with pytest.raises(RuntimeError, match="vt_fused_reduce_next_frame"):
                renderer.render_async(*cams[0])
            with pytest.raises(RuntimeError, match="own buffer"):
                renderer.set_accum_buffer(1 << 20)   # (any non-null pointer: refused before it is ever used)
        finally:
            renderer.fused_reduce_disable()


def test_pipelined_frames_and_async_readback(renderer, scene, assets):
    """vt_render_frame_async + vt_read_color_async: three frames in flight with different cameras land in their own host
    buffers, identical to what render_tick + vt_read_color give one by one; non-pinned destinations are refused."""
    import torch
    t = scene.add(assets["Treasure"])
    scene.set_instances([(glm.identity(), t)])
    w, h, spp = 256, 144, 2
    cams = [scenes.camera(w, h, eye=e) for e in ((1.6, -0.9, 1.2), (0.8, -0.45, 0.6), (-1.2, -0.7, 1.5))]
    renderer.configure(width=w, height=h, mode=abi.MODE_PATHS, spp=spp, bounces=3, seed=7, sample_first=0, sample_stride=1,
                       total_spp=spp, max_frames=0)
    want = []
    for P, V in cams:
        assert renderer.render_tick_raw(P, V)
        want.append(renderer.read_color().copy())
    bufs = [torch.zeros((h, w, 4), dtype=torch.uint8, pin_memory=True).numpy() for _ in cams]
    for (P, V), b in zip(cams, bufs):
        renderer.render_frame_async(P, V)
        renderer.read_color_async(b)
    renderer.read_color_wait()
    for b, c in zip(bufs, want):
        assert np.array_equal(b, c)
    assert np.array_equal(renderer.read_color(), want[-1])  # the synchronous read still sees the newest frame
    with pytest.raises(RuntimeError):
        renderer.read_color_async(np.zeros((h, w, 4), dtype=np.uint8))


def test_config2_full_size_exact(renderer, scene, assets):
    """configs[2] at full size (1080p, 64 spp, 4 bounces) — the bench workload itself — against the
    oracle bit for bit (radiance sums, ray and iteration counters, resolved colour), plus the
    size-independent properties: determinism, ray-count bounds, energy bound, sky pixels exact."""
    t = scene.add(assets["AncientTemple"])
    scene.set_instances([(glm.identity(), t)])
    w, h, spp = 1920, 1080, 64
    P, V = scenes.camera(w, h)
    a1, st1 = scene.check_paths(P, V, w, h, spp=spp, bounces=4, seed=0x5EED, what="configs[2] full size")
    assert renderer.render_tick_raw(P, V)
    a2, st2 = renderer.read_accum(), renderer.stats()
    assert np.array_equal(a1, a2) and st1.rays == st2.rays and st1.iterations == st2.iterations
    n = w * h * spp
    assert n <= st1.rays <= 5 * n
    sky = np.array([int(np.float32(c) / np.float32(100.0) * np.float32(16777216.0)) for c in (53.0, 81.0, 92.0)], dtype=np.uint64)
    assert (a1 <= sky * np.uint64(spp)).all()           # albedo <= 1: nothing brighter than the sky
    assert np.array_equal(a1[0, 0], sky * np.uint64(spp))  # a corner pixel sees only sky


def test_config2_full_size_closeup_exact(scene, assets):
    """The frame-filling camera of the bench's `secondary` block, full size, against the oracle."""
    t = scene.add(assets["AncientTemple"])
    scene.set_instances([(glm.identity(), t)])
    w, h = 1920, 1080
    P, V = scenes.camera(w, h, eye=(0.8, -0.45, 0.6))
    scene.check_paths(P, V, w, h, spp=16, bounces=4, seed=0x5EED, what="configs[2] close-up, 16 spp")


# ---- compiled host over the C ABI ----------------------------------------------------------------

@pytest.mark.parametrize("exe", ["vtrace_headless", "vtrace_headless_static"])
def test_compiled_host_drives_the_abi_like_the_engine(oracle, assets, tmp_path, exe):
    """host/vtrace_headless (C++ mirror of src/main.rs + src/render.rs + src/world.rs' entity grid) runs
    as its own process against librender.so — and, as vtrace_headless_static, with the static librender.a
    the reference's build script would link (build.rs:96-97) inside the executable: one texture upload per tick, instances skipped until their
    texture is resident, frame k rendered with frame k-1's pose.  Its last frame must hash to what the
    oracle renders from the same matrices."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "host")], check=True, capture_output=True)
    env = dict(os.environ, VT_WIDTH="640", VT_HEIGHT="360", VT_MAX_FRAMES="5", VT_MODE="0", VT_FLAGS="0")
    ppm = str(tmp_path / "frame.ppm")
    out = subprocess.run([os.path.join(root, "host", exe), scenes.ASSETS, ppm], env=env, capture_output=True,
                         text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    fields = dict(line.split("=", 1) for line in out.stdout.replace(" ", "\n").splitlines() if "=" in line)
    assert fields["frames"] == "6" and fields["size"] == "640x360"  # 5 rendered + the tick that returned -1
    P = np.array([int(x, 16) for x in fields["P"].split(",")], dtype=np.uint32).view(np.float32)
    V = np.array([int(x, 16) for x in fields["V"].split(",")], dtype=np.uint32).view(np.float32)
    sc = oracle.OracleScene()
    temple = sc.add_texture(assets["AncientTemple"].get_raw(), *assets["AncientTemple"].dims())   # uploaded first
    treasure = sc.add_texture(assets["Treasure"].get_raw(), *assets["Treasure"].dims())
    sc.set_instances(scenes.entity_grid(treasure, temple))
    rec, rgba, _, iters = sc.render_primary(P, V, 640, 360)
    h = 1469598103934665603
    for b in rgba.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert int(fields["iterations"]) == iters
    assert fields["fnv1a"] == f"{h:016x}"
    assert os.path.getsize(ppm) == len("P6\n640 360\n255\n") + 640 * 360 * 3
