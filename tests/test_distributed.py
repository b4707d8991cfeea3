"""world_size-2 CPU test (gloo) of the multi-rank path: each rank traces its shard of the samples
(with the oracle standing in for the GPU), the int64 accumulators are all-reduced exactly as
bench.py does with NCCL, and every rank must end up with the single-rank image bit for bit."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, w, h, spp, out_path):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import oracle_lib
    from tools import scenes
    from vtrace_b200.distributed import reduce_accum, shard_samples

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chunk = scenes.load_asset("AncientTemple")
    sc = oracle_lib.OracleScene()
    sc.add_texture(chunk.get_raw(), *chunk.dims())
    sc.set_instances(scenes.single_instance(0))
    P, V = scenes.camera(w, h, eye=(0.8, -0.45, 0.6))
    first, stride, count = shard_samples(spp, rank, world)
    acc, rays, iters = sc.render_paths(P, V, w, h, spp=count, sample_first=first, sample_stride=stride, threads=2)
    t = torch.from_numpy(acc.view(np.int64).copy())
    reduce_accum(t)
    counters = torch.tensor([rays, iters], dtype=torch.int64)
    dist.all_reduce(counters)
    np.save(f"{out_path}.{rank}.npy", t.numpy())
    np.save(f"{out_path}.{rank}.counters.npy", counters.numpy())
    dist.destroy_process_group()


def test_two_ranks_reduce_to_the_single_rank_image(tmp_path, oracle, assets):
    from tools import scenes
    w, h, spp, world = 80, 45, 5, 2
    out = str(tmp_path / "acc")
    mp.spawn(_worker, args=(world, _free_port(), w, h, spp, out), nprocs=world, join=True)
    sc = oracle.OracleScene()
    sc.add_texture(assets["AncientTemple"].get_raw(), *assets["AncientTemple"].dims())
    sc.set_instances(scenes.single_instance(0))
    P, V = scenes.camera(w, h, eye=(0.8, -0.45, 0.6))
    whole, rays, iters = sc.render_paths(P, V, w, h, spp=spp)
    for rank in range(world):
        got = np.load(f"{out}.{rank}.npy").view(np.uint64)
        assert np.array_equal(got, whole)
        assert tuple(np.load(f"{out}.{rank}.counters.npy")) == (rays, iters)


class _FakeRenderer:
    """Stands in for vtrace_b200.renderer.Renderer in the fused-setup handshake (no GPU here)."""

    def __init__(self, fail_import):
        self.fail_import = fail_import
        self.calls = []

    def fused_reduce_export(self, world):
        self.calls.append(("export", world))
        return bytes(range(64))

    def fused_reduce_import(self, handle, rank, world):
        self.calls.append(("import", bytes(handle), rank, world))
        if self.fail_import:
            raise RuntimeError("no peer access")

    def fused_reduce_disable(self):
        self.calls.append(("disable",))


def _handshake_worker(rank, world, port, fail_rank, out_path):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import json

    import torch
    import torch.distributed as dist

    from vtrace_b200.distributed import setup_fused_reduce

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = _FakeRenderer(fail_import=(rank == fail_rank))
    ok = setup_fused_reduce(r, rank, world, torch.device("cpu"))
    with open(f"{out_path}.{rank}.json", "w") as f:
        json.dump({"ok": bool(ok), "calls": [c[0] for c in r.calls],
                   "handle_ok": all(c[1] == bytes(range(64)) for c in r.calls if c[0] == "import")}, f)
    dist.destroy_process_group()


def test_fused_setup_handshake_all_ranks_or_none(tmp_path):
    """The IPC-handle handshake of the fused accumulation: the root's 64 bytes reach every rank; if one rank
    cannot map the buffer, EVERY rank reports failure (and the ones that succeeded switch the path off again), so
    no rank is left waiting for a flag that will never be raised."""
    import json
    world = 2
    out = str(tmp_path / "ok")
    mp.spawn(_handshake_worker, args=(world, _free_port(), -1, out), nprocs=world, join=True)
    res = [json.load(open(f"{out}.{r}.json")) for r in range(world)]
    assert all(x["ok"] and x["handle_ok"] for x in res)
    assert res[0]["calls"] == ["export"] and res[1]["calls"] == ["import"]
    out = str(tmp_path / "fail")
    mp.spawn(_handshake_worker, args=(world, _free_port(), 1, out), nprocs=world, join=True)
    res = [json.load(open(f"{out}.{r}.json")) for r in range(world)]
    assert not any(x["ok"] for x in res)
    assert res[0]["calls"] == ["export", "disable"]   # the root had succeeded: it backs out
    assert res[1]["calls"] == ["import"]              # the failing rank has nothing to undo


def _rows_worker(rank, world, port, w, h, spp, relief, out_path):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import oracle_lib
    from tools import scenes
    from vtrace_b200.distributed import reduce_accum, row_owner

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chunk = scenes.load_asset("AncientTemple")
    sc = oracle_lib.OracleScene()
    sc.add_texture(chunk.get_raw(), *chunk.dims())
    sc.set_instances(scenes.single_instance(0))
    P, V = scenes.camera(w, h, eye=(0.8, -0.45, 0.6))
    # frames shared by rows of 8x4-pixel tiles (vt_fused_reduce_partition): a rank keeps every sample of the rows it owns
    acc, _, _ = sc.render_paths(P, V, w, h, spp=spp, threads=2)
    acc = acc.reshape(h, w, 3).copy()
    for y in range(h):
        if row_owner(y // 4, world, *relief) != rank:
            acc[y] = 0
    t = torch.from_numpy(acc.view(np.int64).copy())
    reduce_accum(t)
    np.save(f"{out_path}.{rank}.npy", t.numpy())
    dist.destroy_process_group()


def test_two_ranks_sharing_a_frame_by_tile_rows(tmp_path, oracle, assets):
    """The other way of sharing a frame: every pixel is traced (all samples) by exactly one rank — the owner of its row of
    tiles, the root relieved by 3/8 — so the sum over the ranks is the single-rank image."""
    from tools import scenes
    w, h, spp, world = 80, 45, 3, 2
    out = str(tmp_path / "rows")
    mp.spawn(_rows_worker, args=(world, _free_port(), w, h, spp, (3, 8), out), nprocs=world, join=True)
    sc = oracle.OracleScene()
    sc.add_texture(assets["AncientTemple"].get_raw(), *assets["AncientTemple"].dims())
    sc.set_instances(scenes.single_instance(0))
    P, V = scenes.camera(w, h, eye=(0.8, -0.45, 0.6))
    whole, _, _ = sc.render_paths(P, V, w, h, spp=spp)
    for rank in range(world):
        assert np.array_equal(np.load(f"{out}.{rank}.npy").view(np.uint64).reshape(whole.shape), whole)
