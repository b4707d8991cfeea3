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
