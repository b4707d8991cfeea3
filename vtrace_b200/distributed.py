"""Multi-GPU plumbing for the path-tracing extension: one process per GPU, the
scene replicated, the samples of every pixel sharded over the ranks, the fixed-point accumulation
buffers summed with ONE all-reduce (NCCL on GPUs; gloo in the CPU tests).  No other collective:
rays never cross ranks.

Because the accumulators are integers (2^-24 fixed point), the reduced image is bit-identical for
any number of ranks and any reduction order.
"""
from __future__ import annotations


def shard_samples(total_spp: int, rank: int, world: int) -> tuple[int, int, int]:
    """(sample_first, sample_stride, sample_count) of `rank`: global samples s = rank (mod world).

    Works for any total_spp >= 0: the first total_spp % world ranks render one extra sample.
    """
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    count = total_spp // world + (1 if rank < total_spp % world else 0)
    return rank, world, count


def shard_samples_weighted(total_spp: int, rank: int, world: int, root_extra_spp: float = 2.0, root: int = 0) -> tuple[int, int, int]:
    """(sample_first, sample_stride, sample_count) for the fused accumulation, where the root also sums every rank's
    partial sums and encodes the frame: it traces fewer samples, so that all ranks finish a frame together.

    `root_extra_spp` is the root's extra work per frame expressed in samples per pixel (summation + resolve take about as
    long as tracing two samples of the bench frame).  Ranks take CONTIGUOUS sample ranges (stride 1): the root
    round((S - (N-1) e) / N) samples, the others share the rest as evenly as possible.  Every sample is traced exactly
    once, so the reduced image is the same bit pattern as for any other partition."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    if world == 1:
        return 0, 1, total_spp
    c_root = int(round((total_spp - (world - 1) * root_extra_spp) / world))
    c_root = max(0, min(total_spp, c_root))
    rest, others = total_spp - c_root, world - 1
    counts = []
    k = 0
    for r in range(world):
        if r == root:
            counts.append(c_root)
        else:
            counts.append(rest // others + (1 if k < rest % others else 0))
            k += 1
    first = sum(counts[:rank])
    return first, 1, counts[rank]


def reduce_accum(accum, group=None):
    """Sum the (h, w, 3) int64 accumulation tensor over all ranks, in place; returns it."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(accum, op=dist.ReduceOp.SUM, group=group)
    return accum


def setup_fused_reduce(renderer, rank: int, world: int, device, root: int = 0):
    """Wires the library's fused cross-GPU accumulation: the root exports the CUDA IPC handle of its
    partial-sum buffer, torch.distributed broadcasts the 64 bytes, every other rank maps it.  After
    this, each rank streams its sums of the covered rectangle straight into the root's memory over
    NVLink and raises a sequence-number flag there; the root's resolve waits for the flags and adds the
    slots up.  No collective on the data path (VT_FUSED_SYNC=0 switches the flags off; the caller then
    orders the ranks itself, e.g. with stream_barrier).

    Returns False (on every rank, with the fused path switched off again) if any rank could not set it up —
    the caller then falls back to reduce_accum()."""
    if root != 0:
        raise ValueError("the library's fused reduction uses rank 0 as the root")
    import torch
    import torch.distributed as dist

    handle = torch.zeros(64, dtype=torch.uint8, device=device)
    ok = torch.ones(1, dtype=torch.int32, device=device)
    err = None
    if rank == root:
        try:
            handle.copy_(torch.frombuffer(bytearray(renderer.fused_reduce_export(world)), dtype=torch.uint8))
        except RuntimeError as e:  # e.g. no memory for the partial-sum buffer
            err, ok[0] = e, 0
    dist.broadcast(handle, src=root)
    if rank != root:
        try:
            renderer.fused_reduce_import(bytes(handle.cpu().numpy().tobytes()), rank, world)
        except RuntimeError as e:  # e.g. no peer access between this GPU and the root's
            err, ok[0] = e, 0
    # all ranks or none: a rank that could not map the buffer must not leave the others waiting for its flag
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        if err is None:
            renderer.fused_reduce_disable()
        return False
    return True


def stream_barrier(flag):
    """A stream-ordered barrier: a one-element all-reduce completes on a rank only after every rank's
    stream reached it, i.e. after every rank's push kernel finished.  Only needed with VT_FUSED_SYNC=0."""
    import torch.distributed as dist

    dist.all_reduce(flag)


def row_owner(tile_row: int, world: int, relief_num: int = 0, relief_den: int = 8) -> int:
    """Owner of a row of 8x4-pixel tiles when a fused reduction shares frames by tile rows (vt_fused_reduce_partition; the
    device-side rule is RowShare in csrc/kernels.h).  Rows are dealt in cycles of c*(world-1) + (k-c)*world: the first c
    rounds of a cycle skip the root, the other k-c include it — the root owns k-c rows of a cycle, every other rank k."""
    if world <= 1:
        return 0
    c, k = relief_num, relief_den
    skip = c * (world - 1)
    q = tile_row % (skip + (k - c) * world)
    return 1 + q % (world - 1) if q < skip else (q - skip) % world


def rows_per_rank(n_tile_rows: int, world: int, relief_num: int = 0, relief_den: int = 8) -> list[int]:
    counts = [0] * max(world, 1)
    for ty in range(n_tile_rows):
        counts[row_owner(ty, world, relief_num, relief_den)] += 1
    return counts
