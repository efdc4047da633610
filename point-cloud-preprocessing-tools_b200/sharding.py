"""Multi-GPU host logic of the path: frames are independent (BatchMultiBevGen.cpp:727-757 carries no
state between iterations), so they shard by index with NO data-path collective; label rows are split per GPU and
gathered on the host.  Callers: tools/bench_extra.py --sharded (BASELINE configs[2], [3] under torchrun: frame_shard for the
keyframe batch, row_split for the K x M label table) and tests/test_sharding_gloo.py.  The C++ CLI splits the label rows
with the same formula (host/batch_multi_bev_gen.cpp, label stage) and hands frames out per batch as workers become free."""


def frame_shard(n_frames, rank, world):
    """Contiguous block [lo, hi) of the sorted frame list owned by `rank`."""
    return n_frames * rank // world, n_frames * (rank + 1) // world


def row_split(K, world):
    """Label rows [r0, r1) per GPU (getKeyFrameLabel rows, BatchMultiBevGen.cpp:601-632)."""
    return [(K * g // world, K * (g + 1) // world) for g in range(world)]


def max_over_ranks(values, device=None):
    """Max over ranks of a list of floats (timed regions are reported as the slowest rank); identity when not distributed."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]
