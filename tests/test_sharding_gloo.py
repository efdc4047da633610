"""N>1 host logic on CPU: two gloo ranks (127.0.0.1) shard a batch by frame index and split the label rows, run their
share (with the CPU oracle standing in for the GPU — this test checks the sharding/gather logic, not the kernels),
gather on rank 0 and must reproduce the unsharded result exactly (order independence, SURVEY §4 multi-GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from _load_pkg import load_pkg, load_synth, load_oracle
    import importlib
    load_pkg()
    sh = importlib.import_module("pcpt_b200.sharding")
    synth, O = load_synth(), load_oracle()
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    FIELDS = ("x", "y", "z", "intensity", "row", "col", "label")
    n = 5
    frames = [synth.make_frame("HDL_32E", 40 + i) for i in range(n)]
    lo, hi = sh.frame_shard(n, rank, world)
    sp = O.sensor("HDL_32E")
    mine = [O.frame(sp, *[frames[i][k] for k in FIELDS]) for i in range(lo, hi)]
    xyz = synth.make_poses(90, seed=4)
    mi, _ = O.select_major(xyz)                      # the greedy scan is serial: every rank holds all majors
    r0, r1 = sh.row_split(len(xyz), world)[rank]
    lab_full, _, _ = O.labels(xyz, mi)
    part = lab_full[r0:r1]
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(lo=lo, hi=hi, frames=mine, r0=r0, r1=r1, labels=part))
    tmax = sh.max_over_ranks([float(rank + 1)])
    dist.barrier()
    if rank == 0:
        ok = tmax == [float(world)]
        ref = [O.frame(sp, *[frames[i][k] for k in FIELDS]) for i in range(n)]
        got = [None] * n
        for g in gathered:
            for j, i in enumerate(range(g["lo"], g["hi"])):
                got[i] = g["frames"][j]
        for i in range(n):
            for k in ("owner", "label", "single", "multi"):
                ok = ok and got[i] is not None and np.array_equal(got[i][k], ref[i][k])
        lab = np.concatenate([g["labels"] for g in sorted(gathered, key=lambda g: g["r0"])])
        ok = ok and np.array_equal(lab, lab_full) and sum(g["hi"] - g["lo"] for g in gathered) == n
        q.put(bool(ok))
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_shard_formulas():
    from _load_pkg import load_pkg
    import importlib
    load_pkg()
    sh = importlib.import_module("pcpt_b200.sharding")
    for n in (0, 1, 7, 100, 10000):
        for w in (1, 2, 4, 8):
            spans = [sh.frame_shard(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
            assert sh.row_split(n, w) == spans
