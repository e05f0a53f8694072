"""World-size-2 `gloo` test (CPU) of the N>1 path's host logic: batch sharding + all-gather of outputs must
reproduce the unsharded result.  The per-rank compute here is the CPU oracle (test infrastructure)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from oracle import oracle
    from rten_b200 import graphs, shard
    import model_ref

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = oracle.XorShiftRng(5678)
    spec = graphs.make_resnet50(lambda s: rng.uniform(s), num_classes=10, width_mult=0.125)
    x = oracle.XorShiftRng(1234).uniform((5, 3, 32, 32))  # odd batch: uneven shards
    lo, hi = shard.shard_range(rank, world, x.shape[0])
    y = model_ref.resnet50_oracle(oracle, spec, x[lo:hi])
    # uneven shards -> pad to the largest shard for the equal-size all-gather, then trim
    mx = max(shard.shard_range(r, world, x.shape[0])[1] - shard.shard_range(r, world, x.shape[0])[0] for r in range(world))
    ypad = np.zeros((mx,) + y.shape[1:], np.float32)
    ypad[: y.shape[0]] = y
    buf = torch.empty(shard.gather_layout(world, ypad.shape))
    shard.all_gather_outputs(dist, torch.from_numpy(ypad), buf)
    parts = []
    for r in range(world):
        a, b = shard.shard_range(r, world, x.shape[0])
        parts.append(buf[r * mx: r * mx + (b - a)].numpy())
    full = np.concatenate(parts)
    if rank == 0:
        ref = model_ref.resnet50_oracle(oracle, spec, x)
        q.put(bool(np.array_equal(full, ref)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range():
    from rten_b200 import shard
    for batch in (0, 1, 5, 32, 33):
        for world in (1, 2, 3, 8):
            r = [shard.shard_range(k, world, batch) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == batch
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(2, 2, 4)


def test_batch_shard_all_gather_matches_unsharded():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, "sharded + gathered logits differ from the unsharded run"
