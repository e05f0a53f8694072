"""Batch-sharded int8 ResNet-50 on N GPUs (one process per GPU, `python -m torch.distributed.run --nproc-per-node N`):
with the DynamicQuantizeLinear range all-reduced over the ranks (rten_b200.Comm) the gathered pooled features (everything below the f32 classifier) must be
BIT-IDENTICAL to the unsharded CPU oracle's; without the exchange they are not (each shard would pick its own range)."""
import os

os.environ.setdefault("RTEN_B200_F32_MODE", "tf32")  # these tools measure the single-pass TF32 kernels unless told otherwise
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import rten_b200 as rt  # noqa: E402
from rten_b200 import graphs, shard  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    from oracle import oracle  # checker only
    import model_ref
    ctx = rt.Context(local)
    ids = [rt.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = rt.Comm(ctx, ids[0], rank, world)
    rng = oracle.XorShiftRng(5678)
    q = graphs.quantize_resnet50(graphs.make_resnet50(lambda s: rng.uniform(s)))
    per = 2
    x = oracle.XorShiftRng(99).uniform((per * world, 3, 224, 224))
    lo, hi = shard.shard_range(rank, world, per * world)
    xs = ctx.to_device(x[lo:hi], channels_last=True)
    runner = graphs.ResNet50Int8Runner(ctx, q, fuse=True, comm=comm)
    with_x = runner.run(xs, True)[1].numpy()
    without = graphs.ResNet50Int8Runner(ctx, q, fuse=True, comm=None).run(xs, True)[1].numpy()
    # the same exchange through NCCL (the fallback when the peers' memory cannot be opened)
    os.environ["RTEN_B200_NCCL_RANGES"] = "1"
    ids2 = [rt.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids2, src=0)
    comm_nccl = rt.Comm(ctx, ids2[0], rank, world)
    os.environ.pop("RTEN_B200_NCCL_RANGES")
    runner_nccl = graphs.ResNet50Int8Runner(ctx, q, fuse=True, comm=comm_nccl)
    with_nccl = runner_nccl.run(xs, True)[1].numpy()

    def step_us(r, reps=10):
        r.run(xs, True)
        ctx.sync()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            r.run(xs, True)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3 / reps

    t_peer, t_nccl = step_us(runner), step_us(runner_nccl)
    gathered = [None] * world
    dist.all_gather_object(gathered, (with_x, without, with_nccl, comm.uses_peer_memory, comm.timeouts(), comm_nccl.uses_peer_memory))
    ok = True
    if rank == 0:
        ref = model_ref.resnet50_int8_oracle(oracle, q, x)[1]  # pooled features: the exact-arithmetic part of the model
        got = np.concatenate([g[0] for g in gathered], 0)
        got_no = np.concatenate([g[1] for g in gathered], 0)
        got_nccl = np.concatenate([g[2] for g in gathered], 0)
        same = np.array_equal(got.view(np.int32), ref.view(np.int32))
        same_no = np.array_equal(got_no.view(np.int32), ref.view(np.int32))
        same_nccl = np.array_equal(got_nccl.view(np.int32), ref.view(np.int32))
        print(f"world {world}: sharded + range all-reduce bit-identical to the unsharded oracle: {same}; "
              f"without the exchange: {same_no} (max |d| {float(np.abs(got_no - ref).max()):.3e})", flush=True)
        print(f"exchange through peer memory on every rank: {all(g[3] for g in gathered)} (timeouts {sum(g[4] for g in gathered)}); "
              f"NCCL fallback bit-identical: {same_nccl} (peer memory in use there: {any(g[5] for g in gathered)}); "
              f"eager step (batch {per}/rank, 53 exchanges): {t_peer:.0f} us with peer mailboxes, {t_nccl:.0f} us with NCCL", flush=True)
        ok = same and same_nccl and sum(g[4] for g in gathered) == 0
    comm.close()
    comm_nccl.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
