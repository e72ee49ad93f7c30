"""world_size-2 `gloo` test of the N>1 path: slot sharding, the finished-game sample all-gather and the
counter all-reduce that bench.py runs over NCCL — same code, CPU tensors."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REC = 48


def _worker(rank, world, port, q):
    import az_b200_loader
    az_b200_loader.load()
    from alphazero_multi_game_b200 import gather as G
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = G.shard_slots(4097, world, rank)
        cap = 64
        n_local = [5, 0, 17][rank % 3] if rank != 1 else 9
        recs = torch.zeros(cap * REC, dtype=torch.uint8)
        arr = recs.numpy().reshape(cap, REC)
        for i in range(n_local):
            arr[i, 0] = rank; arr[i, 1] = i; arr[i, 2:] = (rank * 31 + i) % 251
        arr[n_local:, :] = 0xEE                                   # garbage beyond the valid records must not travel
        out, counts = G.all_gather_samples(dist, recs, n_local, REC)
        empty, c0 = G.all_gather_samples(dist, recs, 0, REC)      # nothing finished this step
        stats = G.all_reduce_stats(dist, {"simulations": 1000 * (rank + 1), "moves": 7, "games": rank})
        q.put((rank, first, count, counts, out.numpy().copy(), len(empty), c0, stats))
    finally:
        dist.destroy_process_group()


def test_sample_gather_and_sharding_world2():
    world, port = 2, 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    (_, f0, c0, counts0, out0, e0, ec0, st0), (_, f1, c1, counts1, out1, e1, ec1, st1) = res
    assert (f0, c0, f1, c1) == (0, 2049, 2049, 2048)                   # contiguous blocks covering all 4097 slots
    assert counts0 == counts1 == [5, 9] and e0 == e1 == 0 and ec0 == ec1 == [0, 0]
    assert np.array_equal(out0, out1) and out0.shape == (14, REC)
    assert out0[:5, 0].tolist() == [0] * 5 and out0[5:, 0].tolist() == [1] * 9 and out0[:, 1].tolist() == list(range(5)) + list(range(9))
    assert not (out0 == 0xEE).all(1).any()
    assert st0 == st1 == {"games": 1, "moves": 14, "simulations": 3000}
