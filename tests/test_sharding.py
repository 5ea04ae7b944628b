"""N>1 host logic on CPU: world_size-2 gloo processes shard a batch and gather lnL/status."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance():
    from xpsi_b200.sharding import shard_bounds
    for n in (0, 1, 7, 100, 100003):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from xpsi_b200.sharding import gather_blocks, shard_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_total, rank, world)
    theta = np.arange(n_total, dtype=np.float64)
    local = -0.5 * theta[lo:hi] ** 2               # stand-in for the rank's lnL block
    status = (np.arange(lo, hi) % 5 == 0).astype(np.int32)
    lnL, st = gather_blocks(local, status, n_total)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, lnL, st))


def test_gloo_world_size_2_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    n_total, world = 13, 2                          # ragged: 7 + 6
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = -0.5 * np.arange(n_total, dtype=np.float64) ** 2
    for rank, lnL, st in res:
        assert np.array_equal(lnL, expect)
        assert np.array_equal(st, (np.arange(n_total) % 5 == 0).astype(np.int32))
