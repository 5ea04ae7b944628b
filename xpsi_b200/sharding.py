"""Sharding of a parameter-vector batch over ranks (one process per GPU).

The likelihood path partitions by independent parameter vectors -- exactly the reference's MPI pattern
(xpsi/Sample.py:287-336: scatter theta, gather lnL) -- so there is no data-path collective: each rank
evaluates a contiguous block and the only communication is one all_gather of the per-rank lnL/status blocks
(NCCL over NVLink on GPUs, gloo in the CPU tests).  torch.distributed is plumbing here, nothing else.
"""
import numpy as np


def shard_bounds(n_total, rank, world):
    """Contiguous block [lo, hi) of ``n_total`` items owned by ``rank``; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def gather_blocks(local_lnL, local_status, n_total, device=None):
    """all_gather ragged per-rank blocks into full ``lnL[n_total]``, ``status[n_total]`` on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(local_lnL, dtype=np.float64), np.asarray(local_status, dtype=np.int32)
    world, rank = dist.get_world_size(), dist.get_rank()
    width = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world))
    # one fp64 payload per rank: [lnL | status] padded to the widest block
    buf = torch.zeros(2 * width, dtype=torch.float64, device=device)
    n_loc = len(local_lnL)
    buf[:n_loc] = torch.as_tensor(np.asarray(local_lnL, dtype=np.float64), device=device)
    buf[width:width + n_loc] = torch.as_tensor(np.asarray(local_status, dtype=np.float64), device=device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    lnL = np.empty(n_total, dtype=np.float64)
    status = np.empty(n_total, dtype=np.int32)
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        blk = out[r].cpu().numpy()
        lnL[lo:hi] = blk[:hi - lo]
        status[lo:hi] = blk[width:width + hi - lo].astype(np.int32)
    return lnL, status
