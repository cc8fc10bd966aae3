"""Multi-GPU sharding layer: one process per GPU, text split with halo overlap.

Replaces the MPI part of the reference's driver (main.c:464-493 MPI_Scatterv of
overlapping chunks, main.c:656 MPI_Reduce of the count):

* shard geometry is the reference's (``acwm_shard_bounds`` == main.c:467-477): rank r
  scans text[r*c, min((r+1)*c + m_max-1, n)), c = ceil(n / world);
* every rank scans its own shard with the same replicated tables -- no data-path
  collective;
* only the per-rank match COUNT crosses the interconnect.  On GPUs the exchange is fused into
  the scan kernel: ``connect_peers`` gives every rank a mailbox in torch symmetric memory
  (peer-mapped over NVLink) and the kernel's publishing thread stores its 8-byte count into
  all mailboxes and sums its own (``acwm_set_peers``) -- no extra launch per scan.  The NCCL
  ``all_reduce(SUM)`` of a uint64 (as int64) remains as the fallback and as the cross-check
  (gloo in the CPU tests);
* positions stay per rank (shard-local offsets + the shard start = global, already
  sorted); ``gather_positions`` brings them to rank 0's host memory on request.
"""
from __future__ import annotations

import numpy as np

from . import shard_bounds


def shard_of(n: int, world: int, rank: int, m_max: int):
    """(start, length, report_from) of this rank's shard."""
    start, length = shard_bounds(n, world, rank, m_max - 1)
    # ranks > 0 own match ends >= start + m_max - 1 (the previous rank's halo covers the rest)
    return start, length, (m_max - 1 if rank > 0 else 0)


def connect_peers(matcher, device=None) -> bool:
    """Collective: set up the in-kernel count exchange for `matcher` over the default process group.
    Returns False (and leaves the matcher on the NCCL path) if symmetric memory is unavailable."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() <= 1:
        return False
    world, rank = dist.get_world_size(), dist.get_rank()
    ok = True
    try:
        import torch.distributed._symmetric_memory as symm_mem
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        box = symm_mem.empty(4 * world, dtype=torch.int64, device=dev)  # uint64[kPeerRing][world]
        box.zero_()
        hdl = symm_mem.rendezvous(box, dist.group.WORLD)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001 -- any failure means "no peer memory here"
        ok, box, hdl, ptrs = False, None, None, None
        if rank == 0:
            print(f"[sharding] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL all_reduce", flush=True)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # also the barrier that orders the zero-fill before any exchange
    if int(flag.item()) != 1:
        matcher.set_peers(0, 0, None)
        return False
    matcher.set_peers(rank, world, ptrs)
    matcher._mailbox = (box, hdl)  # keep the mapping alive
    return True


def allreduce_count(local_count: int, device=None) -> int:
    """Sum of the per-rank counts (NCCL when `device` is a CUDA device, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([local_count], dtype=torch.int64, device=device if device is not None else "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def allreduce_count_tensor(count_i64):
    """In-place all-reduce of a device-resident int64[1] count (stays on the GPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(count_i64, op=dist.ReduceOp.SUM)
    return count_i64


def gather_positions(local_positions: np.ndarray, shard_start: int, dst: int = 0):
    """Global, sorted positions on rank `dst` (None elsewhere)."""
    import torch.distributed as dist
    glob = local_positions.astype(np.uint64) + np.uint64(shard_start)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return glob
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(glob, out, dst=dst)
    if dist.get_rank() != dst:
        return None
    return np.concatenate(out)  # shards are ascending and disjoint in match ends


def search_sharded_host(matcher, text: np.ndarray, m_max: int, want_positions: bool = True):
    """Every rank holds (or can read) the full host text; scans only its shard.
    Returns (global_count, positions on rank 0 / None)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    start, length, report_from = shard_of(text.size, world, rank, m_max)
    dev = torch.device("cuda", torch.cuda.current_device())
    shard = torch.from_numpy(text[start:start + length]).to(dev)
    matcher.upload(pos_capacity=max(1, length) if want_positions else 0)
    matcher.scan_tensor(shard, want_positions=want_positions, report_from=report_from)
    count, pos, _ = matcher.fetch(cap=length if want_positions else 0,
                                  stream=torch.cuda.current_stream().cuda_stream)
    total = allreduce_count(count, dev)
    return total, (gather_positions(pos, start) if want_positions else None)
