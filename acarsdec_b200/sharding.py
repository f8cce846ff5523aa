"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous only.

The path shards without any data-path collective: channels share nothing but the
read-only matched filter, streams share nothing at all.

  * many streams (BASELINE config 4): stream index ranges per rank, zero communication;
  * one wide stream (configs 3, 5): channel index ranges per rank; every rank needs the same raw
    u8 IQ block, which is the single broadcast the north star allows (NCCL over NVLink on GPUs,
    gloo in the CPU tests) — each rank then channelizes/demodulates only its own channels.

Decoded messages come back per rank and are merged in the reference's emission order
(block-major, then stream, then channel, then time; rtl.c:357-360).
"""
from __future__ import annotations

import os
from typing import Iterable, Sequence


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def split_range(n: int, world: int, rank: int) -> range:
    """Contiguous, balanced (sizes differ by at most 1), order-preserving split of range(n)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def stream_range(total_streams: int, world: int, rank: int) -> range:
    return split_range(total_streams, world, rank)


def channel_range(nch: int, world: int, rank: int) -> range:
    return split_range(nch, world, rank)


def init_process_group(backend: str | None = None):
    """torch.distributed rendezvous from the torchrun environment (MASTER_ADDR must be 127.0.0.1
    style resolvable).  Returns (dist module or None, rank, world, local_rank)."""
    rank, world, local = env_rank_world()
    if world == 1:
        return None, rank, world, local
    import torch
    import torch.distributed as dist
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        torch.cuda.set_device(local)
        kw["device_id"] = torch.device("cuda", local)
    dist.init_process_group(backend, **kw)
    return dist, rank, world, local


def broadcast_iq(dist, tensor, src: int = 0):
    """The one collective of the wide-stream case: the ingest rank's raw uint8 IQ block to all."""
    if dist is not None:
        dist.broadcast(tensor, src=src)
    return tensor


def reduce_scalar(dist, x: float, op: str, device=None) -> float:
    """max / sum of a python float over ranks (timing and sample counts)."""
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM}[op])
    return float(t.item())


def whole_job_throughput(dist, samples_local: float, seconds_local: float, device=None) -> float:
    """Aggregate samples/s: all ranks' samples over the slowest rank's time."""
    total = reduce_scalar(dist, samples_local, "sum", device)
    slowest = reduce_scalar(dist, seconds_local, "max", device)
    return total / slowest


def merge_messages(per_rank: Sequence[Iterable], key=lambda m: (m.block, m.stream, m.chn, m.pos)):
    """Per-rank message lists (each already in emission order, with GLOBAL stream/channel
    indices) -> one list in the reference's emission order."""
    out = [m for lst in per_rank for m in lst]
    out.sort(key=key)
    return out
