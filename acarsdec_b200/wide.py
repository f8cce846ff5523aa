"""One wide IQ stream served by several GPUs, one process per GPU (BASELINE configs 3 and 5).

Every rank owns a contiguous range of the stream's channels.  The ingest rank (0) copies each raw u8 IQ block
host -> device once; ONE broadcast (NCCL over NVLink/NVSwitch; gloo in CPU tests of the plumbing) hands it to
the other ranks; each rank channelizes and demodulates only its own channels straight from the broadcast
buffer (acb_submit_device, ordered behind the broadcast by a CUDA event — no host sync).  No collective inside
the channelizer or the demodulator.  Decoded messages go back to rank 0 and are merged into the reference's
emission order (block, then channel, then time; rtl.c:357-360).

The reference has no counterpart: one thread serves every channel (rtl.c:344-360).
"""
from __future__ import annotations

import numpy as np

from . import api, sharding


class WideStream:
    def __init__(self, dist, rank: int, world: int, local: int, K: int, freqs_hz, fc_hz: int, max_blocks: int,
                 taps: int = 0, wf_all: np.ndarray | None = None, flags: int = 0):
        import torch
        self.torch, self.dist, self.rank, self.world, self.local = torch, dist, rank, world, local
        self.K, self.max_blocks = K, max_blocks
        self.nch_total = len(freqs_hz)
        self.mine = sharding.channel_range(self.nch_total, world, rank)
        self.blk_bytes = 2048 * K
        self.dev = torch.device("cuda", local)
        torch.cuda.set_device(local)
        # two broadcast buffers: submit i+1 is filled while submit i is still being channelized
        self.bufs = [torch.empty(max_blocks * self.blk_bytes, dtype=torch.uint8, device=self.dev) for _ in range(2)]
        self.ev = [torch.cuda.Event() for _ in range(2)]
        self.t0 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        self.t1 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        self.nsub = 0
        self.inflight = 0
        self.bcast_ms = 0.0
        self.ctx = None
        if len(self.mine):
            self.ctx = api.Context(K, 1, len(self.mine), max_blocks, device=local, flags=flags | 1, taps=taps)
            if wf_all is not None:
                self.ctx.set_wf(0, wf_all[self.mine.start:self.mine.stop])
            else:
                self.ctx.set_plan_at(0, list(freqs_hz)[self.mine.start:self.mine.stop], fc_hz)

    def submit(self, iq_host, nblk: int) -> None:
        """iq_host: rank 0's block(s) as a uint8 torch tensor (pinned) or numpy array; ignored elsewhere."""
        torch = self.torch
        b = self.nsub & 1
        if self.inflight == 2:                       # buffer b still feeds the submit two back
            if self.ctx is not None:
                self.ctx.collect()
            self.inflight -= 1
        n = nblk * self.blk_bytes
        buf = self.bufs[b][:n]
        self.t0[b].record()
        if self.rank == 0:
            src = iq_host if isinstance(iq_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(iq_host).reshape(-1))
            buf.copy_(src[:n], non_blocking=True)
        if self.dist is not None:
            self.dist.broadcast(buf, src=0)          # the single collective of the data path
        self.t1[b].record()
        self.ev[b].record()
        if self.ctx is not None:
            self.ctx.wait_event(self.ev[b].cuda_event)
            self.ctx.submit_device(buf.data_ptr(), nblk, n)
        self.nsub += 1
        self.inflight += 1

    def sync(self) -> None:
        if self.ctx is not None:
            self.ctx.sync()
        self.torch.cuda.synchronize()
        self.inflight = 0

    def ingest_ms(self) -> float:
        """H2D (rank 0) + broadcast device time of the last two submits' average (CUDA events, torch stream)."""
        ms = [self.t0[i].elapsed_time(self.t1[i]) for i in range(min(2, self.nsub))]
        return sum(ms) / max(1, len(ms))

    def drain_local(self):
        """This rank's messages as tuples with the GLOBAL channel index:
        (block, stream, chn, pos, len, err, txt, crc, lvl_bits)."""
        out = []
        if self.ctx is not None:
            for m in self.ctx.drain():
                out.append((int(m.block), 0, m.chn + self.mine.start, int(m.pos), m.len, m.err, bytes(m.txt[:m.len]),
                            bytes(m.crc), int(np.float32(m.lvl).view(np.uint32))))
        return out

    def gather(self):
        """All ranks' messages merged in emission order on every rank (object all-gather: a few KB)."""
        mine = self.drain_local()
        parts = [mine]
        if self.dist is not None:
            parts = [None] * self.world
            self.dist.all_gather_object(parts, mine)
        return sharding.merge_messages(parts, key=lambda m: m[:4])

    def close(self) -> None:
        if self.ctx is not None:
            self.ctx.close()
            self.ctx = None
