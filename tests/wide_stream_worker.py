"""Wide-stream multi-GPU case: ONE IQ stream, channels split across
ranks by index, the ingest rank broadcasts the raw u8 block once (NCCL, device to device), every
rank channelizes + demodulates only its own channels from the broadcast buffer (acb_submit_device
on the tensor's storage), frames are gathered and merged in emission order, rank 0 checks them
against the CPU oracle.  Run under torchrun (or alone: world size 1)."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import refs  # noqa: E402
from acarsdec_b200 import api, sharding, synth  # noqa: E402


def main():
    dist, rank, world, local = sharding.init_process_group()
    torch.cuda.set_device(local)
    K, nblk = 192, 4
    fm = tuple(130.000 + 0.025 * i for i in range(24))
    fd, _, fc = api.plan(K, fm)
    wf_all = api.build_wf(K, fm)
    nbytes = nblk * 2048 * K
    buf = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{local}")
    iq = None
    if rank == 0:
        plan = synth.make_plan(K, fm, fc, seconds=nblk * 1024 / 12500, seed=8, msgs_per_chan_per_sec=3.0, text_len=(5, 25), amp=(8.0, 14.0))
        iq = synth.render_blocks(plan, 0, nblk).reshape(-1)
        buf.copy_(torch.from_numpy(iq))
    sharding.broadcast_iq(dist, buf, src=0)                 # the single collective
    torch.cuda.synchronize()
    mine = sharding.channel_range(len(fm), world, rank)
    msgs = []
    if len(mine):
        with api.Context(K, 1, len(mine), nblk, device=local, flags=1) as ctx:
            ctx.set_wf(0, wf_all[mine.start:mine.stop])
            ctx.submit_device(buf.data_ptr(), nblk, nbytes)
            ctx.sync()
            for m in ctx.drain():
                msgs.append((int(m.block), 0, m.chn + mine.start, int(m.pos), m.len, m.err, bytes(m.txt[:m.len]).hex(), bytes(m.crc).hex()))
    gathered = [msgs]
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, msgs)
    if rank == 0:
        merged = sorted((m for lst in gathered for m in lst), key=lambda m: m[:4])
        orc = refs.OracleLib()
        o = refs.OracleStream(orc, K, orc.wf(K, fm))
        o.blocks(iq)
        want = [(m.chn, m.len, m.err, bytes(m.txt[:m.len]).hex(), bytes(m.crc).hex()) for m in o.msgs()]
        got = [(m[2],) + m[4:] for m in merged]
        print(json.dumps({"world": world, "frames": len(got), "match": got == want}), flush=True)
        assert got == want and len(want) >= 8
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
