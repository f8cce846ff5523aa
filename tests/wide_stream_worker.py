"""Wide-stream multi-GPU case through acarsdec_b200.wide.WideStream: ONE IQ
stream, channels split across ranks, one broadcast per submit, frames merged in emission order and checked
against the CPU oracle on rank 0.  Run under torchrun (or alone: world size 1)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import refs  # noqa: E402
from acarsdec_b200 import api, sharding, synth, wide  # noqa: E402


def main():
    dist, rank, world, local = sharding.init_process_group()
    torch.cuda.set_device(local)
    K, nblk = 192, 4
    fm = tuple(130.000 + 0.025 * i for i in range(24))
    fd, _, fc = api.plan(K, fm)
    iq = None
    if rank == 0:
        plan = synth.make_plan(K, fm, fc, seconds=nblk * 1024 / 12500, seed=8, msgs_per_chan_per_sec=3.0, text_len=(5, 25), amp=(8.0, 14.0))
        iq = synth.render_blocks(plan, 0, nblk).reshape(-1)
    ws = wide.WideStream(dist, rank, world, local, K, fd, fc, max_blocks=2)
    half = 2 * 2048 * K
    for i in range(2):                               # two submits: both broadcast buffers, state carried across
        ws.submit(iq[i * half:(i + 1) * half] if rank == 0 else None, 2)
    ws.sync()
    merged = ws.gather()
    ws.close()
    if rank == 0:
        orc = refs.OracleLib()
        o = refs.OracleStream(orc, K, orc.wf(K, fm))
        o.blocks(iq)
        want = [(m.chn, m.len, m.err, bytes(m.txt[:m.len]), bytes(m.crc), int(np.float32(m.lvl).view(np.uint32))) for m in o.msgs()]
        got = [(m[2],) + m[4:] for m in merged]
        print(json.dumps({"world": world, "frames": len(got), "match": got == want}), flush=True)
        assert got == want and len(want) >= 8
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
