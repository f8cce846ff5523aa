"""N>1 host logic on CPU: world_size-2 gloo (127.0.0.1 rendezvous), no GPU."""
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path
from types import SimpleNamespace

import pytest

from acarsdec_b200 import sharding

ROOT = Path(__file__).resolve().parent.parent


def test_split_range_properties():
    for n in (0, 1, 7, 8, 64, 128, 1024, 1031):
        for world in (1, 2, 3, 4, 8):
            parts = [sharding.split_range(n, world, r) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        sharding.split_range(8, 2, 2)
    assert list(sharding.stream_range(128, 8, 3)) == list(range(48, 64))        # BASELINE config 4: 16 streams per GPU
    assert list(sharding.channel_range(64, 8, 7)) == list(range(56, 64))        # config 3: 8 channels per GPU


def test_merge_messages_emission_order():
    mk = lambda b, s, c, p: SimpleNamespace(block=b, stream=s, chn=c, pos=p)
    r0 = [mk(0, 0, 1, 900), mk(1, 0, 0, 1100), mk(1, 1, 3, 1500)]
    r1 = [mk(0, 2, 0, 100), mk(1, 2, 2, 1030), mk(2, 3, 0, 2050)]
    got = sharding.merge_messages([r0, r1])
    assert [(m.block, m.stream, m.chn) for m in got] == [(0, 0, 1), (0, 2, 0), (1, 0, 0), (1, 1, 3), (1, 2, 2), (2, 3, 0)]


WORKER = textwrap.dedent('''
    import sys, json
    sys.path.insert(0, %r)
    import numpy as np, torch
    from acarsdec_b200 import sharding
    dist, rank, world, local = sharding.init_process_group("gloo")
    assert world == 2 and dist is not None
    # wide-stream case: rank 0 ingests the IQ block, one broadcast, each rank takes its channels
    blk = torch.zeros(4096, dtype=torch.uint8)
    if rank == 0:
        blk = torch.from_numpy(np.random.default_rng(3).integers(0, 256, 4096, dtype=np.uint8))
    sharding.broadcast_iq(dist, blk, src=0)
    chans = list(sharding.channel_range(11, world, rank))
    streams = list(sharding.stream_range(5, world, rank))
    # whole-job throughput: sum of samples over the slowest rank's time
    v = sharding.whole_job_throughput(dist, 1000.0 * (rank + 1), 2.0 + rank)
    mx = sharding.reduce_scalar(dist, float(rank), "max")
    print(json.dumps({"rank": rank, "sum": int(blk.sum()), "chans": chans, "streams": streams, "v": v, "mx": mx}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
''')


def test_world_size_2_gloo(tmp_path):
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    script = tmp_path / "w.py"
    script.write_text(WORKER % str(ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=180)
        assert p.returncode == 0, e[-2000:]
        outs.append(__import__("json").loads(o.strip().split("\\n")[-1]))
    outs.sort(key=lambda d: d["rank"])
    assert outs[0]["sum"] == outs[1]["sum"] > 0                        # the broadcast arrived
    assert outs[0]["chans"] + outs[1]["chans"] == list(range(11))
    assert outs[0]["streams"] + outs[1]["streams"] == list(range(5))
    assert outs[0]["v"] == outs[1]["v"] == pytest.approx(3000.0 / 3.0)
    assert outs[0]["mx"] == outs[1]["mx"] == 1.0
