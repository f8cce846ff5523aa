"""Small fixed workload for ncu captures: S streams x B blocks of the bench's synthetic input,
`reps` submits from device-resident input (no CPU baseline, no e2e)."""
import sys, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from acarsdec_b200 import api, synth
sys.path.insert(0, str(ROOT))
from bench import make_pool

S = int(sys.argv[1]) if len(sys.argv) > 1 else 592
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
K = 160
fd, _, fc = api.plan(K, synth.DEFAULT_FREQS_MHZ)
pool = make_pool(K, B, 2, fc)
stride = B * 2048 * K
host = np.empty((S, stride), dtype=np.uint8)
for s in range(S):
    host[s] = pool[s % 2]
ctx = api.Context(K, S, 8, B, flags=1)
for s in range(S):
    ctx.set_plan(s, fd)
d = ctx.device_alloc(S * stride)
ctx.copy_to_device(d, host)
for _ in range(reps):
    ctx.submit_device(d, B, stride)
ctx.sync()
print("frames", len(ctx.drain()))
