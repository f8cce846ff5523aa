import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, time
from acarsdec_b200 import api, synth
api.LIB_PATH = Path(sys.argv[1])
from bench import make_pool
K, B, S = 160, 16, int(sys.argv[2]) if len(sys.argv) > 2 else 592
fd, _, fc = api.plan(K, synth.DEFAULT_FREQS_MHZ)
pool = make_pool(K, B, 2, fc)
stride = B * 2048 * K
host = np.empty((S, stride), dtype=np.uint8)
for s in range(S): host[s] = pool[s % 2]
ctx = api.Context(K, S, 8, B, flags=1)
for s in range(S): ctx.set_plan(s, fd)
d = ctx.device_alloc(S * stride); ctx.copy_to_device(d, host)
for _ in range(3): ctx.submit_device(d, B, stride)
ctx.sync(); n0 = len(ctx.drain_records()); ctx.stats(reset=True)
ctx.mark(0)
for _ in range(8):
    ctx.submit_device(d, B, stride); ctx.drain_records()
ctx.mark(1); ctx.sync()
ms = ctx.elapsed_ms() / 8
st = ctx.stats()
print(Path(sys.argv[1]).name, "S", S, "overlapped ms/step", round(ms, 3), "k1", round(st.chan_ms / st.chan_launches, 3), "k2", round(st.demod_ms / st.demod_launches, 3), "Gsamples/s", round(S * B * 1024 * K / ms / 1e6, 1), "frames", n0)
