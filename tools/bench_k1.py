"""Standalone timing of the u8-IQ channelizer (no overlap with the demod: sync per step), exact or fast
form, S streams x B blocks, K=160, 8 channels.  Prints one JSON line.
  python tools/bench_k1.py [exact|fast] [S] [B]"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from acarsdec_b200 import api, synth
from bench import make_pool

mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 592
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
K = 160
fd, _, fc = api.plan(K, synth.DEFAULT_FREQS_MHZ)
pool = make_pool(K, B, 2, fc)
stride = B * 2048 * K
host = np.empty((S, stride), dtype=np.uint8)
for s in range(S):
    host[s] = pool[s % 2]
ctx = api.Context(K, S, 8, B, flags=1 | (8 if mode == "fast" else 0))
for s in range(S):
    ctx.set_plan(s, fd)
d = ctx.device_alloc(S * stride)
ctx.copy_to_device(d, host)
for _ in range(2):
    ctx.submit_device(d, B, stride); ctx.sync()
ctx.drain_records(); ctx.stats(reset=True)
steps = 5
for _ in range(steps):
    ctx.submit_device(d, B, stride); ctx.sync()
st = ctx.stats()
k1 = st.chan_ms / st.chan_launches
alg = S * B * (2048 * K + 1024 * 8 * 4)
peak = 6571.6
try:
    peak = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"channelizer": mode, "fast_rows_per_lane": __import__("os").environ.get("ACB_FAST_ROWS", "1 (default)"), "fast_warps": __import__("os").environ.get("ACB_FAST_WARPS", "2 (default)"), "fast_launches": int(st.fast_chan_launches), "streams": S, "blocks": B, "K": K,
                  "k_channelize_ms": k1, "k_demod_ms": st.demod_ms / st.demod_launches, "algorithmic_bytes": alg,
                  "achieved_GBs": alg / k1 / 1e6, "peak_GBs": peak, "frac": alg / k1 / 1e6 / peak,
                  "frames_per_step": len(ctx.drain_records()) / steps}))
