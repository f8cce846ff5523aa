"""Roofline measurement of the real-input (air.c) channelizer kernel: S streams of float32 real
samples at rate = K*12500, C channels, B blocks per submit; kernel duration from the library's CUDA
events (no overlap: sync per step).  Prints one JSON line."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, os
from acarsdec_b200 import api, synth
if os.environ.get('ACB_LIB'): api.LIB_PATH = Path(os.environ['ACB_LIB'])
import refs

rate = int(sys.argv[1]) if len(sys.argv) > 1 else 2500000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 296
C = int(sys.argv[3]) if len(sys.argv) > 3 else 8
fast = len(sys.argv) > 4 and sys.argv[4] == "fast"              # ACB_FLAG_FAST_CHANNELIZER: k_channelize_rdft
B = 16
fm = synth.DEFAULT_FREQS_MHZ[:C]
fd, fc, K = api.air_plan(rate, fm)
plan = synth.StreamPlan(K=K, freqs_hz=tuple(fd), fc_hz=fc, seed=3, noise_sigma=1.0)
rng = np.random.default_rng(3)
for ch in range(C):
    plan.bursts.append(synth.Burst(chan=ch, t0=0.02 + 0.1 * ch, frame=synth.frame_bytes(synth.random_text(rng, 30)), amp=15.0))
n = B * 1024 * K
base = synth.render_real(plan, 0, n)
x = np.ascontiguousarray(np.broadcast_to(base, (S, n)))
ctx = api.Context(K, S, C, B, flags=2 | (8 if fast else 0))
for s in range(S):
    ctx.set_plan_air(s, fd)
for _ in range(2):
    ctx.submit_real(x); ctx.sync()
ctx.drain_records(); ctx.stats(reset=True)
steps = 5
for _ in range(steps):
    ctx.submit_real(x); ctx.sync()
st = ctx.stats()
k1 = st.chan_ms / st.chan_launches
alg = S * B * 1024 * (K * 4 + C * 4)
peak = 6571.6
try:
    peak = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
# oracle check of stream 0's frames
orc = refs.OracleLib()
print(json.dumps({"front_end": "air.c real float32", "channelizer": "fast" if fast else "exact",
                  "fast_launches": int(st.fast_chan_launches), "rate": rate, "K": K, "streams": S, "channels": C, "blocks": B,
                  "k_channelize_real_ms": k1, "k_demod_ms": st.demod_ms / st.demod_launches,
                  "Msamples_per_s_kernel": S * n / k1 / 1e3, "algorithmic_bytes": alg,
                  "achieved_GBs": alg / k1 / 1e6, "peak_GBs": peak, "frac": alg / k1 / 1e6 / peak,
                  "fp32_ceiling_frac": (S * n * C / (k1 * 1e-3)) / (148 * 128 * 1.965e9 / 4.0),
                  "frames_per_step": len(ctx.drain_records()) / steps}))
