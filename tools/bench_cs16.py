"""Roofline measurement of the CS16 (soapy.c / sdrplay.c) channelizer kernel: S streams of int16 I,Q
at 2 MS/s (K=160), C channels, B blocks per submit; kernel duration from the library's CUDA events
(sync per step, no overlap).  Prints one JSON line."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from acarsdec_b200 import api, synth

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0          # 0 soapy.c, 1 sdrplay.c
S = int(sys.argv[2]) if len(sys.argv) > 2 else 296
C = int(sys.argv[3]) if len(sys.argv) > 3 else 8
fast = len(sys.argv) > 4 and sys.argv[4] == "fast"              # ACB_FLAG_FAST_CHANNELIZER: the folded DFT form
K, B = 160, 16
fm = synth.DEFAULT_FREQS_MHZ[:C]
fd, _, fc = api.plan(K, fm)
plan = synth.StreamPlan(K=K, freqs_hz=tuple(fd), fc_hz=fc, seed=5, noise_sigma=1.0)
rng = np.random.default_rng(5)
for ch in range(C):
    plan.bursts.append(synth.Burst(chan=ch, t0=0.02 + 0.1 * ch, frame=synth.frame_bytes(synth.random_text(rng, 30)), amp=15.0))
n = B * 1024 * K
base = synth.render_cs16(plan, 0, n)
x = np.ascontiguousarray(np.broadcast_to(base, (S, n, 2)))
ctx = api.Context(K, S, C, B, flags=4 | (8 if fast else 0))
for s in range(S):
    ctx.set_plan_cs16(s, fd, variant)
for _ in range(2):
    ctx.submit_cs16(x); ctx.sync()
ctx.drain_records(); ctx.stats(reset=True)
steps = 5
for _ in range(steps):
    ctx.submit_cs16(x); ctx.sync()
st = ctx.stats()
k1 = st.chan_ms / st.chan_launches
alg = S * B * 1024 * (K * 4 + C * 4)
peak = 6571.6
try:
    peak = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"front_end": ["soapy.c", "sdrplay.c"][variant] + " CS16", "channelizer": "fast" if fast else "exact",
                  "fast_launches": int(st.fast_chan_launches), "K": K, "streams": S, "channels": C, "blocks": B,
                  "k_channelize_cs16_ms": k1, "k_demod_ms": st.demod_ms / st.demod_launches,
                  "Msamples_per_s_kernel": S * n / k1 / 1e3, "algorithmic_bytes": alg,
                  "achieved_GBs": alg / k1 / 1e6, "peak_GBs": peak, "frac": alg / k1 / 1e6 / peak,
                  "fp32_ceiling_frac": (S * n * C / (k1 * 1e-3)) / (148 * 128 * 1.965e9 / 8.0),
                  "frames_per_step": len(ctx.drain_records()) / steps}))
