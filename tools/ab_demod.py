"""Sweep of k_demod's lanes-per-channel / kernel variants (ACB_DEMOD_LANES) against the stream count:
for every (S, lanes) the demod kernel alone (sync per step, no overlap with the channelizer), the fast
channelizer alone, and the pipelined step.  One JSON line per combination.
  python tools/ab_demod.py [S,S,...] [lanes,lanes,...] [exact|fast]"""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from acarsdec_b200 import api, synth
from bench import make_pool

Ss = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [592, 1184, 2368, 4736]
lanes_list = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [-4, -8, 8, 4, 2, 1, 24, 20, 18, 17]
mode = sys.argv[3] if len(sys.argv) > 3 else "fast"
K = 160
fd, _, fc = api.plan(K, synth.DEFAULT_FREQS_MHZ)
for S in Ss:
    B = 16 if S <= 2368 else 8
    pool = make_pool(K, B, 2, fc)
    stride = B * 2048 * K
    d = None
    for lanes in lanes_list:
        os.environ["ACB_DEMOD_LANES"] = str(lanes)
        ctx = api.Context(K, S, 8, B, flags=1 | (8 if mode == "fast" else 0))
        for s in range(S):
            ctx.set_plan(s, fd)
        if d is None:
            d = ctx.device_alloc(S * stride)
            for s in range(S):
                ctx.copy_to_device(d + s * stride, pool[s % 2])
        for _ in range(2):
            ctx.submit_device(d, B, stride); ctx.sync()
        nfr = len(ctx.drain_records()); ctx.stats(reset=True)
        for _ in range(3):
            ctx.submit_device(d, B, stride); ctx.sync()      # sync each step: no K1/K2 overlap
        st = ctx.stats(reset=True)
        ctx.drain_records()
        ctx.mark(0)
        nst = 5
        for _ in range(nst):
            ctx.submit_device(d, B, stride); ctx.drain_records()
        ctx.mark(1); ctx.sync()
        ms = ctx.elapsed_ms() / nst
        st2 = ctx.stats()
        print(json.dumps({"S": S, "B": B, "lanes": lanes, "channelizer": mode, "k2_alone_ms": round(st.demod_ms / st.demod_launches, 4),
                          "k1_alone_ms": round(st.chan_ms / st.chan_launches, 4), "step_ms": round(ms, 4),
                          "Gsamples_s": round(S * B * 1024 * K / ms / 1e6, 1),
                          "k1_in_step_ms": round(st2.chan_ms / st2.chan_launches, 4), "k2_in_step_ms": round(st2.demod_ms / st2.demod_launches, 4),
                          "frames_warm": nfr}), flush=True)
        if lanes == lanes_list[-1]:
            ctx.device_free(d)
        ctx.close()
