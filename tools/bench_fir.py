"""BASELINE configs[4] capture: wideband 20 MS/s u8 IQ (K=1600), 256 channels on a 25 kHz raster,
FIR-tap sweep 65..513 (and config 3: K=192, 64 ch, 165 taps).  Kernel time from the library's CUDA
events; algorithmic bytes = 2*N_out*min(T,K) input + 4*C*N_out envelope."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from acarsdec_b200 import api
sys.path.insert(0, str(ROOT / "tests"))
from test_gpu_fir import fir_tables

peak = 6571.6
try:
    peak = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
rows = []
for (K, C, S, B, sweep) in ((1600, 256, 4, 8, (65, 129, 257, 513)), (192, 64, 64, 16, (165,))):
    rate = K * 12500
    iq = np.random.default_rng(1).integers(0, 256, size=(S, B * 2048 * K), dtype=np.uint8)
    for T in sweep:
        offs = [(-0.45 + 0.9 * i / (C - 1)) * rate / 2 for i in range(C)]
        wf = fir_tables(K, T, offs, rate)
        ctx = api.Context(K, S, C, B, taps=T)
        for s in range(S):
            ctx.set_wf(s, wf)
        for _ in range(2):
            ctx.submit_host(iq, B); ctx.sync()
        ctx.drain_records(); ctx.stats(reset=True)
        for _ in range(3):
            ctx.submit_host(iq, B); ctx.sync()
        st = ctx.stats()
        k1 = st.chan_ms / st.chan_launches
        nout = S * B * 1024
        alg = nout * (2 * min(T, K) + 4 * C)
        rows.append({"K": K, "channels": C, "taps": T, "streams": S, "blocks": B, "k_channelize_ms": k1,
                     "k_demod_ms": st.demod_ms / st.demod_launches, "input_Msamples_per_s": nout * K / k1 / 1e3,
                     "algorithmic_bytes": alg, "achieved_GBs": alg / k1 / 1e6, "frac_of_hbm_peak": alg / k1 / 1e6 / peak,
                     "cmac_per_clk_per_sm": nout * C * ((T + 7) // 8 * 8) / (k1 * 1e-3) / 148 / 1.965e9})
        ctx.close()
print(json.dumps({"peak_GBs": peak, "rows": rows}))
