"""How far the opt-in fast channelizer's intermediates sit from the reference's (CPU; the device kernel is
bit-identical to the restatement used here, tests/test_gpu_fast.py): per-sample relative envelope difference,
demodulator state differences at every 1024-sample chunk boundary, lvl of every decoded message.
  python tools/fast_tolerance.py > profiles/r2_fast_tolerance.json"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import refs
from acarsdec_b200 import api, synth


def compare(orc, K, nch, secs, ref, fast, total):
    """ref / fast: (nch, nout) envelopes of the literal and the fast arithmetic; total: per-row total in-band signal."""
    nblk = ref.shape[1] // 1024
    rel = np.abs(fast.astype(np.float64) - ref) / np.maximum(ref, 1e-30)
    scaled = np.abs(fast.astype(np.float64) - ref) / total[None, :]
    ddf, dclk, dphi, dlvl, nmsg = [], [], [], [], 0
    for c in range(nch):
        a, b = orc.new_chan(c), orc.new_chan(c)
        sa, sb = refs.Sink(), refs.Sink()
        for k in range(nblk):
            orc.demod(a, ref[c, k * 1024:(k + 1) * 1024], sa)
            orc.demod(b, fast[c, k * 1024:(k + 1) * 1024], sb)
            ddf.append(abs(a.MskDf - b.MskDf)); dclk.append(abs(a.MskClk - b.MskClk)); dphi.append(abs(a.MskPhi - b.MskPhi))
        ma, mb = sa.msgs(), sb.msgs()
        assert [m.as_tuple() for m in ma] == [m.as_tuple() for m in mb], "raw frames differ"
        dlvl += [abs(p.lvl - q.lvl) for p, q in zip(ma, mb)]
        nmsg += len(ma)
    q = lambda v, ps: [float(t) for t in np.quantile(v, ps)]
    return {"K": K, "channels": nch, "seconds": secs, "envelope_samples": int(rel.size), "raw_frames_identical": nmsg,
            "envelope_rel_diff": {"p50": q(rel, [0.5])[0], "p90": q(rel, [0.9])[0], "p99": q(rel, [0.99])[0], "p99_9": q(rel, [0.999])[0],
                                  "max": float(rel.max()), "fraction_within_1e-5": float((rel <= 1e-5).mean()),
                                  "fraction_within_1e-4": float((rel <= 1e-4).mean())},
            "envelope_diff_over_total_inband_signal": {"p99": q(scaled, [0.99])[0], "max": float(scaled.max())},
            "MskDf_abs_diff": {"p50": q(ddf, [0.5])[0], "p99": q(ddf, [0.99])[0], "max": float(max(ddf)), "pll_range": 0.0038},
            "MskClk_abs_diff_max": float(max(dclk)), "MskPhi_abs_diff_max": float(max(dphi)),
            "lvl_dB_abs_diff_max": float(max(dlvl)) if dlvl else None}


def study(orc, K, fm, secs, seed):
    fd, _, fc = api.plan(K, fm)
    nblk = synth.blocks_for_seconds(K, secs)
    plan = synth.make_plan(K, fm, fc, seconds=secs, seed=seed, msgs_per_chan_per_sec=2.0)
    iq = synth.render_blocks(plan, 0, nblk).reshape(-1)
    ref = orc.channelize(iq, K, orc.wf(K, fm))
    kbin, tw = orc.fast_plan(K, fd, fc)
    fast = orc.channelize_dft(iq, K, kbin, tw, True)
    x = iq.reshape(-1, K, 2).astype(np.float64) - 127.5
    total = np.hypot(x[..., 0], x[..., 1]).sum(axis=1) / K / 127.5             # total in-band signal per output row
    return compare(orc, K, len(fm), secs, ref, fast, total)


def _bursts(plan, nch, secs, seed):
    rng = np.random.default_rng(seed)
    for ch in range(nch):
        t = 0.01 + 0.03 * ch
        while True:
            fr = synth.frame_bytes(synth.random_text(rng, int(rng.integers(8, 60))))
            if t + len(fr) * 8 / 2400 + 0.01 > secs:
                break
            plan.bursts.append(synth.Burst(chan=ch, t0=t, frame=fr, amp=float(rng.uniform(10, 25)), phase=float(rng.uniform(0, 6))))
            t += len(fr) * 8 / 2400 + 0.2


def study_cs16(orc, variant, K, fm, secs, seed):
    """soapy.c / sdrplay.c input: literal arithmetic vs the folded form (orc_channelize_dft8_cs16 = k_channelize_dft1<CS16>)"""
    fd, _, fc = orc.plan(K, fm)
    plan = synth.StreamPlan(K=K, freqs_hz=tuple(fd), fc_hz=fc, seed=seed, noise_sigma=1.5)
    _bursts(plan, len(fm), secs, seed)
    n = int(secs * plan.rate) // (1024 * K) * 1024 * K
    iq = synth.render_cs16(plan, 0, n)
    ref = orc.channelize_cs16(variant, iq, K, orc.cs16_osc(variant, K, fd, fc))
    kbin, tw = orc.fast_plan_cs16(variant, K, fd, fc)
    fast = orc.channelize_dft8_cs16(iq, K, kbin, tw)
    x = iq.astype(np.float64).reshape(-1, K, 2)
    total = np.hypot(x[..., 0], x[..., 1]).sum(axis=1) / K * (1.0 / 32768 if variant == 0 else 0.25)
    return compare(orc, K, len(fm), secs, ref, fast, total)


def study_air(orc, rate, fm, secs, seed):
    """air.c input: literal arithmetic vs the real-row DFT form (orc_channelize_rdft = k_channelize_rdft)"""
    fd, fc, K = orc.air_plan(rate, fm)
    plan = synth.StreamPlan(K=K, freqs_hz=tuple(fd), fc_hz=fc, seed=seed, noise_sigma=1.0)
    _bursts(plan, len(fm), secs, seed)
    n = int(secs * rate) // (1024 * K) * 1024 * K
    x = synth.render_real(plan, 0, n)
    ref = orc.channelize_real(x, K, orc.air_wf(rate, fm))
    kbin, tw = orc.fast_plan_air(K, fd, fc)
    fast = orc.channelize_rdft(x, K, kbin, tw)
    total = np.abs(x.astype(np.float64)).reshape(-1, K).sum(axis=1) / K
    return compare(orc, K, len(fm), secs, ref, fast, total)


if __name__ == "__main__":
    refs.ensure_built()
    orc = refs.OracleLib()
    out = {"what": "fast channelizer (folded form) vs the reference's envelope, CPU restatements (device kernel bit-identical to the fast one)",
           "north_star_clause": "demod float intermediates within 1e-5 rel: NOT met sample by sample by the fast form (met by the default exact form, bit for bit)",
           "config2": study(orc, 160, synth.DEFAULT_FREQS_MHZ, 4.0, 91),
           "K192_24ch_raster": study(orc, 192, tuple(130.000 + 0.025 * i for i in range(24)), 2.0, 92),
           "cs16_soapy_K160": study_cs16(orc, 0, 160, (131.525, 131.725, 131.825, 131.450, 131.550), 2.0, 93),
           "cs16_sdrplay_K192": study_cs16(orc, 1, 192, (131.525, 131.725, 131.825, 131.450, 131.550), 2.0, 94),
           "air_2p5MSps_K200": study_air(orc, 2500000, synth.DEFAULT_FREQS_MHZ, 2.0, 95)}
    print(json.dumps(out, indent=1))
