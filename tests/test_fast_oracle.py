"""The fast channelizer (ACB_FLAG_FAST_CHANNELIZER) on the CPU: oracle/acars_oracle.c restates the kernel
operation for operation (orc_channelize_dft).  Here that restatement is held to the reference: envelope
within the reference's own table rounding, closer to the exact DFT than the reference is, and — over many
seeded captures, corrupted frames included — the same decoded messages.  tests/test_gpu_fast.py then holds
the GPU kernel to the restatement bit for bit."""
import numpy as np
import pytest

import refs
from acarsdec_b200 import api, synth
from common import msg_tuple

TABLE_EPS = 1e-5   # |dm_fast - dm_ref| <= TABLE_EPS * sum_ind |x[ind] - 127.5| * |w| per output row
IDEAL_EPS = 1e-6   # |dm_fast - exact DFT| <= IDEAL_EPS * the same sum (measured 3e-7; the reference: 1.8e-6)


def _decode(orc, dm):
    """demod + frame sync + FEC of every channel of dm (nch, nout) through the pinned restatement."""
    out = []
    for c in range(dm.shape[0]):
        ch, sink = orc.new_chan(c), refs.Sink()
        orc.demod(ch, dm[c], sink)
        for m in sink.msgs():
            f = orc.fec(m)
            if f is not None:
                out.append(msg_tuple(f)[:-1] + (float(m.lvl),))
    return out


def test_plan_restatement_matches_library(native, oracle):
    for K in (160, 192):
        fd, _, fc = api.plan(K, synth.DEFAULT_FREQS_MHZ)
        k1, tw1 = oracle.fast_plan(K, fd, fc)
        k2, tw2 = api.fast_plan(K, fd, fc)
        assert (k1 == k2).all()
        # value equality (the complex view api.fast_plan returns does not keep the sign of a zero)
        assert np.array_equal(tw1[..., 0], tw2.real) and np.array_equal(tw1[..., 1], tw2.imag)
    fd, _, fc = api.plan(160, (131.4875, 131.725))
    assert oracle.fast_plan(160, fd, fc) is None


@pytest.mark.parametrize("fold8", [False, True])
@pytest.mark.parametrize("K", [160, 192])
def test_envelope_vs_reference_and_exact_dft(oracle, K, fold8):
    fm = synth.DEFAULT_FREQS_MHZ
    fd, fr, fc = oracle.plan(K, fm)
    k, tw = oracle.fast_plan(K, fd, fc)
    wf = oracle.wf(K, fm)
    worst_ref = worst_ideal = worst_ref_ideal = 0.0
    for seed in range(6):
        plan = synth.make_plan(K, fm, fc, seconds=0.3, seed=900 + seed, text_len=(5, 60), msgs_per_chan_per_sec=4.0)
        iq = synth.render_blocks(plan, 0, 3).reshape(-1)
        ref = oracle.channelize(iq, K, wf).astype(np.float64)
        fast = oracle.channelize_dft(iq, K, k, tw, fold8).astype(np.float64)
        x = iq.reshape(-1, K, 2).astype(np.float64)
        bound = TABLE_EPS * np.hypot(x[..., 0] - 127.5, x[..., 1] - 127.5).sum(axis=1) / K / 127.5
        worst_ref = max(worst_ref, (np.abs(fast - ref) / bound[None, :]).max())
        xc = x[..., 0] + 1j * x[..., 1]
        ideal = np.stack([np.abs(xc @ (np.exp(-2j * np.pi * int(kk) * np.arange(K) / K) / K / 127.5)) for kk in k])
        total = bound[None, :] / TABLE_EPS
        worst_ideal = max(worst_ideal, (np.abs(fast - ideal) / total).max())
        worst_ref_ideal = max(worst_ref_ideal, (np.abs(ref - ideal) / total).max())
    assert worst_ref <= 1.0, worst_ref
    assert worst_ideal <= IDEAL_EPS and worst_ideal < worst_ref_ideal, (worst_ideal, worst_ref_ideal)


@pytest.mark.parametrize("K,nseeds,fold8", [(160, 24, False), (192, 8, False), (160, 12, True), (192, 6, True)])
def test_messages_identical_over_many_captures(oracle, K, nseeds, fold8):
    """Same frames — channel, length, error count, text, BCS, in the same order — from the reference's
    envelope and from the fast form's, over seeded captures with clean and corrupted frames, weak and
    strong bursts; lvl (dB) within 0.001."""
    fm = synth.DEFAULT_FREQS_MHZ
    fd, _, fc = oracle.plan(K, fm)
    k, tw = oracle.fast_plan(K, fd, fc)
    wf = oracle.wf(K, fm)
    flips = [[], [(3, 0x04)], [(5, 0x01), (9, 0x80)], [(7, 0x21)], [(-2, 0x10)], [(2, 1), (4, 2), (6, 4), (8, 8)],
             [(1, 0x40), (-3, 0x02)], [(20, 0xFF)], [(0, 0x08), (10, 0x08), (11, 0x08)]]
    total = repaired = 0
    for seed in range(nseeds):
        secs = 0.9
        plan = synth.make_plan(K, fm, fc, seconds=secs, seed=5000 + seed, text_len=(5, 90), msgs_per_chan_per_sec=5.0)
        rng = np.random.default_rng(seed)
        for i, b in enumerate(plan.bursts):
            b.amp = float(rng.uniform(3.0, 30.0))                    # down to marginal SNR
            if seed % 2:
                b.frame = synth.corrupt_frame(b.frame, flips[(i + seed) % len(flips)])
        iq = synth.render_blocks(plan, 0, synth.blocks_for_seconds(K, secs)).reshape(-1)
        a = _decode(oracle, oracle.channelize(iq, K, wf))
        b = _decode(oracle, oracle.channelize_dft(iq, K, k, tw, fold8))
        assert [t[:-1] for t in a] == [t[:-1] for t in b], seed
        for x, y in zip(a, b):
            assert abs(x[-1] - y[-1]) <= 1e-3                          # lvl, in dB
        total += len(a)
        repaired += sum(t[2] > 0 for t in a)
    assert total >= 8 * nseeds and repaired > 0


def test_fast_intermediates_stated_bounds(oracle):
    """What the fast form does NOT meet, with numbers: the north star's "demod float intermediates within 1e-5 rel" holds
    for most envelope samples, not all (the difference is bounded by the total in-band signal — the reference's own table
    rounding — not by the sample), and the demodulator state follows.  The distributions are asserted so that a change of
    the kernel's arithmetic shows up here; the same figures for longer captures are in profiles/r2_fast_tolerance.json
    (tools/fast_tolerance.py).  Messages stay identical (asserted inside study())."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import fast_tolerance
    r = fast_tolerance.study(oracle, 160, synth.DEFAULT_FREQS_MHZ, 1.5, 91)
    e = r["envelope_rel_diff"]
    assert e["fraction_within_1e-5"] >= 0.75 and e["fraction_within_1e-4"] >= 0.96      # measured 0.80 / 0.98
    assert e["p50"] <= 2e-6 and e["p99"] <= 5e-4
    assert e["fraction_within_1e-5"] < 1.0                                              # i.e. the 1e-5 clause is NOT met sample by sample
    assert r["envelope_diff_over_total_inband_signal"]["max"] <= 1e-5                   # the bound that does hold (measured 1.6e-6)
    assert r["MskDf_abs_diff"]["max"] <= 1e-3 and r["MskDf_abs_diff"]["p50"] <= 1e-7    # PLL range is +-3.8e-3
    assert r["lvl_dB_abs_diff_max"] <= 0.05 and r["raw_frames_identical"] >= 8


@pytest.mark.parametrize("variant,K", [(0, 160), (1, 192)])
def test_cs16_fast_restatement(native, oracle, variant, K):
    """The folded form on CS16 input (soapy.c / sdrplay.c tables are sampled exponentials too): the library's plan equals
    the restatement's, the envelope stays within the literal arithmetic's table rounding and closer to the exact DFT
    bin than the literal tables are, and the messages are the same."""
    import ctypes as C
    freqs = (131.525, 131.725, 131.825, 131.450, 131.550)
    fd, _, fc = oracle.plan(K, freqs)
    k, tw = oracle.fast_plan_cs16(variant, K, fd, fc)
    f = np.asarray(fd, dtype=np.uint32)
    k2, tw2 = np.zeros(len(f), dtype=np.int32), np.zeros((len(f), K // 4, 2), dtype=np.float32)
    assert native.acb_fast_plan_cs16(variant, f.ctypes.data, len(f), K, int(fc), k2.ctypes.data, tw2.ctypes.data) == 1
    assert (k == k2).all() and np.array_equal(tw, tw2)
    assert oracle.fast_plan_cs16(variant, K, fd, fc - 12500) is None            # odd bins: no folded form
    plan = synth.StreamPlan(K=K, freqs_hz=tuple(fd), fc_hz=fc, seed=21, noise_sigma=1.5)
    rng = np.random.default_rng(21)
    for ch in range(len(freqs)):
        t = 0.01 + 0.02 * ch
        for _ in range(3):
            fr = synth.frame_bytes(synth.random_text(rng, int(rng.integers(8, 40))))
            plan.bursts.append(synth.Burst(chan=ch, t0=t, frame=fr, amp=float(rng.uniform(10, 25)), phase=float(rng.uniform(0, 6))))
            t += len(fr) * 8 / 2400 + 0.03
    n = int(0.5 * plan.rate) // K * K
    iq = synth.render_cs16(plan, 0, n)
    lit = oracle.channelize_cs16(variant, iq, K, oracle.cs16_osc(variant, K, fd, fc))
    fast = oracle.channelize_dft8_cs16(iq, K, k, tw)
    scale = 1.0 / 32768 if variant == 0 else 0.25
    x = iq.astype(np.float64).reshape(-1, K, 2)
    total = np.hypot(x[..., 0], x[..., 1]).sum(axis=1) / K * scale
    xc = x[..., 0] + 1j * x[..., 1]
    ideal = np.stack([np.abs(xc @ (np.exp(-2j * np.pi * int(kk) * np.arange(K) / K) / K * scale)) for kk in k])
    err_fast = (np.abs(fast - ideal) / total[None, :]).max()
    err_lit = (np.abs(lit - ideal) / total[None, :]).max()
    assert (np.abs(fast.astype(np.float64) - lit) / total[None, :]).max() <= TABLE_EPS
    assert err_fast <= IDEAL_EPS and err_fast < err_lit, (err_fast, err_lit)
    a, b = _decode(oracle, lit), _decode(oracle, fast)
    assert len(a) >= len(freqs) and [m[:-1] for m in a] == [m[:-1] for m in b]
    assert max(abs(x[-1] - y[-1]) for x, y in zip(a, b)) < 0.05                 # lvl, dB


@pytest.mark.parametrize("rate,fm", [(2500000, (131.525, 131.725, 131.825, 131.450)),
                                     (10000000, (131.125, 131.1375, 131.15, 131.1625, 131.55))])
def test_real_input_fast_restatement(native, oracle, rate, fm):
    """The real-input fast form (air.c's table is a sampled exponential of a whole bin number when Fc and the channels sit on
    the 12.5 kHz raster): library plan == restatement plan, envelope within the literal arithmetic's table rounding and
    closer to the exact DFT bin, same messages."""
    fd, fc, K = oracle.air_plan(rate, fm)
    k, tw = oracle.fast_plan_air(K, fd, fc)
    f = np.asarray(fd, dtype=np.uint32)
    k2, tw2 = np.zeros(len(f), dtype=np.int32), np.zeros((len(f), K // 4, 2), dtype=np.float32)
    assert native.acb_fast_plan_air(f.ctypes.data, len(f), K, int(fc), k2.ctypes.data, tw2.ctypes.data) == 1
    assert (k == k2).all() and np.array_equal(tw, tw2) and len(set(int(x) % 4 for x in k)) >= 2
    assert oracle.fast_plan_air(K, [fd[0] + 1000] + list(fd[1:]), fc) is None     # off the raster: no DFT bin
    plan = synth.StreamPlan(K=K, freqs_hz=tuple(fd), fc_hz=fc, seed=31, noise_sigma=1.0)
    rng = np.random.default_rng(31)
    for ch in range(len(fm)):
        t = 0.01 + 0.02 * ch
        for _ in range(2):
            fr = synth.frame_bytes(synth.random_text(rng, int(rng.integers(8, 40))))
            plan.bursts.append(synth.Burst(chan=ch, t0=t, frame=fr, amp=float(rng.uniform(10, 25)), phase=float(rng.uniform(0, 6))))
            t += len(fr) * 8 / 2400 + 0.03
    n = int(0.35 * rate) // K * K
    x = synth.render_real(plan, 0, n)
    lit = oracle.channelize_real(x, K, oracle.air_wf(rate, fm))
    fast = oracle.channelize_rdft(x, K, k, tw)
    xr = x.astype(np.float64).reshape(-1, K)
    total = np.abs(xr).sum(axis=1) / K
    ideal = np.stack([np.abs(xr @ (np.exp(-2j * np.pi * int(kk) * np.arange(K) / K) / K)) for kk in k])
    err_fast = (np.abs(fast - ideal) / total[None, :]).max()
    err_lit = (np.abs(lit - ideal) / total[None, :]).max()
    assert (np.abs(fast.astype(np.float64) - lit) / total[None, :]).max() <= TABLE_EPS
    assert err_fast <= IDEAL_EPS and err_fast < err_lit, (err_fast, err_lit)
    a, b = _decode(oracle, lit), _decode(oracle, fast)
    assert len(a) >= len(fm) and [m[:-1] for m in a] == [m[:-1] for m in b]
    assert max(abs(p[-1] - q[-1]) for p, q in zip(a, b)) < 0.05                 # lvl, dB
