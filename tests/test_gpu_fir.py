"""Generalised FIR of BASELINE configs[2] ("rateMult=192, 64 channels, FIR taps=165") and configs[4]
("wideband 20 MS/s IQ, 256 channels, FIR-tap sweep 65-513"): taps < K weight the first `taps`
samples of every K-sample row with an arbitrary complex table, in the reference's operation order.
The reference has no such mode (SURVEY.md note 1); the definition is oracle/acars_oracle.c's
orc_channelize_fir, which reduces to the pinned orc_channelize at taps == K."""
import numpy as np
import pytest

import refs
from acarsdec_b200 import api, synth
from common import bits_equal

pytestmark = pytest.mark.gpu


fir_tables = synth.fir_tables


def test_fir_equals_reference_arithmetic_at_taps_eq_k(native, oracle):
    K, fm = 160, synth.DEFAULT_FREQS_MHZ[:3]
    wf = oracle.wf(K, fm)
    iq = np.random.default_rng(0).integers(0, 256, size=2 * 2048 * K, dtype=np.uint8)
    assert bits_equal(oracle.channelize_fir(iq, K, K, wf), oracle.channelize(iq, K, wf))


@pytest.mark.parametrize("K,taps,nch,nblk", [
    (192, 165, 64, 2),        # config 3: taps not a multiple of 8 (zero-weight padding to 168)
    (1600, 65, 16, 1),        # config 5 sweep, shortest
    (1600, 257, 24, 1),
    (1600, 513, 8, 1),
    (160, 96, 5, 2),
    (44, 20, 2, 2),           # generic kernel (K % 8 != 0)
])
def test_fir_envelope_bit_exact(native, oracle, K, taps, nch, nblk):
    rate = K * 12500
    offs = [(-0.4 + 0.8 * i / max(nch - 1, 1)) * rate / 2 for i in range(nch)]
    wf = fir_tables(K, taps, offs, rate)
    rng = np.random.default_rng(K + taps)
    iq = rng.integers(0, 256, size=(2, nblk * 2048 * K), dtype=np.uint8)
    with api.Context(K, 2, nch, nblk, taps=taps) as ctx:
        for s in range(2):
            ctx.set_wf(s, wf)
        ctx.submit_host(iq, nblk)
        ctx.sync()
        got = ctx.read_dm(nblk * 1024)
        with pytest.raises(api.AcbError):
            ctx.set_plan(0, [131525000] * nch)          # the reference planner only builds K-tap tables
    for s in range(2):
        assert bits_equal(got[s], oracle.channelize_fir(iq[s], K, taps, wf).T.copy()), s


def test_fir_config3_frames(native, oracle):
    """Config 3 end to end: 64 channels, K=192, 165-tap windowed tables, injected frames: frames and
    demodulator states vs the oracle (channelize_fir -> orc_demod -> FEC)."""
    K, taps, nblk = 192, 165, 5
    fm = tuple(130.000 + 0.025 * i for i in range(64))
    fd, _, fc = api.plan(K, fm)
    rate = K * 12500
    wf = fir_tables(K, taps, [f - fc for f in fd], rate)
    plan = synth.make_plan(K, fm, fc, seconds=nblk * 1024 / 12500, seed=33, msgs_per_chan_per_sec=2.5, text_len=(5, 25), amp=(6.0, 10.0))
    iq = synth.render_blocks(plan, 0, nblk).reshape(1, -1)
    dm = oracle.channelize_fir(iq[0], K, taps, wf)
    sink = refs.Sink()
    want, chans = [], []
    for c in range(64):
        ch = oracle.new_chan(c)
        oracle.demod(ch, dm[c], sink)
        chans.append(ch)
    for m in sink.msgs():
        f = oracle.fec(m)
        if f is not None:
            want.append((f.chn, f.len, f.err, bytes(f.txt[:f.len]), bytes(f.crc)))
    with api.Context(K, 1, 64, nblk, taps=taps) as ctx:
        ctx.set_wf(0, wf)
        ctx.submit_host(iq, nblk)
        ctx.sync()
        got = [m.as_tuple() for m in ctx.drain()]
        for c in (0, 9, 40, 63):
            assert ctx.get_state(0, c).vec() == chans[c].vec(), c
    assert sorted(got) == sorted(want) and len(want) >= 10


@pytest.mark.parametrize("taps", [65, 513])
def test_fir_config5_full_width(native, oracle, taps):
    """BASELINE config 5 at its full width: K=1600 (20 MS/s), 256 channels on a 25 kHz raster, shortest and longest
    tap count of the sweep: envelope bit-exact for all 256 channels (32 channel groups in grid.z), and frames +
    demodulator state vs the oracle (channelize_fir -> orc_demod -> FEC) on injected messages."""
    K, nch, nblk = 1600, 256, 3
    fm = tuple(126.000 + 0.025 * i for i in range(nch))
    fd, _, fc = api.plan(K, fm)
    rate = K * 12500
    wf = fir_tables(K, taps, [f - fc for f in fd], rate)
    plan = synth.StreamPlan(K=K, freqs_hz=tuple(fd), fc_hz=fc, seed=5, noise_sigma=1.5)
    rng = np.random.default_rng(5)
    carriers = (0, 7, 8, 100, 255)                                   # first / last channel, both sides of a group boundary
    for i, ch in enumerate(carriers):
        fr = synth.frame_bytes(synth.random_text(rng, 6 + 2 * i))
        plan.bursts.append(synth.Burst(chan=ch, t0=0.008 + 0.01 * i, frame=fr, amp=9.0, phase=0.5 * i))
    iq = synth.render_blocks(plan, 0, nblk).reshape(1, -1)
    dm = oracle.channelize_fir(iq[0], K, taps, wf)                   # (256, nblk*1024)
    with api.Context(K, 1, nch, nblk, taps=taps) as ctx:
        ctx.set_wf(0, wf)
        ctx.submit_host(iq, nblk)
        ctx.sync()
        got = [m.as_tuple() for m in ctx.drain()]
        env = ctx.read_dm(nblk * 1024)
        states = {c: ctx.get_state(0, c).vec() for c in carriers + (1, 254)}
    assert bits_equal(env[0], dm.T.copy())
    # 65 taps at 20 MS/s pass ~300 kHz: a burst is decoded on its neighbours 25 kHz away too — every channel counts
    sink = refs.Sink()
    want = []
    for c in range(nch):
        ch = oracle.new_chan(c)
        oracle.demod(ch, dm[c], sink)
        if c in states:
            assert states[c] == ch.vec(), c
    for m in sink.msgs():
        f = oracle.fec(m)
        if f is not None:
            want.append((f.chn, f.len, f.err, bytes(f.txt[:f.len]), bytes(f.crc)))
    assert sorted(got) == sorted(want) and len(want) >= 3
