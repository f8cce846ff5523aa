"""GPU parity for the CS16 front-ends (soapy.c:232-254, sdrplay.c:215-236):
int16 I,Q input (interleaved, or planar like the SDRplay callback), any submit length with the
remainder carried, vs the CPU oracle's LITERAL arithmetic (pinned against both references compiled
in place, tests/test_cs16_oracle.py)."""
import numpy as np
import pytest

import refs
from acarsdec_b200 import api, synth
from common import bits_equal, msg_tuple

pytestmark = pytest.mark.gpu
FLAG_CS16 = 4
FREQS = (131.525, 131.725, 131.825, 131.450, 131.550)
LVL_FIELD = 5          # msg_tuple: (chn, len, err, txt, crc, lvl bits) — the fast form moves lvl by < 0.05 dB


def _capture(oracle, K, seconds, seed):
    fd, _, fc = oracle.plan(K, FREQS)
    plan = synth.StreamPlan(K=K, freqs_hz=tuple(fd), fc_hz=fc, seed=seed, noise_sigma=1.5)
    rng = np.random.default_rng(seed)
    for ch in range(len(FREQS)):
        t = 0.01 + 0.03 * ch
        while True:
            fr = synth.frame_bytes(synth.random_text(rng, int(rng.integers(8, 40))))
            dur = len(fr) * 8 / 2400
            if t + dur + 0.01 > seconds:
                break
            plan.bursts.append(synth.Burst(chan=ch, t0=t, frame=fr, amp=float(rng.uniform(10, 25)), phase=float(rng.uniform(0, 6))))
            t += dur + 0.04
    n = int(seconds * plan.rate) + 77
    return synth.render_cs16(plan, 0, n), fd, fc


@pytest.mark.parametrize("variant,K,planar", [(0, 160, False), (0, 192, False), (1, 160, True), (1, 160, False)])
def test_cs16_envelope_frames_states(native, oracle, variant, K, planar):
    iq, fd, fc = _capture(oracle, K, 0.45, seed=3 + variant)
    osc = oracle.cs16_osc(variant, K, fd, fc)
    nout_all = len(iq) // K
    want_dm = oracle.channelize_cs16(variant, iq[: nout_all * K], K, osc)
    max_blocks = nout_all // 1024 + 2
    rng = np.random.default_rng(2)
    with api.Context(K, 1, len(FREQS), max_blocks, flags=FLAG_CS16) as ctx:
        assert ctx.set_plan_cs16(0, fd, variant) == fc
        sizes, pos, done, i, frames = [K // 2, 3 * K + 5, 1024 * K + 9, 1, 2048 * K + 100], 0, 0, 0, []
        while pos < len(iq):
            n = sizes[i] if i < len(sizes) else int(rng.integers(1, 1200 * K))
            n = min(n, len(iq) - pos, (max_blocks * 1024 - 1) * K)
            i += 1
            part = np.ascontiguousarray(iq[None, pos:pos + n])
            m = ctx.submit_cs16_planar(np.ascontiguousarray(part[:, :, 0]), np.ascontiguousarray(part[:, :, 1])) if planar else ctx.submit_cs16(part)
            pos += n
            ctx.sync()
            frames += [msg_tuple(f) for f in ctx.drain()]
            if m:
                assert bits_equal(ctx.read_dm(m)[0], want_dm[:, done:done + m].T.copy()), (pos, m)
                done += m
        assert done == nout_all
        states = [ctx.get_state(0, c).vec() for c in range(len(FREQS))]
    sink = refs.Sink()
    want = []
    for c in range(len(FREQS)):
        ch = oracle.new_chan(c)
        oracle.demod(ch, want_dm[c], sink)
        assert states[c] == ch.vec(), c
    for msg in sink.msgs():
        f = oracle.fec(msg)
        if f is not None:
            want.append(msg_tuple(f))
    assert sorted(frames) == sorted(want) and len(want) >= len(FREQS)


def test_cs16_user_centre_frequency_and_errors(native, oracle):
    """soapy.c:132-133 honours a user-supplied centre frequency instead of chooseFc."""
    K = 160
    fd, _, fc = oracle.plan(K, FREQS)
    user_fc = fc - 12500
    iq = np.random.default_rng(9).integers(-2000, 2000, size=(1, 2048 * K, 2), dtype=np.int16)
    with api.Context(K, 1, len(FREQS), 3, flags=FLAG_CS16) as ctx:
        assert ctx.set_plan_cs16(0, fd, 0, user_fc) == user_fc
        assert ctx.submit_cs16(iq) == 2048
        ctx.sync()
        got = ctx.read_dm(2048)[0]
        with pytest.raises(api.AcbError):
            ctx.submit_real(np.zeros((1, 100), dtype=np.float32))
        with pytest.raises(api.AcbError):
            ctx.set_plan_cs16(0, fd, 7)
    osc = oracle.cs16_osc(0, K, fd, user_fc)
    assert bits_equal(got, oracle.channelize_cs16(0, iq[0], K, osc).T.copy())
    with pytest.raises(api.AcbError):
        api.Context(K, 1, 2, 2, flags=FLAG_CS16 | 2)


@pytest.mark.parametrize("variant,K,planar,warps", [(0, 160, False, 2), (0, 192, False, 2), (1, 160, True, 2), (1, 192, False, 1), (0, 160, False, 1)])
def test_cs16_fast_form(native, oracle, monkeypatch, variant, K, planar, warps):
    """ACB_FLAG_FAST_CHANNELIZER on a CS16 context: the folded DFT form of the u8 path with the int16 converter.  Held to
    (a) its CPU restatement bit for bit (orc_channelize_dft8_cs16), ragged submits and partial blocks included,
    (b) the literal arithmetic within 1e-5 of the row's total in-band signal, (c) the same decoded messages."""
    monkeypatch.setenv("ACB_FAST_WARPS", str(warps))
    iq, fd, fc = _capture(oracle, K, 0.45, seed=11 + variant)
    osc = oracle.cs16_osc(variant, K, fd, fc)
    nout_all = len(iq) // K
    exact_dm = oracle.channelize_cs16(variant, iq[: nout_all * K], K, osc)
    kbin, tw = oracle.fast_plan_cs16(variant, K, fd, fc)
    want_dm = oracle.channelize_dft8_cs16(iq[: nout_all * K], K, kbin, tw)
    x = iq[: nout_all * K].astype(np.float64).reshape(nout_all, K, 2)
    total = np.hypot(x[..., 0], x[..., 1]).sum(axis=1) / K * (1.0 / 32768 if variant == 0 else 0.25)
    assert (np.abs(want_dm.astype(np.float64) - exact_dm) / total[None, :]).max() <= 1e-5
    max_blocks = nout_all // 1024 + 2
    rng = np.random.default_rng(5)
    with api.Context(K, 1, len(FREQS), max_blocks, flags=FLAG_CS16 | 8) as ctx:
        assert ctx.set_plan_cs16(0, fd, variant) == fc
        sizes, pos, done, i, frames = [K // 2, 3 * K + 5, 1024 * K + 9, 1, 2048 * K + 100, 33 * K, 31 * K + 7], 0, 0, 0, []
        while pos < len(iq):
            n = sizes[i] if i < len(sizes) else int(rng.integers(1, 1200 * K))
            n = min(n, len(iq) - pos, (max_blocks * 1024 - 1) * K)
            i += 1
            part = np.ascontiguousarray(iq[None, pos:pos + n])
            m = ctx.submit_cs16_planar(np.ascontiguousarray(part[:, :, 0]), np.ascontiguousarray(part[:, :, 1])) if planar else ctx.submit_cs16(part)
            pos += n
            ctx.sync()
            frames += [msg_tuple(f) for f in ctx.drain()]
            if m:
                assert bits_equal(ctx.read_dm(m)[0], want_dm[:, done:done + m].T.copy()), (pos, m)
                done += m
        assert done == nout_all
        st = ctx.stats()
        assert st.fast_chan_launches == st.chan_launches > 0
    sink = refs.Sink()
    want = []
    for c in range(len(FREQS)):
        ch = oracle.new_chan(c)
        oracle.demod(ch, exact_dm[c], sink)           # the LITERAL envelope: the fast form must decode the same messages
    for msg in sink.msgs():
        f = oracle.fec(msg)
        if f is not None:
            want.append(msg_tuple(f))
    strip = lambda t: tuple(v for i, v in enumerate(t) if i != LVL_FIELD)
    assert sorted(map(strip, frames)) == sorted(map(strip, want)) and len(want) >= len(FREQS)


def test_cs16_fast_needs_the_raster(native, oracle):
    """A user centre frequency off the 25 kHz raster leaves the context on the exact kernel."""
    K = 160
    fd, _, fc = oracle.plan(K, FREQS)
    iq = np.random.default_rng(9).integers(-2000, 2000, size=(1, 1024 * K, 2), dtype=np.int16)
    with api.Context(K, 1, len(FREQS), 2, flags=FLAG_CS16 | 8) as ctx:
        ctx.set_plan_cs16(0, fd, 0, fc - 12500)                    # odd bin numbers
        assert ctx.submit_cs16(iq) == 1024
        ctx.sync()
        got = ctx.read_dm(1024)[0]
        assert ctx.stats().fast_chan_launches == 0
    assert bits_equal(got, oracle.channelize_cs16(0, iq[0], K, oracle.cs16_osc(0, K, fd, fc - 12500)).T.copy())
