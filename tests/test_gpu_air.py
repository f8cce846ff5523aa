"""GPU parity for the Airspy front-end: float32 real input at IF = rate/4,
arbitrary submit lengths with the remainder carried, vs the CPU oracle (itself pinned against the
reference's air.c compiled in place, tests/test_air_oracle.py)."""
import numpy as np
import pytest

import refs
from acarsdec_b200 import api, synth
from common import bits_equal, msg_tuple

pytestmark = pytest.mark.gpu
FLAG_REAL = 2


def _plan(oracle, rate, fm, seconds, seed, nmsg_scale=1.0):
    fd, fc, K = oracle.air_plan(rate, fm)
    plan = synth.StreamPlan(K=K, freqs_hz=tuple(fd), fc_hz=fc, seed=seed, noise_sigma=1.0)
    rng = np.random.default_rng(seed)
    for ch in range(len(fm)):
        t = 0.01 + 0.03 * ch
        while True:
            fr = synth.frame_bytes(synth.random_text(rng, int(rng.integers(8, 40))))
            dur = len(fr) * 8 / 2400
            if t + dur + 0.01 > seconds:
                break
            plan.bursts.append(synth.Burst(chan=ch, t0=t, frame=fr, amp=float(rng.uniform(10, 25)), phase=float(rng.uniform(0, 6))))
            t += dur + 0.05 / nmsg_scale
    return plan, fd, fc, K


def _oracle_run(oracle, x, K, wf):
    """frames and final states for one stream's complete sample array (rows of K)."""
    dm = oracle.channelize_real(x[: (len(x) // K) * K], K, wf)
    chans = [oracle.new_chan(c) for c in range(wf.shape[0])]
    return dm, chans


@pytest.mark.parametrize("rate,fm", [
    (2500000, (131.525, 131.725, 131.825, 131.450)),                 # K=200
    (3000000, synth.DEFAULT_FREQS_MHZ),                              # K=240, 8 channels
    (6000000, (129.125, 130.025, 131.550)),                          # K=480, 3 channels (scalar stores)
])
def test_real_input_envelope_and_frames(native, oracle, rate, fm):
    secs = 0.5
    plan, fd, fc, K = _plan(oracle, rate, fm, secs, seed=rate // 100000)
    total = int(secs * rate)
    x = synth.render_real(plan, 0, total)
    wf = oracle.air_wf(rate, fm)
    assert bits_equal(api.build_wf_air(rate, fm), wf)
    nout_all = total // K
    max_blocks = nout_all // 1024 + 2
    rng = np.random.default_rng(1)
    with api.Context(K, 1, len(fm), max_blocks, flags=FLAG_REAL) as ctx:
        assert ctx.set_plan_air(0, fd) == fc
        # awkward submit sizes: below one row, a few rows, several thousand rows (pipeline + generic kernels)
        sizes, pos = [K // 3, 5 * K + 7, 1024 * K + 13, 3, 2 * 1024 * K + 999], 0
        dm_got, frames = [], []
        want_dm = oracle.channelize_real(x[: nout_all * K], K, wf)           # (nch, nout)
        done = 0
        i = 0
        while pos < total:
            n = sizes[i] if i < len(sizes) else int(rng.integers(1, 1500 * K))
            n = min(n, total - pos, (max_blocks * 1024 - 1) * K)
            i += 1
            m = ctx.submit_real(x[None, pos:pos + n])
            pos += n
            ctx.sync()
            frames += [msg_tuple(f) for f in ctx.drain()]
            if m:
                got = ctx.read_dm(m)[0]                                      # (m, nch)
                assert bits_equal(got, want_dm[:, done:done + m].T.copy()), (pos, m)
                done += m
        assert done == nout_all
        states = [ctx.get_state(0, c).vec() for c in range(len(fm))]
    # oracle demod over the whole envelope; per-submit emission groups make the global order
    # (submit, channel, time), so compare per channel
    sink = refs.Sink()
    want = []
    for c in range(len(fm)):
        ch = oracle.new_chan(c)
        oracle.demod(ch, want_dm[c], sink)
        assert states[c] == ch.vec(), c
    for msg in sink.msgs():
        f = oracle.fec(msg)
        if f is not None:
            want.append(msg_tuple(f))
    assert sorted(frames) == sorted(want) and len(want) >= len(fm)
    for c in range(len(fm)):
        assert [f for f in frames if f[0] == c] == [f for f in want if f[0] == c]


@pytest.mark.parametrize("rate,fm,warps", [
    (2500000, (131.525, 131.725, 131.825, 131.450), 2),              # K=200, odd residues
    (2500000, synth.DEFAULT_FREQS_MHZ, 4),                           # K=200, 8 channels, four warps per CTA (the default)
    (6000000, (129.125, 130.025, 131.550), 2),                       # K=480
    (3000000, synth.DEFAULT_FREQS_MHZ, 4),                           # K=240 (Airspy Mini)
    (10000000, (131.125, 131.1375, 131.15, 131.1625, 131.55, 131.825, 131.85, 131.475, 131.525, 131.725), 2),   # K=800, two groups, every residue
    (5000000, (131.45, 131.4625, 131.55), 2),                        # K=400 (the kernel only: air.c's 5 MS/s tuner-filter offset is not modelled)
])
def test_real_input_fast_form(native, oracle, monkeypatch, rate, fm, warps):
    """ACB_FLAG_FAST_CHANNELIZER on a real-input context (k_channelize_rdft): held to its CPU restatement bit for bit
    (orc_channelize_rdft; ragged submits, partial blocks), to the literal arithmetic within 1e-5 of the row's total
    signal, and to the same decoded messages."""
    monkeypatch.setenv("ACB_FAST_WARPS", str(warps))
    secs = 0.4
    plan, fd, fc, K = _plan(oracle, rate, fm, secs, seed=rate // 100000 + 1)
    total = int(secs * rate)
    x = synth.render_real(plan, 0, total)
    nout_all = total // K
    lit = oracle.channelize_real(x[: nout_all * K], K, oracle.air_wf(rate, fm))
    kbin, tw = oracle.fast_plan_air(K, fd, fc)
    want_dm = oracle.channelize_rdft(x[: nout_all * K], K, kbin, tw)
    tot = np.abs(x[: nout_all * K].astype(np.float64)).reshape(nout_all, K).sum(axis=1) / K
    assert (np.abs(want_dm.astype(np.float64) - lit) / tot[None, :]).max() <= 1e-5
    max_blocks = nout_all // 1024 + 2
    rng = np.random.default_rng(4)
    with api.Context(K, 1, len(fm), max_blocks, flags=FLAG_REAL | 8) as ctx:
        assert ctx.set_plan_air(0, fd) == fc
        sizes, pos, done, i, frames = [K // 3, 5 * K + 7, 1024 * K + 13, 3, 2 * 1024 * K + 999, 33 * K, 31 * K + 1], 0, 0, 0, []
        while pos < total:
            n = sizes[i] if i < len(sizes) else int(rng.integers(1, 1500 * K))
            n = min(n, total - pos, (max_blocks * 1024 - 1) * K)
            i += 1
            m = ctx.submit_real(x[None, pos:pos + n])
            pos += n
            ctx.sync()
            frames += [msg_tuple(f) for f in ctx.drain()]
            if m:
                assert bits_equal(ctx.read_dm(m)[0], want_dm[:, done:done + m].T.copy()), (pos, m)
                done += m
        assert done == nout_all
        st = ctx.stats()
        assert st.fast_chan_launches == st.chan_launches > 0
    sink, want = refs.Sink(), []
    for c in range(len(fm)):
        oracle.demod(oracle.new_chan(c), lit[c], sink)         # the LITERAL envelope: same messages expected
    for msg in sink.msgs():
        f = oracle.fec(msg)
        if f is not None:
            want.append(msg_tuple(f))
    strip = lambda t: t[:5]                                    # (chn, len, err, txt, crc): lvl moves by < 0.05 dB
    assert sorted(map(strip, frames)) == sorted(map(strip, want)) and len(want) >= 3


def test_real_input_multistream_chunking_independence(native, oracle):
    rate, fm = 2500000, (131.525, 131.725, 131.825)
    fd, fc, K = oracle.air_plan(rate, fm)
    secs, nstreams = 0.25, 3
    total = int(secs * rate)
    xs = np.stack([synth.render_real(_plan(oracle, rate, fm, secs, seed=50 + s)[0], 0, total) for s in range(nstreams)])
    outs = []
    for split in ([total], [100001, 50000, total - 150001], [7] * 10 + [total - 70]):
        with api.Context(K, nstreams, 3, total // K // 1024 + 2, flags=FLAG_REAL) as ctx:
            for s in range(nstreams):
                ctx.set_plan_air(s, fd)
            pos = 0
            for n in split:
                ctx.submit_real(np.ascontiguousarray(xs[:, pos:pos + n]))
                pos += n
            ctx.sync()
            msgs = sorted((m.stream,) + msg_tuple(m) for m in ctx.drain())
            st = [[ctx.get_state(s, c).vec() for c in range(3)] for s in range(nstreams)]
            outs.append((msgs, st))
    assert outs[0] == outs[1] == outs[2]
    assert len(outs[0][0]) >= 3


def test_real_input_argument_errors(native):
    with api.Context(200, 1, 2, 2, flags=FLAG_REAL) as ctx:
        with pytest.raises(api.AcbError):
            ctx.submit_host(np.zeros((1, 2048 * 200), dtype=np.uint8), 1)
        with pytest.raises(api.AcbError):
            ctx.set_plan(0, [131525000, 131725000])
        with pytest.raises(api.AcbError):
            ctx.submit_real(np.zeros((1, 3 * 1024 * 200), dtype=np.float32))
    with api.Context(160, 1, 2, 2) as ctx:
        with pytest.raises(api.AcbError):
            ctx.submit_real(np.zeros((1, 4096), dtype=np.float32))
