"""CS16 front-ends (soapy.c interleaved int16 with /32768.0 in double, sdrplay.c planar int16 with
envelope/4): restatement vs the unmodified references compiled in place, and the product's folded
tables (powers of two moved into the table) vs the literal arithmetic.  CPU only."""
import numpy as np
import pytest

import refs
from acarsdec_b200 import api, synth
from common import bits_equal, msg_tuple

FREQS = (131.525, 131.725, 131.825, 131.450)


@pytest.fixture(params=["soapy", "sdrplay"])
def cs16ref(request):
    refs.ensure_built()
    if not (refs.ORACLE_DIR / "_ref" / f"libacarsref_{request.param}_O2.so").exists():
        pytest.skip("oracle/_ref CS16 builds absent")
    r = refs.RefCs16Lib(request.param)
    yield r
    r.close()


def test_cs16_plan_tables_carry_and_messages(oracle, cs16ref):
    variant = 0 if cs16ref.which == "soapy" else 1
    K = 160
    cs16ref.open(FREQS, K)
    fd, _, fc = oracle.plan(K, FREQS)
    assert cs16ref.fc == fc
    osc = oracle.cs16_osc(variant, K, fd, fc)
    for c in range(len(FREQS)):
        assert bits_equal(osc[c], cs16ref.osc(c)), c
    plan = synth.StreamPlan(K=K, freqs_hz=tuple(fd), fc_hz=fc, seed=12, noise_sigma=1.5)
    rng = np.random.default_rng(12)
    for ch in range(4):
        plan.bursts.append(synth.Burst(chan=ch, t0=0.01 + 0.04 * ch, frame=synth.frame_bytes(synth.random_text(rng, 15 + 8 * ch)),
                                       amp=15.0 + 4 * ch, phase=0.7 * ch))
    nblock = 1024 if variant == 0 else 512                   # demodMSK chunk: soapy.c:246 / sdrplay.c:228
    total = (int(0.5 * plan.rate) // (K * nblock)) * K * nblock + 37 * K + 11      # not a whole number of rows
    iq = synth.render_cs16(plan, 0, total)
    sizes = [4000, 163840, 7, 159, 161, 100000] if variant == 0 else [336, 252, 1000, 7, 159, 161]
    cs16ref.feed(iq, sizes)
    nout = total // K
    dm = oracle.channelize_cs16(variant, iq[: nout * K], K, osc)
    # the reference demodulates in chunks of nblock outputs; what is left sits in dm_buffer / D
    done = (nout // nblock) * nblock
    chans = [oracle.new_chan(c) for c in range(4)]
    sink = refs.Sink()
    got = []
    for b in range(0, done, nblock):
        for c in range(4):
            oracle.demod(chans[c], dm[c, b:b + nblock], sink)
        for m in sink.msgs():
            f = oracle.fec(m)
            if f is not None:
                got.append(msg_tuple(f))
        sink.c.nmsg = 0
    for c in range(4):
        assert cs16ref.counter(c) == nout - done
        assert bits_equal(cs16ref.dm(c, nout - done), dm[c, done:]), c
        assert cs16ref.state(c).vec() == chans[c].vec(), c
    assert [msg_tuple(m) for m in cs16ref.msgs()] == got and len(got) == 4
    # the product's tables fold soapy's /32768.0 and sdrplay's envelope/4 into the oscillator
    wf = api.build_wf_cs16(variant, K, fd, fc)
    scale = np.float32(1 / 32768) if variant == 0 else np.float32(0.25)
    assert bits_equal(wf, osc * scale)


def test_cs16_folded_tables_equal_literal_arithmetic(oracle):
    """fl32(fl64(D + p/32768)) == fl32(D + p') with p' computed from the pre-scaled table, and
    hypot(D)/4 == hypot(D') for the quarter-scaled table: same envelope bits."""
    K = 160
    fd, _, fc = oracle.plan(K, FREQS)
    iq = np.random.default_rng(5).integers(-32768, 32768, size=(64 * K, 2), dtype=np.int16)
    iq[:K] = 32767
    iq[K:2 * K] = -32768
    for variant in (0, 1):
        osc = oracle.cs16_osc(variant, K, fd, fc)
        lit = oracle.channelize_cs16(variant, iq, K, osc)
        scale = np.float32(1 / 32768) if variant == 0 else np.float32(0.25)
        plain = oracle.channelize_cs16(2, iq, K, osc * scale)      # plain D += v*w ; |D|
        assert bits_equal(lit, plain), variant
