"""Pins the oracle: the C restatement (oracle/acars_oracle.c) against the unmodified reference
compiled in place (oracle/_ref) and against the committed golden vectors.  CPU only."""
import hashlib

import numpy as np
import pytest

import refs
from acarsdec_b200 import synth
from common import (bits_equal, load_synth_k16, load_testwav, msg_tuple, msg_tuple_from_json,
                    state_tuple_from_json)


def test_tables_match_reference(oracle, reflib):
    syn, crc, nb = reflib.tables()
    assert len(syn) == 1936
    assert all(oracle.lib.orc_syndrome(i & 7, i >> 3) == syn[i] for i in range(len(syn)))
    assert all(oracle.lib.orc_crc_step(0, i) == crc[i] for i in range(256))
    assert all(oracle.lib.orc_odd_parity(i) == (nb[i] & 1) for i in range(256))


@pytest.mark.parametrize("K,freqs", [
    (160, synth.DEFAULT_FREQS_MHZ),
    (192, (131.525, 131.725, 131.825)),
    (200, (129.125, 130.025, 130.450, 131.125, 131.550)),
    (160, (131.5125, 131.7375, 131.2625)),          # off the 8 Hz float grid: exercises rtl.c:255
    (16, (131.525, 131.550, 131.475)),
])
def test_plan_and_tables_match_reference(oracle, reflib, K, freqs):
    reflib.open_rtl(K, freqs)
    fd, fr, fc = oracle.plan(K, freqs)
    assert fc == reflib.fc
    assert fr == [reflib.chan_freq(i) for i in range(len(freqs))]
    wf = oracle.wf(K, freqs)
    for i in range(len(freqs)):
        assert bits_equal(wf[i], reflib.wf(i))


def test_testwav_golden_restatement(oracle):
    """The 7 known messages of test.wav, final states and bit count, through the
    restatement, using the soundfile.c chunking."""
    x, exp = load_testwav()
    chans = [oracle.new_chan(c) for c in range(4)]
    sink = refs.Sink()
    got = []
    for s in range(0, len(x), 4096):
        for c in range(4):
            oracle.demod(chans[c], x[s:s + 4096, c], sink)
        for m in sink.msgs():
            f = oracle.fec(m)
            if f is not None:
                got.append(msg_tuple(f))
        sink.c.nmsg = 0
    assert got == [msg_tuple_from_json(j) for j in exp["messages"]]
    assert len(got) == 7
    assert [chans[c].vec() for c in range(4)] == [state_tuple_from_json(j) for j in exp["final_state"]]
    assert sum(c.nbit_total for c in chans) == exp["putbit_calls"]


def test_testwav_state_trace_vs_reference(oracle, reflib):
    """Per-chunk state equality, reference vs restatement, with awkward chunk sizes."""
    x, _ = load_testwav()
    reflib.open_audio(4)
    chans = [oracle.new_chan(c) for c in range(4)]
    rng = np.random.default_rng(5)
    pos = 0
    n = 20000
    while pos < n:
        step = int(rng.integers(1, 700))
        for c in range(4):
            seg = x[pos:pos + step, c]
            reflib.audio(c, seg)
            oracle.demod(chans[c], seg)
            assert reflib.state(c).vec() == chans[c].vec(), (pos, c)
        pos += step


def test_synth_k16_golden(oracle, reflib):
    iq, exp = load_synth_k16()
    K = exp["K"]
    assert hashlib.sha256(iq.tobytes()).hexdigest() == exp["iq_sha256"]
    wf = oracle.wf(K, exp["freqs_mhz"])
    assert hashlib.sha256(wf.tobytes()).hexdigest() == exp["wf_sha256"]
    st = refs.OracleStream(oracle, K, wf)
    sha = hashlib.sha256()
    for b in range(exp["nblk"]):
        st.blocks(iq[b])
        for c in range(3):
            sha.update(st.dm(c).tobytes())
    assert sha.hexdigest() == exp["dm_sha256"]
    assert [msg_tuple(m) for m in st.msgs()] == [msg_tuple_from_json(j) for j in exp["messages"]]
    assert [st.chan(c).vec() for c in range(3)] == [state_tuple_from_json(j) for j in exp["final_state"]]
    # and live against the reference
    reflib.open_rtl(K, exp["freqs_mhz"])
    for b in range(exp["nblk"]):
        reflib.block(iq[b])
    assert [msg_tuple(m) for m in reflib.msgs()] == [msg_tuple_from_json(j) for j in exp["messages"]]


@pytest.mark.parametrize("K,seed", [(160, 3), (192, 4)])
def test_full_path_synthetic_vs_reference(oracle, reflib, K, seed):
    """Seeded multi-channel IQ with injected messages: dm, state after every block and messages,
    restatement vs reference, bit for bit."""
    fm = synth.DEFAULT_FREQS_MHZ
    reflib.open_rtl(K, fm)
    _, _, fc = oracle.plan(K, fm)
    plan = synth.make_plan(K, fm, fc, seconds=0.7, seed=seed)
    nblk = synth.blocks_for_seconds(K, 0.7)
    iq = synth.render_blocks(plan, 0, nblk)
    st = refs.OracleStream(oracle, K, oracle.wf(K, fm))
    for b in range(nblk):
        reflib.block(iq[b])
        st.blocks(iq[b])
        for c in range(len(fm)):
            assert bits_equal(reflib.dm(c), st.dm(c)), (b, c)
            assert reflib.state(c).vec() == st.chan(c).vec(), (b, c)
    rm, om = reflib.msgs(), st.msgs()
    assert [msg_tuple(m) for m in rm] == [msg_tuple(m) for m in om]
    assert len(rm) >= len(plan.bursts) - 1 and len(rm) > 0       # the generator's frames decode


def _fec_cases(rng, n):
    """Random pre-FEC blocks: valid frames with 0..5 flipped bits in text and/or BCS, plus junk."""
    out = []
    for i in range(n):
        body = synth.frame_bytes(synth.random_text(rng, int(rng.integers(0, 200))), prekey=0)[5:-1]
        txt, crc = bytearray(body[:-2]), bytearray(body[-2:])
        kind = int(rng.integers(0, 8))
        nflip = [0, 1, 1, 2, 2, 3, 4, 5][kind]
        for _ in range(nflip):
            where = int(rng.integers(0, len(txt) + 2))
            if kind == 4 and _ == 1:                       # second flip in the same byte: fixdberr's case
                where = last
            last = where
            bit = 1 << int(rng.integers(0, 8))
            if where < len(txt):
                txt[where] ^= bit
            else:
                crc[where - len(txt)] ^= bit
        if kind == 7 and rng.random() < 0.5:
            txt = bytearray(rng.integers(0, 256, size=int(rng.integers(5, 240)), dtype=np.uint8).tobytes())
        out.append((i % 16, bytes(txt), bytes(crc)))
    return out


def test_block_fec_fuzz_vs_reference(oracle, reflib):
    """acars.c:93-215 (parity, CRC, fixprerr, fixdberr, drops): reference blk_thread vs restatement."""
    reflib.open_audio(1)
    rng = np.random.default_rng(99)
    cases = _fec_cases(rng, 1500)
    want = []
    for chn, txt, crc in cases:
        reflib.push_block(chn, txt, crc)
        m = refs.Msg()
        m.chn, m.len = chn, len(txt)
        m.txt[:len(txt)] = txt
        m.crc[:] = crc
        f = oracle.fec(m)
        if f is not None:
            want.append(f.as_tuple())
    got = [m.as_tuple() for m in reflib.msgs()]
    assert got == want
    assert 200 < len(got) < len(cases)              # both repairs and drops happened
    assert any(t[2] > 0 for t in got)


def test_corrupted_frames_on_air_vs_reference(oracle, reflib):
    """Bit errors injected into the transmitted frames, so the FEC paths run behind the demod."""
    K = 160
    fm = (131.525, 131.725, 131.825)
    reflib.open_rtl(K, fm)
    _, _, fc = oracle.plan(K, fm)
    plan = synth.make_plan(K, fm, fc, seconds=1.5, seed=21, msgs_per_chan_per_sec=6.0, text_len=(5, 40))
    flips = [[(3, 0x04)], [(5, 0x01), (9, 0x80)], [(7, 0x21)], [(-2, 0x10)], [(2, 1), (4, 2), (6, 4), (8, 8)],
             [(1, 0x40), (-3, 0x02)], [], [(20, 0xFF)], [(0, 0x08), (10, 0x08), (11, 0x08)]]
    for i, b in enumerate(plan.bursts):
        b.frame = synth.corrupt_frame(b.frame, [f for f in flips[i % len(flips)] if f[0] < len(b.frame) - 30])
    nblk = synth.blocks_for_seconds(K, 1.5)
    iq = synth.render_blocks(plan, 0, nblk)
    st = refs.OracleStream(oracle, K, oracle.wf(K, fm))
    for b in range(nblk):
        reflib.block(iq[b])
        st.blocks(iq[b])
    rm = reflib.msgs()
    assert [msg_tuple(m) for m in rm] == [msg_tuple(m) for m in st.msgs()]
    assert [reflib.state(c).vec() for c in range(3)] == [st.chan(c).vec() for c in range(3)]
    assert any(m.err > 0 for m in rm) and 0 < len(rm) < len(plan.bursts)
