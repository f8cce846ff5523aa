"""CPU check of the demodulator loop the CUDA kernel k_demod2 runs (acarsdec_b200/csrc/demod_core.h).

tests/host/demod_emul.cpp compiles that header for the host as a single lane (test infrastructure, not
product code) and this file holds it to the oracle (oracle/acars_oracle.c: orc_demod, pinned against the
unmodified msk.c/acars.c) bit for bit: every state variable after every chunk, every pre-FEC frame.
What it covers: the fast-path / general-path selection, the 5-or-6 sample bit period, the doubled ring,
the one-division decision with its guard, the phase-index guard, both bit-clock rounding forms.  What it
cannot cover — which lane of a channel's group evaluates which mixer sample — is in tests/test_gpu_parity.py.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import refs
from acarsdec_b200 import api, synth
from common import load_testwav

HOST = Path(__file__).resolve().parent / "host"
ROOT = HOST.parent.parent
EMUL = HOST / "libdemod_emul.so"
SRC = [HOST / "demod_emul.cpp", ROOT / "acarsdec_b200" / "csrc" / "demod_core.h",
       ROOT / "acarsdec_b200" / "csrc" / "frame_sm.h", ROOT / "acarsdec_b200" / "csrc" / "acb_internal.h"]


class EmulFrame(C.Structure):
    _fields_ = [("len", C.c_int), ("err", C.c_int), ("bitcount", C.c_int), ("pad", C.c_int), ("lvlsum", C.c_double),
                ("pos", C.c_uint64), ("soh_pos", C.c_uint64), ("crc", C.c_ubyte * 2), ("txt", C.c_ubyte * 250)]


@pytest.fixture(scope="module")
def emul(native):
    lib = ROOT / "acarsdec_b200" / "libacars_b200.so"
    if not EMUL.exists() or any(p.stat().st_mtime > EMUL.stat().st_mtime for p in SRC + [lib]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-o", str(EMUL), str(SRC[0]), "-L", str(lib.parent), "-lacars_b200",
                        "-Wl,-rpath," + str(lib.parent), "-lm"], check=True)
    L = C.CDLL(str(EMUL))
    L.demod_emul.argtypes = [C.POINTER(api.ChanState), C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(EmulFrame), C.c_int]
    return L


def fresh_state() -> api.ChanState:
    st = api.ChanState()
    st.nbits = 8                      # initMsk msk.c:34-40 zeroes; initAcars acars.c:230-234
    return st


def run_both(emul, oracle, x: np.ndarray, chunks, f2f: int):
    """x: float32 envelope of one channel.  Feeds the same chunks to the oracle and to the emulation and
    compares state after every chunk and the frames at the end."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    oc = oracle.new_chan(0)
    sink = refs.Sink()
    st = fresh_state()
    frames = (EmulFrame * 64)()
    got, pos = [], 0
    for n in chunks:
        seg = x[pos:pos + n]
        if len(seg) == 0:
            break
        oracle.demod(oc, seg, sink)
        k = emul.demod_emul(C.byref(st), seg.ctypes.data, len(seg), 1, f2f, frames, 64)
        assert k <= 64
        for i in range(k):
            f = frames[i]
            lvl = np.float32(10 * np.log10(f.lvlsum / f.bitcount)) if f.bitcount else np.float32(0)
            got.append((f.len, f.err, bytes(f.txt[:f.len]), bytes(f.crc), int(lvl.view(np.uint32)), int(f.pos)))
        pos += len(seg)
        assert st.vec() == oc.vec(), f"state differs after {pos} samples"
        assert st.pos == pos
    want = [(m.len, m.err, bytes(m.txt[:m.len]), bytes(m.crc), int(np.float32(m.lvl).view(np.uint32))) for m in sink.msgs()]
    assert [g[:5] for g in got] == want
    return got


@pytest.mark.parametrize("f2f", [0, 1])
def test_testwav_state_trace_bit_exact(emul, oracle, f2f):
    """test.wav (BASELINE config 1), all 4 channels, soundfile.c's 4096-frame chunks: the seven known messages'
    raw frames and a bit-identical state after every chunk."""
    x, exp = load_testwav()
    total = 0
    for c in range(4):
        total += len(run_both(emul, oracle, x[:, c], [4096] * 14, f2f))
    assert total == 7 == len(exp["messages"])


def test_chunking_independence_and_odd_lengths(emul, oracle):
    """Chunks of 1..3000 samples, so launches end in the middle of bit periods (general path) everywhere."""
    x, _ = load_testwav()
    rng = np.random.default_rng(11)
    for c in range(4):
        chunks = [int(v) for v in rng.integers(1, 3000, size=200)]
        run_both(emul, oracle, x[:, c], chunks, 0)
    run_both(emul, oracle, x[:, 0], [1] * 700 + [7] * 300 + [100000], 0)


def test_synthetic_envelope_with_noise_and_silence(emul, oracle):
    """The oracle's own channelizer output for a seeded synthetic capture (frames, noise, gaps), plus pure
    zeros (lvl = 0: the division guard's slow path, vs = 0) and a large-amplitude stretch."""
    K = 160
    fm = synth.DEFAULT_FREQS_MHZ
    _, _, fc = api.plan(K, fm)
    plan = synth.make_plan(K, fm, fc, seconds=1.0, seed=21)
    nblk = synth.blocks_for_seconds(K, 1.0)
    iq = synth.render_blocks(plan, 0, nblk).reshape(-1)
    dm = oracle.channelize(iq, K, oracle.wf(K, fm))          # [nch][nsamp]
    nframes = 0
    for c in range(dm.shape[0]):
        nframes += len(run_both(emul, oracle, dm[c], [1024] * nblk, 0))
    assert nframes >= 3
    z = np.concatenate([np.zeros(3000, np.float32), dm[0][:4000] * np.float32(1e7), np.zeros(2000, np.float32),
                        dm[1][:3000] * np.float32(1e-30)])
    run_both(emul, oracle, z, [1024] * 12, 0)
    run_both(emul, oracle, z, [999] * 13, 1)


def test_foreign_state_takes_the_general_path(emul, oracle):
    """A state no reset produces (bit clock far ahead, big MskDf): the loop must follow the reference anyway."""
    x, _ = load_testwav()
    seg = np.ascontiguousarray(x[:6000, 2])
    for clk, df in [(4.9, 0.0), (7.5, 0.02), (-3.0, -0.03), (0.1, 0.3)]:
        oc = oracle.new_chan(0)
        oc.MskClk = clk
        oc.MskDf = df
        st = fresh_state()
        st.MskClk = clk
        st.MskDf = df
        frames = (EmulFrame * 8)()
        oracle.demod(oc, seg, refs.Sink())
        emul.demod_emul(C.byref(st), seg.ctypes.data, len(seg), 1, 0, frames, 8)
        assert st.vec() == oc.vec(), (clk, df)
