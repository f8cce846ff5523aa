"""CPU-only tests of the product's host side: the C-ABI library loads and exports every symbol
include/acars_b200.h declares, the host planning functions and the block FEC agree with the
oracle bit for bit, and (no GPU here) creating a context fails loudly instead of falling back."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import refs
from acarsdec_b200 import api, synth
from common import bits_equal
from conftest import HAVE_GPU
from test_oracle_vs_reference import _fec_cases

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(native):
    header = (ROOT / "include" / "acars_b200.h").read_text()
    declared = set(re.findall(r"\b(acb_[a-z0-9_]+)\s*\(", header))
    bound = {name for name, _, _ in api.ABI}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(native, name)
    assert b"sm_100a" in native.acb_version()


def test_struct_sizes_match_header(native):
    # acb_msg_t / acb_chan_state_t layouts as the C compiler sees them
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "acars_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n",sizeof(acb_msg_t),sizeof(acb_chan_state_t),sizeof(acb_stats_t),sizeof(acb_config_t));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "s.c").write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", f"{d}/s", f"{d}/s.c"], check=True)
        out = subprocess.run([f"{d}/s"], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [C.sizeof(api.Msg), C.sizeof(api.ChanState), C.sizeof(api.Stats), C.sizeof(api.Config)]


@pytest.mark.parametrize("K,freqs", [
    (160, synth.DEFAULT_FREQS_MHZ),
    (192, (131.525, 131.725, 131.825)),
    (200, (129.125, 130.025, 130.450, 131.125, 131.550)),
    (160, (131.5125, 131.7375, 131.2625)),
    (320, (118.000, 119.500, 121.000)),
    (16, (131.525, 131.550, 131.475)),
])
def test_planning_matches_oracle(native, oracle, K, freqs):
    assert api.plan(K, freqs) == oracle.plan(K, freqs)
    assert bits_equal(api.build_wf(K, freqs), oracle.wf(K, freqs))


def test_plan_rejects_wide_span(native, oracle):
    f = np.array([118000000, 121000000], dtype=np.uint32)
    assert native.acb_choose_fc(f.ctypes.data, 2, 160) == 0
    assert oracle.lib.orc_choose_fc(f.ctypes.data, 2, 160) == 0


def test_matched_filter_matches_oracle(native, oracle):
    assert bits_equal(api.build_h(), oracle.h)


def test_crc_and_syndrome_tables(native, oracle):
    for i in range(256):
        for crc in (0, 0x1234, 0xFFFF):
            assert native.acb_crc_update(crc, i) == oracle.lib.orc_crc_step(crc, i)
    for i in range(1936):
        assert native.acb_syndrome(i) == oracle.lib.orc_syndrome(i & 7, i >> 3)


def test_block_fec_fuzz_matches_oracle(native, oracle):
    rng = np.random.default_rng(1234)
    n_out = n_fixed = 0
    for chn, txt, crc in _fec_cases(rng, 3000):
        a, o = api.Msg(), refs.Msg()
        for m in (a, o):
            m.chn, m.len = chn, len(txt)
            m.txt[:len(txt)] = txt
            m.crc[:] = crc
        fa, fo = api.block_fec(a), oracle.fec(o)
        assert (fa is None) == (fo is None)
        if fa is not None:
            assert fa.as_tuple() == fo.as_tuple()
            n_out += 1
            n_fixed += fa.err > 0
    assert n_out > 500 and n_fixed > 100


def test_frame_state_machine_host_build(oracle):
    """frame_sm.h compiled for the host (the compat shim's decodeAcars uses it) against the
    restatement's orc_decode_byte on random byte streams with embedded frames."""
    import subprocess, tempfile
    src = r'''
#include <stdio.h>
#include <string.h>
#include "frame_sm.h"
struct Acc {
  int st, nb, bc, len, err; unsigned S; double df, ls; unsigned char txt[256], crc[2]; int emitted; unsigned long h;
  int &state(){return st;} int &nbits(){return nb;} int &bitcount(){return bc;} int &blk_len(){return len;} int &blk_err(){return err;}
  unsigned &msk_s(){return S;} double &msk_df(){return df;} double &lvlsum(){return ls;}
  void txt_put(int i,unsigned char r){txt[i]=r;} unsigned char txt_get(int i){return txt[i];} void crc_put(int i,unsigned char r){crc[i]=r;}
  bool frame_begin(){return true;}
  void frame_emit(){emitted++; for(int i=0;i<len;i++) h=h*131+txt[i]; h=h*131+crc[0]; h=h*131+crc[1]; h=h*131+len;}
};
int main(){ Acc a; memset(&a,0,sizeof(a)); a.nb=8; int c; 
  while((c=getchar())!=EOF){ a.df=1.0; acb::frame_byte(a,(unsigned char)c); printf("%d %d %u %d %d %d %lu %d\n",a.st,a.nb,a.S,a.len,a.err,a.emitted,a.h,a.df==0.0); }
  return 0; }
'''
    rng = np.random.default_rng(8)
    stream = bytearray()
    for i in range(300):
        kind = i % 5
        fr = synth.frame_bytes(synth.random_text(rng, int(rng.integers(0, 224))), prekey=0, etb=bool(i & 1))
        fr = bytearray(fr[2:])                         # SYN SYN SOH ... BCS DEL
        if kind == 1:
            fr = bytearray(b ^ 0xFF for b in fr)        # inverted polarity (~SYN)
            fr[2:] = bytes(b ^ 0xFF for b in fr[2:])    # only the SYNs inverted: decoder flips MskS
        if kind == 2:
            for _ in range(6):
                fr[int(rng.integers(3, len(fr)))] ^= 1 << int(rng.integers(0, 8))
        if kind == 3 and len(fr) > 40:
            fr[-4] = 0x55                               # destroy ETX: the DLE path (acars.c:324)
        stream += fr + bytes(rng.integers(0, 256, size=int(rng.integers(0, 12)), dtype=np.uint8))
    stream += bytes([0x16, 0x16, 0x01]) + bytes([0x31] * 260)      # too long (acars.c:334)
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "f.cpp").write_text(src)
        subprocess.run(["g++", "-std=c++17", "-I", str(ROOT / "acarsdec_b200" / "csrc"), "-o", f"{d}/f", f"{d}/f.cpp"], check=True)
        out = subprocess.run([f"{d}/f"], input=bytes(stream), capture_output=True, check=True).stdout.decode().split("\n")
    c = oracle.new_chan(0)
    sink = refs.Sink()
    h, emitted = 0, 0
    for i, byte in enumerate(stream):
        c.outbits = byte
        c.MskDf = 1.0
        oracle.lib.orc_decode_byte(C.byref(c), C.byref(sink.c))
        if sink.c.nmsg > emitted:
            m = sink.c.msgs[emitted]
            for k in range(m.len):
                h = (h * 131 + m.txt[k]) % 2**64
            for v in (m.crc[0], m.crc[1], m.len):
                h = (h * 131 + v) % 2**64
            emitted += 1
        st, nb, S, ln, err, em, hh, dfz = (int(x) for x in out[i].split())
        assert (st, nb, S, em, hh, dfz) == (c.state, c.nbits, c.MskS, emitted, h, int(c.MskDf == 0.0)), i
        if st == 3:
            assert (ln, err) == (c.blk.len, c.blk.err), i
    assert emitted > 150


@pytest.mark.skipif(HAVE_GPU, reason="only meaningful without a GPU")
def test_no_cpu_fallback(native):
    with pytest.raises(api.AcbError):
        api.Context(K=160, nstreams=1, nch=1, max_blocks=1)


@pytest.mark.parametrize("rate,freqs", [
    (2500000, (131.525, 131.725, 131.825)),
    (3000000, synth.DEFAULT_FREQS_MHZ),
    (6000000, (129.125, 130.025, 131.550, 131.725)),
    (10000000, (131.525, 136.900)),
])
def test_air_planning_matches_oracle(native, oracle, rate, freqs):
    assert api.air_plan(rate, freqs) == oracle.air_plan(rate, freqs)
    assert bits_equal(api.build_wf_air(rate, freqs), oracle.air_wf(rate, freqs))


@pytest.mark.parametrize("K", [160, 192])
def test_fast_plan_factorises_the_reference_table(native, oracle, K):
    """ACB_FLAG_FAST_CHANNELIZER's planning step, on the CPU: the reference's mixer table (rtl.c:283-286,
    through the pinned restatement) must be the fast form's factorisation (-j)^(r*n1) * T[n2] up to the
    table's own float phase rounding — this is the whole mathematical content of k_channelize_dft."""
    fm = synth.DEFAULT_FREQS_MHZ
    fd, fr, fc = api.plan(K, fm)
    res = api.fast_plan(K, fd, fc)
    assert res is not None
    k, tw = res
    assert [int(x) for x in k] == [round((float(np.float32(f)) - float(np.float32(fc))) / 12500) for f in fr]
    assert all(x % 2 == 0 and x != 0 for x in k)
    wf = oracle.wf(K, fm)
    N2 = K // 4
    for c in range(len(fm)):
        ref = wf[c, 0::2].astype(np.float64) + 1j * wf[c, 1::2]
        r = int(k[c]) % 4
        fact = np.concatenate([(-1j) ** (r * n1) * tw[c].astype(np.complex128) for n1 in range(4)])
        # the reference's phases are float(AMFreq*ind): up to ~1.5e-5 rad of rounding at ind ~ K
        assert np.abs(fact - ref).max() <= 3e-5 * np.abs(ref).max()
        # and the twiddles themselves are the double-precision values rounded once
        want = np.exp(-2j * np.pi * ((int(k[c]) * np.arange(N2)) % K) / K) / K / 127.5
        assert np.abs(tw[c] - want).max() <= 1e-7 * np.abs(want).max()


def test_fast_plan_rejects_off_raster_channels(native):
    K = 160
    fd, _, fc = api.plan(K, (131.525, 131.725, 131.825))
    assert api.fast_plan(K, fd, fc) is not None
    # 131.4875 MHz is on the 12.5 kHz raster but its float image (what rtl.c:255 stores) is 4 Hz off
    fd2, _, fc2 = api.plan(K, (131.4875, 131.725))
    assert api.fast_plan(K, fd2, fc2) is None
    # a centre frequency off the raster (fc - 1 would not do: rtl.c:283 mixes with (float)Fc, and the float
    # image of fc - 1 is fc)
    assert api.fast_plan(K, fd, fc - 100) is None
    assert api.fast_plan(K, fd, fc - 1) is not None
    # a channel at the centre (bin 0) or beyond the band edge
    assert api.fast_plan(K, [fc], fc) is None
    assert api.fast_plan(K, [fc - 80 * 12500], fc) is None


def test_fast_plan_coverage_of_real_channel_sets(native):
    """Which real-world plans take the fast channelizer: all sets of 25 kHz-raster channels below 2^27 Hz
    (the 129-132 MHz ACARS channels); in the 136 MHz band (float ulp 16 Hz) the float image of the channel
    or of Fc is 8 Hz off, the reference then mixes off-raster and the plan is refused (exact kernel)."""
    import random
    low = [129.125, 129.35, 129.525, 130.025, 130.425, 130.45, 130.825, 130.85, 131.125, 131.25, 131.45, 131.475,
           131.525, 131.55, 131.725, 131.825, 131.85, 131.95]
    rnd = random.Random(3)
    for K in (160, 192):
        span = (K * 12500 - 4 * 12500) / 1e6
        for n in (1, 2, 3, 5, 8):
            for _ in range(40):
                base = rnd.choice(low)
                cand = [f for f in low if abs(f - base) <= span / 2]
                sel = rnd.sample(cand, min(n, len(cand)))
                fd, _, fc = api.plan(K, sel)
                assert fc != 0 and api.fast_plan(K, fd, fc) is not None, sel
    # 136.975 MHz = 136975000 Hz is not a multiple of 16: its float image is 136975008
    fd, fr, fc = api.plan(160, (136.975,))
    assert float(np.float32(fr[0])) != fd[0] and api.fast_plan(160, fd, fc) is None
    fd, fr, fc = api.plan(160, (136.8,))              # 136800000 is a multiple of 16, its Fc = 136825000 is not
    assert float(np.float32(fr[0])) == fd[0] and float(np.float32(fc)) != fc and api.fast_plan(160, fd, fc) is None


def test_band_planner_covers_wide_sets(native, oracle):
    """acb_plan_bands: what chooseFc refuses (span > rate - 4*INTRATE, rtl.c:149-152) is cut into the fewest bands,
    each centred by chooseFc itself — the whole 130-137 MHz raster scan.sh walks in 280 five-minute steps fits
    eight 2 MS/s receivers at once (on a dense 25 kHz raster chooseFc's "nobody within 25 kHz of DC" rule pushes the
    centre outside the band, so one receiver takes 39 channels = 950 kHz, not the full 1.95 MHz)."""
    K = 160
    span = 12500 * K - 4 * 12500
    scan = [130_000_000 + 25_000 * i for i in range(280)]                 # scan.sh:3-14: 130.000 .. 136.975
    rng = np.random.default_rng(0)
    for freqs in (scan, list(rng.permutation(scan)), [131_525_000, 131_725_000], [118_000_000, 137_000_000, 127_500_000],
                  [129_125_000, 131_550_000, 136_900_000, 136_925_000, 131_125_000]):
        grp, fc = api.plan_bands(K, freqs)
        f = np.asarray(freqs)
        for g in range(len(fc)):
            mine = f[grp == g]
            assert mine.max() - mine.min() <= span
            assert int(fc[g]) == native.acb_choose_fc(np.sort(mine).astype(np.uint32).ctypes.data, len(mine), K) != 0
            if len(mine) <= 64:                                                       # the restatement's table size
                assert int(fc[g]) == oracle.lib.orc_choose_fc(np.sort(mine).astype(np.uint32).ctypes.data, len(mine), K)
            d = np.abs(mine.astype(np.int64) - int(fc[g]))
            assert np.all(d >= 2 * 12500) and np.all(d <= 12500 * K // 2 - 2 * 12500)   # nobody on DC, everybody in band
        # fewest bands: the first member of every band could not have joined the previous one
        srt = np.sort(f)
        lo = [f[grp == g].min() for g in range(len(fc))]
        hi = [f[grp == g].max() for g in range(len(fc))]
        for g in range(1, len(fc)):
            nxt = srt[srt > hi[g - 1]].min()
            assert nxt == lo[g]
            if nxt - lo[g - 1] <= span:              # it fitted by span: then chooseFc must have had no valid centre for it
                m = np.sort(np.append(f[grp == g - 1], nxt)).astype(np.uint32)
                c = native.acb_choose_fc(m.ctypes.data, len(m), K)
                dd = np.abs(m.astype(np.int64) - int(c))
                mirror = any(int(c) - int(m[i - 1]) == int(m[i]) - int(c) for i in range(1, len(m)))
                assert c == 0 or dd.min() < 2 * 12500 or dd.max() > 12500 * K // 2 - 2 * 12500 or mirror
    grp, fc = api.plan_bands(K, scan)
    assert len(fc) == 8 and [int((grp == g).sum()) for g in range(8)] == [39] * 7 + [7]
    grp1, fc1 = api.plan_bands(K, [131_525_000, 131_725_000, 131_825_000])
    assert len(fc1) == 1 and int(fc1[0]) == api.plan(K, (131.525, 131.725, 131.825))[2]      # one band = chooseFc
    with pytest.raises(api.AcbError):
        api.plan_bands(K, scan, max_groups=7)
