"""Synthetic RTL-SDR style input for the channelize -> demodMSK -> frame-sync path.

Host-side tooling for tests and bench.py (SURVEY.md §8(d) recipe): builds ACARS frames,
MSK-modulates them on the 1200/2400 Hz audio sub-carrier, AM-modulates each onto its
channel offset from the tuner centre, adds noise and quantises to interleaved uint8 I/Q
exactly the way the reference expects to receive it (rtl.c:334-342).

Everything is float64 numpy so the uint8 output is reproducible across hosts (a 1-ulp
libm difference would have to land within 1e-14 of a quantisation boundary to matter).
Expected decoder output always comes from the oracle run on these bytes, never from here.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

INTRATE = 12500          # acarsdec.h:31
OUTBLK = 1024            # RTLOUTBUFSZ, rtl.c:49
BAUD = 2400

SYN, SOH, STX, ETX, ETB, DEL, NAK = 0x16, 0x01, 0x02, 0x03, 0x17, 0x7F, 0x15

# the usual European/US VHF ACARS set; all on the 12.5 kHz raster (rtl.c:245-247)
DEFAULT_FREQS_MHZ = (131.125, 131.450, 131.475, 131.525, 131.550, 131.725, 131.825, 131.850)


def odd_parity(b: int) -> int:
    """7-bit character -> byte with bit 7 set so that the number of ones is odd (acars.c:138)."""
    b &= 0x7F
    return b | (0x80 if bin(b).count("1") % 2 == 0 else 0)


def crc16_kermit(data: bytes) -> int:
    """Reflected CCITT CRC, poly 0x8408, init 0 (syndrom.h:15-49)."""
    crc = 0
    for c in data:
        crc ^= c
        for _ in range(8):
            crc = (crc >> 1) ^ 0x8408 if crc & 1 else crc >> 1
    return crc


def frame_bytes(text: bytes, *, mode=b"2", addr=b".N123AB", ack=NAK, label=b"H1", bid=b"3",
                prekey: int = 16, etb: bool = False) -> bytes:
    """One ACARS downlink frame as transmitted, parity and BCS included."""
    assert len(addr) == 7 and len(label) == 2 and len(mode) == 1 and len(bid) == 1
    body = mode + addr + bytes([ack]) + label + bid + bytes([STX]) + text + bytes([ETB if etb else ETX])
    assert len(body) <= 238, "ACARS text too long"
    body = bytes(odd_parity(c) for c in body)
    crc = crc16_kermit(body)
    head = bytes([0xFF] * prekey) + bytes([odd_parity(ord("+")), odd_parity(ord("*")), SYN, SYN, SOH])
    return head + body + bytes([crc & 0xFF, crc >> 8, DEL])


def corrupt_frame(frame: bytes, flips, prekey: int = 16) -> bytes:
    """Flip bits after the BCS was computed, to exercise the block FEC (acars.c:39-215).
    `flips` is a list of (index, xor_mask); index 0 is the first byte after SOH (the mode
    character), negative indices count from the end of the frame (-1 = DEL, -2/-3 = BCS)."""
    f = bytearray(frame)
    body0 = prekey + 5
    for idx, mask in flips:
        f[body0 + idx if idx >= 0 else len(f) + idx] ^= mask
    return bytes(f)


def frame_bits(frame: bytes) -> np.ndarray:
    """LSB-first bit stream (msk.c:55-58)."""
    return np.unpackbits(np.frombuffer(frame, dtype=np.uint8), bitorder="little")


def msk_audio_phase(bits: np.ndarray):
    """Per-bit start phase and tone for the ACARS MSK sub-carrier: 2400 Hz when the bit equals
    the previous one, 1200 Hz when it differs (initial 'previous' = 1); phase-continuous."""
    prev = np.concatenate(([1], bits[:-1]))
    f = np.where(bits == prev, 2400.0, 1200.0)
    dtheta = 2 * np.pi * f / BAUD
    theta0 = np.concatenate(([0.0], np.cumsum(dtheta)[:-1]))
    return theta0, f


@dataclass
class Burst:
    chan: int
    t0: float            # seconds from stream start (carrier key-on is 2 ms earlier)
    frame: bytes
    amp: float = 20.0
    depth: float = 0.8
    phase: float = 0.0


@dataclass
class StreamPlan:
    K: int                                   # rtlMult: input rate = K * 12500 (rtl.c:213-214)
    freqs_hz: tuple                          # per channel, CLI order
    fc_hz: int                               # tuner centre (chooseFc, rtl.c:131-168)
    bursts: list = field(default_factory=list)
    noise_sigma: float = 1.5                 # LSB, per I/Q component
    seed: int = 7

    @property
    def rate(self) -> int:
        return self.K * INTRATE

    @property
    def block_bytes(self) -> int:
        return OUTBLK * self.K * 2


def burst_samples(plan: StreamPlan, b: Burst, n_lo: int, n_hi: int):
    """Complex baseband contribution of one burst on sample indices [n_lo, n_hi)."""
    fs = plan.rate
    guard = 0.002
    bits = frame_bits(b.frame)
    dur = len(bits) / BAUD
    s_on = int(np.floor((b.t0 - guard) * fs))
    s_off = int(np.ceil((b.t0 + dur + guard) * fs))
    lo, hi = max(n_lo, s_on), min(n_hi, s_off)
    if lo >= hi:
        return None
    n = np.arange(lo, hi, dtype=np.float64)
    t = n / fs - b.t0
    theta0, f = msk_audio_phase(bits)
    k = np.floor(t * BAUD).astype(np.int64)
    inside = (k >= 0) & (k < len(bits))
    kc = np.clip(k, 0, len(bits) - 1)
    audio = np.where(inside, np.cos(theta0[kc] + 2 * np.pi * f[kc] * (t - kc / BAUD)), 0.0)
    off = float(plan.freqs_hz[b.chan] - plan.fc_hz)
    carrier = np.exp(1j * (2 * np.pi * off * (n / fs) + b.phase))
    return lo, hi, b.amp * (1.0 + b.depth * audio) * carrier


def render_blocks(plan: StreamPlan, blk0: int, nblk: int) -> np.ndarray:
    """uint8 interleaved I/Q for blocks [blk0, blk0+nblk) of the stream, shape (nblk, 1024*K*2).
    Noise is drawn per block from (seed, block index) so any block range renders identically."""
    spb = OUTBLK * plan.K
    out = np.empty((nblk, spb * 2), dtype=np.uint8)
    for i in range(nblk):
        blk = blk0 + i
        n_lo, n_hi = blk * spb, (blk + 1) * spb
        x = np.zeros(spb, dtype=np.complex128)
        for b in plan.bursts:
            r = burst_samples(plan, b, n_lo, n_hi)
            if r is not None:
                lo, hi, v = r
                x[lo - n_lo:hi - n_lo] += v
        rng = np.random.default_rng([plan.seed, blk])
        g = rng.standard_normal(2 * spb) * plan.noise_sigma
        iq = np.empty(2 * spb, dtype=np.float64)
        iq[0::2] = x.real
        iq[1::2] = x.imag
        iq += g
        out[i] = np.clip(np.floor(iq + 128.0), 0, 255).astype(np.uint8)   # round(x + 127.5)
    return out


_TEXT_ALPHABET = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 /.,-#", dtype=np.uint8)


def random_text(rng, n: int) -> bytes:
    return bytes(_TEXT_ALPHABET[rng.integers(0, len(_TEXT_ALPHABET), size=n)])


def make_plan(K: int = 160, freqs_mhz=DEFAULT_FREQS_MHZ, fc_hz: int | None = None, *, seconds: float = 1.0,
              msgs_per_chan_per_sec: float = 1.0, seed: int = 7, noise_sigma: float = 1.5,
              amp=(12.0, 30.0), text_len=(20, 120)) -> StreamPlan:
    """Channel plan + random message schedule.  `fc_hz` must come from the caller's chooseFc
    (library or oracle) when parity on the frequency plan matters; the default reproduces the
    rule for the common case (highest channel + 2*INTRATE)."""
    freqs = tuple(((int(1000000 * f + INTRATE / 2)) // INTRATE) * INTRATE for f in freqs_mhz)
    if fc_hz is None:
        fc_hz = max(freqs) + 2 * INTRATE
    plan = StreamPlan(K=K, freqs_hz=freqs, fc_hz=int(fc_hz), noise_sigma=noise_sigma, seed=seed)
    rng = np.random.default_rng([seed, 0xACA5])
    for ch in range(len(freqs)):
        t = 0.01 + rng.uniform(0, 0.2)
        while True:
            n = int(rng.integers(text_len[0], text_len[1] + 1))
            txt = random_text(rng, n)
            addr = b"." + random_text(rng, 6).replace(b" ", b"-")
            fr = frame_bytes(txt, addr=addr, label=random_text(rng, 2).replace(b" ", b"_"),
                             bid=bytes([int(rng.integers(0x30, 0x3A))]))
            dur = len(fr) * 8 / BAUD
            if t + dur + 0.01 > seconds:
                break
            plan.bursts.append(Burst(chan=ch, t0=t, frame=fr, amp=float(rng.uniform(*amp)),
                                     phase=float(rng.uniform(0, 2 * np.pi))))
            t += dur + 0.02 + rng.exponential(1.0 / max(msgs_per_chan_per_sec, 1e-6))
    return plan


def fir_tables(K: int, taps: int, offsets_hz, rate: int) -> np.ndarray:
    """Generalised-FIR mixer tables for BASELINE configs 3/5 (taps < K; the reference has no such mode, SURVEY.md
    note 1): Hamming-windowed w[t] = hamming(t) * exp(-j 2 pi f t / rate) / sum / 127.5, (nch, 2*taps) float32."""
    t = np.arange(taps)
    win = 0.54 - 0.46 * np.cos(2 * np.pi * t / max(taps - 1, 1))
    out = np.empty((len(offsets_hz), 2 * taps), dtype=np.float32)
    for i, f in enumerate(offsets_hz):
        w = win * np.exp(-2j * np.pi * f * t / rate) / win.sum() / 127.5
        out[i, 0::2] = w.real
        out[i, 1::2] = w.imag
    return out


def blocks_for_seconds(K: int, seconds: float) -> int:
    return int(np.ceil(seconds * INTRATE / OUTBLK))


# ----------------------------------------------------------------------------- Airspy-style input

def render_real(plan: StreamPlan, n0: int, nsamples: int, *, scale: float = 1.0 / 256) -> np.ndarray:
    """float32 REAL samples [n0, n0+nsamples) at plan.rate for the air.c front-end: every channel
    sits at the intermediate frequency Fc - Fr + rate/4 (air.c:278), AM-modulated like the IQ
    case; noise from (seed, n0) so any range renders reproducibly when cut at the same points."""
    fs = plan.rate
    x = np.zeros(nsamples, dtype=np.float64)
    n = np.arange(n0, n0 + nsamples, dtype=np.float64)
    for b in plan.bursts:
        bits = frame_bits(b.frame)
        dur = len(bits) / BAUD
        guard = 0.002
        lo = max(n0, int(np.floor((b.t0 - guard) * fs)))
        hi = min(n0 + nsamples, int(np.ceil((b.t0 + dur + guard) * fs)))
        if lo >= hi:
            continue
        nn = n[lo - n0:hi - n0]
        t = nn / fs - b.t0
        theta0, f = msk_audio_phase(bits)
        k = np.floor(t * BAUD).astype(np.int64)
        inside = (k >= 0) & (k < len(bits))
        kc = np.clip(k, 0, len(bits) - 1)
        audio = np.where(inside, np.cos(theta0[kc] + 2 * np.pi * f[kc] * (t - kc / BAUD)), 0.0)
        f_if = float(plan.fc_hz - plan.freqs_hz[b.chan] + fs // 4)
        x[lo - n0:hi - n0] += b.amp * (1.0 + b.depth * audio) * np.cos(2 * np.pi * f_if * (nn / fs) + b.phase)
    rng = np.random.default_rng([plan.seed, 0xA1, n0])
    x += rng.standard_normal(nsamples) * plan.noise_sigma
    return (x * scale).astype(np.float32)


def render_cs16(plan: StreamPlan, n0: int, nsamples: int, *, gain: float = 64.0) -> np.ndarray:
    """int16 (nsamples, 2) I,Q for the SoapySDR / SDRplay front-ends: the same complex baseband as
    render_blocks (before u8 quantisation), scaled by `gain` LSB per unit and rounded."""
    fs = plan.rate
    x = np.zeros(nsamples, dtype=np.complex128)
    for b in plan.bursts:
        r = burst_samples(plan, b, n0, n0 + nsamples)
        if r is not None:
            lo, hi, v = r
            x[lo - n0:hi - n0] += v
    rng = np.random.default_rng([plan.seed, 0xC5, n0])
    g = rng.standard_normal(2 * nsamples) * plan.noise_sigma
    out = np.empty((nsamples, 2), dtype=np.float64)
    out[:, 0] = x.real + g[0::2]
    out[:, 1] = x.imag + g[1::2]
    return np.clip(np.floor(out * gain + 0.5), -32768, 32767).astype(np.int16)
