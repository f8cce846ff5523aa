"""Regenerates the golden fixtures in this directory.  Run in the build container, where
/root/reference exists:   python tests/golden/make_golden.py

testwav_pcm16.npz     the reference's only fixture (test.wav: 4 ch, PCM16, 12.5 kHz, 53 843 frames)
                      re-encoded as a compressed int16 array so that it can travel to the GPU box
                      (the reference tree does not).
testwav_expected.json what the UNMODIFIED reference (oracle/_ref, -O2 -ffp-contract=off) decodes from
                      it through the soundfile.c:58-81 loop: the 7 messages of SURVEY.md §4 with the
                      bit pattern of lvl, the final channel_t state of each channel, putbit count.
synth_k16.npz/.json   a small seeded synthetic u8 IQ capture (K=16, 3 channels, 4 blocks) with the
                      reference's dm checksums, messages and final states for the RTL path.
"""
import hashlib
import json
import struct
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import refs  # noqa: E402
from acarsdec_b200 import synth  # noqa: E402


def read_wav(path):
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE"
    pos, fmt, data = 12, None, None
    while pos < len(raw):
        cid, sz = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        if cid == b"fmt ":
            fmt = raw[pos + 8:pos + 8 + sz]
        if cid == b"data":
            data = raw[pos + 8:pos + 8 + sz]
            break
        pos += 8 + sz + (sz & 1)
    nch, rate, bits = struct.unpack("<H", fmt[2:4])[0], struct.unpack("<I", fmt[4:8])[0], struct.unpack("<H", fmt[14:16])[0]
    assert bits == 16
    return np.frombuffer(data, dtype="<i2").reshape(-1, nch).copy(), rate


def msg_json(m):
    return {"chn": m.chn, "len": m.len, "err": m.err, "lvl_bits": int(np.float32(m.lvl).view(np.uint32)),
            "lvl": float(m.lvl), "crc": bytes(m.crc).hex(), "txt": bytes(m.txt[:m.len]).hex()}


def state_json(s):
    v = s.vec()
    return {"MskPhi": v[0].hex(), "MskDf": v[1].hex(), "MskLvlSum": v[2].hex(), "MskClk": float(v[3]).hex(),
            "MskBitCount": v[4], "MskS": v[5], "idx": v[6], "nbits": v[7], "state": v[8], "outbits": v[9],
            "inb": [float(x).hex() for x in v[10]]}


def main():
    refs.ensure_built()
    pcm, rate = read_wav("/root/reference/test.wav")
    assert rate == 12500 and pcm.shape == (53843, 4)
    np.savez_compressed(HERE / "testwav_pcm16.npz", pcm=pcm)
    x = pcm.astype(np.float32) / np.float32(32768.0)     # libsndfile's sf_read_float scaling
    ref = refs.RefLib("O2")
    ref.open_audio(4)
    for s in range(0, len(x), 4096):                      # soundfile.c:65-77 chunking
        for ch in range(4):
            ref.audio(ch, x[s:s + 4096, ch])
    msgs = ref.msgs()
    exp = {"source": "test.wav via oracle/_ref (-O2 -ffp-contract=off), soundfile.c loop, 4096-frame chunks",
           "messages": [msg_json(m) for m in msgs],
           "final_state": [state_json(ref.state(ch)) for ch in range(4)]}
    ref.close()
    # putbit() call count over the file, from the restatement (pinned against the reference
    # state-for-state by tests/test_oracle_vs_reference.py)
    orc = refs.OracleLib()
    total = 0
    for ch in range(4):
        c = orc.new_chan(ch)
        orc.demod(c, x[:, ch])
        total += c.nbit_total
    exp["putbit_calls"] = int(total)
    json.dump(exp, open(HERE / "testwav_expected.json", "w"), indent=1)
    print("test.wav:", len(msgs), "messages,", total, "bits")

    # ---- small synthetic RTL-path capture ----
    K = 16
    fm = (131.525, 131.550, 131.475)
    ref.open_rtl(K, fm)
    fc = ref.fc
    plan = synth.StreamPlan(K=K, freqs_hz=tuple(ref.chan_freq(i) for i in range(3)), fc_hz=fc, seed=11, noise_sigma=1.0)
    plan.bursts.append(synth.Burst(chan=0, t0=0.012, frame=synth.frame_bytes(b"GOLDEN VECTOR 0", addr=b".N00001"), amp=25.0, phase=0.3))
    plan.bursts.append(synth.Burst(chan=2, t0=0.030, frame=synth.frame_bytes(b"CH2/GOLDEN", addr=b".N00002", label=b"Q0"), amp=18.0, phase=1.1))
    plan.bursts.append(synth.Burst(chan=1, t0=0.005, frame=synth.frame_bytes(b"B", addr=b".N00003", label=b"_d"), amp=30.0, phase=2.0))
    nblk = 4
    iq = synth.render_blocks(plan, 0, nblk)
    np.savez_compressed(HERE / "synth_k16.npz", iq=iq)
    dm_sha = hashlib.sha256()
    for b in range(nblk):
        ref.block(iq[b])
        for ch in range(3):
            dm_sha.update(ref.dm(ch).tobytes())
    msgs = ref.msgs()
    exp = {"K": K, "freqs_mhz": list(fm), "fc": fc, "nblk": nblk, "iq_sha256": hashlib.sha256(iq.tobytes()).hexdigest(),
           "dm_sha256": dm_sha.hexdigest(), "wf_sha256": hashlib.sha256(b"".join(ref.wf(i).tobytes() for i in range(3))).hexdigest(),
           "messages": [msg_json(m) for m in msgs], "final_state": [state_json(ref.state(ch)) for ch in range(3)]}
    ref.close()
    json.dump(exp, open(HERE / "synth_k16.json", "w"), indent=1)
    print("synth_k16:", len(msgs), "messages")


if __name__ == "__main__":
    main()
