"""bench.py's host logic on CPU: the step sizing rule and the frame checker that holds the timed region to the
oracle (it must accept the oracle's own frames replicated over tiled streams and reject a perturbed record)."""
import numpy as np

import bench
from acarsdec_b200 import api, synth


def test_blocks_per_step_rule():
    assert [bench.blocks_for(s, 16) for s in bench.SWEEP] == [16, 16, 16, 8]
    assert [bench.blocks_for(s, 16, bench.STEP_CAP // 2) for s in bench.SWEEP] == [16, 16, 8, 4]      # host cannot pin 12.4 GB
    assert bench.blocks_for(592, 8) == 8 and bench.blocks_for(100000, 16) == 1
    for s in bench.SWEEP:
        assert s * bench.blocks_for(s, 16) * 2048 * 160 <= 12.5e9         # a step's input stays <= 12.4 GB


def test_check_frames_accepts_oracle_and_rejects_a_flipped_bit(oracle):
    K, B, reps, S = 160, 10, 2, 5
    fd, _, fc = api.plan(K, synth.DEFAULT_FREQS_MHZ)
    pool = bench.make_pool(K, B, 2, fc, seed0=400)
    want = bench.oracle_frames(K, pool, reps)
    assert sum(len(w) for w in want) >= 4
    recs = []
    for s in range(S):                                  # streams 0, 2, 4 carry pool[0]; 1, 3 carry pool[1]
        for (chn, ln, err, txt, crc, lvl) in want[s % 2]:
            r = np.zeros(1, dtype=api.MSG_DTYPE)
            r["stream"], r["chn"], r["len"], r["err"] = s, chn, ln, err
            r["lvl"] = np.array([lvl], dtype=np.uint32).view(np.float32)
            r["txt"][0, :ln] = np.frombuffer(txt, dtype=np.uint8)
            r["crc"][0] = np.frombuffer(crc, dtype=np.uint8)
            recs.append(r)
    recs = np.concatenate(recs)
    ok = bench.check_frames(K, pool, reps, recs, True, S)
    assert ok["bit_exact"] and ok["replicas_identical_counts"] and ok["frames"] == sum(len(w) for w in want)
    bad = recs.copy()
    i = int(np.flatnonzero(bad["stream"] == 1)[0])
    bad["txt"][i, 2] ^= 1
    assert not bench.check_frames(K, pool, reps, bad, True, S)["bit_exact"]
    fewer = recs[:-1]                                   # the last replica lost a frame
    assert not bench.check_frames(K, pool, reps, fewer, True, S)["replicas_identical_counts"]
    lvl = recs.copy()
    lvl["lvl"][0] += np.float32(0.01)                   # within the fast form's lvl tolerance, not bit-exact
    assert not bench.check_frames(K, pool, reps, lvl, True, S)["bit_exact"]
    assert bench.check_frames(K, pool, reps, lvl, False, S)["bit_exact"]
