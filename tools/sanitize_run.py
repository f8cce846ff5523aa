"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): u8 path with 11 channels
(two groups), real-input path with an awkward length, envelope path; checks frames against the oracle."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from acarsdec_b200 import api, synth
import refs
from common import msg_tuple, load_testwav

orc = refs.OracleLib()
K, fm, nblk = 160, synth.DEFAULT_FREQS_MHZ, 3
fd, _, fc = api.plan(K, fm)
plan = synth.make_plan(K, fm, fc, seconds=0.24, seed=5, text_len=(5, 20))
iq = synth.render_blocks(plan, 0, nblk).reshape(1, -1)
with api.Context(K, 2, 8, nblk) as ctx:
    for s in range(2): ctx.set_plan(s, fd)
    ctx.submit_host(np.concatenate([iq, iq]), nblk); ctx.submit_host(np.concatenate([iq, iq]), nblk); ctx.sync()
    n1 = len(ctx.drain())
rate, fma = 2500000, (131.525, 131.725, 131.825)
fda, fca, Ka = api.air_plan(rate, fma)
x = (np.random.default_rng(1).standard_normal((1, 1024 * Ka + 777)) * 0.01).astype(np.float32)
with api.Context(Ka, 1, 3, 3, flags=2) as ctx:
    ctx.set_plan_air(0, fda); ctx.submit_real(x); ctx.submit_real(x[:, :5000]); ctx.sync()
xw, exp = load_testwav()
with api.Context(160, 1, 4, 4, flags=1) as ctx:
    ctx.submit_dm(xw[None, :4096]); ctx.submit_dm(xw[None, 4096:8000]); ctx.sync()
    n3 = len(ctx.drain())
# fast channelizer (bulk copies + mbarrier), K=160 and K=192, partial channel group
for Kf, fmf in ((160, fm), (192, (131.525, 131.725, 131.825))):
    fdf, _, fcf = api.plan(Kf, fmf)
    pf = synth.make_plan(Kf, fmf, fcf, seconds=0.24, seed=6, text_len=(5, 20))
    iqf = synth.render_blocks(pf, 0, 2).reshape(1, -1)
    with api.Context(Kf, 2, len(fmf), 2, flags=8) as ctx:
        for s in range(2): ctx.set_plan(s, fdf)
        ctx.submit_host(np.concatenate([iqf, iqf]), 2); ctx.submit_host(np.concatenate([iqf, iqf]), 2); ctx.sync()
        assert ctx.stats().fast_chan_launches == 2
        n4 = len(ctx.drain())
# CS16 front-end with a ragged length, and the batched device FEC
iqc = synth.render_cs16(plan, 0, 1024 * K + 333)[None]
with api.Context(K, 1, 8, 2, flags=4) as ctx:
    ctx.set_plan_cs16(0, fd, 0); ctx.submit_cs16(iqc); ctx.submit_cs16(iqc[:, :4000]); ctx.sync()
    m = api.Msg(); fr = synth.frame_bytes(b"SANITIZER", prekey=0)[5:-1]
    m.len = len(fr) - 2; m.txt[:m.len] = fr[:-2]; m.crc[:] = fr[-2:]
    m.txt[3] ^= 0x04
    out = ctx.block_fec_batch([m] * 70)
    assert all(o is not None and o.err == 1 for o in out)
# the fast forms of the streaming front-ends (partial blocks, tiles past the end, the 32 rows of slack), four submits so that
# every one of the three pipeline slots is reused
with api.Context(K, 1, 8, 2, flags=4 | 8) as ctx:
    ctx.set_plan_cs16(0, fd, 1)
    for part in (iqc, iqc[:, :4000], iqc[:, :33 * K + 5], iqc):
        ctx.submit_cs16(part)
    ctx.sync()
    assert ctx.stats().fast_chan_launches >= 3
    n5 = len(ctx.drain())
xr = synth.render_real(synth.StreamPlan(K=Ka, freqs_hz=tuple(fda), fc_hz=fca, seed=2, noise_sigma=1.0), 0, 1024 * Ka + 777)[None]
with api.Context(Ka, 1, 3, 3, flags=2 | 8) as ctx:
    ctx.set_plan_air(0, fda)
    for part in (xr, xr[:, :5000], xr[:, :17 * Ka + 3], xr):
        ctx.submit_real(part)
    ctx.sync()
    assert ctx.stats().fast_chan_launches >= 3
print("sanitize run ok", n1, n3, n4, n5)
