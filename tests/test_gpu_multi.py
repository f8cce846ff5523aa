"""acb_multi_*: one process, several GPUs (or several contexts on one GPU: the merge logic is the same).
Stream-split and channel-split must give exactly what ONE context gives: same messages, same order."""
import os
import subprocess

import numpy as np
import pytest

import refs
from acarsdec_b200 import api, synth
from common import msg_tuple

pytestmark = pytest.mark.gpu


def _devices(n):
    import ctypes
    cnt = ctypes.c_int(0)
    ctypes.CDLL("libcudart.so.12").cudaGetDeviceCount(ctypes.byref(cnt))
    return [i % max(1, cnt.value) for i in range(n)]


def _key(m):
    return (m.stream,) + msg_tuple(m) + (int(m.block), int(m.pos), int(m.soh_pos))


@pytest.mark.parametrize("mode,nstreams,nch,nparts", [(0, 5, 8, 2), (0, 3, 8, 4), (1, 1, 11, 3), (1, 2, 8, 2)])
def test_multi_equals_single_context(native, oracle, mode, nstreams, nch, nparts):
    K = 160
    fm = tuple(130.000 + 0.025 * i for i in range(nch))
    fd, _, fc = api.plan(K, fm)
    secs = 0.45
    nblk = synth.blocks_for_seconds(K, secs)
    plans = [synth.make_plan(K, fm, fc, seconds=secs, seed=70 + s) for s in range(nstreams)]
    iq = np.stack([synth.render_blocks(p, 0, nblk).reshape(-1) for p in plans])
    half = nblk // 2
    bb = 2048 * K

    def run(ctx):
        for s in range(nstreams):
            ctx.set_plan(s, fd)
        ctx.submit_host(np.ascontiguousarray(iq[:, :half * bb]), half)
        ctx.submit_host(np.ascontiguousarray(iq[:, half * bb:]), nblk - half)
        ctx.sync()
        return [_key(m) for m in ctx.drain()]

    with api.Context(K, nstreams, nch, nblk) as one:
        want = run(one)
        st_one = one.get_state(nstreams - 1, nch - 1).vec()
    with api.MultiContext(K, nstreams, nch, nblk, _devices(nparts), mode=mode) as multi:
        assert multi.parts() == min(nparts, nstreams if mode == 0 else nch)
        got = run(multi)
        st_multi = multi.get_state(nstreams - 1, nch - 1).vec()
    assert got == want and len(want) >= 4
    assert st_multi == st_one
    # and the single context is the oracle's (emission order included)
    o = refs.OracleStream(oracle, K, oracle.wf(K, fm))
    o.blocks(iq[0])
    assert [k[1:7] for k in want if k[0] == 0] == [msg_tuple(m) for m in o.msgs()]


def test_unmodified_host_over_several_devices(tmp_path):
    """The reference's own main() linked to the shim, ACARSDEC_B200_DEVICES naming several contexts (two real
    GPUs when the box has them): byte-identical stdout to the reference program."""
    from test_compat import REFBIN, _strip_time
    if not (REFBIN / "acarsdec_b200").exists() or not (REFBIN / "acarsdec_ref").exists():
        pytest.skip("oracle/_ref programs absent")
    orc = refs.OracleLib()
    K, fm = 160, (131.525, 131.725, 131.825, 131.450, 131.550)
    _, _, fc = orc.plan(K, fm)
    plan = synth.make_plan(K, fm, fc, seconds=1.2, seed=78, text_len=(10, 80), msgs_per_chan_per_sec=3.0)
    cap = tmp_path / "cap.iq"
    synth.render_blocks(plan, 0, synth.blocks_for_seconds(K, 1.2)).tofile(cap)
    freqs = [str(f) for f in fm]
    ref = subprocess.run([str(REFBIN / "acarsdec_ref"), "-o", "2", "-m", str(K), "-r", "0", *freqs],
                         env=dict(os.environ, ACARSDEC_STUB_IQ=str(cap)), capture_output=True, text=True, timeout=120)
    devs = ",".join(str(d) for d in _devices(3))
    mine = subprocess.run([str(REFBIN / "acarsdec_b200"), "-o", "2", "-m", str(K), "-r", str(cap), *freqs],
                          env=dict(os.environ, ACARSDEC_B200_DEVICES=devs, ACARSDEC_B200_BLOCKS="4"), capture_output=True, text=True, timeout=120)
    assert mine.returncode == 0, mine.stderr
    a, b = _strip_time(ref.stdout), _strip_time(mine.stdout)
    assert a.count("<time>") >= 6 and a == b
