"""A/B timing of k_demod builds (tooling): loads each variant library and times the demod kernel
alone (no overlap) for 1 and 592 streams; prints demod ms per 16-block submit."""
import sys, glob, os, subprocess
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) > 1 and sys.argv[1] == "one":
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    import numpy as np
    from acarsdec_b200 import api, synth
    api.LIB_PATH = Path(sys.argv[2])
    from bench import make_pool
    K, B = 160, 16
    fd, _, fc = api.plan(K, synth.DEFAULT_FREQS_MHZ)
    pool = make_pool(K, B, 2, fc)
    out = []
    for S in (1, 592):
        stride = B * 2048 * K
        host = np.empty((S, stride), dtype=np.uint8)
        for s in range(S): host[s] = pool[s % 2]
        ctx = api.Context(K, S, 8, B, flags=1)
        for s in range(S): ctx.set_plan(s, fd)
        d = ctx.device_alloc(S * stride); ctx.copy_to_device(d, host)
        for _ in range(2):
            ctx.submit_device(d, B, stride); ctx.sync()
        ctx.drain_records(); ctx.stats(reset=True)
        for _ in range(4):
            ctx.submit_device(d, B, stride); ctx.sync()      # sync each step: no K1/K2 overlap
        st = ctx.stats()
        out.append((S, st.demod_ms / st.demod_launches, st.chan_ms / st.chan_launches, len(ctx.drain_records())))
        ctx.close()
    print(Path(sys.argv[2]).name, out, flush=True)
else:
    for lib in sorted(glob.glob(str(ROOT / "acarsdec_b200/build/variants/lib_*.so"))):
        subprocess.run([sys.executable, __file__, "one", lib])
