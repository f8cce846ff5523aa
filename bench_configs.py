"""bench.py's `configs` object: the literal BASELINE.json configs 2..5, short runs, each with the kernel that
limits it named and (rank 0) its first pass checked against the CPU port of the reference (oracle/).

  2  synthetic 2 MS/s uint8 IQ, 8 ACARS channels, ONE stream (what a single receiver's backlog decodes at)
  3  rateMult=192 (2.4 MS/s), 64 channels across 1.6 MHz, FIR taps=165 — one wide stream; at N>1 the channels are
     split over the ranks and the raw block is broadcast once per step (NCCL), inside the timed region
  4  1024 channels = 128 streams x 8 ch TOTAL, sharded by stream index over the N GPUs (strong scaling)
  5  wideband 20 MS/s IQ (K=1600), 256 channels, FIR-tap sweep 65..513 — one wide stream, channel-split like 3

One stream is a serial recurrence per channel in the demodulator (msk.c:67-137): configs 2, 3 and 5 are bound by
the demod kernel's latency per bit whatever N is, config 4 (16 streams per GPU at N=8) likewise; they are
reported as measured, next to the saturated headline.  Only bench.py imports this (it uses oracle/ as the checker).
"""
from __future__ import annotations

import time

import numpy as np

from acarsdec_b200 import api, sharding, synth


def _reduce(dist, x, op, dev):
    return sharding.reduce_scalar(dist, x, op, dev)


def _msgs_key(m):
    return (m[2],) + tuple(m[4:])          # (chn, len, err, txt, crc, lvl_bits) of wide.WideStream tuples


class _LocalWide:
    """world == 1 stand-in for wide.WideStream without torch: host buffer -> acb_submit_host."""

    def __init__(self, local, K, nch, max_blocks, taps, wf_all, flags):
        self.ctx = api.Context(K, 1, nch, max_blocks, device=local, flags=flags, taps=taps)
        self.ctx.set_wf(0, wf_all)
        self.bufs = [api.PinnedBuffer(max_blocks * 2048 * K) for _ in range(2)]
        self.holds = [None, None]
        self.n = 0

    def submit(self, iq, nblk):
        i = self.n & 1                             # two submits in flight at most: the buffer of two back is free
        v = self.bufs[i].array[:iq.size]
        if self.holds[i] is not iq:                # the capture already sits in this pinned buffer (a receiver's DMA target)
            v[:] = iq.reshape(-1)
            self.holds[i] = iq
        self.ctx.submit_host(v.reshape(1, -1), nblk)
        self.n += 1

    def sync(self):
        self.ctx.sync()

    def ingest_ms(self):
        return None

    def gather(self):
        return [(int(m.block), 0, m.chn, int(m.pos), m.len, m.err, bytes(m.txt[:m.len]), bytes(m.crc),
                 int(np.float32(m.lvl).view(np.uint32))) for m in self.ctx.drain()]

    def close(self):
        self.ctx.close()
        for b in self.bufs:
            b.close()


def _oracle_fir_frames(K, taps, wf, iq, chans):
    """channelize_fir -> orc_demod -> block FEC for the listed channels (the reference has no taps < K mode; the
    definition is oracle/acars_oracle.c: orc_channelize_fir, equal to the pinned orc_channelize at taps = K)."""
    import refs
    refs.ensure_built()
    orc = refs.OracleLib()
    dm = orc.channelize_fir(iq, K, taps, wf[chans])
    out = []
    for i, c in enumerate(chans):
        sink = refs.Sink()
        ch = orc.new_chan(int(c))
        orc.demod(ch, dm[i], sink)
        for m in sink.msgs():
            f = orc.fec(m)
            if f is not None:
                out.append((f.chn, f.len, f.err, bytes(f.txt[:f.len]), bytes(f.crc), int(np.float32(f.lvl).view(np.uint32))))
    return out


def _wide(dist, rank, world, local, dev, K, fm_mhz, taps, B, plan_kw, steps, do_check, bursts=None):
    """One wide stream: `steps` timed submits of the same B blocks (the ingest rank's H2D + the broadcast inside the
    timed region), frames of the first pass checked on rank 0."""
    fd, _, fc = api.plan(K, fm_mhz)
    rate = K * 12500
    wf = synth.fir_tables(K, taps, [f - fc for f in fd], rate)
    iq = None
    if rank == 0:
        if bursts is None:
            plan = synth.make_plan(K, fm_mhz, fc, seconds=B * 1024 / 12500, **plan_kw)
        else:
            plan = synth.StreamPlan(K=K, freqs_hz=tuple(fd), fc_hz=fc, seed=plan_kw.get("seed", 7), noise_sigma=1.5)
            plan.bursts = bursts(fd)
        iq = synth.render_blocks(plan, 0, B).reshape(-1)
    if world == 1:
        ws = _LocalWide(local, K, len(fd), B, taps, wf, 0)
    else:
        import torch
        from acarsdec_b200 import wide
        ws = wide.WideStream(dist, rank, world, local, K, fd, fc, B, taps=taps, wf_all=wf)
        if rank == 0:
            iq_host = iq
            iq = torch.from_numpy(iq).pin_memory()       # the ingest rank's source is pinned host memory, like a receiver's DMA target
    ws.submit(iq, B)
    ws.sync()
    first = ws.gather()
    check = None
    if do_check and rank == 0:
        chans = sorted({m[2] for m in first} | ({b.chan for b in plan.bursts} if rank == 0 else set()) | {0, len(fd) - 1})
        want = _oracle_fir_frames(K, taps, wf, iq if world == 1 else iq_host, np.array(chans))
        got = [_msgs_key(m) for m in first]
        check = {"frames": len(want), "bit_exact": sorted(got) == sorted(want), "channels_checked": len(chans)}
    for _ in range(2):
        ws.submit(iq, B)
    ws.sync()
    ws.gather()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        ws.submit(iq, B)
    ws.sync()
    t1 = time.perf_counter()
    ms = _reduce(dist, (t1 - t0) * 1e3, "max", dev) / steps
    nfr = len(ws.gather())
    st = ws.ctx.stats() if ws.ctx is not None else None
    k1 = _reduce(dist, (st.chan_ms / max(1, st.chan_launches)) if st else 0.0, "max", dev)
    k2 = _reduce(dist, (st.demod_ms / max(1, st.demod_launches)) if st else 0.0, "max", dev)
    ing = ws.ingest_ms()
    ws.close()
    samples = B * 1024 * K
    return {"K": K, "channels": len(fd), "taps": taps, "blocks_per_step": B, "ms_per_step": ms, "value": samples / ms / 1e3, "unit": "Msamples/s",
            "real_time_factor": samples / ms / 1e3 / (rate / 1e6),
            "k_channelize_ms": k1, "k_demod_and_fec_ms": k2, "limited_by": "k_demod (latency per bit of one channel's serial recurrence)" if k2 >= k1 else "k_channelize",
            "ingest_ms_per_step": ing, "ingest_bytes_per_step": B * 2048 * K,
            "ingest": ("rank 0 H2D + one NCCL broadcast per step, inside the timed region" if world > 1 else "H2D from pinned host memory, inside the timed region"),
            "channels_per_gpu": -(-len(fd) // world), "frames_timed": nfr, "checked": check,
            "timing": "host wall clock between full syncs (ingest included), max over ranks"}


def run(want, dist, rank, world, local, dev, fastflag, do_check):
    out = {}
    K = 160
    fm = synth.DEFAULT_FREQS_MHZ
    fd, _, fc = api.plan(K, fm)
    B = 16
    stride = B * 2048 * K

    def streams_run(nstreams_total, label):
        """device-resident, streams sharded by index (config 4) or one stream on rank 0 (config 2)"""
        mine = sharding.stream_range(nstreams_total, world, rank)
        ms_local, k1, k2, check = 0.0, 0.0, 0.0, None
        if len(mine):
            plans = [synth.make_plan(K, fm, fc, seconds=B * 1024 / 12500, seed=3000 + (s % 4)) for s in range(mine.start, min(mine.stop, mine.start + 4))]
            pool = [synth.render_blocks(p, 0, B).reshape(-1) for p in plans]
            with api.Context(K, len(mine), 8, B, device=local, flags=1 | fastflag) as c:
                for s in range(len(mine)):
                    c.set_plan(s, fd)
                d = c.device_alloc(len(mine) * stride)
                for s in range(len(mine)):
                    c.copy_to_device(d + s * stride, pool[s % len(pool)])
                c.submit_device(d, B, stride)
                c.sync()
                first = c.drain_records()
                if do_check and rank == 0:
                    import bench
                    w = bench.oracle_frames(K, pool[:1], 1)[0]
                    g = [bench.rec_tuple(m) for m in first[first["stream"] == 0]]
                    ok = g == w if fastflag == 0 else [t[:-1] for t in g] == [t[:-1] for t in w]
                    check = {"frames": len(w), "bit_exact": bool(ok)}
                for _ in range(2):
                    c.submit_device(d, B, stride)
                c.sync(); c.drain_records(); c.stats(reset=True)
                if dist is not None:
                    dist.barrier()
                c.mark(0)
                nst = 6
                for _ in range(nst):
                    c.submit_device(d, B, stride)
                    c.drain_records()
                c.mark(1)
                c.sync(); c.drain_records()
                ms_local = c.elapsed_ms() / nst
                st = c.stats()
                k1, k2 = st.chan_ms / max(1, st.chan_launches), st.demod_ms / max(1, st.demod_launches)
                c.device_free(d)
        elif dist is not None:
            dist.barrier()
        ms = _reduce(dist, ms_local, "max", dev)
        k1, k2 = _reduce(dist, k1, "max", dev), _reduce(dist, k2, "max", dev)
        samples = nstreams_total * B * 1024 * K
        return {"workload": label, "streams_total": nstreams_total, "streams_per_gpu": -(-nstreams_total // world), "blocks_per_step": B,
                "ms_per_step": ms, "value": samples / ms / 1e3, "unit": "Msamples/s", "k_channelize_ms": k1, "k_demod_and_fec_ms": k2,
                "limited_by": "k_demod (latency per bit of one channel's serial recurrence)" if k2 >= k1 else "k_channelize",
                "checked": check, "timing": "CUDA events on the library's streams, max over ranks"}

    if "2" in want:
        out["2"] = streams_run(1, "configs[1] as written: ONE 2 MS/s stream, 8 channels (rank 0 only at N>1)")
    if "4" in want:
        out["4"] = streams_run(128, "configs[3]: 1024 channels = 128 streams x 8 ch TOTAL, sharded by stream index (strong scaling)")
        out["4"]["scaling"] = "strong"
    if "3" in want:
        fm3 = tuple(130.000 + 0.025 * i for i in range(64))
        out["3"] = _wide(dist, rank, world, local, dev, 192, fm3, 165, 16,
                         dict(seed=33, msgs_per_chan_per_sec=1.0, text_len=(5, 25), amp=(6.0, 10.0)), 6, do_check)
    if "5" in want:
        fm5 = tuple(126.000 + 0.025 * i for i in range(256))
        rng = np.random.default_rng(55)

        def bursts(fd5):
            bl = []
            for i, ch in enumerate((3, 40, 97, 128, 200, 251)):
                fr = synth.frame_bytes(synth.random_text(rng, 8 + i))
                bl.append(synth.Burst(chan=ch, t0=0.01 + 0.012 * i, frame=fr, amp=9.0, phase=0.7 * i))
            return bl

        rows = []
        for taps in (65, 129, 257, 513):
            rng = np.random.default_rng(55)
            rows.append(_wide(dist, rank, world, local, dev, 1600, fm5, taps, 4, dict(seed=55), 5,
                              do_check and taps in (65, 513), bursts=bursts))
        out["5"] = {"workload": "configs[4]: wideband 20 MS/s uint8 IQ (K=1600), 256 channels on a 25 kHz raster, one stream, FIR-tap sweep",
                    "sweep": rows}
    return out
