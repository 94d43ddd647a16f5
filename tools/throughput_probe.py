#!/usr/bin/env python
"""Kernel-only throughput of the sweep on a large batch (many segments), for several launch geometries."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from stringdecomposer_b200 import synth, Decomposer
from stringdecomposer_b200.hostpipe import segment_reads
nreads = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rn, reads, mn, mons = synth.config3(n_reads=nreads, read_len=100_000)
segs, _ = segment_reads(reads, 5000, 500)
blob = "".join(segs).encode(); off = np.zeros(len(segs) + 1, dtype=np.int64); np.cumsum([len(s) for s in segs], out=off[1:])
cells = sum(len(s) for s in segs) * 2 * sum(len(m) for m in mons)
for geom in sys.argv[2:] or [""]:
    for ll in (os.environ.get("PROBE_LL", "0,1").split(",")):
        os.environ["SD_GEOM"] = geom; os.environ["SD_LOWLAT"] = ll
        d = Decomposer(mons, devices=[0])
        d.stage((blob, off)); d.run_staged(); d.reset_stats()
        ms = min(d.run_staged() for _ in range(3))
        st = d.stats()
        print("segments=%d geom=%s lowlat=%s -> C=%d T=%d NS=%d NT=%d  kernels %.3f ms  sweep %.3f tb %.3f  %.0f GCUPS (actual cells)" % (
            len(segs), geom, ll, st["C"], st["T"], st["NS"], st["NT"], ms, st["sweep_ms"] / 3, st["traceback_ms"] / 3, cells / ms / 1e6), flush=True)
        d.close()
