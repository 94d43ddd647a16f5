#!/usr/bin/env python
"""Kernel-only time of the sweep on the first N segments of the config-2 array for a list of planner settings:
each argument after N is a comma-free spec  lat:C:T:warps  (lat = 0 classic / 1 deferred-jump / a = planner's choice;
C:T:warps optional).  Used to calibrate plan.cpp's choice between the classic and the deferred-jump sweep."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from stringdecomposer_b200 import synth, Decomposer
from stringdecomposer_b200.hostpipe import segment_reads

nseg = int(sys.argv[1])
rn, reads, mn, mons = synth.config2()
segs, _ = segment_reads(reads, 5000, 500)
segs = segs[:nseg]
blob = "".join(segs).encode(); off = np.zeros(len(segs) + 1, dtype=np.int64); np.cumsum([len(s) for s in segs], out=off[1:])
cells = sum(len(s) for s in segs) * 2 * sum(len(m) for m in mons)
base = None
for spec in sys.argv[2:] or ["a"]:
    f = spec.split(":")
    env = {}
    if f[0] != "a":
        env["SD_LAT"] = f[0]
    if len(f) >= 3:
        env["SD_GEOM"] = "%s,%s,%s" % (f[1], f[2], f[4] if len(f) > 4 else "1")
    if len(f) >= 4 and f[3]:
        env["SD_LAT_WARPS"] = f[3]
    os.environ.update(env)
    try:
        d = Decomposer(mons, devices=[0])
        d.stage((blob, off)); d.run_staged(); d.run_staged(); d.reset_stats()
        reps = 5
        ms = min(d.run_staged() for _ in range(reps))
        st = d.stats()
        recs, roff = d.fetch_staged()
        if base is None:
            base = (recs, roff)
        same = bool(len(recs) == len(base[0]) and (recs == base[0]).all() and (roff == base[1]).all())
        print("segments=%d %-16s -> lat=%d C=%d T=%d NS=%d NT=%d NG=%d W=%d  kernels %.3f ms  sweep %.3f tb %.3f  %.0f GCUPS  same=%s" % (
            len(segs), spec, st["lat"], st["C"], st["T"], st["NS"], st["NT"], st["NG"], st["scanw"], ms, st["sweep_ms"] / reps,
            st["traceback_ms"] / reps, cells / ms / 1e6, same), flush=True)
        d.close()
    except Exception as e:
        print("segments=%d %-16s -> FAILED: %s" % (len(segs), spec, str(e)[:200]), flush=True)
    for k in env:
        del os.environ[k]
