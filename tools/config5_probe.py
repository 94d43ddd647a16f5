#!/usr/bin/env python
"""Throughput of the group sweep on a config-5-like workload: 1000 monomers, part 20000 / overlap 500."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from stringdecomposer_b200 import synth, Decomposer
from stringdecomposer_b200.hostpipe import segment_reads
nm = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
total = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
rn, reads, mn, mons = synth.config5(n_monomers=nm, total=total)
segs, _ = segment_reads(reads, 20000, 500)
cells = sum(len(s) for s in segs) * 2 * sum(len(m) for m in mons)
print("monomers", len(mons), "segments", len(segs), "cells %.3e" % cells, flush=True)
d = Decomposer(mons, devices=[0])
t0 = time.perf_counter(); recs, off = d.decompose(segs); dt = time.perf_counter() - t0
st = d.stats()
print("geometry", {k: st[k] for k in ("C", "T", "NS", "NT", "packed")}, "sweep %.1f ms traceback %.1f ms wall %.1f ms -> %.0f GCUPS kernels" % (
    st["sweep_ms"], st["traceback_ms"], dt * 1e3, cells / ((st["sweep_ms"] + st["traceback_ms"]) * 1e6)), "records", len(recs), flush=True)
