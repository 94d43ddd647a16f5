#!/usr/bin/env python
"""Feasibility probe: config 2's 400 segments as two concurrent launches on one GPU -- k*148 segments through the classic
sweep (k CTAs per SM) and the remainder through the cluster sweep with small CTAs that fill the gaps -- against the single
classic launch.  Wall time of both run_staged() calls issued from two host threads."""
import os, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from stringdecomposer_b200 import synth, Decomposer
from stringdecomposer_b200.hostpipe import segment_reads

rn, reads, mn, mons = synth.config2()
segs, _ = segment_reads(reads, 5000, 500)

def pack(ss):
    blob = "".join(ss).encode(); off = np.zeros(len(ss) + 1, dtype=np.int64); np.cumsum([len(s) for s in ss], out=off[1:])
    return blob, off

def make(ss, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    d = Decomposer(mons, devices=[0]); d.stage(pack(ss)); d.run_staged(); d.run_staged()
    for k, v in old.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    return d

def timed(ds, reps=20):
    best = 1e9
    for _ in range(reps):
        ths = [threading.Thread(target=d.run_staged) for d in ds[1:]]
        t0 = time.perf_counter()
        for t in ths: t.start()
        ds[0].run_staged()
        for t in ths: t.join()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3

whole = make(segs, {"SD_LAT": "0"})
print("400 classic alone: wall %.3f ms" % timed([whole]), whole.stats()["sweep_ms"] / 22)
nmain = int(sys.argv[1]) if len(sys.argv) > 1 else 296
for spec in sys.argv[2:] or ["24,8,1:1", "24,8,1:3", "12,16,1:1", "12,16,1:3", "6,32,1:1", "6,32,1:4"]:
    geom, warps = spec.split(":")
    a = make(segs[:nmain], {"SD_LAT": "0"})
    b = make(segs[nmain:], {"SD_LAT": "1", "SD_GEOM": geom, "SD_LAT_WARPS": warps})
    w = timed([a, b])
    sa, sb = a.stats(), b.stats()
    print("split %d classic + %d cluster %s warps/CTA %s (NG=%d): wall %.3f ms   [A sweep %.3f tb %.3f | B sweep %.3f tb %.3f]" % (
        nmain, len(segs) - nmain, geom, warps, sb["NG"], w, sa["sweep_ms"] / 22, sa["traceback_ms"] / 22, sb["sweep_ms"] / 22, sb["traceback_ms"] / 22))
    a.close(); b.close()
