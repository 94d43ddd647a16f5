import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stringdecomposer_b200 import synth, Decomposer
from stringdecomposer_b200.hostpipe import segment_reads
rn, reads, mn, mons = synth.config2()
segs, _ = segment_reads(reads, 5000, 500)
blob = "".join(segs).encode(); off = np.zeros(len(segs) + 1, dtype=np.int64); np.cumsum([len(s) for s in segs], out=off[1:])
d = Decomposer(mons, devices=[0])
for it in range(4):
    t0 = time.perf_counter(); r, o = d.decompose((blob, off)); print("python wall %.3f ms" % (1e3 * (time.perf_counter() - t0)), file=sys.stderr)
