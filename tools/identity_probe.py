#!/usr/bin/env python
"""Throughput of the identity rescoring kernel on a config-2-shaped batch: every decomposed interval of
a 2 Mb cenX-like array against all 24 monomer rows, plain and homopolymer-collapsed (what --second-best asks for),
next to the reference's own edlib (oracle/_ref/libedlib_ref.so, one host core) on a sample of the same pairs.
    python tools/identity_probe.py [n_intervals]"""
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import stringdecomposer_b200 as sd  # noqa: E402
from stringdecomposer_b200 import synth, convert as cv  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 11700
    r = random.Random(2)
    _, mons = sd.read_fasta(os.path.join(ROOT, "tests", "golden", "DXZ1_star_monomers.fa"))
    rows = mons + [synth.revcomp(m) for m in mons]
    rng = np.random.default_rng(2)
    qs = []
    for _ in range(n):
        m = synth.mutate(rows[r.randrange(len(rows))], 0.016, 0.002, 0.002, rng)
        qs.append(m if isinstance(m, str) else bytes(m).decode())
    qb, qo = cv._pack(qs)
    tb, to = cv._pack(rows)
    out = {"intervals": n, "targets": len(rows)}
    for label, (q, t) in (("plain", ((qb, qo), (tb, to))), ("homopolymer", (cv._collapse(qb, qo), cv._collapse(tb, to)))):
        cells = float(np.diff(q[1]).sum()) * float(np.diff(t[1]).sum())
        best, wall = 1e30, 1e30
        for rep in range(4):
            t0 = time.perf_counter()
            res = sd.nw_identity(q, t)
            wall = min(wall, time.perf_counter() - t0)
            best = min(best, res["kernel_ms"])
        out[label] = {"pairs": len(res["matches"]), "cells": cells, "kernel_ms": round(best, 3), "call_ms": round(wall * 1e3, 2),
                      "kernel_gcups": round(cells / best / 1e6, 1), "call_gcups": round(cells / wall / 1e9, 1),
                      "pairs_per_s_call": round(len(res["matches"]) / wall)}
    try:
        import sd_oracle as O
        samp = [(qs[r.randrange(n)], rows[r.randrange(len(rows))]) for _ in range(20000)]
        t0 = time.perf_counter()
        for q, t in samp:
            O.ref_nw_path_counts(q, t)
        dt = time.perf_counter() - t0
        cells = sum(len(q) * len(t) for q, t in samp)
        out["reference_edlib_1core"] = {"pairs": len(samp), "s": round(dt, 3), "pairs_per_s": round(len(samp) / dt), "gcups": round(cells / dt / 1e9, 3),
                                        "note": "ctypes call overhead included (about 1 us per pair)"}
    except Exception as e:  # the compiled reference is absent
        out["reference_edlib_1core"] = {"unavailable": str(e)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
