#!/usr/bin/env python
"""Whole reference command line (DP + identity rescoring + final TSV) on BASELINE config 2, default and --second-best:
    python tools/pipeline_probe.py [profile]
prints one JSON line with the stage times main.py logs; `profile` adds a cProfile of the --second-best rescoring."""
import cProfile
import io
import json
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stringdecomposer_b200 import synth, cli as sdmain, convert as cv  # noqa: E402


def main():
    td = tempfile.mkdtemp()
    rn, reads, mn, mons = synth.config2()
    synth.write_fasta(os.path.join(td, "reads.fa"), rn, reads)
    synth.write_fasta(os.path.join(td, "monomers.fa"), mn, mons)
    out = {}
    for label, extra in (("light", []), ("second_best", ["--second-best"]), ("second_best_again", ["--second-best"])):
        t0 = time.perf_counter()
        sdmain.main([os.path.join(td, "reads.fa"), os.path.join(td, "monomers.fa"), "-o", os.path.join(td, label)] + extra)
        wall = time.perf_counter() - t0
        log = open(os.path.join(td, label, "stringdecomposer.log")).read().splitlines()
        out[label] = {"wall_s": round(wall, 3), "stages": [ln.split(" - ")[-1] for ln in log if "decomposition " in ln][-1],
                      "rows": sum(1 for _ in open(os.path.join(td, label, "final_decomposition.tsv")))}
    print(json.dumps(out))
    if len(sys.argv) > 1:
        raw = open(os.path.join(td, "light", "final_decomposition_raw.tsv")).read()
        rd = cv.load_fasta(os.path.join(td, "reads.fa"), "map")
        mm = cv.add_rc_monomers(cv.load_fasta(os.path.join(td, "monomers.fa")))
        pr = cProfile.Profile()
        pr.enable()
        cv.convert_tsv(raw, rd, mm, os.path.join(td, "p.tsv"), 0, False)
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
        print(s.getvalue())


if __name__ == "__main__":
    main()
