#!/usr/bin/env python
"""Generates tests/golden/nw_pairs.json: (query, target) pairs with the edit distance, '=' count and alignment length
that the reference's OWN vendored edlib reports for edlibAlign(query, target, NW, PATH) -- i.e. what main.py:29-60
gets from python-edlib.  The library is oracle/_ref/libedlib_ref.so, compiled by oracle/Makefile from
/root/reference/stringdecomposer/src/edlib.cpp where it lies.  Run in the build container:
    python tests/golden/make_golden_identity.py
The JSON is committed; the GPU box never needs the reference tree."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import sd_oracle as O  # noqa: E402
from stringdecomposer_b200.hostpipe import read_fasta  # noqa: E402


def mutate(rnd, s, rate, alpha):
    out = []
    for c in s:
        r = rnd.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append(rnd.choice(alpha))
            continue
        if r < rate:
            out.append(rnd.choice(alpha))
        out.append(c)
    return "".join(out)


def main():
    rnd = random.Random(11)
    _, mons = read_fasta(os.path.join(HERE, "DXZ1_star_monomers.fa"))
    _, reads = read_fasta(os.path.join(HERE, "config1_read.fa"))
    read = reads[0]
    pairs = []

    def add(q, t, note):
        d, m, c = O.ref_nw_path_counts(q, t)
        pairs.append({"q": q, "t": t, "distance": d, "matches": m, "columns": c, "note": note})

    for q, t in (("A", "A"), ("A", "C"), ("A", "AAAA"), ("AAAA", "A"), ("ACGT", "TGCA"), ("AC", "CA"), ("AAC", "ACA"),
                 ("ACGTACGT", "ACGT"), ("GATTACA", "GCATGCU"), ("N", "N"), ("ANNA", "ANA")):
        add(q, t, "tiny")
    for k in range(60):                                   # monomer against noisy copies: the production shape
        m = rnd.choice(mons)
        add(mutate(rnd, m, rnd.choice((0.02, 0.1, 0.3)), "ACGT"), rnd.choice(mons) if k % 3 == 0 else m, "monomer-like")
    for k in range(40):                                   # real read intervals against monomers
        a = rnd.randrange(0, len(read) - 400)
        add(read[a:a + rnd.randint(120, 260)], rnd.choice(mons), "read interval")
    for alpha in ("A", "AC", "ACGT", "ACGTN"):            # low-complexity text: many co-optimal paths
        for k in range(25):
            q = "".join(rnd.choice(alpha) for _ in range(rnd.randint(1, 90)))
            t = mutate(rnd, q, 0.4, alpha) or alpha[0] if k % 2 else "".join(rnd.choice(alpha) for _ in range(rnd.randint(1, 90)))
            add(q, t, "alphabet " + alpha)
    for n in (63, 64, 65, 127, 128, 129, 191, 192, 193, 255, 256, 257, 300, 511, 513, 700, 1100):   # strip / tile edges
        q = "".join(rnd.choice("ACGT") for _ in range(n))
        add(q, mutate(rnd, q, 0.15, "ACGT"), "query length %d" % n)
        add(mutate(rnd, q, 0.15, "ACGT"), q, "target length %d" % n)
    add("ACGT" * 300, "ACG" * 350, "periodic")
    add("A" * 500, "A" * 320, "homopolymer")
    out = os.path.join(HERE, "nw_pairs.json")
    with open(out, "w") as f:
        json.dump({"generator": "tests/golden/make_golden_identity.py", "source": "reference edlib.cpp (NW, PATH)", "pairs": pairs}, f)
    print("wrote", out, len(pairs), "pairs")


if __name__ == "__main__":
    main()
