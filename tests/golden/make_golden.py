#!/usr/bin/env python
"""Generates tests/golden/edge_cases.json by running the UNMODIFIED reference binary (oracle/_ref/dp, compiled from
/root/reference by oracle/Makefile) on small inputs that exercise the behaviours catalogued in SURVEY.md App. B.
Run in the build container (where /root/reference is mounted):   python tests/golden/make_golden.py
The JSON (inputs as FASTA text, argv tail, exit status, stdout, stderr) is committed; the GPU box never needs the
reference tree."""
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "dp")
sys.path.insert(0, ROOT)
from stringdecomposer_b200.hostpipe import read_fasta  # noqa: E402
from stringdecomposer_b200 import synth  # noqa: E402


def fasta(names, seqs, width=0, eol="\n"):
    out = []
    for n, s in zip(names, seqs):
        out.append(">" + n + eol)
        if width:
            out += [s[i:i + width] + eol for i in range(0, len(s), width)]
        else:
            out.append(s + eol)
    return "".join(out)


def run(reads_txt, mono_txt, argv_tail):
    with tempfile.TemporaryDirectory() as td:
        rp, mp = os.path.join(td, "reads.fa"), os.path.join(td, "monomers.fa")
        open(rp, "w", newline="").write(reads_txt)
        open(mp, "w", newline="").write(mono_txt)
        p = subprocess.run([REF, rp, mp] + [str(a) for a in argv_tail], stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=td)
        return p.returncode, p.stdout.decode(), p.stderr.decode().replace(td + "/", "")


def main():
    rn, rs = read_fasta(os.path.join(HERE, "config1_read.fa"))
    mn, ms = read_fasta(os.path.join(HERE, "DXZ1_star_monomers.fa"))
    read = rs[0]
    mono_txt = fasta(mn, ms)
    rnd = random.Random(7)
    cases = []

    def add(name, reads_txt, mtxt, tail, note=""):
        st, out, err = run(reads_txt, mtxt, tail)
        cases.append({"name": name, "note": note, "reads_fa": reads_txt, "monomers_fa": mtxt, "argv_tail": [str(a) for a in tail],
                      "status": st, "stdout": out, "stderr": err})
        print("%-34s status=%3d rows=%d" % (name, st, out.count("\n")))

    # read lengths around the segmentation rule (main.cpp:74): 1, < overlap, part+overlap-1 / = / +1, 2 parts
    for L in (1, 2, 300, 499, 500, 501, 1299, 1300, 1301, 2000, 2600):
        add("len_%d_part1000_ov300" % L, fasta(["r%d" % L], [read[1000:1000 + L]]), mono_txt, [2, 1000, 300],
            "segment boundary rule, part 1000 overlap 300")
    add("len_5499_default", fasta(["a"], [read[:5499]]), mono_txt, [2, 5000, 500])
    add("len_5500_default", fasta(["a"], [read[:5500]]), mono_txt, [2, 5000, 500])
    add("len_5501_default", fasta(["a"], [read[:5501]]), mono_txt, [2, 5000, 500])
    add("len_10000_default", fasta(["a"], [read[20000:30000]]), mono_txt, [2, 5000, 500])
    add("multi_read", fasta(["r1", "r2 with description", "r3"], [read[:3000], read[40000:47000], read[90000:]], width=60), mono_txt,
        [3, 2000, 400], "three reads, wrapped lines, header with description (main.cpp:321-325)")
    add("part700_ov50", fasta(["x"], [read[5000:9000]]), mono_txt, [1, 700, 50], "PostProcessing with a small overlap")
    add("part1500_ov1400", fasta(["x"], [read[5000:11000]]), mono_txt, [1, 1500, 1400], "overlap nearly as long as the part")
    for sc in ((-2, -2, -3, 1), (-3, -2, -4, 2), (-1, -3, -2, 3), (0, -1, -1, 1), (-1, 0, -1, 1), (-6, -6, -6, 1), (-1, -1, -1, 5)):
        add("scoring_%s" % "_".join(map(str, sc)), fasta(["s"], [read[30000:33500]]), mono_txt, [2, 1000, 300] + list(sc),
            "custom scoring, argc == 10 (main.cpp:381-386)")
    add("argc11_scores_ignored", fasta(["s"], [read[30000:32000]]), mono_txt, [2, 1000, 300, -2, -2, -3, 1, -1],
        "argc == 11: scores ignored, ed_thr = -1 (main.cpp:388-391)")
    # N handling: fifth symbol that matches itself
    rn_read = list(read[50000:53000])
    for p in range(100, 3000, 97):
        rn_read[p] = "N"
    add("N_in_read", fasta(["n"], ["".join(rn_read)]), mono_txt, [2, 1000, 300], "N in the read (main.cpp:330,342-344)")
    ms_n = list(ms)
    ms_n[3] = ms_n[3][:50] + "NNN" + ms_n[3][53:]
    add("N_in_monomer", fasta(["n"], ["".join(rn_read)]), fasta(mn, ms_n), [2, 1000, 300], "N in a monomer and in the read: N==N matches")
    # duplicates -> arg-max ties must resolve to the lowest row (main.cpp:212, :230-236)
    add("dup_monomers_fwd", fasta(["d"], [read[60000:63000]]), fasta(["X1", "X2", "B", "X3"], [ms[0], ms[0], ms[1], ms[0]]), [2, 1000, 300])
    add("dup_monomers_rev", fasta(["d"], [read[60000:63000]]), fasta(["X3", "B", "X2", "X1"], [ms[0], ms[1], ms[0], ms[0]]), [2, 1000, 300])
    pal = "ACGTTGCAACGT" * 5 + "ACGTTGCAACGT"[::-1].translate(str.maketrans("ACGT", "TGCA")) * 5
    add("palindromic_monomer", fasta(["p"], [(pal * 12)[:900]]), fasta(["pal", "other"], [pal, ms[2]]), [1, 400, 100],
        "monomer equal to its own reverse complement: forward row wins ties")
    add("single_monomer", fasta(["one"], [read[70000:72000]]), fasta(mn[:1], ms[:1]), [1, 1000, 300])
    add("short_monomers", fasta(["sm"], ["".join(rnd.choice("ACGT") for _ in range(700))]),
        fasta(["m1", "m2", "m3", "m4"], ["A", "AC", "GTT", "ACGTACG"]), [1, 300, 100], "monomers of length 1, 2, 3, 7")
    add("len1_monomer_only", fasta(["l1"], ["ACGTTTGACA" * 8]), fasta(["a", "c"], ["A", "C"]), [1, 50, 10])
    add("tiny_alphabet_ties", fasta(["t"], ["".join(rnd.choice("AC") for _ in range(1200))]),
        fasta(["p", "q", "r"], ["ACCA" * 9, "CACA" * 11, "ACCA" * 9]), [2, 500, 100], "two-letter alphabet: thousands of ties")
    add("poly_A", fasta(["pa"], ["A" * 1500]), fasta(["ma", "mb"], ["A" * 60, "AAAT" * 20]), [2, 500, 100])
    # error contract
    add("lowercase_rejected", fasta(["lc"], [read[:500].lower()]), mono_txt, [1, 1000, 300], "exit 255 (main.cpp:29,333-336)")
    add("crlf_rejected", fasta(["cr"], [read[:500]], eol="\r\n"), mono_txt, [1, 1000, 300], "\\r is an undefined symbol")
    add("illegal_symbol_in_monomer", fasta(["ok"], [read[:500]]), fasta(["bad"], ["ACGTRYACGT"]), [1, 1000, 300])
    add("blank_lines", ">b\n" + read[:700] + "\n\n" + read[700:1500] + "\n\n", mono_txt, [1, 1000, 300], "blank lines are harmless")
    add("empty_read_alone", ">e\n", mono_txt, [1, 1000, 300], "no output, exit 0")
    st, out, err = (subprocess.run([REF, "a", "b", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE).returncode, None, None)
    p = subprocess.run([REF, "a", "b", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    cases.append({"name": "argc_lt_5", "note": "usage on stdout, status 255 (main.cpp:375-379)", "reads_fa": None, "monomers_fa": None,
                  "argv_raw": ["a", "b", "1"], "status": p.returncode, "stdout": p.stdout.decode(), "stderr": p.stderr.decode()})
    # --ed_thr pre-filter (FilterMonomersForRead, main.cpp:135-149): argc == 11, scores ignored, rows re-ordered per segment
    for thr in (0, 5, 12, 30, 200):
        add("ed_thr_%d" % thr, fasta(["e1", "e2"], [read[10000:14000], read[47000:49500]]), mono_txt, [2, 1000, 300, -1, -1, -1, 1, thr],
            "argc == 11 with ed_thr = %d" % thr)
    add("ed_thr_dups", fasta(["d"], [read[60000:63000]]), fasta(["X3", "B", "X2", "X1"], [ms[0], ms[1], ms[0], ms[0]]), [2, 1000, 300, -1, -1, -1, 1, 20],
        "duplicate monomers: equal distances are ordered by row index")
    add("ed_thr_short_monomers", fasta(["sm"], ["".join(rnd.choice("ACGT") for _ in range(700))]),
        fasta(["m1", "m2", "m3", "m4"], ["A", "AC", "GTT", "ACGTACG"]), [1, 300, 100, -1, -1, -1, 1, 1])
    for seed in range(12):
        al = ["AT", "AC", "ACGT", "ACGTN"][seed % 4]
        rnn, rr, mnn, mm = synth.random_case(2000 + seed, alphabet=al)
        add("ed_thr_fuzz_%02d_%s" % (seed, al), fasta(rnn, rr), fasta(mnn, mm), [2, [50, 120, 300][seed % 3], [10, 30, 100][seed % 3], -1, -1, -1, 1, [0, 1, 2, 4, 8, 16][seed % 6]],
            "random case seed %d with the pre-filter" % (2000 + seed))
    # random fuzz cases (tiny alphabets, duplicates, odd scoring)
    scorings = [(-1, -1, -1, 1), (-2, -2, -3, 1), (-3, -2, -4, 2), (0, -1, -1, 1), (-1, 0, -2, 2), (-4, -1, -1, 1), (-1, -4, -1, 2)]
    geoms = [(50, 10), (120, 30), (300, 100), (5000, 500)]
    for seed in range(40):
        al = ["A", "AT", "AC", "ACGT", "ACGTN"][seed % 5]
        rnn, rr, mnn, mm = synth.random_case(1000 + seed, alphabet=al)
        sc = scorings[seed % len(scorings)]
        part, ov = geoms[seed % len(geoms)]
        tail = [2, part, ov] + (list(sc) if sc != (-1, -1, -1, 1) else [])
        add("fuzz_%02d_%s" % (seed, al), fasta(rnn, rr), fasta(mnn, mm), tail, "random case seed %d" % (1000 + seed))
    json.dump({"generator": "tests/golden/make_golden.py", "reference": "ablab/stringdecomposer v1.1.2 src/main.cpp built by oracle/Makefile",
               "cases": cases}, open(os.path.join(HERE, "edge_cases.json"), "w"), indent=0)
    print("wrote %d cases" % len(cases))


if __name__ == "__main__":
    main()
