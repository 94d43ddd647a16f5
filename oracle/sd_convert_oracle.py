"""CPU oracle for the final-TSV stage of the reference (SURVEY §8 f1 + f4).  TEST INFRASTRUCTURE ONLY.

Restates stringdecomposer/main.py:87-184 (add_rc_monomers, convert_to_homo, classify, convert_read, print_read,
convert_tsv) in plain Python over the C identity oracle (sd_oracle.identity).  Biopython / python-edlib /
pandas are not needed.  Parity: PINNED -- tests/test_identity_oracle.py regenerates all 12 columns of the
reference's golden file test_data/final_decomposition_fc89af8.tsv from the golden raw TSV (both vendored under
tests/golden/) and compares them byte for byte.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sd_oracle as O  # noqa: E402

# stringdecomposer/models/ont_logreg_model.txt, read at main.py:25-26: intercept, identity, identity difference
LOGREG = (-31.48494996, 0.41784018, 0.69186882)
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def fasta_records(path):
    """(id, upper-cased sequence) per record; id = first word of the title (Bio.SeqIO's record.id, main.py:63-73)."""
    out, name, chunks = [], None, []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    out.append((name, "".join(chunks).upper()))
                words = line[1:].split()
                name, chunks = (words[0] if words else ""), []
            elif name is not None:
                chunks.append("".join(line.split()))
    if name is not None:
        out.append((name, "".join(chunks).upper()))
    return out


def revcomp(s):
    return "".join(_COMP.get(c, c) for c in reversed(s))


def with_rc(monomers):
    """main.py:80-85: every monomer is followed by its reverse complement, named with a trailing quote."""
    out = []
    for name, seq in monomers:
        out.append((name, seq))
        out.append((name + "'", revcomp(seq)))
    return out


def squeeze(seq):
    """convert_to_homo, main.py:88-93: collapse homopolymer runs."""
    out = []
    for c in seq:
        if not out or out[-1] != c:
            out.append(c)
    return "".join(out)


def rescore(records, read_seq, monomers, light):
    """convert_read + classify, main.py:96-150.  records: [(monomer name, start, end)]; monomers: with_rc() list."""
    rows = []
    for mono, start, end in records:
        piece = read_seq[start:end + 1]
        if light:
            score = None
            for name, seq in monomers:
                if name == mono:
                    score = O.identity(piece, seq)
            rows.append(dict(m=mono, start=start, end=end, score=score, second="None", second_score=-1,
                             homo="None", homo_score=-1, homo2="None", homo2_score=-1, alt={}))
            continue
        scores = {}
        for name, seq in monomers:
            scores[name] = O.identity(piece, seq)
        second, second_score = None, -1
        for name, val in scores.items():
            if name != mono and (not second or second_score < val):
                second, second_score = name, val
        hp = squeeze(piece)
        homo = [(name, O.identity(hp, squeeze(seq))) for name, seq in monomers]
        homo.sort(key=lambda x: -x[1])                        # stable, like sorted() at main.py:143
        rows.append(dict(m=mono, start=start, end=end, score=scores[mono], second=str(second), second_score=second_score,
                         homo=homo[0][0], homo_score=homo[0][1], homo2=homo[1][0], homo2_score=homo[1][1], alt=scores))
    for r in rows:                                            # classify, main.py:96-105
        z = LOGREG[0] * 1 + LOGREG[1] * r["score"] + LOGREG[2] * (r["score"] - r["second_score"])
        r["q"] = "+" if z > 0 else "?"
    return rows


def final_tsv(raw_text, reads, monomers, min_identity=0, light=True):
    """convert_tsv + print_read, main.py:153-184.  reads: {id: sequence}; monomers: forward [(name, seq)].
    Returns (final TSV text, alt TSV text)."""
    mons = with_rc(monomers)
    main, alt = [], []

    def flush(read, recs):
        for r in rescore(recs, reads[read], mons, light):
            if r["score"] >= min_identity:
                main.append("\t".join([read, r["m"], str(r["start"]), str(r["end"]), "%.2f" % r["score"], r["second"],
                                       "%.2f" % r["second_score"], r["homo"], "%.2f" % r["homo_score"], r["homo2"],
                                       "%.2f" % r["homo2_score"], r["q"]]) + "\n")
                for name, val in r["alt"].items():
                    alt.append("\t".join([read, name, str(r["start"]), str(r["end"]), "%.2f" % val,
                                          "*" if name == r["m"] else "-"]) + "\n")

    cur, prev = [], None
    for line in raw_text.split("\n")[:-1]:
        read, mono, start, end = line.split("\t")[:4]
        read, mono = read.split()[0], mono.split()[0]
        if prev is not None and read != prev:
            flush(prev, cur)
            cur = []
        prev = read
        cur.append((mono, int(start), int(end)))
    if cur:
        flush(prev, cur)
    return "".join(main), "".join(alt)
