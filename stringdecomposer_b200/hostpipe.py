"""Host mirror of the reference's `dp` main()/AlignReadsSet/SaveBatch (main.cpp:67-122, 272-285, 374-402) on top
of the C ABI: in-memory reads and monomers in, raw TSV text out.  Used by the tests and bench.py."""
import numpy as np

from ._lib import Decomposer, segment_read, postprocess, RECORD_DTYPE


def read_fasta(path):
    """Minimal FASTA reader with the reference's naming rule (first whitespace token, main.cpp:321-325)."""
    names, seqs = [], []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                tok = line[1:].split()
                names.append(tok[0] if tok else "")
                seqs.append([])
            elif seqs:
                seqs[-1].append(line)
    return names, ["".join(s) for s in seqs]


def segment_reads(reads, part_size, overlap, flavour="cuda"):
    """-> (segment strings, list of (read index, offset)) in read order."""
    segs, where = [], []
    for p, r in enumerate(reads):
        for off, ln in segment_read(len(r), part_size, overlap, flavour=flavour):
            segs.append(r[off:off + ln])
            where.append((p, off))
    return segs, where


def format_raw_tsv(read_name, monomer_names, records):
    """SaveBatch (main.cpp:272-285): 7 tab-separated columns, score as %f of a float."""
    M = len(monomer_names)
    out, prev_end = [], 0
    for r in records:
        row = int(r["row"])
        name = monomer_names[row] if row < M else monomer_names[row - M] + "'"
        out.append("%s\t%s\t%d\t%d\t%f\t%d\t%d\n" % (read_name, name, r["start"], r["end"], float(np.float32(r["score"])),
                                                      int(r["start"]) - prev_end, int(r["end"]) - int(r["start"])))
        prev_end = int(r["end"])
    return "".join(out)


def decompose_reads(read_names, reads, monomer_names, monomers, part_size=5000, overlap=500, scoring=(-1, -1, -1, 1),
                    devices=None, flavour="cuda", decomposer=None):
    """The whole `dp` run on in-memory sequences -> raw TSV text (what the reference writes to stdout)."""
    dec = decomposer or Decomposer(monomers, scoring[0], scoring[1], scoring[2], scoring[3], devices=devices, flavour=flavour)
    segs, where = segment_reads(reads, part_size, overlap, flavour=flavour)
    if not segs:
        return ""
    recs, offs = dec.decompose(segs)
    text = []
    s = 0
    for p, name in enumerate(read_names):
        parts = []
        while s < len(where) and where[s][0] == p:
            r = recs[offs[s]:offs[s + 1]].copy()
            r["start"] += where[s][1]
            r["end"] += where[s][1]
            parts.append(r)
            s += 1
        if not parts:
            continue
        allr = np.concatenate(parts) if parts else np.zeros(0, dtype=RECORD_DTYPE)
        if len(allr) == 0:
            continue
        text.append(format_raw_tsv(name, monomer_names, postprocess(allr, flavour=flavour)))
    if decomposer is None:
        dec.close()
    return "".join(text)
