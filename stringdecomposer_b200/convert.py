"""Host mirror of the reference's final-TSV stage (stringdecomposer/main.py:29-184) on top of the C ABI.

The reference rescans the raw `dp` output line by line and calls ``edlib.align`` once per (interval, monomer) pair --
once per line in the default ("light") mode, ``4 * monomers`` times per line with ``--second-best``.  Here all pairs of
a chunk of lines go to the device in one ``sd_identity`` call (identity_kernels.cu) and the rest -- second best,
homopolymer-collapsed ranking, the logistic-regression reliability flag, formatting -- is numpy on the returned counts.
Same function names and argument meaning as the reference; there is no CPU path: without the CUDA library or a GPU
the identity call raises.
"""
import numpy as np

from ._lib import nw_identity, convert_raw, SdError

# stringdecomposer/models/ont_logreg_model.txt as read at main.py:22-26: intercept, identity, identity - second best
LR_MODEL_COEF = [-31.48494996, 0.41784018, 0.69186882]

_COMPLEMENT = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")
MAX_PAIRS_PER_CALL = 1 << 24
MAX_LINES_PER_CALL = 1 << 21


def load_fasta(filename, tp="list"):
    """main.py:63-73: records as (id, upper-cased sequence); ``tp="map"`` gives {id: sequence} and, like
    Bio.SeqIO.to_dict, refuses duplicate ids.  The id is the first word of the title line."""
    recs, name, chunks = [], None, []
    with open(filename) as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    recs.append((name, "".join(chunks).upper()))
                words = line[1:].split()
                name, chunks = (words[0] if words else ""), []
            elif name is not None:
                chunks.append("".join(line.split()))
    if name is not None:
        recs.append((name, "".join(chunks).upper()))
    if tp != "map":
        return recs
    out = {}
    for name, seq in recs:
        if name in out:
            raise ValueError("Duplicate key '%s'" % name)
        out[name] = seq
    return out


def add_rc_monomers(monomers):
    """main.py:80-85: each monomer followed by its reverse complement, named with a trailing quote."""
    res = []
    for name, seq in monomers:
        res.append((name, seq))
        res.append((name + "'", seq.encode().translate(_COMPLEMENT)[::-1].decode()))
    return res


def convert_to_homo(seq):
    """main.py:88-93: homopolymer runs collapsed to one character."""
    b = np.frombuffer(seq.encode(), dtype=np.uint8)
    if len(b) == 0:
        return ""
    keep = np.ones(len(b), dtype=bool)
    keep[1:] = b[1:] != b[:-1]
    return b[keep].tobytes().decode()


def _strip_star(s):
    return s[:-1] if s.endswith("*") else s          # main.py:38-41


def _collapse(blob, off):
    """convert_to_homo of every sequence of a packed (blob, offsets) batch at once."""
    b = np.frombuffer(blob, dtype=np.uint8)
    keep = np.ones(len(b), dtype=bool)
    keep[1:] = b[1:] != b[:-1]
    starts = off[:-1][off[:-1] < len(b)]
    keep[starts] = True                               # a run never continues across two sequences
    csum = np.concatenate(([0], np.cumsum(keep, dtype=np.int64)))
    return b[keep].tobytes(), csum[off]


def _pack(seqs):
    """Sequences -> (blob, int64 offsets), each without the one trailing '*' that aai() drops (main.py:38-41)."""
    lens = np.fromiter(map(len, seqs), dtype=np.int64, count=len(seqs))
    blob = "".join(seqs).encode()
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    if len(blob) != off[-1]:                                  # non-ASCII text: byte lengths differ from str lengths
        bs = [s.encode() for s in seqs]
        blob = b"".join(bs)
        np.cumsum([len(x) for x in bs], out=off[1:])
    if b"*" in blob:
        b = np.frombuffer(blob, dtype=np.uint8)
        ends = off[1:] - 1
        star = (off[1:] > off[:-1]) & (b[np.maximum(ends, 0)] == ord("*"))
        if star.any():
            keep = np.ones(len(b), dtype=bool)
            keep[ends[star]] = False
            blob = b[keep].tobytes()
            off = off - np.concatenate(([0], np.cumsum(star)))
    return blob, off


def _percent(res):
    m = res["matches"].astype(np.float64)
    c = res["columns"].astype(np.float64)
    out = np.zeros(len(m), dtype=np.float64)
    nz = c > 0
    out[nz] = m[nz] / c[nz] * 100                     # aai /= total_length; aai * 100  (main.py:58-60)
    return out


def aai(ar, device=0, flavour="cuda"):
    """main.py:36-60 for one pair: percent identity of edlib's global alignment of ar[0] against ar[1]; 0 when
    either is empty."""
    q, t = _strip_star(str(ar[0])), _strip_star(str(ar[1]))
    return float(_percent(nw_identity([q], [t], device=device, flavour=flavour))[0])


def classify(score, second_best_score):
    """main.py:96-105: the ONT logistic-regression reliability flag, '+' or '?'."""
    score = np.asarray(score, dtype=np.float64)
    x = np.empty((len(score), 3), dtype=np.float64)
    x[:, 0] = 1
    x[:, 1] = score
    x[:, 2] = score - np.asarray(second_best_score, dtype=np.float64)
    return np.where(x.dot(np.asarray(LR_MODEL_COEF)) > 0, "+", "?") if len(score) else np.zeros(0, dtype="<U1")


def _score(meta, pieces, monomers, light, device, flavour, stats):
    """The numeric body of convert_read + classify (main.py:96-150) for any number of raw lines at once.  meta:
    [(monomer name, start, end)]; pieces: the read intervals of those lines.  Returns plain lists / arrays:
    light: {"score", "q"};  else also "second", "second_score", "homo" / "homo_score" (two best, by name), "uniq"
    (the keys of the reference's `scores` dict, in its order) and "table" (n x len(uniq) identities)."""
    n = len(meta)
    names = [m[0] for m in monomers]
    qblob, qoff = _pack(pieces)                       # aai() drops one trailing '*' of either side (main.py:38-41)
    tblob, toff = _pack([m[1] for m in monomers])

    def run(q, t, pairs=None):
        r = nw_identity(q, t, pairs=pairs, device=device, flavour=flavour)
        if stats is not None:
            stats["pairs"] = stats.get("pairs", 0) + len(r["matches"])
            stats["cells"] = stats.get("cells", 0) + int(_cells(q[1], t[1], pairs))
            stats["kernel_ms"] = stats.get("kernel_ms", 0.0) + r["kernel_ms"]
            stats["hirschberg_pairs"] = stats.get("hirschberg_pairs", 0) + r["hirschberg_pairs"]
        return _percent(r)

    if light:
        last = {nm: i for i, nm in enumerate(names)}              # the loop at main.py:113-116 keeps the last match
        cols = np.array([last[m[0]] for m in meta], dtype=np.int32)          # KeyError like scores[monomer]
        score = run((qblob, qoff), (tblob, toff), pairs=(np.arange(n, dtype=np.int32), cols))
        return {"score": score.tolist(), "q": classify(score, np.full(n, -1.0)).tolist()}

    nm = len(names)
    plain = run((qblob, qoff), (tblob, toff)).reshape(n, nm)
    homo = run(_collapse(qblob, qoff), _collapse(tblob, toff)).reshape(n, nm)
    # `scores` of main.py:123-126 is a dict: a repeated name keeps its first position and its last value
    uniq, col = [], {}
    for i, name in enumerate(names):
        if name not in col:
            uniq.append(name)
        col[name] = i
    table = plain[:, np.array([col[u] for u in uniq], dtype=np.int64)]
    upos = {u: i for i, u in enumerate(uniq)}
    mine = np.array([upos[m[0]] for m in meta], dtype=np.int64)
    rows = np.arange(n)
    best = table[rows, mine]
    if len(uniq) > 1:
        masked = table.copy()
        masked[rows, mine] = -np.inf
        second = np.argmax(masked, axis=1)                          # first of the largest, like main.py:131-135
        second_score = masked[rows, second]
        second_name = [uniq[i] for i in second.tolist()]
        second_out = second_score.tolist()
    else:
        second_score = np.full(n, -1.0)
        second_name = ["None"] * n                                  # str(None), main.py:145
        second_out = [-1] * n
    order = np.argsort(-homo, axis=1, kind="stable")[:, :2]          # sorted(..., key=-score), main.py:143
    return {"score": best.tolist(), "q": classify(best, second_score).tolist(), "second": second_name,
            "second_score": second_out,
            "homo": [[names[a], names[b]] for a, b in order.tolist()],
            "homo_score": np.take_along_axis(homo, order, axis=1).tolist(),
            "uniq": uniq, "table": table.tolist()}


def _rescore(meta, pieces, monomers, light, device, flavour, stats):
    """_score() in the shape convert_read returns: the reference's list of dicts (same keys, main.py:117-147)."""
    if not meta:
        return []
    sc = _score(meta, pieces, monomers, light, device, flavour, stats)
    res = []
    for k, (mono, start, end) in enumerate(meta):
        if light:
            res.append({"m": mono, "start": str(start), "end": str(end), "score": sc["score"][k],
                        "second_best": "None", "second_best_score": -1, "homo_best": "None", "homo_best_score": -1,
                        "homo_second_best": "None", "homo_second_best_score": -1, "alt": {}, "q": sc["q"][k]})
        else:
            res.append({"m": mono, "start": str(start), "end": str(end), "score": sc["score"][k],
                        "second_best": sc["second"][k], "second_best_score": sc["second_score"][k],
                        "homo_best": sc["homo"][k][0], "homo_best_score": sc["homo_score"][k][0],
                        "homo_second_best": sc["homo"][k][1], "homo_second_best_score": sc["homo_score"][k][1],
                        "alt": dict(zip(sc["uniq"], sc["table"][k])), "q": sc["q"][k]})
    return res


def _cells(qoff, toff, pairs):
    ql, tl = np.diff(qoff), np.diff(toff)
    if pairs is None:
        return ql.sum() * tl.sum()
    return (ql[pairs[0]] * tl[pairs[1]]).sum()


def convert_read(decomposition, read, monomers, light=False, device=0, flavour="cuda", stats=None):
    """main.py:107-150.  decomposition: [{"m", "start", "end"}]; read: the read's sequence; monomers: the
    add_rc_monomers() list.  Returns the reference's list of dicts (same keys)."""
    meta = [(d["m"], d["start"], d["end"]) for d in decomposition]
    return _rescore(meta, [read[s:e + 1] for _, s, e in meta], monomers, light, device, flavour, stats)


def _write(fout, fout_alt, read_names, dec, identity_th):
    for name, d in zip(read_names, dec):
        if d["score"] >= identity_th:
            fout.write("\t".join([name, d["m"], d["start"], d["end"], "{:.2f}".format(d["score"]),
                                  d["second_best"], "{:.2f}".format(d["second_best_score"]),
                                  d["homo_best"], "{:.2f}".format(d["homo_best_score"]),
                                  d["homo_second_best"], "{:.2f}".format(d["homo_second_best_score"]), d["q"]]) + "\n")
            for a, val in d["alt"].items():
                fout_alt.write("\t".join([name, a, d["start"], d["end"], "{:.2f}".format(val),
                                          "*" if a == d["m"] else "-"]) + "\n")


def print_read(fout, fout_alt, dec, read, monomers, identity_th, light, read_name=None, device=0, flavour="cuda",
               stats=None):
    """main.py:153-166: one line per alignment with identity >= identity_th, plus the per-monomer lines of the
    `_alt` file in --second-best mode.  ``read`` is the sequence, ``read_name`` what column 1 shows."""
    out = convert_read(dec, read, monomers, light, device=device, flavour=flavour, stats=stats)
    _write(fout, fout_alt, [read_name] * len(out), out, identity_th)


def _write_chunk(fout, fout_alt, names, meta, sc, identity_th, light):
    """print_read's two loops (main.py:155-166) straight from the arrays of _score(): same text as _write() on the
    dicts, without building them (the formatting is what this stage costs on the host)."""
    if light:
        fout.write("".join(
            f"{nm}\t{m}\t{s}\t{e}\t{v:.2f}\tNone\t-1.00\tNone\t-1.00\tNone\t-1.00\t{q}\n"
            for nm, (m, s, e), v, q in zip(names, meta, sc["score"], sc["q"]) if v >= identity_th))
        return
    uniq = sc["uniq"]
    main, alt = [], []
    for nm, (m, s, e), v, q, sn, sv, hn, hv, row in zip(names, meta, sc["score"], sc["q"], sc["second"], sc["second_score"],
                                                        sc["homo"], sc["homo_score"], sc["table"]):
        if v >= identity_th:
            main.append(f"{nm}\t{m}\t{s}\t{e}\t{v:.2f}\t{sn}\t{sv:.2f}\t{hn[0]}\t{hv[0]:.2f}\t{hn[1]}\t{hv[1]:.2f}\t{q}\n")
            head = f"{nm}\t"
            tail = f"\t{s}\t{e}\t"
            alt.append("".join(f"{head}{u}{tail}{x:.2f}\t{'*' if u == m else '-'}\n" for u, x in zip(uniq, row)))
    fout.write("".join(main))
    fout_alt.write("".join(alt))


def convert_tsv(decomposition, reads, monomers, outfile, identity_th, light, device=0, flavour="cuda", stats=None):
    """main.py:168-184: raw `dp` text -> ``outfile`` and ``outfile[:-4] + "_alt.tsv"``.  reads: {id: sequence};
    monomers: the add_rc_monomers() list.  Every raw line is rescored independently of the others (the reference's
    per-read grouping only selects the read to cut from), so lines of many reads share one device call."""
    per_line = 1 if light else 2 * max(1, len(monomers))
    chunk = max(1, min(MAX_PAIRS_PER_CALL // per_line, MAX_LINES_PER_CALL))
    lines = decomposition.split("\n")[:-1]
    first_word = {}                   # `x.split()[0]` of main.py:175-176, once per distinct name instead of per line
    with open(outfile[:-len(".tsv")] + "_alt.tsv", "w") as fout_alt, open(outfile, "w") as fout:
        for lo in range(0, len(lines), chunk):
            names, meta, pieces = [], [], []
            for ln in lines[lo:lo + chunk]:
                read, monomer, start, end = ln.split("\t")[:4]
                if read not in first_word:
                    first_word[read] = read.split()[0]
                if monomer not in first_word:
                    first_word[monomer] = monomer.split()[0]
                read, start, end = first_word[read], int(start), int(end)
                names.append(read)
                meta.append((first_word[monomer], start, end))
                pieces.append(reads[read][start:end + 1])
            _write_chunk(fout, fout_alt, names, meta, _score(meta, pieces, monomers, light, device, flavour, stats),
                         identity_th, light)


def convert_tsv_native(decomposition, reads, monomers, outfile, identity_th, light, device=0, flavour="cuda", stats=None):
    """convert_tsv() through the library's sd_convert (csrc/convert.cpp): same two files, byte for byte, without the
    per-line Python work -- the path the command line takes."""
    with open(outfile[:-len(".tsv")] + "_alt.tsv", "w") as fout_alt, open(outfile, "w") as fout:
        fout.flush(); fout_alt.flush()
        st = convert_raw(decomposition, reads, monomers, fout.fileno(), fout_alt.fileno(), identity_th, light,
                         device=device, flavour=flavour)
    if stats is not None:
        for k in ("pairs", "kernel_ms", "hirschberg_pairs"):
            stats[k] = stats.get(k, 0) + st[k]
    return st


def raw_chunks(path, max_bytes=32 << 20):
    """The raw `dp` file in pieces of whole reads (the lines of a read are consecutive, main.py:173-184), so that the final
    TSV can be produced without holding the whole raw text in memory the way main.py:195-197 does."""
    with open(path, "r") as f:
        buf, size, last = [], 0, None
        for ln in f:
            name = ln.split("\t", 1)[0].split()[0] if ln.strip() else last
            if size >= max_bytes and name != last:
                yield "".join(buf)
                buf, size = [], 0
            buf.append(ln)
            size += len(ln)
            last = name
        if buf:
            yield "".join(buf)


def convert_raw_file_native(raw_path, reads, monomers, outfile, identity_th, light, device=0, flavour="cuda", stats=None,
                            max_bytes=32 << 20):
    """convert_tsv_native() fed from the raw file in bounded chunks of whole reads; same two output files."""
    tot = {}
    with open(outfile[:-len(".tsv")] + "_alt.tsv", "w") as fout_alt, open(outfile, "w") as fout:
        fout.flush(); fout_alt.flush()
        for chunk in raw_chunks(raw_path, max_bytes):
            st = convert_raw(chunk, reads, monomers, fout.fileno(), fout_alt.fileno(), identity_th, light, device=device, flavour=flavour)
            for k, v in st.items():
                tot[k] = tot.get(k, 0) + v
    if stats is not None:
        for k in ("pairs", "kernel_ms", "hirschberg_pairs"):
            stats[k] = stats.get(k, 0) + tot.get(k, 0)
    return tot


__all__ = ["LR_MODEL_COEF", "load_fasta", "add_rc_monomers", "convert_to_homo", "aai", "classify", "convert_read",
           "print_read", "convert_tsv", "convert_tsv_native", "convert_raw_file_native", "raw_chunks", "SdError"]
