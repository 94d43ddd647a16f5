"""ctypes wrapper of the CPU oracle (oracle/sd_oracle.c).  TEST INFRASTRUCTURE ONLY -- imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product package."""
import ctypes as C
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libsd_oracle.so")
BIN = os.path.join(HERE, "_build", "oracle_dp")
REF_BIN = os.path.join(HERE, "_ref", "dp")


class Rec(C.Structure):
    _fields_ = [("row", C.c_int), ("start", C.c_int), ("end", C.c_int), ("score", C.c_float)]


def build():
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.sdo_align_segment.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.POINTER(Rec), C.c_int]
        _lib.sdo_align_segment.restype = C.c_int
        _lib.sdo_segment_read.argtypes = [C.c_long, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
        _lib.sdo_segment_read.restype = C.c_int
        _lib.sdo_postprocess.argtypes = [C.POINTER(Rec), C.c_int, C.POINTER(Rec)]
        _lib.sdo_postprocess.restype = C.c_int
    return _lib


_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def rows_of(monomers):
    """forward monomers, then all reverse complements (main.cpp:364-371)"""
    return list(monomers) + ["".join(_COMP[c] for c in reversed(m)) for m in monomers]


def align_segment(seg, monomers, scoring=(-1, -1, -1, 1)):
    """AlignPartClassicDP on one segment -> list of (row, start, end, score)"""
    rows = rows_of(monomers)
    off = (C.c_int * (len(rows) + 1))()
    for i, r in enumerate(rows):
        off[i + 1] = off[i] + len(r)
    out = (Rec * (len(seg) + 1))()
    n = lib().sdo_align_segment(seg.encode(), len(seg), "".join(rows).encode(), off, len(rows), scoring[0], scoring[1],
                                scoring[2], scoring[3], out, len(seg) + 1)
    if n < 0:
        raise RuntimeError("oracle failed (%d)" % n)
    return [(out[i].row, out[i].start, out[i].end, out[i].score) for i in range(n)]


def segment_read(read_len, part, overlap):
    n = lib().sdo_segment_read(read_len, part, overlap, None, None, 0)
    offs, lens = (C.c_int * max(n, 1))(), (C.c_int * max(n, 1))()
    lib().sdo_segment_read(read_len, part, overlap, offs, lens, n)
    return [(offs[i], lens[i]) for i in range(n)]


def postprocess(recs):
    a = (Rec * max(len(recs), 1))(*[Rec(*r) for r in recs])
    b = (Rec * max(len(recs), 1))()
    n = lib().sdo_postprocess(a, len(recs), b)
    return [(b[i].row, b[i].start, b[i].end, b[i].score) for i in range(n)]


def run_cli(binary, reads_fa, monomers_fa, threads=1, part=5000, overlap=500, scoring=None, extra=()):
    """Run a dp-compatible binary; returns (status, stdout bytes, stderr bytes)."""
    cmd = [binary, reads_fa, monomers_fa, str(threads), str(part), str(overlap)]
    if scoring is not None:
        cmd += [str(x) for x in scoring]
    cmd += list(extra)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return p.returncode, p.stdout, p.stderr


def decompose_reads(read_names, reads, monomer_names, monomers, part_size=5000, overlap=500, scoring=(-1, -1, -1, 1),
                    binary=None, threads=None):
    """Whole `dp` run of the oracle port (or of `binary`, e.g. the compiled reference) on in-memory sequences."""
    if not os.path.exists(BIN):
        build()
    with tempfile.TemporaryDirectory() as td:
        rp, mp = os.path.join(td, "r.fa"), os.path.join(td, "m.fa")
        with open(rp, "w") as f:
            for n, s in zip(read_names, reads):
                f.write(">%s\n%s\n" % (n, s))
        with open(mp, "w") as f:
            for n, s in zip(monomer_names, monomers):
                f.write(">%s\n%s\n" % (n, s))
        st, out, err = run_cli(binary or BIN, rp, mp, threads or min(8, os.cpu_count() or 1), part_size, overlap,
                               None if tuple(scoring) == (-1, -1, -1, 1) else scoring)
        if st != 0:
            raise RuntimeError("oracle dp exited with %d: %s" % (st, err.decode()[-300:]))
        return out.decode()


# ---- identity rescoring: sd_identity_oracle.c and the reference's own edlib ----------------
REF_EDLIB = os.path.join(HERE, "_ref", "libedlib_ref.so")
_ref_edlib = None


def _counts(fn, q, t):
    qb, tb = q.encode(), t.encode()
    m, c = C.c_int(), C.c_int()
    d = fn(qb, len(qb), tb, len(tb), C.byref(m), C.byref(c))
    return d, m.value, c.value


def nw_path_counts(q, t):
    """(edit distance, '=' columns, alignment columns) of edlib's NW path; distance -1 when either is empty."""
    f = lib().sdo_nw_path_counts
    f.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return _counts(f, q, t)


def nw_uses_traceback(qlen, tlen):
    return bool(lib().sdo_nw_uses_traceback(int(qlen), int(tlen)))


def ref_nw_path_counts(q, t):
    """The same triple from the reference's vendored edlib.cpp (oracle/_ref/libedlib_ref.so); non-empty inputs only."""
    global _ref_edlib
    if _ref_edlib is None:
        _ref_edlib = C.CDLL(REF_EDLIB)
        _ref_edlib.ref_nw_path_counts.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return _counts(_ref_edlib.ref_nw_path_counts, q, t)


def identity(q, t):
    """aai of main.py:36-60: percent of '=' columns in edlib's NW alignment; 0 when either string is empty."""
    q = q[:-1] if q.endswith("*") else q                      # main.py:38-41
    t = t[:-1] if t.endswith("*") else t
    d, m, c = nw_path_counts(q, t)
    if d == -1:
        return 0
    return m / c * 100
