"""Seeded synthetic alpha-satellite workloads (BASELINE.json configs 2-5) and edge cases."""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DXZ1_FASTA = os.path.join(os.path.dirname(_HERE), "tests", "golden", "DXZ1_star_monomers.fa")
_ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    _COMP[a] = b


def load_dxz1():
    from .hostpipe import read_fasta
    return read_fasta(DXZ1_FASTA)


def revcomp(s):
    return _COMP[np.frombuffer(s.encode(), dtype=np.uint8)][::-1].tobytes().decode()


def mutate(seq, sub, ins, dele, rng):
    """Independent per-base substitutions / insertions (before the base) / deletions, uniform bases."""
    a = np.frombuffer(seq.encode(), dtype=np.uint8).copy() if isinstance(seq, str) else np.asarray(seq, dtype=np.uint8).copy()
    n = len(a)
    u = rng.random(n)
    is_sub = u < sub
    if is_sub.any():
        # substitute with a different base
        cur = np.searchsorted(_ALPHA, a[is_sub])
        cur = np.where(_ALPHA[np.clip(cur, 0, 3)] == a[is_sub], cur, 0)
        a[is_sub] = _ALPHA[(cur + rng.integers(1, 4, is_sub.sum())) % 4]
    keep = rng.random(n) >= dele
    n_ins = (rng.random(n) < ins).astype(np.int64)
    reps = keep.astype(np.int64) + n_ins
    out = np.repeat(a, reps)
    # positions of inserted symbols: the first copy of every base that has an insertion
    starts = np.cumsum(reps) - reps
    ins_pos = starts[n_ins > 0]
    out[ins_pos] = _ALPHA[rng.integers(0, 4, len(ins_pos))]
    return out


def hor_array(monomers, length, divergence, seed, order=None):
    """Concatenate the monomers (HOR unit) until `length` bp, every copy independently mutated (80% of the
    divergence as substitutions, 10% insertions, 10% deletions); upper-case ACGT, trimmed to exactly `length`."""
    rng = np.random.default_rng(seed)
    unit = "".join(monomers[i] for i in (order if order is not None else range(len(monomers))))
    parts, tot = [], 0
    while tot < length + len(unit):
        m = mutate(unit, 0.8 * divergence, 0.1 * divergence, 0.1 * divergence, rng)
        parts.append(m)
        tot += len(m)
    return np.concatenate(parts)[:length].tobytes().decode()


def config2(length=2_000_000, seed=2):
    """One 2 Mb cenX-like DXZ1 HOR contig, ~2 % divergence (BASELINE config 2)."""
    names, mons = load_dxz1()
    return ["cenX_synth_%d" % length], [hor_array(mons, length, 0.02, seed)], names, mons


def sample_reads(n_reads, read_len, err_sub, err_ins, err_del, seed, array_len=None):
    names, mons = load_dxz1()
    rng = np.random.default_rng(seed)
    array_len = array_len or max(4 * read_len, 2_000_000)
    arr = np.frombuffer(hor_array(mons, array_len, 0.02, seed).encode(), dtype=np.uint8)
    reads, rnames = [], []
    for r in range(n_reads):
        st = int(rng.integers(0, array_len - read_len - read_len // 5))
        frag = arr[st:st + read_len + read_len // 5]
        m = mutate(frag, err_sub, err_ins, err_del, rng)[:read_len]
        s = m.tobytes().decode()
        if rng.random() < 0.5:
            s = revcomp(s)
        reads.append(s)
        rnames.append("read_%d" % r)
    return rnames, reads, names, mons


def config3(n_reads=2000, read_len=100_000, seed=3):
    """ONT-like reads, ~10 % indel-heavy error (BASELINE config 3)."""
    return sample_reads(n_reads, read_len, 0.02, 0.04, 0.04, seed)


def config4(n_reads=20000, read_len=15_000, seed=4):
    """HiFi-like reads, 0.5 % error; scoring -2,-2,-3,1 is applied by the caller (BASELINE config 4)."""
    return sample_reads(n_reads, read_len, 0.003, 0.001, 0.001, seed)


def config5(n_monomers=1000, total=10_000_000, seed=5):
    """~1000 synthetic monomers (families mutated 10-35 % off the 12 DXZ1 monomers) vs 5 x 2 Mb HOR arrays."""
    names, mons = load_dxz1()
    rng = np.random.default_rng(seed)
    fam, fnames = [], []
    for j in range(n_monomers):
        base = mons[j % len(mons)]
        d = rng.uniform(0.10, 0.35)
        m = mutate(base, 0.8 * d, 0.1 * d, 0.1 * d, rng)
        L = int(np.clip(len(m), 165, 180))
        m = m[:L] if len(m) >= L else np.concatenate([m, _ALPHA[rng.integers(0, 4, L - len(m))]])
        fam.append(m.tobytes().decode())
        fnames.append("M%04d" % j)
    reads, rnames = [], []
    per = total // 5
    for c in range(5):
        k = int(rng.integers(8, 21))
        sub = rng.choice(n_monomers, k, replace=False)
        reads.append(hor_array(fam, per, 0.02, seed * 100 + c, order=list(sub)))
        rnames.append("contig_%d" % c)
    return rnames, reads, fnames, fam


def random_case(seed, alphabet="ACGT", n_monomers=(1, 8), mono_len=(1, 100), n_reads=(1, 3), read_len=(1, 900), dup=True):
    """Small random reads/monomers for parity fuzzing (tiny alphabets force arg-max ties)."""
    rng = np.random.default_rng(seed)
    al = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    M = int(rng.integers(n_monomers[0], n_monomers[1] + 1))
    mons = []
    for _ in range(M):
        if dup and mons and rng.random() < 0.25:
            mons.append(mons[int(rng.integers(0, len(mons)))])      # exact duplicate -> ties between rows
        else:
            mons.append(al[rng.integers(0, len(al), int(rng.integers(mono_len[0], mono_len[1] + 1)))].tobytes().decode())
    R = int(rng.integers(n_reads[0], n_reads[1] + 1))
    reads = []
    for _ in range(R):
        L = int(rng.integers(read_len[0], read_len[1] + 1))
        if rng.random() < 0.6 and mons:
            # read made of noisy monomer copies, so that real alignments exist
            parts, tot = [], 0
            while tot < L:
                m = mons[int(rng.integers(0, M))]
                if rng.random() < 0.5:
                    m = revcomp(m)
                mm = mutate(m, 0.05, 0.03, 0.03, rng)
                mm = np.where(np.isin(mm, al), mm, al[mm % len(al)])
                parts.append(mm)
                tot += len(mm)
            reads.append(np.concatenate(parts)[:L].tobytes().decode())
        else:
            reads.append(al[rng.integers(0, len(al), L)].tobytes().decode())
    return ["r%d" % i for i in range(R)], reads, ["m%d" % j for j in range(M)], mons


def write_fasta(path, names, seqs, width=0):
    with open(path, "w") as f:
        for n, s in zip(names, seqs):
            f.write(">%s\n" % n)
            if width:
                for i in range(0, len(s), width):
                    f.write(s[i:i + width] + "\n")
            else:
                f.write(s + "\n")
