#!/usr/bin/env python
"""Differential fuzzing on CPU: random reads / monomer sets / argv through the UNMODIFIED reference binary (oracle/_ref/dp,
build container only) and through the host emulator of the kernels (tests/emu/_build/dp_emu -- same host pipeline, same
per-lane functions as the CUDA kernels), with random launch geometries, chunk sizes and device counts forced through the
environment.  Compares stdout, stderr and exit status.  A soak tool, not part of the test-suites:
    python tools/fuzz_vs_reference.py --seed 1 --cases 500"""
import argparse
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stringdecomposer_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "dp")
EMU = os.environ.get("SD_FUZZ_BINARY") or os.path.join(ROOT, "tests", "emu", "_build", "dp_emu")     # e.g. a sanitizer build of dp_emu
SCORINGS = [(-1, -1, -1, 1), (-2, -2, -3, 1), (-3, -2, -4, 2), (0, -1, -1, 1), (-1, 0, -1, 1), (-1, -1, -1, 5), (-6, -6, -6, 1), (-1, -2, 0, 0)]
GEOMS = ["", "", "8,32,1", "12,16,2", "19,10,1", "24,8,3", "16,4,2", "48,2,1", "20,10,2", "24,8,1", "48,4,2", "32,8,1"]
LATS = ["", "", "0", "6,32,4", "6,32,1", "12,16,2", "24,8,1", "12,32,1"]


def fasta(names, seqs, width):
    out = []
    for n, s in zip(names, seqs):
        out.append(">" + n + "\n")
        if width:
            out += [s[i:i + width] + "\n" for i in range(0, len(s), width)]
        else:
            out.append(s + "\n")
    return "".join(out)


def realistic(rng, seed):
    """alpha-satellite-like input: a subset of the DXZ1 monomers, reads cut from a noisy higher-order-repeat array"""
    mn, mm = synth.load_dxz1()
    keep = sorted(rng.sample(range(len(mm)), rng.randint(2, len(mm))))
    mn, mm = [mn[i] for i in keep], [mm[i] for i in keep]
    arr = synth.hor_array(mm, rng.randint(3000, 14000), rng.choice([0.0, 0.02, 0.1, 0.25]), seed)
    cuts = sorted(rng.sample(range(1, len(arr)), rng.randint(0, 2)))
    rr = [arr[a:b] for a, b in zip([0] + cuts, cuts + [len(arr)])]
    return ["read%d" % i for i in range(len(rr))], rr, mn, mm


def one(rng, seed, real=False):
    if real:
        rn, rr, mn, mm = realistic(rng, seed)
        part = rng.choice([5000, 5000, 2000, 1000])
        overlap = rng.choice([500, 500, 300, 0])
    else:
        alphabet = rng.choice(["A", "AC", "ACG", "ACGT", "ACGT", "ACGTN"])
        nmon = (1, 8) if rng.random() < 0.8 else (20, 70)            # now and then a large set: many slots, slot groups
        rn, rr, mn, mm = synth.random_case(seed, alphabet=alphabet, n_monomers=nmon, mono_len=(1, 90), n_reads=(1, 4), read_len=(1, 1500))
        part = rng.choice([50, 137, 300, 700, 1000])
        overlap = rng.choice([0, 10, 40, part // 3, part - 1])
    tail = [str(rng.randint(1, 4)), str(part), str(overlap)]
    mode = rng.random()
    if mode < 0.35:
        tail += [str(x) for x in rng.choice(SCORINGS)]                                  # argc == 10: scores honoured
    elif mode < 0.6:
        tail += [str(x) for x in rng.choice(SCORINGS)] + [str(rng.choice([-1, 0, 3, 10, 25, 60, 500]))]   # argc == 11: ed_thr
    env = dict(os.environ)
    geom, lat = rng.choice(GEOMS), rng.choice(LATS)
    lmax = max(len(m) for m in mm)
    if geom:
        c, t, _ = map(int, geom.split(","))
        if c * t < lmax or len(mm) > 12:         # a forced geometry that cannot hold the set is a loud error, not a case
            geom = ""
    if "," in lat:
        c, t, w = lat.split(",")
        if int(c) * int(t) >= lmax and len(mm) <= 12:            # a forced cluster shape must be able to hold the set (else: loud error)
            env.update({"SD_LAT": "1", "SD_GEOM": "%s,%s,1" % (c, t), "SD_LAT_WARPS": w})
    else:
        if geom:
            env["SD_GEOM"] = geom
        if lat:
            env["SD_LAT"] = lat
    if rng.random() < 0.5:
        env["SD_CHUNK_BASES"] = str(rng.choice([1, 77, 500, 4000]))
    if rng.random() < 0.3:
        env["SD_DEVICES"] = str(rng.choice([2, 3, 4]))
    if rng.random() < 0.3:
        env["SD_FORCE_S32"] = "1"
    if rng.random() < 0.25 and "SD_LAT_WARPS" not in env:
        env["SD_GROUP_SLOTS"] = rng.choice(["1", "2", "4"])           # group sweep: a segment's slots spread over several CTAs
    knobs = {k: env[k] for k in ("SD_GEOM", "SD_LAT", "SD_LAT_WARPS", "SD_CHUNK_BASES", "SD_DEVICES", "SD_FORCE_S32", "SD_GROUP_SLOTS") if k in env}
    with tempfile.TemporaryDirectory() as td:
        rp, mp = os.path.join(td, "reads.fa"), os.path.join(td, "monomers.fa")
        open(rp, "w").write(fasta(rn, rr, rng.choice([0, 0, 60, 7])))
        open(mp, "w").write(fasta(mn, mm, rng.choice([0, 0, 50])))
        a = subprocess.run([REF, rp, mp] + tail, stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=td)
        b = subprocess.run([EMU, rp, mp] + tail, stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=td, env=env)
        berr = b"".join(ln for ln in b.stderr.splitlines(True) if not ln.startswith(b"[sd_b200]"))
        if b"Sanitizer" in b.stderr or b"runtime error" in b.stderr:
            print("SANITIZER REPORT seed=%d\n%s" % (seed, b.stderr.decode(errors="replace")[:3000]), flush=True)
        same = a.returncode == b.returncode and a.stdout == b.stdout and a.stderr == berr
        if not same:
            keep = tempfile.mkdtemp(prefix="fuzz_fail_%d_" % seed)
            for p in (rp, mp):
                os.replace(p, os.path.join(keep, os.path.basename(p)))
            open(os.path.join(keep, "ref.tsv"), "wb").write(a.stdout)
            open(os.path.join(keep, "ours.tsv"), "wb").write(b.stdout)
            open(os.path.join(keep, "ours.err"), "wb").write(b.stderr)
            print("MISMATCH seed=%d tail=%s knobs=%s status %d/%d kept in %s" % (seed, tail, knobs, a.returncode, b.returncode, keep), flush=True)
        return same, a.stdout.count(b"\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--realistic", action="store_true", help="DXZ1 monomers and noisy HOR arrays instead of small random strings")
    args = ap.parse_args()
    rng = random.Random(args.seed)
    bad = rows = 0
    for i in range(args.cases):
        ok, n = one(rng, args.seed * 100000 + i, args.realistic)
        bad += not ok
        rows += n
        if (i + 1) % 100 == 0:
            print("seed %d: %d cases, %d mismatches, %d reference rows" % (args.seed, i + 1, bad, rows), flush=True)
    print("seed %d done: %d cases, %d mismatches, %d reference rows" % (args.seed, args.cases, bad, rows))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
