"""Hardware check of the process-wide block cache (ctypes only, no torch): results against the oracle with the cache on
and off, config-1 golden MD5 through sd_run_files, and the cost of an engine's lifetime with and without it."""
import hashlib, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
t00 = time.time()
import sd_oracle
from stringdecomposer_b200 import Decomposer, decompose_reads, synth, _lib
FL = os.environ.get("SD_CHECK_FLAVOUR", "cuda")      # a path = the host emulator (dry run of this script without a GPU)

def life(seed):
    rn, rr, mn, mm = synth.random_case(seed, n_monomers=(1, 6), read_len=(50, 900))
    want = sd_oracle.decompose_reads(rn, rr, mn, mm, part_size=300, overlap=60)
    t0 = time.time()
    got = decompose_reads(rn, rr, mn, mm, part_size=300, overlap=60, flavour=FL)
    return got == want, time.time() - t0

def golden():
    g = os.path.join(ROOT, "tests", "golden")
    with tempfile.TemporaryFile("w+b") as fo, tempfile.TemporaryFile("w+b") as fe:
        st = _lib.run_files(os.path.join(g, "config1_read.fa"), os.path.join(g, "DXZ1_star_monomers.fa"), 1, 5000, 500, (-1, -1, -1, 1), -1,
                            out_fd=fo.fileno(), err_fd=fe.fileno(), flavour=FL)
        fo.seek(0)
        return st, hashlib.md5(fo.read()).hexdigest()

ok, t = life(0); print("first engine (context, module load): ok=%s %.3f s (imports %.1f s)" % (ok, t, time.time() - t00 - t), flush=True)
for mode in ("1", "0", "1", "0"):
    os.environ["SD_NO_BUFFER_CACHE"] = mode
    res = [life(100 + s) for s in range(8)]
    print("cache %s: %d/%d identical to the oracle, engine lifetime median %.1f ms max %.1f ms" % (
        "off" if mode == "1" else "on ", sum(r[0] for r in res), len(res), 1e3 * sorted(r[1] for r in res)[len(res) // 2], 1e3 * max(r[1] for r in res)), flush=True)
    print("   config 1 through sd_run_files:", golden(), "(want 3acf5a26a9b6006ec5573f47d517e232)", flush=True)
print("total %.1f s" % (time.time() - t00))
