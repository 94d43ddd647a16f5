"""CPU suite for the host side of the product: the C ABI surface, segmentation / overlap resolution / FASTA
contract, and the kernel algebra run through the host emulator (libsd_emu.so -- the same per-lane functions the
sm_100a kernels execute) against the oracle.  No CUDA compute happens here."""
import ctypes
import os
import re

import numpy as np
import pytest

import cases
import sd_oracle
from stringdecomposer_b200 import _lib, synth, Decomposer, SdError, decompose_reads, segment_read, postprocess, RECORD_DTYPE


def test_cuda_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(cases.ROOT, "include", "sd_b200.h")).read()
    declared = set(re.findall(r"\b(sd_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.library_path("cuda"))
    for s in declared:
        assert hasattr(lib, s), s
    assert "sm_100a" in _lib.load_library("cuda").sd_version().decode()
    # the timing harness of bench.py's end-to-end leg: a loop around the C-ABI call, nothing else
    bl = ctypes.CDLL(os.path.join(os.path.dirname(_lib.library_path("cuda")), "libsd_bench.so"))
    assert hasattr(bl, "sd_bench_decompose")


def test_product_library_has_no_cpu_path():
    # without a GPU the CUDA library must refuse, not fall back
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(SdError) as e:
        Decomposer(["ACGT"], flavour="cuda")
    assert e.value.status == 2
    import subprocess
    p = subprocess.run([cases.DP_CUDA, os.path.join(cases.GOLDEN, "config1_read.fa"), os.path.join(cases.GOLDEN, "DXZ1_star_monomers.fa"),
                        "1", "5000", "500"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0 and p.stdout == b"" and b"no CUDA device" in p.stderr


@pytest.mark.parametrize("L,part,ov", [(1, 1000, 300), (299, 1000, 300), (300, 1000, 300), (1299, 1000, 300), (1300, 1000, 300),
                                       (1301, 1000, 300), (94871, 5000, 500), (5499, 5000, 500), (5500, 5000, 500), (5501, 5000, 500),
                                       (10000, 5000, 500), (700, 700, 50), (20000, 20000, 500), (123456, 777, 1000)])
def test_segmentation_matches_oracle(L, part, ov):
    assert segment_read(L, part, ov, flavour=cases.EMU_LIB) == sd_oracle.segment_read(L, part, ov)
    assert segment_read(L, part, ov, flavour="cuda") == sd_oracle.segment_read(L, part, ov)


def test_postprocess_matches_oracle():
    rng = np.random.default_rng(3)
    for trial in range(200):
        n = int(rng.integers(0, 40))
        starts = np.sort(rng.integers(0, 3000, n))
        recs = np.zeros(n, dtype=RECORD_DTYPE)
        recs["start"] = starts
        recs["end"] = starts + rng.integers(0, 400, n)
        recs["row"] = rng.integers(0, 24, n)
        recs["score"] = rng.integers(-50, 200, n)
        got = postprocess(recs, flavour=cases.EMU_LIB)
        want = sd_oracle.postprocess([(int(r["row"]), int(r["start"]), int(r["end"]), float(r["score"])) for r in recs])
        assert [(int(r["row"]), int(r["start"]), int(r["end"]), float(r["score"])) for r in got] == want


@pytest.mark.parametrize("case", cases.load_cases(), ids=lambda c: c["name"])
def test_emulated_kernels_match_reference_on_edge_cases(case):
    cases.check_case(cases.DP_EMU, case)


@pytest.mark.parametrize("geom", ["8,32,1", "16,16,1", "24,8,1", "24,8,2", "32,8,1", "48,4,2", "48,4,3", "20,10,1", "20,10,2", "12,16,1", "19,10,1", "19,10,3"])
@pytest.mark.parametrize("force32", ["0", "1"])
def test_emulated_kernels_every_geometry(geom, force32):
    # same answer whatever the lane layout, packed s16x2 and s32 (the emulator traps 16-bit overflow)
    picked = [c for c in cases.load_cases() if c["name"] in ("multi_read", "scoring_-3_-2_-4_2", "N_in_monomer", "dup_monomers_rev")]
    for case in picked:
        cases.check_case(cases.DP_EMU, case, env={"SD_GEOM": geom, "SD_FORCE_S32": force32})


@pytest.mark.parametrize("geom", ["8,1,4", "8,2,2", "16,1,1", "24,1,2", "32,2,1", "48,1,1", "8,4,3", "12,10,2", "20,2,5"])
def test_emulated_kernels_small_slots(geom):
    C, T, _ = map(int, geom.split(","))
    for seed in range(6):
        rn, rr, mn, mm = synth.random_case(200 + seed, mono_len=(1, C * T), read_len=(1, 400))
        want = sd_oracle.decompose_reads(rn, rr, mn, mm, part_size=150, overlap=40)
        os.environ["SD_GEOM"] = geom
        try:
            got = decompose_reads(rn, rr, mn, mm, part_size=150, overlap=40, flavour=cases.EMU_LIB)
        finally:
            del os.environ["SD_GEOM"]
        assert got == want


@pytest.mark.parametrize("case", cases.load_cases(), ids=lambda c: c["name"])
def test_emulated_deferred_jump_sweep_on_edge_cases(case):
    # the latency form of the recurrence (sweep_core.cuh: lat_*) forced on: same raw TSV, stderr and status
    cases.check_case(cases.DP_EMU, case, env={"SD_LAT": "1"})


@pytest.mark.parametrize("geom,warps", [("6,32,1", "4"), ("6,32,1", "1"), ("12,16,1", "3"), ("12,16,1", "1"), ("24,8,1", "0"), ("24,8,1", "1"),
                                        ("12,32,1", "2"), ("24,16,1", "1"), ("24,32,1", "2"), ("48,32,1", "1")])
@pytest.mark.parametrize("force32", ["0", "1"])
def test_emulated_deferred_jump_sweep_every_geometry(geom, warps, force32):
    # cluster shapes from one CTA per segment to one warp per CTA; the emulator also checks the windowed deletion carry
    # against the full prefix maximum and traps 16-bit overflow
    picked = [c for c in cases.load_cases() if c["name"] in ("multi_read", "scoring_-3_-2_-4_2", "N_in_monomer", "dup_monomers_rev")]
    for case in picked:
        cases.check_case(cases.DP_EMU, case, env={"SD_LAT": "1", "SD_GEOM": geom, "SD_LAT_WARPS": warps, "SD_FORCE_S32": force32})


@pytest.mark.parametrize("alphabet", ["A", "AT", "ACGT", "ACGTN"])
def test_emulated_deferred_jump_sweep_low_complexity(alphabet):
    # long runs without a matching symbol stretch the carry window (plan.cpp: scan_window) to its maximum
    for seed in range(5):
        rn, rr, mn, mm = synth.random_case(900 + seed, alphabet=alphabet, mono_len=(1, 190), read_len=(1, 700))
        for sc in ((-1, -1, -1, 1), (-1, -1, 1, -1), (-2, -3, 0, 0)):
            want = sd_oracle.decompose_reads(rn, rr, mn, mm, part_size=300, overlap=80, scoring=sc)
            os.environ["SD_LAT"] = "1"
            try:
                got = decompose_reads(rn, rr, mn, mm, part_size=300, overlap=80, scoring=sc, flavour=cases.EMU_LIB)
            finally:
                del os.environ["SD_LAT"]
            assert got == want


def test_emulated_config1_full_golden():
    st, out, err = sd_oracle.run_cli(cases.DP_EMU, os.path.join(cases.GOLDEN, "config1_read.fa"),
                                     os.path.join(cases.GOLDEN, "DXZ1_star_monomers.fa"))
    assert st == 0
    assert out == open(os.path.join(cases.GOLDEN, "config1_raw_default.tsv"), "rb").read()


@pytest.mark.parametrize("chunk", ["1", "700", "5000", "100000"])
def test_streamed_chunks_do_not_change_the_result(chunk):
    # run_files streams the reads file in chunks (reader -> device -> writer); a read may straddle chunks
    picked = [c for c in cases.load_cases() if c["name"] in ("multi_read", "len_10000_default", "part700_ov50", "blank_lines", "ed_thr_12")]
    assert len(picked) == 5
    for case in picked:
        cases.check_case(cases.DP_EMU, case, env={"SD_CHUNK_BASES": chunk})


def test_waves_do_not_change_the_result():
    rn, rr, mn, mm = synth.random_case(77, read_len=(2000, 3000), n_reads=(3, 3))
    want = sd_oracle.decompose_reads(rn, rr, mn, mm, part_size=300, overlap=100)
    os.environ["SD_WAVE_BYTES"] = "200000"
    try:
        got = decompose_reads(rn, rr, mn, mm, part_size=300, overlap=100, flavour=cases.EMU_LIB)
    finally:
        del os.environ["SD_WAVE_BYTES"]
    assert got == want


def test_score_range_switch():
    # wide penalties must leave the packed s16x2 domain and still be exact (the emulator traps overflow)
    names, mons = synth.load_dxz1()
    seg = synth.hor_array(mons, 1500, 0.05, seed=9)
    for sc, packed in (((-1, -1, -1, 1), 1), ((-2, -2, -3, 1), 1), ((-9, -30, -9, 2), 0), ((1, -1, -1, 1), 0)):
        d = Decomposer(mons, *sc, flavour=cases.EMU_LIB)
        recs, off = d.decompose([seg])
        assert d.stats()["packed"] == packed
        want = sd_oracle.align_segment(seg, mons, sc)
        assert [(int(r["row"]), int(r["start"]), int(r["end"]), float(r["score"])) for r in recs] == want
        d.close()


def test_unsupported_inputs_fail_loudly():
    with pytest.raises(SdError):
        Decomposer(["ACGU"], flavour=cases.EMU_LIB)
    d = Decomposer(["ACGT"], flavour=cases.EMU_LIB)
    with pytest.raises(SdError):
        d.decompose(["ACGTX"])
    with pytest.raises(SdError):
        d.decompose(["ACGT", ""])
    d.close()


def test_hw_distance_against_plain_dp():
    # the bit-vector infix edit distance of the --ed_thr pre-filter (main.cpp:128-133) vs the textbook DP
    from stringdecomposer_b200 import hw_distance
    rng = np.random.default_rng(11)
    for trial in range(150):
        al = list("ACGTN"[:int(rng.integers(1, 6))])
        m = int(rng.integers(1, 64 if trial % 3 else 400)); n = int(rng.integers(1, 500))
        pat = "".join(rng.choice(al, m)); txt = "".join(rng.choice(al, n))
        if trial % 4 == 0 and n > m:                      # plant a noisy copy so that small distances occur
            k = int(rng.integers(0, n - m))
            txt = txt[:k] + pat + txt[k + m:]
        prev = [0] * (n + 1)
        for i in range(1, m + 1):
            cur = [i] + [0] * n
            for j in range(1, n + 1):
                cur[j] = min(prev[j - 1] + (pat[i - 1] != txt[j - 1]), prev[j] + 1, cur[j - 1] + 1)
            prev = cur
        assert hw_distance(pat, txt, flavour=cases.EMU_LIB) == min(prev)
        assert hw_distance(pat, txt, flavour="cuda") == min(prev)


def test_ed_thr_filter_through_the_api():
    names, mons = synth.load_dxz1()
    seg = synth.hor_array(mons, 1200, 0.03, seed=3)
    d = Decomposer(mons, flavour=cases.EMU_LIB)
    base, _ = d.decompose([seg])
    d.set_ed_thr(10 ** 6)                                  # nothing filtered, but rows re-ordered by distance
    allr, _ = d.decompose([seg])
    d.set_ed_thr(0)                                        # only the closest row survives
    one, _ = d.decompose([seg])
    d.set_ed_thr(-1)
    again, _ = d.decompose([seg])
    assert (again == base).all()
    assert len(set(one["row"])) == 1
    assert allr["start"][0] == 0 and allr["end"][-1] == len(seg) - 1
    d.close()


def test_pinned_host_buffer_gives_the_same_records():
    # sd_host_alloc / sd_host_free: the text may live in page-locked memory; results do not depend on where it lives
    from stringdecomposer_b200._lib import HostBuffer
    _, segs, _, mons = synth.random_case(11, n_reads=(3, 3))
    dec = Decomposer(mons, flavour=cases.EMU_LIB)
    a = dec.decompose(segs)
    buf, off = HostBuffer.pack(segs, cases.EMU_LIB)
    assert len(buf) == sum(len(s) for s in segs)
    b = dec.decompose((buf, off))
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    dec.close()


def test_block_cache_policy():
    # csrc/block_cache.h (best fit within 4x, bounds per device and kind, flush, the off switch) on opaque pointers
    import subprocess
    p = subprocess.run([os.path.join(cases.HERE, "emu", "_build", "block_cache_test")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()


@pytest.mark.parametrize("N", [1, 5, 13, 32, 33, 100, 256, 1000, 2047])
def test_planner_covers_monomer_set_shapes(N):
    # from one row of one symbol to 2047 monomers and rows of 1536 bp: the planner must find a geometry (single CTA,
    # cluster or slot groups, s16x2 or s32) and the emulated kernels must reproduce the oracle through it
    import numpy as np
    rng = np.random.default_rng(N)
    al = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rs(n):
        return al[rng.integers(0, 4, n)].tobytes().decode()
    for L in (1, 7, 48, 171, 340, 700, 1536):
        if N * L > 200000:
            continue
        mons = [rs(int(rng.integers(max(1, L // 2), L + 1))) for _ in range(N)]
        mons[0] = rs(L)
        seg = rs(40)
        d = Decomposer(mons, flavour=cases.EMU_LIB)
        recs, off = d.decompose([seg])
        d.close()
        got = [(int(r["row"]), int(r["start"]), int(r["end"]), float(r["score"])) for r in recs]
        assert got == sd_oracle.align_segment(seg, mons, (-1, -1, -1, 1)), (N, L)
