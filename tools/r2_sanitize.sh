#!/bin/bash
# compute-sanitizer memcheck over the kernels that are new in round 2 (cluster sweep in several cluster shapes, --ed_thr
# filter kernels, gather, wave slots), on small golden cases through the dp binary.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; OUT=gpurun_out/r2_sanitize.txt; rm -f $OUT
python - <<'PY'
import json
cases = {c["name"]: c for c in json.load(open("tests/golden/edge_cases.json"))["cases"]}
for n in ("multi_read", "ed_thr_12", "short_monomers", "N_in_monomer"):
    open("/tmp/san_%s_reads.fa" % n, "w").write(cases[n]["reads_fa"]); open("/tmp/san_%s_mons.fa" % n, "w").write(cases[n]["monomers_fa"])
    open("/tmp/san_%s.argv" % n, "w").write(" ".join(cases[n]["argv_tail"])); open("/tmp/san_%s.out" % n, "w").write(cases[n]["stdout"])
PY
DP=stringdecomposer_b200/build/bin/dp
run() {  # label env... -- case
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  local c=$1
  env "${envs[@]}" compute-sanitizer --tool memcheck --error-exitcode 99 --log-file /tmp/san.log $DP /tmp/san_${c}_reads.fa /tmp/san_${c}_mons.fa $(cat /tmp/san_$c.argv) > /tmp/san.out 2> /tmp/san.err
  local rc=$?
  local same=no; cmp -s /tmp/san.out /tmp/san_$c.out && same=yes
  echo "$label case=$c rc=$rc output_identical=$same $(grep -E "ERROR SUMMARY" /tmp/san.log | tail -1)" >> $OUT
  grep -E "Invalid|out of bounds|misaligned|Race" /tmp/san.log | head -5 >> $OUT
}
run classic SD_LAT=0 -- multi_read
run cluster_6_32_ng3 SD_LAT=1 SD_GEOM=6,32,1 SD_LAT_WARPS=4 -- multi_read
run cluster_6_32_ng6 SD_LAT=1 SD_GEOM=6,32,1 SD_LAT_WARPS=2 -- multi_read
run cluster_12_16_ng2 SD_LAT=1 SD_GEOM=12,16,1 SD_LAT_WARPS=3 -- multi_read
run cluster_24_8_ng1 SD_LAT=1 SD_GEOM=24,8,1 -- multi_read
run cluster_s32 SD_LAT=1 SD_GEOM=12,16,1 SD_FORCE_S32=1 -- N_in_monomer
run filter_classic SD_LAT=0 -- ed_thr_12
run filter_cluster SD_LAT=1 -- ed_thr_12
run group SD_GROUP_SLOTS=4 -- short_monomers
run waves SD_WAVE_BYTES=300000 SD_CHUNK_BASES=3000 -- multi_read
cat $OUT
