#!/bin/bash
# config 3 / config 4 through the dp binary at several chunk sizes: wall, SD_PROFILE lines, MD5 (must not depend on the chunk size)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/r2_chunks.log; rm -f $LOG
DP=stringdecomposer_b200/build/bin/dp
python - <<PY
import sys; sys.path.insert(0, ".")
from stringdecomposer_b200 import synth
rn, r, mn, m = synth.config3(n_reads=2000, read_len=100_000)
synth.write_fasta("/tmp/c3.fa", rn, r, width=80); synth.write_fasta("/tmp/mons.fa", mn, m)
rn, r, mn, m = synth.config4(n_reads=20000, read_len=15_000)
synth.write_fasta("/tmp/c4.fa", rn, r, width=80)
PY
stamp() { date +%s.%N; }
run() {  # label chunk args...
  local label=$1 chunk=$2; shift; shift
  local t0=$(stamp)
  SD_CHUNK_BASES=$chunk SD_PROFILE=1 SD_VERBOSE=1 $DP "$@" > /tmp/ours.tsv 2> /tmp/ours.err
  local rc=$? t1=$(stamp)
  echo "$label chunk=$chunk: rc=$rc wall $(python -c "print('%.3f' % ($t1 - $t0))") s md5 $(md5sum < /tmp/ours.tsv | cut -d' ' -f1)" >> $LOG
  grep "run_files\|sd_b200\] devices" /tmp/ours.err | tail -2 >> $LOG
}
run warmup 33554432 tests/golden/config1_read.fa tests/golden/DXZ1_star_monomers.fa 1 5000 500
for c in ${CHUNKS3:-268435456 67108864 33554432 16777216 8388608}; do run config3 $c /tmp/c3.fa /tmp/mons.fa 1 5000 500; done
for c in ${CHUNKS4:-536870912 67108864 33554432 16777216}; do run config4 $c /tmp/c4.fa /tmp/mons.fa 1 5000 500 -2 -2 -3 1; done
./tools/probes/ctx_time >> $LOG 2>&1
cat $LOG
