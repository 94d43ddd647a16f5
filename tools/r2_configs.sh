#!/bin/bash
# Kept evidence for BASELINE configs 1, 3, 4, 5 through the drop-in `dp` binary: MD5 of the raw TSV next to the unmodified
# reference's (configs 1, 3 in full, config 4 on its first 3,000 reads) and the SD_PERF_JSON line of every run.
# Writes gpurun_out/r2_configs.log and gpurun_out/r2_configs_perf.jsonl
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/r2_configs.log; PERF=$PWD/gpurun_out/r2_configs_perf.jsonl; rm -f $LOG $PERF
DP=stringdecomposer_b200/build/bin/dp; REF=oracle/_ref/dp; NP=$(nproc)
G=tests/golden
stamp() { date +%s.%N; }
run() {  # label reads monomers part overlap [scores...]
  local label=$1; shift
  local t0=$(stamp)
  SD_PERF_JSON=$PERF SD_VERBOSE=1 $DP "$@" > /tmp/ours.tsv 2> /tmp/ours.err
  local rc=$? t1=$(stamp)
  echo "$label: ours rc=$rc wall $(python -c "print('%.3f' % ($t1 - $t0))") s lines $(wc -l < /tmp/ours.tsv) md5 $(md5sum < /tmp/ours.tsv | cut -d' ' -f1)" >> $LOG
  grep "sd_b200\]" /tmp/ours.err | tail -1 >> $LOG
}
ref() {
  local label=$1; shift
  local t0=$(stamp)
  $REF "$@" > /tmp/ref.tsv 2> /dev/null
  local rc=$? t1=$(stamp)
  echo "$label: reference (-t $NP) rc=$rc wall $(python -c "print('%.3f' % ($t1 - $t0))") s lines $(wc -l < /tmp/ref.tsv) md5 $(md5sum < /tmp/ref.tsv | cut -d' ' -f1)" >> $LOG
  cmp -s /tmp/ours.tsv /tmp/ref.tsv && echo "$label: IDENTICAL to the reference binary" >> $LOG || echo "$label: DIFFERENT" >> $LOG
}
echo "host: $NP cores; $(nvidia-smi -L | head -1)" >> $LOG
# config 1: the shipped read (expected MD5 of SURVEY 8c: 3acf5a26a9b6006ec5573f47d517e232)
run config1 $G/config1_read.fa $G/DXZ1_star_monomers.fa 1 5000 500
run config1_again $G/config1_read.fa $G/DXZ1_star_monomers.fa 1 5000 500
ref config1 $G/config1_read.fa $G/DXZ1_star_monomers.fa $NP 5000 500
python - <<PY
import sys; sys.path.insert(0, ".")
from stringdecomposer_b200 import synth
rn, r, mn, m = synth.config3(n_reads=2000, read_len=100_000)
synth.write_fasta("/tmp/c3.fa", rn, r, width=80); synth.write_fasta("/tmp/mons.fa", mn, m)
rn, r, mn, m = synth.config4(n_reads=20000, read_len=15_000)
synth.write_fasta("/tmp/c4.fa", rn, r, width=80); synth.write_fasta("/tmp/c4_3000.fa", rn[:3000], r[:3000], width=80)
rn, r, mn, m = synth.config5()
synth.write_fasta("/tmp/c5.fa", rn, r, width=80); synth.write_fasta("/tmp/c5_mons.fa", mn, m)
PY
run config3_full /tmp/c3.fa /tmp/mons.fa 1 5000 500
run config3_full_again /tmp/c3.fa /tmp/mons.fa 1 5000 500
ref config3_full /tmp/c3.fa /tmp/mons.fa $NP 5000 500
run config4_full /tmp/c4.fa /tmp/mons.fa 1 5000 500 -2 -2 -3 1
run config4_3000 /tmp/c4_3000.fa /tmp/mons.fa 1 5000 500 -2 -2 -3 1
ref config4_3000 /tmp/c4_3000.fa /tmp/mons.fa $NP 5000 500 -2 -2 -3 1
run config5_full /tmp/c5.fa /tmp/c5_mons.fa 1 20000 500
run config5_full_again /tmp/c5.fa /tmp/c5_mons.fa 1 20000 500
cat $LOG
