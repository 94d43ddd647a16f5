#!/bin/bash
# where does the wall time of the drop-in `dp` process go on the config-2 contig?
mkdir -p gpurun_out; out=gpurun_out/r2_startup.txt; rm -f $out
python - <<'PY'
import sys; sys.path.insert(0, '.')
from stringdecomposer_b200 import synth
rn, reads, mn, mons = synth.config2()
synth.write_fasta('/tmp/c2_reads.fa', rn, reads); synth.write_fasta('/tmp/c2_mons.fa', mn, mons)
PY
for i in 1 2 3; do tools/probes/ctx_time >> $out 2>&1; done
DP=stringdecomposer_b200/build/bin/dp
for i in 1 2 3; do
  s=$(date +%s.%N); SD_PROFILE=1 $DP /tmp/c2_reads.fa /tmp/c2_mons.fa 1 5000 500 > /tmp/o.tsv 2>> $out; e=$(date +%s.%N); echo "dp wall $(python -c "print('%.3f' % ($e - $s))") s" >> $out
done
for i in 1 2; do
  s=$(date +%s.%N); CUDA_MODULE_LOADING=EAGER $DP /tmp/c2_reads.fa /tmp/c2_mons.fa 1 5000 500 > /tmp/o.tsv 2>/dev/null; e=$(date +%s.%N); echo "dp EAGER wall $(python -c "print('%.3f' % ($e - $s))") s" >> $out
done
s=$(date +%s.%N); oracle/_ref/dp /tmp/c2_reads.fa /tmp/c2_mons.fa 16 5000 500 > /tmp/o_ref.tsv 2>/dev/null; e=$(date +%s.%N); echo "ref dp -t 16 wall $(python -c "print('%.3f' % ($e - $s))") s" >> $out
cmp /tmp/o.tsv /tmp/o_ref.tsv && echo "outputs identical" >> $out
ls -la stringdecomposer_b200/libsd_b200.so >> $out
