#!/bin/bash
# end-to-end wall time of the dp binary on a config-3-like input (N reads x 100 kb), ours vs the reference
cd "$(dirname "$0")/.."
N=${1:-200}
python - <<PY
import sys; sys.path.insert(0, ".")
from stringdecomposer_b200 import synth
rn, r, mn, m = synth.config3(n_reads=$N, read_len=100_000)
synth.write_fasta("/tmp/reads.fa", rn, r, width=80); synth.write_fasta("/tmp/mons.fa", mn, m)
PY
ls -la /tmp/reads.fa
for i in 1 2; do echo "ours:"; time (SD_VERBOSE=1 SD_PROFILE=1 stringdecomposer_b200/build/bin/dp /tmp/reads.fa /tmp/mons.fa 1 5000 500 > /tmp/ours.tsv 2>/tmp/ours.err); grep sd_b200 /tmp/ours.err | tail -3 | cut -c1-250; done
grep -c . /tmp/ours.tsv
if [ -n "$2" ]; then echo "reference -t $(nproc):"; time oracle/_ref/dp /tmp/reads.fa /tmp/mons.fa $(nproc) 5000 500 > /tmp/ref.tsv 2>/dev/null; cmp /tmp/ours.tsv /tmp/ref.tsv && echo IDENTICAL; fi
