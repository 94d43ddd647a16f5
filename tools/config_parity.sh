#!/bin/bash
# Full-output parity of the dp binary against the UNMODIFIED reference binary (oracle/_ref/dp) on a prefix of BASELINE
# config 3 or 4:   tools/config_parity.sh <3|4> <n_reads>
cd "$(dirname "$0")/.."
CFG=${1:-4}; N=${2:-2000}
python - <<PY
import sys; sys.path.insert(0, ".")
from stringdecomposer_b200 import synth
rn, r, mn, m = (synth.config3(n_reads=$N, read_len=100_000) if $CFG == 3 else synth.config4(n_reads=$N, read_len=15_000))
synth.write_fasta("/tmp/reads.fa", rn, r, width=80); synth.write_fasta("/tmp/mons.fa", mn, m)
PY
SC=""; if [ "$CFG" = 4 ]; then SC="-2 -2 -3 1"; fi
echo "ours:"; time (SD_VERBOSE=1 stringdecomposer_b200/build/bin/dp /tmp/reads.fa /tmp/mons.fa 1 5000 500 $SC > /tmp/ours.tsv 2>/tmp/ours.err); grep sd_b200 /tmp/ours.err | tail -1 | cut -c1-220
echo "reference -t $(nproc):"; time oracle/_ref/dp /tmp/reads.fa /tmp/mons.fa $(nproc) 5000 500 $SC > /tmp/ref.tsv 2>/dev/null
wc -l /tmp/ours.tsv /tmp/ref.tsv; cmp /tmp/ours.tsv /tmp/ref.tsv && echo IDENTICAL
