#!/bin/bash
# 8-GPU box: multi-GPU parity tests, the strong-scaling bench lines and the host-side profile of sd_decompose over 8 devices
mkdir -p gpurun_out
NLIST="${NLIST:-2 4 8}" bash tools/r2_multigpu.sh 8
tail -3 gpurun_out/r2_multigpu_pytest.txt
for n in ${NLIST:-2 4 8}; do python -c "
import json; d=json.load(open('gpurun_out/r2_multigpu_bench_n$n.json')); print($n, d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e'].get('python_caller_ms_per_step'))"; done
SD_PROFILE=1 python - > gpurun_out/r2_e2e8_profile.txt 2>&1 <<'PY'
import bench
from stringdecomposer_b200 import Decomposer
from stringdecomposer_b200._lib import HostBuffer
rnames, reads, mnames, mons, segs = bench.workload(0)
dec = Decomposer(mons, *bench.SCORING, devices=list(range(8)))
pinned = HostBuffer.pack(bench.pack_segments(segs))
for i in range(5): dec.decompose(pinned)
PY
tail -9 gpurun_out/r2_e2e8_profile.txt
