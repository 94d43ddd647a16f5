#!/bin/bash
# end-to-end leg: bench line + where the host time of one sd_decompose call goes
mkdir -p gpurun_out
timeout 800 python bench.py > gpurun_out/r2_e2e.json 2> gpurun_out/r2_e2e.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_e2e.json')); print(d['value'], d['ms_per_step'], json.dumps(d['e2e']))"
timeout 300 python -m pytest tests/test_gpu.py -x -q -k "pinned or c_abi or golden" 2>&1 | tail -2
SD_PROFILE=1 python - 2>&1 <<'PY' | tail -8
import bench
from stringdecomposer_b200 import Decomposer
from stringdecomposer_b200._lib import HostBuffer
rnames, reads, mnames, mons, segs = bench.workload(0)
dec = Decomposer(mons, *bench.SCORING, devices=[0])
packed = bench.pack_segments(segs)
pinned = HostBuffer.pack(packed)
for i in range(4): dec.decompose(packed)
print("--- pinned")
import sys; sys.stdout.flush()
for i in range(4): dec.decompose(pinned)
PY
