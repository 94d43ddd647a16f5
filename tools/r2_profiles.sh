#!/bin/bash
# ncu evidence of round 2 (one GPU): launch list of the bench step, full captures of the kernels that changed
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r02_launches.log 2>&1
timeout 300 $NCU -k regex:sweep_kernel -s 2 -c 1 -f -o gpurun_out/r02_sweep_cfg2 python tools/lat_probe.py 400 0 > gpurun_out/r02_ncu.log 2>&1
timeout 300 $NCU -k regex:traceback_kernel -s 2 -c 1 -f -o gpurun_out/r02_traceback_cfg2 python tools/lat_probe.py 400 0 >> gpurun_out/r02_ncu.log 2>&1
timeout 300 $NCU -k regex:sweep_lat -s 2 -c 1 -f -o gpurun_out/r02_lat_12_16_50seg python tools/lat_probe.py 50 a >> gpurun_out/r02_ncu.log 2>&1
timeout 300 $NCU -k regex:sweep_lat -s 2 -c 1 -f -o gpurun_out/r02_lat_6_32_19seg python tools/lat_probe.py 19 a >> gpurun_out/r02_ncu.log 2>&1
cat > /tmp/edthr.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np
from stringdecomposer_b200 import synth, Decomposer
from stringdecomposer_b200.hostpipe import segment_reads
rn, reads, mn, mons = synth.config2()
segs, _ = segment_reads(reads, 5000, 500)
d = Decomposer(mons, devices=[0]); d.set_ed_thr(40)
for _ in range(3): d.decompose(segs)
print(d.stats())
PY
timeout 300 $NCU -k regex:"hw_distance_rows|filter_rank" -s 2 -c 2 -f -o gpurun_out/r02_filter python /tmp/edthr.py >> gpurun_out/r02_ncu.log 2>&1
# saturated launch (4000 segments of ONT-like reads): plain timing with and without the windowed carry, then a capture
timeout 300 python tools/throughput_probe.py 200 "" > gpurun_out/r02_saturated.txt 2>&1
SD_FULL_SCAN=1 timeout 300 python tools/throughput_probe.py 200 "" >> gpurun_out/r02_saturated.txt 2>&1
PROBE_LL=0 timeout 300 $NCU -k regex:sweep_kernel -s 1 -c 1 -f -o gpurun_out/r02_sweep_saturated python tools/throughput_probe.py 200 "" >> gpurun_out/r02_ncu.log 2>&1
cat gpurun_out/r02_saturated.txt
ls -la gpurun_out/*.ncu-rep
