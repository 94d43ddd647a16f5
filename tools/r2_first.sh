#!/bin/bash
# first GPU pass of round 2: deferred-jump sweep parity + calibration sweep (writes gpurun_out/r2_first.*)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_first.gpus 2>&1
timeout 900 python -m pytest tests/test_gpu.py -x -q -m gpu -k "deferred or forced or native or config1_golden" > gpurun_out/r2_first.pytest 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_first.pytest
for n in 19 50 100 200 400; do
  timeout 300 python tools/lat_probe.py $n 0 1:6:32:4 1:6:32:2 1:6:32:1 1:12:16:3 1:12:16:2 1:12:16:1 1:12:16:6 1:24:8:3 1:24:8:1 1:6:32:12 a >> gpurun_out/r2_first.probe 2>&1
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_first.bench 2> gpurun_out/r2_first.bench.err
echo "bench rc=$?" >> gpurun_out/r2_first.bench.err
