#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/throughput_probe.py 200 "" > gpurun_out/r2_check6.sat 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_check6.bench 2> gpurun_out/r2_check6.bench.err
timeout 2400 python -m pytest tests/ -x -q -m gpu --durations=25 > gpurun_out/r2_check6.pytest 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_check6.pytest
