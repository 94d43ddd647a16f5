#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu --durations=40 > gpurun_out/r2_check7.pytest 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_check7.pytest
timeout 200 python tools/lat_probe.py 19 a >> gpurun_out/r2_check7.txt 2>&1
timeout 200 python tools/lat_probe.py 50 a >> gpurun_out/r2_check7.txt 2>&1
