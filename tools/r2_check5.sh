#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_check5.pytest 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_check5.pytest
timeout 600 python bench.py > gpurun_out/r2_check5.bench 2> gpurun_out/r2_check5.bench.err
echo "bench rc=$?" >> gpurun_out/r2_check5.bench.err
