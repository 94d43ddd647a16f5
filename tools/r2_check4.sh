#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py tests/test_identity.py -x -q -m gpu -k "not every_geometry and not edge_cases_through and not deferred_jump_sweep_every" > gpurun_out/r2_check4.pytest 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_check4.pytest
SD_PROFILE=1 timeout 200 python tools/e2e_probe.py > gpurun_out/r2_check4.e2e 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_check4.bench 2> gpurun_out/r2_check4.bench.err
