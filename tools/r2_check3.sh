#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_check3.txt
timeout 600 python -m pytest tests/test_gpu.py -x -q -m gpu -k "deferred or forced or every_geometry or fuzz or small_and_large or group_mode" > gpurun_out/r2_check3.pytest 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_check3.pytest
for n in 19 50 100 148 200 400; do
  timeout 300 python tools/lat_probe.py $n a 0 >> gpurun_out/r2_check3.txt 2>&1
done
SD_FULL_SCAN=1 timeout 300 python tools/lat_probe.py 400 0 >> gpurun_out/r2_check3.txt 2>&1
timeout 300 python tools/lat_probe.py 400 0:24:8::1 0:12:16::1 0:20:10::1 0:16:12::1 >> gpurun_out/r2_check3.txt 2>&1
timeout 300 python tools/throughput_probe.py 40 "" "24,8,4" "19,10,4" "24,8,3" >> gpurun_out/r2_check3.txt 2>&1
