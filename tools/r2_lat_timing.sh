#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_lat_timing.txt
SD_LAT_TIMING=1 timeout 200 python tools/lat_probe.py 50 1:6:32:4 2>&1 | tail -2 >> gpurun_out/r2_lat_timing.txt
for n in 19 50 100; do
  timeout 300 python tools/lat_probe.py $n 0 1:6:32:4 1:12:16:3 1:24:8:3 2:12:16:3 2:24:8:3 >> gpurun_out/r2_lat_timing.txt 2>&1
done
for n in 148 296; do
  timeout 300 python tools/lat_probe.py $n 0 0:19:10::1 0:24:8::1 2:24:8:3 2:19:10:4 2:12:16:6 2:12:16:3 1:12:16:3 2:6:32:12 >> gpurun_out/r2_lat_timing.txt 2>&1
done
