#!/bin/bash
# kernel times of the deferred-jump (cluster) sweep against the classic one for a range of segment counts and cluster
# shapes, plus its per-phase cycle counters (SD_LAT_TIMING=1)
mkdir -p gpurun_out
rm -f gpurun_out/r2_lat_timing.txt
for spec in 1:6:32:4 1:12:16:3; do
  SD_LAT_TIMING=1 timeout 200 python tools/lat_probe.py 19 $spec 2>&1 | tail -2 >> gpurun_out/r2_lat_timing.txt
done
for n in 19 50 75 100 125 148; do
  timeout 300 python tools/lat_probe.py $n a 0 1:6:32:4 1:6:32:6 1:12:16:3 1:12:16:6 1:24:8:3 >> gpurun_out/r2_lat_timing.txt 2>&1
done
