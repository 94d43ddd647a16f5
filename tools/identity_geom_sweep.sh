#!/bin/bash
# identity kernel: lanes per pair x rows per lane (SD_NW_GEOM=L,R), kernel ms for the plain and the collapsed batch
cd "$(dirname "$0")/.."
for g in 8,24 8,16 16,12 32,6; do echo "geom $g"; SD_NW_GEOM=$g python tools/identity_probe.py 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['plain']['kernel_ms'], d['homopolymer']['kernel_ms'])"; done
