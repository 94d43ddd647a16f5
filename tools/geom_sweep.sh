#!/bin/bash
# config-2 sweep time for a list of launch geometries C,T,NS (bench.py --geom): value GCUPS, sweep ms
cd "$(dirname "$0")/.."
for g in "$@"; do
  python bench.py --geom $g --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$g', round(d['value'],1), 'GCUPS  sweep', round(d['roofline']['kernel_ms'],3), 'ms  tb', round(d['roofline']['traceback_ms'],3), 'ms  NT', d['config']['geometry']['NT'])"
done
