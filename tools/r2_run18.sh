#!/bin/bash
# Drop-in call on the whole database (N = 1): how many slices?
set -u
out=gpurun_out/r2run18; mkdir -p $out
for s in 2 3 6 8; do
  OPAL_B200_SLICES=$s timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > $out/n1_slices$s.json
  python -c "import json;b=json.load(open('$out/n1_slices$s.json'));print('slices $s', round(b['value']), round(b['e2e']['value']), round(b['e2e']['ms_per_step']))"
done
