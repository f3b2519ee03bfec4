#!/bin/bash
# Drop-in call on an eighth of BASELINE configs[2] (what a rank of an 8-GPU run sees): slices and host threads.
set -u
out=gpurun_out/r2run14; mkdir -p $out
for s in 1 2 3 4; do
  OPAL_B200_SLICES=$s timeout 200 python bench.py --steps 2 --warmup 3 --shard-of 8 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > $out/shard8_slices$s.json
done
OPAL_B200_HOST_THREADS=2 timeout 200 python bench.py --steps 2 --warmup 3 --shard-of 8 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > $out/shard8_threads2.json
OPAL_B200_HOST_THREADS=2 OPAL_B200_SLICES=3 timeout 200 python bench.py --steps 2 --warmup 3 --shard-of 8 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > $out/shard8_threads2_slices3.json
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2run14/*.json')):
    try:
        b = json.loads(open(f).read())
        print(f.split('/')[-1], round(b['value']), round(b['e2e']['value']), round(b['e2e']['ms_per_step']))
    except Exception as e:
        print(f, 'failed', e)
P
