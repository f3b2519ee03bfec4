#!/bin/bash
# Drop-in call on a quarter / a half of BASELINE configs[2]: where do slices start to pay?
set -u
out=gpurun_out/r2run15; mkdir -p $out
for m in 4 2; do for s in 1 2 3; do
  OPAL_B200_SLICES=$s timeout 200 python bench.py --steps 2 --warmup 3 --shard-of $m --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > $out/shard${m}_slices$s.json
done; done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2run15/*.json')):
    try:
        b = json.loads(open(f).read())
        print(f.split('/')[-1], round(b['value']), round(b['e2e']['value']), round(b['e2e']['ms_per_step']), round(b['ms_per_step']))
    except Exception as e:
        print(f, 'failed', e)
P
