#!/bin/bash
# Round-2 GPU run 3: NW pad fix + finer strip heights + device top-k; batch order / in-flight sweep on an eighth of the database; e2e phases.
set -u
out=gpurun_out/r2run3; mkdir -p $out
timeout 1200 python -m pytest tests -x -q -m gpu > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 600 python bench.py --steps 2 --warmup 3 --in-flight 4 > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
for f in 4 8 12; do for o in desc interleave; do
  timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 8 --in-flight $f --order $o --no-cpu-baseline --no-extras > $out/bench_shard8_f${f}_$o.json 2> $out/bench_shard8_f${f}_$o.err
done; done
timeout 300 python bench.py --steps 2 --warmup 3 --shard-of 8 --no-cpu-baseline > $out/bench_shard8_extras.json 2> $out/bench_shard8_extras.err
timeout 300 python bench.py --steps 2 --warmup 3 --shard-of 2 --in-flight 8 --no-cpu-baseline --no-extras > $out/bench_shard2_f8.json 2> $out/bench_shard2_f8.err
OPAL_B200_TRACE=1 timeout 300 python tools/e2e_probe.py config3 > $out/e2e_probe.txt 2>&1
ls $out
