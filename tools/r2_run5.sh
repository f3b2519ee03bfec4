#!/bin/bash
# Round-2 GPU run 5: pipelined slices on the drop-in path, column-wise profile build.
set -u
out=gpurun_out/r2run5; mkdir -p $out
timeout 1500 python -m pytest tests -x -q -m gpu > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 300 python bench.py --workload config2 --steps 100 --warmup 5 --no-cpu-baseline > $out/bench_config2.json 2> $out/bench_config2.err
timeout 600 python bench.py --steps 2 --warmup 3 --in-flight 8 --no-cpu-baseline --no-extras > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
OPAL_B200_SLICES=1 timeout 600 python bench.py --steps 2 --warmup 3 --in-flight 8 --no-cpu-baseline --no-extras > $out/bench_n1_noslices.json 2> $out/bench_n1_noslices.err
timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 8 --in-flight 12 --no-cpu-baseline --no-extras > $out/bench_shard8_f12.json 2> $out/bench_shard8_f12.err
OPAL_B200_TRACE=1 timeout 300 python tools/e2e_probe.py config3 > $out/e2e_probe.txt 2>&1
for k in 2 3 6 8; do OPAL_B200_SLICES=$k timeout 300 python tools/e2e_probe.py config3 2>&1 | grep "^call" | tail -2 | sed "s/^/slices=$k /" >> $out/e2e_slices.txt; done
cat $out/e2e_slices.txt
