#!/bin/bash
# Round-2 GPU run 4: layout cache + fused records + full-size parity; plans and launch list of an eighth of the database.
set -u
out=gpurun_out/r2run4; mkdir -p $out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -25 $out/pytest.log
timeout 600 python bench.py --steps 2 --warmup 3 --in-flight 8 > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 8 --in-flight 12 --no-cpu-baseline --no-extras > $out/bench_shard8_f12.json 2> $out/bench_shard8_f12.err
OPAL_B200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 3 --shard-of 8 --in-flight 12 --no-cpu-baseline --no-extras 2>&1 | grep "group type" | sort | uniq -c | sort -rn > $out/plans_shard8.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_shard8.csv \
    python bench.py --steps 1 --warmup 3 --shard-of 8 --in-flight 12 --no-cpu-baseline --no-extras > /dev/null 2>&1
OPAL_B200_TRACE=1 timeout 300 python tools/e2e_probe.py config3 > $out/e2e_probe.txt 2>&1
ls $out
