#!/bin/bash
set -u
out=gpurun_out/r2run11; mkdir -p $out
timeout 600 python -m pytest tests -x -q -m gpu --timeout 120 --durations=8 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -16 $out/pytest.log
timeout 200 python bench.py --steps 2 --warmup 3 --shard-of 8 --no-cpu-baseline > $out/bench_shard8.json 2> $out/bench_shard8.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
timeout 120 python bench.py --workload config2 --steps 100 --warmup 5 --no-cpu-baseline > $out/bench_config2.json 2> $out/bench_config2.err
