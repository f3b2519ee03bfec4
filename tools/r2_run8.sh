#!/bin/bash
set -u
out=gpurun_out/r2run8; mkdir -p $out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
for m in 2 4 8; do
timeout 300 python bench.py --steps 3 --warmup 3 --shard-of $m --no-cpu-baseline --no-extras > $out/bench_shard$m.json 2> $out/bench_shard$m.err
done
timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 8 --in-flight 16 --no-cpu-baseline --no-extras > $out/bench_shard8_f16.json 2> $out/bench_shard8_f16.err
timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 8 --in-flight 8 --no-cpu-baseline --no-extras > $out/bench_shard8_f8.json 2> $out/bench_shard8_f8.err
timeout 300 python bench.py --workload config2 --steps 100 --warmup 5 --no-cpu-baseline > $out/bench_config2.json 2> $out/bench_config2.err
