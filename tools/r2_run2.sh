#!/bin/bash
# Round-2 GPU run 2: parity suite (range tracking, multi-device, drop-in programs, concurrency), bench, sanitizer.
set -u
out=gpurun_out/r2run2; mkdir -p $out
timeout 1200 python -m pytest tests -x -q -m gpu > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -15 $out/pytest.log
timeout 600 python bench.py --steps 2 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 8 --no-cpu-baseline > $out/bench_shard8.json 2> $out/bench_shard8.err
bash tools/sanitize.sh > $out/sanitize.txt 2>&1
cp -r gpurun_out/sanitize $out/ 2>/dev/null
tail -12 $out/sanitize.txt
