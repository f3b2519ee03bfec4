#!/bin/bash
set -u
out=gpurun_out/r2run13; mkdir -p $out
timeout 700 python -m pytest tests -x -q -m gpu --timeout 120 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -6 $out/pytest.log
timeout 200 python bench.py --steps 2 --warmup 3 --shard-of 8 --no-cpu-baseline > $out/bench_shard8.json 2> $out/bench_shard8.err
for c in 8 32 8 32; do s=$(date +%s.%N); CUDA_DEVICE_MAX_CONNECTIONS=$c opal_b200/cli/opal_aligner_b200 -s tests/golden/cli/query.fasta tests/golden/cli/db_clean.fasta > /dev/null 2>&1; e=$(date +%s.%N); echo "conn=$c $(echo "$e - $s" | bc) s" >> $out/startup.txt; done
cat $out/startup.txt
