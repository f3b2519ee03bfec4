#!/bin/bash
set -u
out=gpurun_out/r2run9; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_chained.py -x -q > $out/pytest_chain.log 2>&1; echo "pytest chain rc=$?" >> $out/pytest_chain.log
tail -15 $out/pytest_chain.log
timeout 1500 python -m pytest tests -x -q -m gpu > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 300 python bench.py --steps 2 --warmup 3 --shard-of 8 --no-cpu-baseline > $out/bench_shard8.json 2> $out/bench_shard8.err
OPAL_B200_NO_CHAIN=1 timeout 300 python bench.py --steps 2 --warmup 3 --shard-of 8 --no-cpu-baseline --no-extras > $out/bench_shard8_nochain.json 2> $out/bench_shard8_nochain.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
OPAL_B200_TRACE=1 QLEN=5478 timeout 300 python tools/one_search.py 570000 HW 1 2 2>&1 | grep -E "group type|GCUPS" | tail -6 > $out/plan_hw5478.txt
cat $out/plan_hw5478.txt
