#!/bin/bash
set -u
out=gpurun_out/r2run7; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_concurrency.py tests/test_gpu_edge_cases.py -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
for k in 1 3 4 6; do OPAL_B200_SLICES=$k timeout 300 python tools/e2e_probe.py config3 2>&1 | grep "^call" | tail -2 | sed "s/^/slices=$k /" >> $out/e2e_slices.txt; done
cat $out/e2e_slices.txt
OPAL_B200_TRACE=1 timeout 300 python tools/e2e_probe.py config3 > $out/e2e_probe.txt 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --in-flight 8 --no-cpu-baseline --no-extras > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
# where does an eighth of the database lose time?  hardware queues, and the same shard without its heavy tail
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 8 --in-flight 12 --no-cpu-baseline --no-extras > $out/bench_shard8_conn32.json 2> $out/bench_shard8_conn32.err
timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 8 --in-flight 12 --no-cpu-baseline --no-extras > $out/bench_shard8.json 2> $out/bench_shard8.err
timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 8 --in-flight 16 --no-cpu-baseline --no-extras > $out/bench_shard8_f16.json 2> $out/bench_shard8_f16.err
