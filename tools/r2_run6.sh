#!/bin/bash
set -u
out=gpurun_out/r2run6; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_concurrency.py tests/test_gpu_parity.py tests/test_gpu_handle_api.py -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
for k in 1 2 4 8; do OPAL_B200_SLICES=$k timeout 300 python tools/e2e_probe.py config3 2>&1 | grep "^call" | tail -2 | sed "s/^/slices=$k /" >> $out/e2e_slices.txt; done
cat $out/e2e_slices.txt
OPAL_B200_TRACE=1 timeout 300 python tools/e2e_probe.py config3 > $out/e2e_probe.txt 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --in-flight 8 --no-cpu-baseline --no-extras > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --workload config2 --steps 100 --warmup 5 --no-cpu-baseline > $out/bench_config2.json 2> $out/bench_config2.err
