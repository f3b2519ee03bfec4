#!/bin/bash
# Two real GPUs: the several-devices tests on distinct devices, then the strong-scaling bench under torchrun (N = 2).
set -u
out=gpurun_out/r2scale; mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $out/smi_n2.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi_device.py -x -q -m gpu --timeout 120 > $out/pytest_multi_device_n2.log 2>&1; echo "pytest rc=$?" >> $out/pytest_multi_device_n2.log
tail -3 $out/pytest_multi_device_n2.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 3 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err
echo "rc=$?"; tail -c 1500 $out/bench_n2.json; tail -5 $out/bench_n2.err
