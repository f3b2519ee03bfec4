#!/bin/bash
# Strong scaling on real GPUs: one process per GPU under torchrun, BASELINE configs[2] dealt over the ranks.
# usage (on the GPU box): bash tools/r2_scale.sh N
set -u
N=${1:-8}
out=gpurun_out/r2scale; mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $out/smi_n$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 > $out/bench_n$N.json 2> $out/bench_n$N.err
echo "rc=$?"; tail -c 600 $out/bench_n$N.json; tail -5 $out/bench_n$N.err
