#!/bin/bash
# Four real GPUs: the strong-scaling bench under torchrun (N = 4), final build of the round.
set -u
out=gpurun_out/r2scale; mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $out/smi_n4.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 4 --steps 3 --warmup 3 > $out/bench_n4.json 2> $out/bench_n4.err
echo "rc=$?"; python -c "
import json
b=json.loads(open('$out/bench_n4.json').read().strip().splitlines()[-1])
print(b['value'], b['ms_per_step'], b['e2e']['value'], b['roofline']['frac'], b.get('one_process_multi_gpu'))"; tail -3 $out/bench_n4.err
