#!/bin/bash
# Round-2 GPU run 1: parity suite, the configs[2] bench at N=1, the per-GPU share of an 8-way deal, ncu of the bulk kernels.
set -u
out=gpurun_out/r2run1; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 600 python bench.py --steps 2 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
for f in 1 4 8; do
  timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 8 --in-flight $f --no-cpu-baseline --no-extras > $out/bench_shard8_f$f.json 2> $out/bench_shard8_f$f.err
done
timeout 300 python bench.py --steps 2 --warmup 3 --shard-of 8 --no-cpu-baseline > $out/bench_shard8_extras.json 2> $out/bench_shard8_extras.err
timeout 300 python bench.py --workload config2 --steps 50 --warmup 5 > $out/bench_config2.json 2> $out/bench_config2.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
# ncu: launch list of one sweep on the 1/8 shard, then full captures of the bulk launches of one SW and one NW search on the whole database
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_shard8.csv \
    python bench.py --steps 1 --warmup 3 --shard-of 8 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 2 -f -o $out/ncu_sw513 \
    python tools/one_search.py 570000 SW 1 1 > $out/ncu_sw513.log 2>&1
QLEN=2005 timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 3 -f -o $out/ncu_nw2005 \
    python tools/one_search.py 570000 NW 1 1 > $out/ncu_nw2005.log 2>&1
for r in sw513 nw2005; do
  ncu -i $out/ncu_$r.ncu-rep --page raw > $out/ncu_$r.raw.txt 2>&1
  ncu -i $out/ncu_$r.ncu-rep --page details > $out/ncu_$r.details.txt 2>&1
done
ls -la $out
