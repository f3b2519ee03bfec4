#!/bin/bash
# Collects the evidence kept under profiles/ (run on the GPU box through gpurun; outputs land in gpurun_out/profiles/).
# Bench values are never taken under a profiler: the bench lines come from plain runs, ncu is used for shares,
# pipe utilisation and DRAM bytes only.  QUICK=1 skips the slow probes (step times, large-database batch).
set -u
out=gpurun_out/profiles; mkdir -p $out
python bench.py --steps 200 --warmup 10 2>/dev/null | tail -1 > $out/r1_bench_n1.json
python bench.py --impl reference --steps 20 --warmup 2 2>/dev/null | tail -1 > $out/r1_bench_reference.json
python bench.py --workload config3 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $out/r1_bench_n1_config3.json
QLENS=144,375,1000,2005,5478 MODES=SW,NW,HW,OV python tools/quick_perf.py config3 570000 > $out/r1_config3_table.txt 2>&1
python tools/quick_perf.py config2 2>&1 | grep "type=\|DPX" > $out/r1_config2_table.txt
OPAL_B200_TRACE=1 python tools/one_search.py config2 SW 1 2 2>&1 | grep "group" | tail -2 > $out/r1_config2_plan.txt
python tools/batch_probe.py config2 SW 1 32 > $out/r1_batch_probe.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/r1_launches_bench_config2.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 6 -c 2 -f -o $out/r1_search_kernel \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu -i $out/r1_search_kernel.ncu-rep --page raw 2>/dev/null | grep -E "search_kernel|Block Size|Grid Size|dram__bytes|gpu__dram_throughput|gpu__time_duration|l1tex__data_bank_conflicts_pipe_lsu_mem_shared|l1tex__t_sector_hit_rate|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|launch__occupancy|sm__inst_executed_pipe_(alu|fma|fmaheavy|lsu|xu|uniform)|smsp__inst_executed.sum |smsp__issue_active|sm__warps_active|sm__cycles_elapsed.sum |smsp__cycles_active.avg |sm__throughput|smsp__inst_executed_pipe_(alu|fma|lsu)|smsp__average_warps_issue_stalled_.*_per_issue_active" > $out/r1_ncu_search_kernel.txt
if [ "${QUICK:-0}" != "1" ]; then
  python tools/batch_probe.py 570000 SW 1 20 >> $out/r1_batch_probe.txt 2>&1
  KS=1,2,3 python tools/steptime_probe.py 2000 > $out/r1_steptime_probe.txt 2>&1
fi
ls -la $out
