#!/bin/bash
# Collects the round-2 evidence kept under profiles/ (run on the GPU box through gpurun; outputs land in
# gpurun_out/profiles_r2/ and the summaries are copied into profiles/ by hand).  Bench values are never taken under a
# profiler: the bench lines come from plain runs, ncu is used for shares, pipe utilisation and DRAM bytes only.
set -u
out=gpurun_out/profiles_r2; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/r2_smi.txt 2>&1
# --- bench lines: the headline (BASELINE configs[2]) at N = 1, the reference arm, the per-GPU share of 4/8-way deals, configs[1]
timeout 600 python bench.py --steps 5 --warmup 3 2> $out/r2_bench_n1.err | tail -1 > $out/r2_bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $out/r2_bench_reference.json
timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 4 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > $out/r2_bench_shard_of_4.json
timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 8 --no-cpu-baseline 2>/dev/null | tail -1 > $out/r2_bench_shard_of_8.json
timeout 300 python bench.py --workload config2 --steps 200 --warmup 10 2>/dev/null | tail -1 > $out/r2_bench_config2.json
echo "bench lines done $(date +%T)"
# --- every kernel launch of one sweep (N = 1) with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/r2_launches_config3_sweep.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
echo "launch list done $(date +%T)"
# --- ncu --set full of the launches of single searches: SW Q=513 (R=33, G=16), NW Q=2005 (R=32, two passes), HW Q=5478
#     (chained latency class + six bulk passes)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 2 -f -o $out/ncu_sw513 python tools/one_search.py 570000 SW 1 1 > $out/ncu_sw513.log 2>&1
QLEN=2005 timeout 400 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 3 -f -o $out/ncu_nw2005 python tools/one_search.py 570000 NW 1 1 > $out/ncu_nw2005.log 2>&1
QLEN=5478 timeout 400 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 3 -f -o $out/ncu_hw5478 python tools/one_search.py 570000 HW 1 1 > $out/ncu_hw5478.log 2>&1
pat="search_kernel|Block Size|Grid Size|dram__bytes_(read|write).sum |gpu__dram_throughput|gpu__time_duration.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum |l1tex__t_sector_hit_rate|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|launch__occupancy_limit|sm__inst_executed_pipe_(alu|fma|fmaheavy|fmalite|lsu|xu|uniform).avg.pct_of_peak_sustained_active|smsp__inst_executed.sum |smsp__issue_active.avg.pct|sm__warps_active.avg.pct|smsp__warps_active.avg.per_cycle_active|sm__cycles_elapsed.sum |sm__throughput.avg.pct|smsp__average_warps_issue_stalled_.*_per_issue_active"
for r in sw513 nw2005 hw5478; do
  ncu -i $out/ncu_$r.ncu-rep --page raw 2>/dev/null | grep -E "$pat" > $out/r2_ncu_$r.txt
done
rm -f $out/ncu_sw513.ncu-rep $out/ncu_hw5478.ncu-rep   # (one report is kept for the source page; the others would exceed the transfer limit)
echo "ncu done $(date +%T)"
# --- plans of single searches (the planner's trace), drop-in call phases, BASELINE configs[3] and [4] protocols, sanitizer
for ql in 144 850 2005 5478; do
  QLEN=$ql OPAL_B200_TRACE=1 python tools/one_search.py 570000 NW 1 3 2>&1 | grep -E "group type|GCUPS" | tail -6 > $out/r2_plan_nw$ql.txt
done
OPAL_B200_TRACE=1 timeout 200 python tools/e2e_probe.py config3 2>&1 | tail -40 > $out/r2_e2e_phases.txt
timeout 200 python tools/config4_probe.py config3 > $out/r2_config4_probe.txt 2>&1
timeout 300 python tools/config5_probe.py 1000000 check > $out/r2_config5_probe.txt 2>&1
echo "probes done $(date +%T)"
bash tools/sanitize.sh > $out/r2_sanitizer.txt 2>&1
cp gpurun_out/sanitize/summary.txt $out/r2_sanitizer_summary.txt
ls -la $out
