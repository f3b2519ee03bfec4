#!/bin/bash
# Final build of the round: GPU suite, bench lines, ncu of the tightened bulk kernel, forced-geometry probes.
set -u
out=gpurun_out/r2run17; mkdir -p $out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 2> $out/r2_bench_n1.err | tail -1 > $out/r2_bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $out/r2_bench_reference.json
timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 4 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > $out/r2_bench_shard_of_4.json
timeout 300 python bench.py --steps 3 --warmup 3 --shard-of 8 --no-cpu-baseline 2>/dev/null | tail -1 > $out/r2_bench_shard_of_8.json
timeout 300 python bench.py --workload config2 --steps 200 --warmup 10 2>/dev/null | tail -1 > $out/r2_bench_config2.json
echo "bench done $(date +%T)"
QLEN=2005 timeout 400 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 3 -f -o $out/ncu_nw2005 python tools/one_search.py 570000 NW 1 1 > $out/ncu_nw2005.log 2>&1
pat="search_kernel|Block Size|Grid Size|dram__bytes_(read|write).sum |gpu__dram_throughput|gpu__time_duration.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum |l1tex__t_sector_hit_rate|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|launch__occupancy_limit|sm__inst_executed_pipe_(alu|fma|fmaheavy|fmalite|lsu|xu|uniform).avg.pct_of_peak_sustained_active|smsp__inst_executed.sum |smsp__issue_active.avg.pct|sm__warps_active.avg.pct|smsp__warps_active.avg.per_cycle_active|sm__cycles_elapsed.sum |sm__throughput.avg.pct|smsp__average_warps_issue_stalled_.*_per_issue_active"
ncu -i $out/ncu_nw2005.ncu-rep --page raw 2>/dev/null | grep -E "$pat" > $out/r2_ncu_nw2005_tight.txt
echo "ncu done $(date +%T)"
for g in "" "16,32,3" "16,33,3" "32,32,2" "8,32,3"; do
  echo "HW 2005 geometry [$g]: $(env ${g:+OPAL_B200_GEOMETRY=$g} QLEN=2005 python tools/one_search.py 570000 HW 1 3 2>/dev/null | tail -1 | cut -c1-200)" >> $out/geometry.txt
done
for g in "" "16,33,3" "32,33,3" "32,29,2" "16,29,3"; do
  echo "HW 5478 geometry [$g]: $(env ${g:+OPAL_B200_GEOMETRY=$g} QLEN=5478 python tools/one_search.py 570000 HW 1 3 2>/dev/null | tail -1 | cut -c1-200)" >> $out/geometry.txt
done
cat $out/geometry.txt
for m in 8; do SHARD_OF=$m QLEN=2005 MODE=NW OPAL_B200_TRACE=1 timeout 120 python tools/e2e_probe.py config3 2>&1 | tail -14 > $out/e2e_phases_shard$m.txt; done
python - <<'P'
import json
for f in ('r2_bench_n1','r2_bench_reference','r2_bench_shard_of_4','r2_bench_shard_of_8','r2_bench_config2'):
    b=json.loads(open('gpurun_out/r2run17/'+f+'.json').read())
    print(f, round(b['value'],1), round(b['ms_per_step'],2), round(b.get('e2e',{}).get('value',0),1), b.get('roofline',{}).get('frac'))
P
