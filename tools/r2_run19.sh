#!/bin/bash
# Launch list of the final build (every kernel launch of 3 warm-up sweeps + 1, with its device time).
set -u
out=gpurun_out/r2run19; mkdir -p $out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/r2_launches_config3_sweep.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python tools/summarise_launches.py $out/r2_launches_config3_sweep.csv | head -12
