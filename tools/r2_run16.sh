#!/bin/bash
# A/B of the tightened NW / HW / OV step (new = in-tree library, base = variants/libopal_b200_base.so), then the GPU suite.
set -u
out=gpurun_out/r2run16; mkdir -p $out
base=$PWD/opal_b200/csrc/variants/libopal_b200_base.so
for v in base new base new; do
  [ $v = base ] && export OPAL_B200_LIB=$base || unset OPAL_B200_LIB
  for cfg in "NW 2005" "HW 5478" "OV 1000" "NW 144" "SW 513"; do set -- $cfg
    echo "$v $1 $2: $(QLEN=$2 python tools/one_search.py 570000 $1 1 4 2>/dev/null | tail -2 | awk '{print $2, $4}' | tr '\n' ' ')" >> $out/ab_single.txt
  done
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > $out/bench_$v.json
  python -c "import json;b=json.load(open('$out/bench_$v.json'));print('$v sweep', round(b['value'],1), round(b['ms_per_step'],1), 'e2e', round(b['e2e']['value'],1))" >> $out/ab_single.txt
done
unset OPAL_B200_LIB
cat $out/ab_single.txt
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
