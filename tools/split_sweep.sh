#!/bin/bash
# Forced latency-class splits (development aid): tools/split_sweep.sh "m,SMs,folded,k ..." -> device time and the planner's estimate
for s in ${1:-32,8,0,2 8,4,1,2 4,2,1,2}; do
  echo -n "$s: "
  OPAL_B200_TRACE=1 OPAL_B200_SPLIT=$s python tools/one_search.py ${DB:-config2} ${MODE:-SW} ${ST:-1} 4 2>&1 | grep "group\|GCUPS" | tail -3 | sed 's/.*\(G=[0-9]* R=[0-9]* k=[0-9]*\).*est=\([0-9]*\).*/\1 est \2;/; s/^0 \([0-9.]*\) ms.*/ actual \1/' | tr '\n' ' '
  echo
done
