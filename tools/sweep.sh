#!/bin/bash
# geometry sweep (development aid): usage tools/sweep.sh <config2|N> <mode> <type> "G,R,k ..."
which=$1; mode=$2; st=$3; shift 3
for g in "$@"; do
  echo -n "$g: "; OPAL_B200_GEOMETRY=$g python tools/one_search.py $which $mode $st 3 | tail -1
done
