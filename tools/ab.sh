#!/bin/bash
# A/B of two builds of the library on one box (development aid): tools/ab.sh <baseline.so> <args of quick_perf.py...>
base=$1; shift
echo "== baseline ($base)"; OPAL_B200_LIB=$base python tools/quick_perf.py "$@" 2>&1 | grep "type=" | cut -c1-75
echo "== candidate";        python tools/quick_perf.py "$@" 2>&1 | grep "type=" | cut -c1-75
