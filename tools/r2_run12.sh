#!/bin/bash
set -u
out=gpurun_out/r2run12; mkdir -p $out
# (1) NW single searches that were sporadically slow: plans and times, with and without chained passes
OPAL_B200_TRACE=1 QLEN=850 timeout 120 python tools/one_search.py 570000 NW 1 6 2>&1 | grep -E "group type|GCUPS" > $out/nw850.txt
OPAL_B200_NO_CHAIN=1 OPAL_B200_TRACE=1 QLEN=850 timeout 120 python tools/one_search.py 570000 NW 1 6 2>&1 | grep -E "group type|GCUPS" > $out/nw850_nochain.txt
CUDA_DEVICE_MAX_CONNECTIONS=8 QLEN=850 timeout 120 python tools/one_search.py 570000 NW 1 6 2>&1 | grep -E "GCUPS" > $out/nw850_conn8.txt
# (2) process start-up cost of 32 hardware queues
for c in 8 32; do /usr/bin/time -f "conn=$c %e s" env CUDA_DEVICE_MAX_CONNECTIONS=$c opal_b200/cli/opal_aligner_b200 -s tests/golden/cli/query.fasta tests/golden/cli/db_clean.fasta > /dev/null 2>> $out/startup.txt; done
for c in 8 32; do /usr/bin/time -f "conn=$c %e s (second run)" env CUDA_DEVICE_MAX_CONNECTIONS=$c opal_b200/cli/opal_aligner_b200 -s tests/golden/cli/query.fasta tests/golden/cli/db_clean.fasta > /dev/null 2>> $out/startup.txt; done
cat $out/startup.txt; tail -8 $out/nw850.txt; tail -3 $out/nw850_nochain.txt; tail -3 $out/nw850_conn8.txt
