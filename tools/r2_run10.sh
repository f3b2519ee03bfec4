#!/bin/bash
# chained passes, with tight timeouts
set -u
out=gpurun_out/r2run10; mkdir -p $out
timeout 120 python -m pytest tests/test_gpu_chained.py -x -q --timeout 40 > $out/pytest_chain.log 2>&1; echo "pytest chain rc=$?" >> $out/pytest_chain.log
tail -25 $out/pytest_chain.log
