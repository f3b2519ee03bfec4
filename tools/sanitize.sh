#!/bin/bash
# compute-sanitizer over the small cases of tools/sanitize_cases.py: memcheck, racecheck,
# synccheck and initcheck of every kernel path.  Run on the GPU box; summaries land in gpurun_out/sanitize/ and the
# ones kept as evidence are copied to profiles/.
set -u
out=gpurun_out/sanitize; mkdir -p $out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --log-file $out/$tool.log python tools/sanitize_cases.py > $out/$tool.out 2>&1
  echo "$tool rc=$? $(tail -1 $out/$tool.out)" | tee -a $out/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|error" $out/$tool.log | tail -3 | tee -a $out/summary.txt
done
