#!/bin/bash
# compute-sanitizer over tools/sanitizer_smoke.py (under gpurun); one bounded run per tool, logs under gpurun_out/
O=gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitizer_smoke.py > $O/san_$tool.log 2>&1
  echo "$tool rc=$? $(grep -c 'ERROR SUMMARY' $O/san_$tool.log) $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $O/san_$tool.log | tail -1)"
done
