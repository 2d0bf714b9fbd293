#!/bin/bash
# compute-sanitizer over the two-level scenes of tools/sanitize_driver.py
mkdir -p gpurun_out
python tools/sanitize_driver.py two_level > gpurun_out/sanitize_two_plain.log 2>&1 || { tail -30 gpurun_out/sanitize_two_plain.log; exit 1; }
for tool in memcheck initcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_driver.py two_level > gpurun_out/sanitize_two_$tool.log 2>&1
  echo "== $tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|driver done|^ok " gpurun_out/sanitize_two_$tool.log | tail -8
  grep -E "=========     at |Invalid|Uninitialized|hazard" gpurun_out/sanitize_two_$tool.log | sort | uniq -c | sort -rn | head -10
done
