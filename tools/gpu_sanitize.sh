#!/bin/bash
# compute-sanitizer passes over tools/sanitize_driver.py; logs under gpurun_out/
set -x
mkdir -p gpurun_out
python tools/sanitize_driver.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -30 gpurun_out/sanitize_plain.log; exit 1; }
for tool in memcheck initcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_driver.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|driver done|^ok " gpurun_out/sanitize_$tool.log | tail -12
  grep -E "=========     at |Invalid|Uninitialized|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | sort -rn | head -20
done
