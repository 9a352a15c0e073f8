#!/bin/bash
# compute-sanitizer evidence: memcheck, racecheck, synccheck, initcheck over tools/sanitize_driver.py.
# Usage (1 GPU): gpurun --timeout 900 -- tools/sanitize.sh
export B200MPI_NO_AUTOBUILD=1
mkdir -p gpurun_out/sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== compute-sanitizer --tool $tool ==="
  timeout 280 compute-sanitizer --tool $tool --launch-timeout 0 python tools/sanitize_driver.py > gpurun_out/sanitizer/compute_sanitizer_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize driver done|Error|hazard" gpurun_out/sanitizer/compute_sanitizer_$tool.log | sort | uniq -c | head -8
done
