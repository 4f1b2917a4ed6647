#!/bin/bash
# compute-sanitizer over every kernel family (run under gpurun). Output: gpurun_out/sanitize/*.log
mkdir -p gpurun_out/sanitize
for TOOL in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $TOOL --error-exitcode 9 python scripts/sanitize_cases.py > gpurun_out/sanitize/$TOOL.log 2>&1
  echo "$TOOL exit=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize/$TOOL.log | head -5
done
