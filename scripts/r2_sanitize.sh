#!/bin/bash
# compute-sanitizer memcheck + racecheck over the hot path (SURVEY section 5 / VERDICT r1 item 6); summaries -> gpurun_out/
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  SAN_L=${SAN_L:-2048} timeout ${SAN_TIMEOUT:-500} compute-sanitizer --tool $tool --print-limit 20 \
      python scripts/r2_sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_TARGET_DONE|^ok|Error|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -20
done
