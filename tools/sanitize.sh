#!/bin/bash
# compute-sanitizer over the smoke set (run on the GPU box): memcheck, racecheck (shared-memory hazards: the st.async
# cluster exchange, the bulk-copy ring), synccheck.  Logs go to gpurun_out/sanitizer_*.log; the summaries are committed
# under profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_set.py "$@" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_$tool.log | tail -3
done
