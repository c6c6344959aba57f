#!/bin/bash
# Dev tool (GPU box): compute-sanitizer memcheck + racecheck over tools/sanitize_case.py; summaries -> gpurun_out/
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py > gpurun_out/r2_sanitizer_$tool.txt 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_case|hazard" gpurun_out/r2_sanitizer_$tool.txt | tail -5
done
