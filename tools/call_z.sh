#!/bin/bash
mkdir -p gpurun_out
for tool in racecheck synccheck memcheck; do
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize.py > gpurun_out/r02z_sanitize_$tool.log 2>&1
  grep -E "ok|SUMMARY|ERROR" gpurun_out/r02z_sanitize_$tool.log | tail -12
done
