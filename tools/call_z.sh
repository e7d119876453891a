#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02z_pytest_gpu.txt; cat gpurun_out/r02z_pytest_gpu.txt
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py > gpurun_out/r02z_sanitize_$tool.log 2>&1
  grep -E "ok|SUMMARY|ERROR" gpurun_out/r02z_sanitize_$tool.log | tail -12
done
