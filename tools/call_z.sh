#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_robustness.py -x -q -m gpu 2>&1 | tail -3
for tool in racecheck synccheck memcheck; do
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize.py > gpurun_out/r02g_sanitize_$tool.log 2>&1
  grep -E "ok|SUMMARY|ERROR|terminate" gpurun_out/r02g_sanitize_$tool.log | tail -13
done
