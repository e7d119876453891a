#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize.py > gpurun_out/r02w_sanitize_racecheck.log 2>&1; tail -12 gpurun_out/r02w_sanitize_racecheck.log | cut -c1-200
