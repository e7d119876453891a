#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02d_pytest.txt
cat gpurun_out/r02d_pytest.txt
timeout 900 python tools/integration_bench.py 3000 > gpurun_out/r02d_integration_bench.md 2> gpurun_out/r02d_integration_bench.err
cat gpurun_out/r02d_integration_bench.md; tail -5 gpurun_out/r02d_integration_bench.err
