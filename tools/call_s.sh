#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_integration.py tests/test_gpu_rawbuild.py -x -q 2>&1 | tail -8
timeout 900 python tools/integration_bench.py 3000 > gpurun_out/r02s_integration_bench.md 2> gpurun_out/r02s_integration_bench.err
cat gpurun_out/r02s_integration_bench.md; tail -3 gpurun_out/r02s_integration_bench.err
