#!/bin/bash
# Round-2 evidence pass on the GPU box: tests, ncu --set full of the cluster kernel and the two-slot giant kernel
# (DRAM bytes of the benchmarked shape), compute-sanitizer on all tiers. Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02a_pytest.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:em_cluster_kernel -c 2 -f -o gpurun_out/r02a_cluster_big python tools/prof.py human 20000 > gpurun_out/r02a_cluster_big.log 2>&1
timeout 600 $NCU -k regex:em_cluster_kernel -s 5 -c 4 -f -o gpurun_out/r02a_cluster_cs1 python tools/prof.py human 20000 > gpurun_out/r02a_cluster_cs1.log 2>&1
timeout 600 $NCU -k regex:dual -c 2 -f -o gpurun_out/r02a_grid_dual python tools/prof.py giant 1000000 8 > gpurun_out/r02a_grid_dual.log 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize.py > gpurun_out/r02a_sanitize_$tool.log 2>&1
  tail -3 gpurun_out/r02a_sanitize_$tool.log
done
ls -la gpurun_out
