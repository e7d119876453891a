#!/bin/bash
# Giant-locus kernel evidence (1 GPU): per-phase cycle counters of a -DSBQ_G6_PHASES build (build/variants/libsbq_phases.so:
#   nvcc ... -DSBQ_G6_PHASES -o build/variants/libsbq_phases.so strawberry_b200/csrc/sbq.cu strawberry_b200/csrc/sbq_builder.cpp)
# and ncu --set full of the layout, packing and EM kernels on the benchmarked shape (1 M rows x 48 non-zeros).
mkdir -p gpurun_out
SBQ_LIB_PATH=build/variants/libsbq_phases.so timeout 300 python tools/prof.py giant 1000000 40 2>&1 | grep -E "G6PHASES|grid GB" | tail -4
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:dual -c 3 -f -o gpurun_out/r02t_grid_dual python tools/prof.py giant 1000000 8 > gpurun_out/r02t_grid_dual.log 2>&1; tail -3 gpurun_out/r02t_grid_dual.log
