#!/bin/bash
set -x
mkdir -p gpurun_out
SBQ_LIB_PATH=build/variants/libsbq_trace.so python tools/prof.py human 20000 > gpurun_out/r02c_trace.txt 2>&1
python tools/trace_timeline.py < gpurun_out/r02c_trace.txt > gpurun_out/r02c_timeline.txt
cat gpurun_out/r02c_timeline.txt
grep -v TRACE gpurun_out/r02c_trace.txt | tail -12
for rows in 125000 250000 500000 1000000; do python tools/prof.py giant $rows 40 2>&1 | grep -E "grid GB|em_grid" ; done > gpurun_out/r02c_giant_rows.txt
cat gpurun_out/r02c_giant_rows.txt
gzip -f gpurun_out/r02c_trace.txt
