#!/bin/bash
set -x
mkdir -p gpurun_out
SBQ_LIB_PATH=build/variants/libsbq_trace.so python tools/prof.py human 20000 > gpurun_out/r02i_trace.txt 2>&1
python tools/trace_timeline.py < gpurun_out/r02i_trace.txt > gpurun_out/r02i_timeline.txt
cat gpurun_out/r02i_timeline.txt
grep -v TRACE gpurun_out/r02i_trace.txt | head -3
for th in 12,24,48,112 8,16,32,80 16,32,64,150 6,12,24,60 12,24,48,200; do echo "== thresholds $th"; SBQ_CS_THRESH=$th python tools/prof.py human 20000 2>&1 | grep -E "^\{'n_loci'" | cut -c1-120; done > gpurun_out/r02i_thresh.txt 2>&1
cat gpurun_out/r02i_thresh.txt
gzip -f gpurun_out/r02i_trace.txt
