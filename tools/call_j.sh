#!/bin/bash
mkdir -p gpurun_out
for ord in 0 1 2 3; do for th in 12,24,48,112 16,32,64,150 14,28,56,130; do echo "== order $ord thresholds $th"; SBQ_ORDER=$ord SBQ_CS_THRESH=$th python tools/prof.py human 20000 2>&1 | grep -E "^\{'n_loci'" | cut -c40-100; done; done > gpurun_out/r02j_order.txt 2>&1
cat gpurun_out/r02j_order.txt
