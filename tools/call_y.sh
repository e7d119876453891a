#!/bin/bash
mkdir -p gpurun_out
SBQ_LIB_PATH=build/variants/libsbq_cphase.so timeout 300 python tools/cphase.py 2 0 > gpurun_out/r02y_cphase_top2.txt 2>&1
SBQ_LIB_PATH=build/variants/libsbq_cphase.so timeout 300 python tools/cphase.py 3 40 > gpurun_out/r02y_cphase_mid.txt 2>&1
cat gpurun_out/r02y_cphase_top2.txt gpurun_out/r02y_cphase_mid.txt | grep -v "^setup" | cut -c1-420
