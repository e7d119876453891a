#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:em_bias -c 6 -f -o gpurun_out/r02i_bias python tools/bias_small.py > gpurun_out/r02i_bias.log 2>&1; tail -3 gpurun_out/r02i_bias.log
ls -la gpurun_out/r02i_bias.ncu-rep
