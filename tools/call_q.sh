#!/bin/bash
mkdir -p gpurun_out
for nc in 12 14; do echo "== forced NC $nc, T in 500..620"; SBQ_DUAL_NC=$nc ISO_LO=500 ISO_HI=620 timeout 200 python tools/giant_multi.py 3 1000000 40 2>&1 | head -1 | cut -c1-400; done
echo "== default, T in 500..620"; ISO_LO=500 ISO_HI=620 timeout 200 python tools/giant_multi.py 3 1000000 1000 2>&1 | cut -c1-400
timeout 300 python -m pytest tests/test_gpu_synth.py tests/test_gpu_em.py -x -q 2>&1 | tail -2
timeout 300 python tools/grid_sweep.py 2>&1 | tail -2
