#!/bin/bash
for o in 0 2 3; do for c in 0 16 24 32 48; do SBQ_ORDER=$o SBQ_WARP_CTAS=$c timeout 120 python tools/ab_quick.py order $o warp_ctas $c 2>&1 | tail -1; done; done
