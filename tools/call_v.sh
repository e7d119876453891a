#!/bin/bash
echo "== pre-pack build, racecheck 0 40"; SBQ_LIB_PATH=build/variants/libsbq_prepack.so timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python tools/rc_test.py one 0 40 2>&1 | grep -E "tier 3|illegal|hazards" | head -4 | cut -c1-250
