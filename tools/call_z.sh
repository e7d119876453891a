#!/bin/bash
SBQ_LIB_PATH=build/variants/libsbq_nofence.so timeout 300 python tools/prof.py giant 1000000 40 2>&1 | grep -E "grid GB"
SBQ_LIB_PATH=build/variants/libsbq_phases.so timeout 300 python tools/prof.py giant 1000000 40 2>&1 | grep -E "G6PHASES|grid GB" | tail -2
