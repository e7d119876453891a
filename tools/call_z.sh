#!/bin/bash
timeout 600 python -m pytest tests/test_bias.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python tools/bias_profile.py 2>&1 | grep -E "^all|nnz in \[0"
