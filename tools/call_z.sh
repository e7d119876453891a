#!/bin/bash
timeout 600 python -m pytest tests/test_bias.py -x -q -m gpu 2>&1 | tail -15
timeout 900 python tools/bias_profile.py 2>&1 | tail -24
