#!/bin/bash
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_rawbuild.py -x -q 2>&1 | tail -5
