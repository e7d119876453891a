#!/bin/bash
set -x
mkdir -p gpurun_out
SBQ_TIMING=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-giant --no-cpu-baseline > gpurun_out/r02h_bench1.json 2> gpurun_out/r02h_bench1.err; grep plan_ms gpurun_out/r02h_bench1.err | tail -4; timeout 300 python -m pytest tests/test_gpu_em.py -x -q 2>&1 | tail -2
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_bench1.json').read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"])
PY
