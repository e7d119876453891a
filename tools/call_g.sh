#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_em.py tests/test_gpu_synth.py tests/test_gpu_locus.py tests/test_bias.py -x -q 2>&1 | tail -6 > gpurun_out/r02g_pytest.txt
cat gpurun_out/r02g_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-giant --no-cpu-baseline > gpurun_out/r02g_bench1.json 2> gpurun_out/r02g_bench1.err; tail -3 gpurun_out/r02g_bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench1.json').read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"], d["config"]["tiers_rank0"])
for l in d["roofline"]["launches"]: print(l)
PY
