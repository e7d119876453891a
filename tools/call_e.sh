#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02e_pytest.txt
cat gpurun_out/r02e_pytest.txt
timeout 300 python tools/prof.py human 20000 > gpurun_out/r02e_prof_human.txt 2>&1; tail -14 gpurun_out/r02e_prof_human.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-giant --no-cpu-baseline > gpurun_out/r02e_bench1.json 2> gpurun_out/r02e_bench1.err; tail -3 gpurun_out/r02e_bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench1.json').read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["config"]["tiers_rank0"])
for l in d["roofline"]["launches"]: print(l)
PY
