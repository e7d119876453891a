#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02i_pytest_gpu.txt; cat gpurun_out/r02i_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r02i_bench1.json 2> gpurun_out/r02i_bench1.err; tail -2 gpurun_out/r02i_bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i_bench1.json').read().strip().splitlines()[-1])
print("N", d["n_gpus"], "ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
print("cpu_baseline", d["cpu_baseline"]["value"], "bias", d["bias"]["ms_per_step"], "giant", d["giant"]["ms_per_step"], d["giant"]["roofline"]["frac"], d["giant"]["roofline"]["real_bytes_frac"], "burst", d["roofline_giant"]["frac"], d["roofline_giant"]["real_bytes_frac"])
PY
