#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02h_bench1.json 2> gpurun_out/r02h_bench1.err; tail -2 gpurun_out/r02h_bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_bench1.json').read().strip().splitlines()[-1])
print("N", d["n_gpus"], "ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
print("cpu_baseline", d["cpu_baseline"]["value"], "bias", d["bias"]["ms_per_step"], "giant", d["giant"]["ms_per_step"], d["giant"]["roofline"]["frac"], d["giant"]["roofline"]["real_bytes_frac"], d["giant"]["wave_ms_per_pass"], "burst", d["roofline_giant"]["frac"], d["roofline_giant"]["real_bytes_frac"])
print(d["clocks"], d["giant"]["clocks"])
PY
