#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-bias --no-cpu-baseline > gpurun_out/r02m_bench1.json 2> gpurun_out/r02m_bench1.err; grep -E "Error|error" gpurun_out/r02m_bench1.err | tail -5
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02m_bench1.json').read().strip().splitlines()[-1])
g=d.get("giant",{}); print("giant", {k:g.get(k) for k in ("value","ms_per_step","em_ms_per_step","generate_ms","em_iters_total","error","wave_ms_per_pass","clocks")}, g.get("roofline",{}).get("frac"))
r=d.get("roofline_giant",{}); print("burst", {k:r.get(k) for k in ("achieved","frac","real_bytes_frac","kernel_ms","workload")})
PY
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
