#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_em.py tests/test_gpu_synth.py tests/test_two_slot_layout.py tests/test_gpu_robustness.py -x -q 2>&1 | tail -6
timeout 300 python tools/grid_sweep.py 2>&1 | tail -3
for n in 1 3; do timeout 200 python tools/giant_multi.py $n 1000000 40; done 2>&1 | grep -v iters
timeout 900 python bench.py --steps 3 --warmup 3 --no-bias --no-cpu-baseline > gpurun_out/r02p_bench1.json 2> gpurun_out/r02p_bench1.err; grep -E "Error|error" gpurun_out/r02p_bench1.err | tail -5
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02p_bench1.json').read().strip().splitlines()[-1])
g=d.get("giant",{}); print("giant", {k:g.get(k) for k in ("value","ms_per_step","em_ms_per_step","generate_ms","em_iters_total","error","wave_ms_per_pass")}, g.get("roofline",{}).get("frac"))
r=d.get("roofline_giant",{}); print("burst", {k:r.get(k) for k in ("achieved","frac","real_bytes_frac","kernel_ms")})
PY
