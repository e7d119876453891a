#!/bin/bash
mkdir -p gpurun_out
for th in 16,32,64,150 18,36,72,170 20,40,80,200 24,48,96,250 16,32,64,300; do echo "== thresholds $th"; SBQ_CS_THRESH=$th python tools/prof.py human 20000 2>&1 | grep -E "^\{'n_loci'" | cut -c40-100; done > gpurun_out/r02k_thresh.txt 2>&1
cat gpurun_out/r02k_thresh.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02k_pytest.txt
cat gpurun_out/r02k_pytest.txt
/usr/bin/time -v timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench1.json 2> gpurun_out/r02k_bench1.err; grep -E "Elapsed|Error|error" gpurun_out/r02k_bench1.err | tail -5
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02k_bench1.json').read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
print("bias", json.dumps(d.get("bias"))[:900])
g=d.get("giant",{}); print("giant", {k:g.get(k) for k in ("value","ms_per_step","em_ms_per_step","generate_ms","wall_ms_incl_generation","em_iters_total","error")}, g.get("roofline",{}).get("frac"), g.get("roofline",{}).get("real_bytes_frac"))
print("cpu", d.get("cpu_baseline"))
PY
