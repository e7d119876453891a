#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s.%N)
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02l_bench1.json 2> gpurun_out/r02l_bench1.err; grep -E "Error|error" gpurun_out/r02l_bench1.err | tail -5
t1=$(date +%s.%N); echo "bench wall seconds: $(echo "$t1 - $t0" | bc)"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02l_bench1.json').read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
print("bias", json.dumps(d.get("bias"))[:1200])
g=d.get("giant",{}); print("giant", {k:g.get(k) for k in ("value","ms_per_step","em_ms_per_step","generate_ms","wall_ms_incl_generation","em_iters_total","error")}, g.get("roofline",{}).get("frac"), g.get("roofline",{}).get("real_bytes_frac"))
print("cpu", d.get("cpu_baseline"))
PY
t0=$(date +%s.%N)
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02l_bench_ref.json 2> gpurun_out/r02l_bench_ref.err
t1=$(date +%s.%N); echo "reference arm wall seconds: $(echo "$t1 - $t0" | bc)"; cut -c1-400 gpurun_out/r02l_bench_ref.json
