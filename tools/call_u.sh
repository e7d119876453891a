#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize.py > gpurun_out/r02u_sanitize_$tool.log 2>&1
  tail -4 gpurun_out/r02u_sanitize_$tool.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02u_pytest.txt; cat gpurun_out/r02u_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02u_bench1.json 2> gpurun_out/r02u_bench1.err; grep -iE "error|Traceback" gpurun_out/r02u_bench1.err | tail -3
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02u_bench_ref.json 2> gpurun_out/r02u_bench_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02u_bench1.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r02u_bench_ref.json').read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "ref", r["ms_per_step"], "e2e ratio", d["e2e"]["value"]/r["value"], "resident ratio", d["value"]/r["value"])
g=d.get("giant",{}); print("giant", {k:g.get(k) for k in ("value","ms_per_step","error")}, g.get("roofline",{}).get("frac"), g.get("roofline",{}).get("real_bytes_frac"))
print("burst", {k:d.get("roofline_giant",{}).get(k) for k in ("frac","real_bytes_frac","kernel_ms")})
print("bias", {k:d.get("bias",{}).get(k) for k in ("value","ms_per_step","error")})
PY
