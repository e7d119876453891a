#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5 > gpurun_out/r02x_pytest_multi.txt; cat gpurun_out/r02x_pytest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02x_bench8.json 2> gpurun_out/r02x_bench8.err
tail -c 1500 gpurun_out/r02x_bench8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02x_bench8.json').read().strip().splitlines()[-1])
print("N", d["n_gpus"], "ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
print("strong", json.dumps(d.get("strong"))[:1500])
print("bias", {k:d.get("bias",{}).get(k) for k in ("value","ms_per_step","error")})
g=d.get("giant",{}); print("giant", {k:g.get(k) for k in ("value","ms_per_step","em_ms_per_step","generate_ms","wall_ms_incl_generation","em_iters_total","error","loci_per_rank")}, g.get("roofline",{}).get("frac"))
print("burst", {k:d.get("roofline_giant",{}).get(k) for k in ("frac","real_bytes_frac")})
PY
