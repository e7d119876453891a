#!/bin/bash
# usage: bash tools/call_n.sh N TAG   - bench.py on N GPUs under torchrun (+ the C-ABI multi-GPU tests when N >= 2)
N=$1; TAG=$2
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench$N.json 2> gpurun_out/${TAG}_bench$N.err
tail -c 600 gpurun_out/${TAG}_bench$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench$N.json').read().strip().splitlines()[-1])
print("N", d["n_gpus"], "ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
s=d.get("strong",{}); print("strong", s.get("ms_per_step"), s.get("e2e_ms_per_step"), s.get("partition_invariance"))
print("bias", {k:d.get("bias",{}).get(k) for k in ("value","ms_per_step","error")})
g=d.get("giant",{}); print("giant", {k:g.get(k) for k in ("value","ms_per_step","em_ms_per_step","generate_ms","error","loci_per_rank","wave_ms_per_pass")}, g.get("roofline",{}).get("frac"), g.get("roofline",{}).get("real_bytes_frac"))
PY
