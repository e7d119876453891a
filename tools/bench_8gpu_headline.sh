#!/bin/bash
# Headline + strong legs of bench.py on 8 GPUs (no giant / bias legs): what changed with the sub-grid class for small giants.
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 --no-giant --no-bias > gpurun_out/r02g_bench8_headline.json 2> gpurun_out/r02g_bench8_headline.err
tail -c 300 gpurun_out/r02g_bench8_headline.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02g_bench8_headline.json').read().strip().splitlines()[-1])
print("N", d["n_gpus"], "ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
s=d.get("strong",{}); print("strong", s.get("ms_per_step"), s.get("e2e_ms_per_step"), s.get("partition_invariance"))
PY
