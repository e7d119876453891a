#!/bin/bash
# Final round-2 evidence pass (1 GPU): full GPU test suite, bench (ours + reference arm), ncu launch list of the bench command,
# ncu --set full of the final cluster kernel (16- and 8-CTA launches) and of the bias kernels. Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02f_pytest_gpu.txt; cat gpurun_out/r02f_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r02f_bench1.json 2> gpurun_out/r02f_bench1.err; tail -2 gpurun_out/r02f_bench1.err
timeout 600 python bench.py --impl reference > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 1 --no-giant --no-bias --no-cpu-baseline > gpurun_out/r02f_bench_under_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:em_cluster_kernel -c 2 -f -o gpurun_out/r02f_cluster_big python tools/prof.py human 20000 > gpurun_out/r02f_cluster_big.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench1.json').read().strip().splitlines()[-1])
print("N", d["n_gpus"], "ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
print("cpu_baseline", d["cpu_baseline"]["value"], "bias", d["bias"]["ms_per_step"], "giant", d["giant"]["ms_per_step"], d["giant"]["roofline"]["frac"], d["giant"]["roofline"]["real_bytes_frac"], "burst", d["roofline_giant"]["frac"], d["roofline_giant"]["real_bytes_frac"])
r=json.loads(open('gpurun_out/r02f_bench_ref.json').read().strip().splitlines()[-1]); print("ref", r["value"], r["ms_per_step"], "ratio e2e", d["e2e"]["value"]/r["value"], "resident", d["value"]/r["value"])
PY
ls -la gpurun_out | tail -12
