#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_em.py tests/test_gpu_synth.py tests/test_gpu_locus.py tests/test_two_slot_layout.py -x -q 2>&1 | tail -8 > gpurun_out/r02f_pytest.txt
cat gpurun_out/r02f_pytest.txt
for n in 1 2 3 6; do timeout 200 python tools/giant_multi.py $n 1000000 40; done > gpurun_out/r02f_giant_multi.txt 2>&1
cat gpurun_out/r02f_giant_multi.txt
timeout 200 python tools/giant_multi.py 6 1000000 1000 >> gpurun_out/r02f_giant_multi.txt 2>&1; tail -2 gpurun_out/r02f_giant_multi.txt
timeout 300 python tools/grid_sweep.py > gpurun_out/r02f_grid_sweep.txt 2>&1; tail -3 gpurun_out/r02f_grid_sweep.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-giant --no-cpu-baseline > gpurun_out/r02f_bench1.json 2> gpurun_out/r02f_bench1.err; tail -3 gpurun_out/r02f_bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench1.json').read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["config"]["tiers_rank0"])
for l in d["roofline"]["launches"]: print(l)
PY
