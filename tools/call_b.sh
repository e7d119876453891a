#!/bin/bash
# 2-GPU validation: GPU tests (multi-GPU ones included), bench.py at N=2 through torchrun and at N=1 (short configs)
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02b_pytest.txt
cat gpurun_out/r02b_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --giant-loci 8 --giant-rows 200000 > gpurun_out/r02b_bench2.json 2> gpurun_out/r02b_bench2.err
tail -c 3000 gpurun_out/r02b_bench2.err
timeout 600 python bench.py --steps 5 --warmup 3 --giant-loci 4 --giant-rows 1000000 --no-cpu-baseline > gpurun_out/r02b_bench1.json 2> gpurun_out/r02b_bench1.err
tail -c 2000 gpurun_out/r02b_bench1.err
wc -c gpurun_out/r02b_bench*.json
