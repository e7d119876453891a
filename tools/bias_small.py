"""A small bias-mode batch through both bias kernels (for ncu): python tools/bias_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strawberry_b200 import api, synth  # noqa: E402

b = synth.human_shaped(n_loci=4000, total_fragments=2_000_000, seed=2)
q = api.Quantifier(bias_mode=1)
q.submit_flat(b)
q.set_covariates(synth.covariates(b, seed=3))
q.run(b["total_mapped_reads"])
st = q.stats()
print({k: st[k] for k in ("n_loci", "nnz", "solve_ms", "em_iters_total", "kernel_launches")})
