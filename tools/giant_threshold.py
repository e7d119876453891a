"""Solve time of a human-shaped batch that contains 'small giants' (300 k - 500 k non-zeros) for a given tier threshold:
   SBQ_GRID_MIN_NNZ=... python tools/giant_threshold.py SEED [SEED...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strawberry_b200 import api, synth  # noqa: E402

for seed in [int(x) for x in sys.argv[1:]] or [7]:
    b = synth.human_shaped(seed=seed)
    q = api.Quantifier()
    q.submit_flat(b)
    q.upload()
    best = 1e9
    for _ in range(4):
        q.solve(b["total_mapped_reads"])
        best = min(best, q.stats()["em_ms"])
    st = q.stats()
    print("seed", seed, "SBQ_GRID_MIN_NNZ", os.environ.get("SBQ_GRID_MIN_NNZ"), "best em_ms %.3f" % best, "grid loci", st["loci_grid"], "grid_em_ms %.3f" % st["grid_em_ms"])
    for r in q.launch_stats():
        if r["cluster_size"] in (0, 16) or r["kernel"].startswith("em_grid"):
            print("   ", {k: (round(r[k], 3) if isinstance(r[k], float) else r[k]) for k in ("kernel", "cluster_size", "n_loci", "start_ms", "ms", "max_iters")})
    q.close()
