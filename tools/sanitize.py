"""Small all-tier workload for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strawberry_b200 import api, synth  # noqa: E402

b = synth.concat([synth.human_shaped(n_loci=60, total_fragments=30_000, seed=5, max_rows=400),
                  synth.giant(n_loci=1, rows_per_locus=3000, seed=4)])
for tier, cs, it in ((0, 0, 30), (2, 1, 10), (2, 4, 10), (2, 16, 10), (3, 0, 6)):
    q = api.Quantifier(max_iter=it)
    q.set_plan(tier, cs)
    q.submit_flat(b)
    q.run(b["total_mapped_reads"])
    r = q.results()
    print("tier", tier, "cs", cs, "ok", np.isfinite(r["theta"]).all(), q.stats()["kernel_launches"])
    q.close()
