import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from strawberry_b200 import api, synth
mode = sys.argv[1]
if mode == "smallT":
    b = synth.giant(n_loci=1, rows_per_locus=3000, seed=4, iso_lo=20, iso_hi=30)
elif mode == "fewrows":
    b = synth.giant(n_loci=1, rows_per_locus=5, seed=4)
elif mode == "one":
    b = synth.human_shaped(n_loci=60, total_fragments=30_000, seed=5, max_rows=400)
    from strawberry_b200 import partition
    b, _ = partition.take(b, np.arange(int(sys.argv[2]), int(sys.argv[3])))
else:
    b = synth.human_shaped(n_loci=60, total_fragments=30_000, seed=5, max_rows=400)
q = api.Quantifier(max_iter=6)
q.set_plan(3, 0)
q.submit_flat(b)
q.run(b["total_mapped_reads"])
print("tier 3 ok", mode, np.isfinite(q.results()["theta"]).all(), [(r["kernel"], r["n_loci"]) for r in q.launch_stats()], q.stats()["kernel_launches"], "R", int(b["loc_row_off"][-1]), "T", int(b["loc_iso_off"][-1]), "nnz", int(b["row_ptr"][-1]))
