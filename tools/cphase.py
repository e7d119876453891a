"""Per-phase cycle counters of the cluster tier on the largest loci of the headline batch (needs a -DSBQ_PHASE_TIMING build):
   SBQ_LIB_PATH=build/variants/libsbq_cphase.so python tools/cphase.py [n_top] [skip]
prints one line per locus (rank 0, threads 0 and NT-32): cycles per iteration in E / or / col / csync1 / own / d2 / csync2 / chk."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strawberry_b200 import api, partition, synth  # noqa: E402

n_top = int(sys.argv[1]) if len(sys.argv) > 1 else 6
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
b = synth.human_shaped(n_loci=20000, total_fragments=10_000_000, seed=2)
nnz = np.diff(np.asarray(b["row_ptr"])[np.asarray(b["loc_row_off"])])
order = np.argsort(-nnz)[skip:skip + n_top]
print("loci", order.tolist(), "nnz", nnz[order].tolist())
sub, _ = partition.take(b, np.sort(order))
q = api.Quantifier()
q.submit_flat(sub)
q.upload()
q.solve(b["total_mapped_reads"])
print(q.stats()["em_ms"], "ms")
