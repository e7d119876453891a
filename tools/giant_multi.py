"""Several device-generated giant loci in one launch (the three-slot phase rotation of em_grid_dual_kernel):
   python tools/giant_multi.py [n_loci] [rows] [max_iter]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strawberry_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
q = api.Quantifier(max_iter=max_iter)
iso_lo, iso_hi = int(os.environ.get("ISO_LO", 500)), int(os.environ.get("ISO_HI", 800))
q.synth_giant(list(range(n)), rows, iso_lo=iso_lo, iso_hi=iso_hi)
for _ in range(2):
    q.solve(rows * n)
q.finalize_tpm(q.fpkm_sum())
q.download()
st = q.stats()
passes = st["em_iters_total"] + n
print(dict(n_loci=n, rows=rows, nnz=st["nnz"], generate_ms=st["upload_ms"], grid_em_ms=st["grid_em_ms"], iters=st["em_iters_total"],
           ms_per_pass=st["grid_em_ms"] / passes, alg_GBps=st["grid_alg_bytes"] / st["grid_em_ms"] / 1e6,
           real_GBps=10.63 / 12 * st["grid_alg_bytes"] / st["grid_em_ms"] / 1e6))
r = q.results()
print("iters", r["iters"], "status", r["status"])
