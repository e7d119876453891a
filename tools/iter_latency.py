"""Per-iteration latency of each tier on single loci (theta_tol = 0 forces max_iter iterations)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strawberry_b200 import api, synth  # noqa: E402


def one_locus(T, R, k_mean, seed=0):
    rng = np.random.default_rng(seed)
    rows, cols = [], []
    rp = [0]
    for i in range(R):
        k = int(min(T, max(1, rng.poisson(k_mean))))
        c = np.sort(rng.choice(T, k, replace=False))
        cols.append(c)
        rp.append(rp[-1] + k)
    col = np.concatenate(cols).astype(np.int32)
    return dict(loc_row_off=np.array([0, R], np.int64), loc_iso_off=np.array([0, T], np.int64), row_ptr=np.array(rp, np.int64), col=col,
                alpha=10.0 ** rng.uniform(-4, -1.5, len(col)), count=rng.integers(1, 100, R).astype(np.int32),
                iso_len=np.full(T, 1000, np.int32), total_mapped_reads=1000000)


ITERS = 400
q = api.Quantifier(max_iter=ITERS, theta_tol=0.0)
shapes = [(4, 4, 2, 1, 0), (4, 4, 2, 2, 1), (8, 30, 3, 1, 0), (30, 150, 12, 1, 0), (30, 150, 12, 2, 1), (34, 101, 13, 2, 1), (60, 300, 20, 2, 1),
          (60, 300, 20, 2, 2), (145, 1000, 60, 2, 4), (145, 2000, 60, 2, 8), (145, 4362, 63, 2, 16), (200, 3000, 90, 2, 16), (145, 4362, 63, 3, 0), (200, 3000, 90, 3, 0), (145, 2000, 60, 3, 0), (500, 20000, 25, 3, 0), (500, 20000, 25, 2, 16)]
for T, R, k, tier, cs in shapes:
    b = one_locus(T, R, k)
    q.clear()
    q.set_plan(tier, cs)
    q.submit_flat(b)
    q.upload()
    for _ in range(3):
        q.solve(b["total_mapped_reads"])
    st = q.stats()
    q.finalize_tpm(q.fpkm_sum())
    q.download()
    it = int(q.results()["iters"][0])
    print(f"T={T:4d} R={R:5d} nnz={len(b['col']):7d} tier={tier} cs={cs:2d}: {st['em_ms']:.3f} ms / {it} iters = {1e3 * st['em_ms'] / max(it, 1):.2f} us/iter")
