"""Which loci bound the bias-in-EM leg: theta iterations / outer rounds of the largest and of the longest-running loci, and the
solve time when only one size class is submitted.   python tools/bias_profile.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strawberry_b200 import api, partition, synth  # noqa: E402

one = synth.human_shaped(seed=2)
X = synth.covariates(one, seed=3)
lro = np.asarray(one["loc_row_off"])
nnz = np.diff(np.asarray(one["row_ptr"])[lro])
T = np.diff(one["loc_iso_off"])


def run(idx, tag):
    idx = np.sort(np.asarray(idx))
    sub, _ = partition.take(one, idx)
    rows = np.concatenate([np.arange(lro[l], lro[l + 1]) for l in idx])
    q = api.Quantifier(bias_mode=1)
    q.submit_flat(sub)
    q.set_covariates(X[rows])
    q.upload()
    for _ in range(2):
        q.solve(one["total_mapped_reads"])
    ms = q.stats()["solve_ms"]
    q.download()
    res = q.results()
    _, outer = q.bias_results()
    q.close()
    print("%-28s loci %6d  solve %9.3f ms  theta iters max %7d sum %9d  outer max %4d" % (tag, len(idx), ms, res["iters"].max(), res["iters"].sum(), outer.max()))
    return res["iters"], outer, idx


it, outer, idx = run(np.arange(len(nnz)), "all")
top = np.argsort(-it)[:12]
print("longest-running loci: (locus, nnz, T, rows, theta iters, outer)")
for k in top:
    l = idx[k]
    print("   ", l, nnz[l], T[l], lro[l + 1] - lro[l], it[k], outer[k])
for lo, hi in ((0, 257), (257, 6001), (6001, 16001), (16001, 40001), (40001, 100001), (100001, 10**9)):
    sel = np.nonzero((nnz >= lo) & (nnz < hi))[0]
    if len(sel):
        run(sel, "nnz in [%d, %d)" % (lo, hi))
