"""A/B of two builds of libsbq.so on the headline batch (seed 2, 20 000 loci, 10 M fragments):
   SBQ_LIB_PATH=<lib> python tools/ab_compare.py run <tag>      solves the batch 8 times, prints the best resident solve time and
                                                               the per-launch times, writes gpurun_out/ab_<tag>.npz
   python tools/ab_compare.py diff <tagA> <tagB>               bitwise comparison of the outputs"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if sys.argv[1] == "run":
    from strawberry_b200 import api, synth
    tag = sys.argv[2]
    b = synth.human_shaped(n_loci=20000, total_fragments=10_000_000, seed=2)
    q = api.Quantifier()
    q.submit_flat(b)
    q.upload()
    best = 1e9
    for _ in range(8):
        q.solve(b["total_mapped_reads"])
        best = min(best, q.stats()["em_ms"])
    print(tag, "best em_ms", best)
    for r in q.launch_stats():
        print("  ", {k: r[k] for k in ("kernel", "cluster_size", "threads", "n_loci", "ms", "start_ms", "max_iters") if k in r})
    q.finalize_tpm(q.fpkm_sum())
    q.download()
    res = q.results()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez(os.path.join(ROOT, "gpurun_out", "ab_%s.npz" % tag), **{k: np.asarray(v) for k, v in res.items()})
else:
    a = np.load(os.path.join(ROOT, "gpurun_out", "ab_%s.npz" % sys.argv[2]))
    b = np.load(os.path.join(ROOT, "gpurun_out", "ab_%s.npz" % sys.argv[3]))
    for k in a.files:
        x, y = a[k], b[k]
        same = x.tobytes() == y.tobytes()
        extra = ""
        if not same and x.dtype.kind == "f":
            d = np.abs(x - y) / np.maximum(np.abs(x), 1e-300)
            extra = " max rel dev %.3g, differing %d of %d" % (np.nanmax(d), int((x != y).sum()), x.size)
        elif not same:
            extra = " differing %d of %d" % (int((x != y).sum()), x.size)
        print(k, "bitwise equal" if same else "DIFFERENT" + extra)
