"""Best resident solve time of the headline batch under the current environment knobs: python tools/ab_quick.py [label]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strawberry_b200 import api, synth  # noqa: E402

b = synth.human_shaped(n_loci=20000, total_fragments=10_000_000, seed=2)
q = api.Quantifier()
q.submit_flat(b)
q.upload()
best = 1e9
for _ in range(6):
    q.solve(b["total_mapped_reads"])
    best = min(best, q.stats()["em_ms"])
ls = q.launch_stats()
w = [r for r in ls if r["kernel"] == "em_warp_kernel"]
print(" ".join(sys.argv[1:]), "best em_ms %.3f" % best, "warp start %.3f ms %.3f" % (w[0]["start_ms"], w[0]["ms"]) if w else "",
      "last end %.3f" % max(r["start_ms"] + r["ms"] for r in ls))
