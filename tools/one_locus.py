"""One synthetic locus through a forced tier (for ncu): python tools/one_locus.py T R k_mean tier cs iters"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from strawberry_b200 import api  # noqa: E402
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("il", os.path.join(ROOT, "tools", "iter_latency.py"))
src = open(os.path.join(ROOT, "tools", "iter_latency.py")).read().split("ITERS = 400")[0]
ns = {"__file__": os.path.join(ROOT, "tools", "iter_latency.py")}
exec(compile(src, "iter_latency_head", "exec"), ns)
T, R, k, tier, cs, iters = (int(x) for x in sys.argv[1:7])
b = ns["one_locus"](T, R, k)
q = api.Quantifier(max_iter=iters, theta_tol=0.0)
q.set_plan(tier, cs)
q.submit_flat(b)
q.upload()
for _ in range(3):
    q.solve(b["total_mapped_reads"])
print(q.stats()["em_ms"], "ms")
