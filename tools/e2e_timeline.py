"""Where the end-to-end step (host buffers -> results) spends its time: host wall clock of every call + the launch timeline.
   python tools/e2e_timeline.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strawberry_b200 import api, synth  # noqa: E402

b = synth.human_shaped(n_loci=20000, total_fragments=10_000_000, seed=2)
pinned = api.pinned_batch(b)
q = api.Quantifier()
tot = b["total_mapped_reads"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
best = None
for it in range(8):
    flush.fill_(1)
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    q.clear(); t.append(time.perf_counter())
    q.submit_flat(pinned); t.append(time.perf_counter())
    q.upload_begin(); t.append(time.perf_counter())
    q.solve(tot); t.append(time.perf_counter())
    q.finalize_tpm(q.fpkm_sum()); t.append(time.perf_counter())
    q.download(); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    if it >= 2 and (best is None or t[-1] - t[0] < best[0]):
        best = (t[-1] - t[0], d, q.stats(), q.launch_stats())
print("best e2e ms %.3f: clear %.3f submit %.3f upload_begin %.3f solve %.3f tpm %.3f download %.3f sync %.3f" % ((best[0] * 1e3,) + tuple(best[1])))
st = best[2]
print({k: st[k] for k in st if k.endswith("_ms")})
for r in best[3]:
    print("  ", {k: (round(r[k], 3) if isinstance(r[k], float) else r[k]) for k in ("kernel", "cluster_size", "threads", "n_loci", "start_ms", "ms")})
