"""Small profiling driver (run under ncu on the GPU box):
   python tools/prof.py giant [rows] [max_iter] [iso_lo iso_hi]   one giant locus through the grid tier
   python tools/prof.py human [n_loci]            human-shaped batch through all tiers
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strawberry_b200 import api, synth  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "giant"
if mode == "giant":
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    iso = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (500, 800)     # optional: T range
    b = synth.giant(n_loci=1, rows_per_locus=rows, seed=4, iso_lo=iso[0], iso_hi=iso[1])
    q = api.Quantifier(max_iter=max_iter)
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    b = synth.human_shaped(n_loci=n, total_fragments=500 * n, seed=2)
    q = api.Quantifier()
q.submit_flat(b)
q.upload()
L = api.lib()
for i in range(3):
    if i == 2 and hasattr(L, "sbq_debug_trace"):      # -DSBQ_TRACE build: drop the records of the warm-up solves
        import ctypes
        import numpy as np
        buf = np.zeros(4 << 18, np.uint64)
        L.sbq_debug_trace(buf.ctypes.data_as(ctypes.c_void_p), 1 << 18)
    q.solve(b["total_mapped_reads"])
if hasattr(L, "sbq_debug_trace"):
    n = L.sbq_debug_trace(buf.ctypes.data_as(ctypes.c_void_p), 1 << 18)
    for r in buf[:4 * n].reshape(n, 4):
        cs, nt, locus = int(r[0]) >> 48, (int(r[0]) >> 32) & 0xffff, int(np.int32(np.uint32(int(r[0]) & 0xffffffff)))
        rank, sm, it = int(r[1]) >> 48, (int(r[1]) >> 32) & 0xffff, int(r[1]) & 0xffffffff
        print(f"TRACE {'warp' if cs == 0 else 'c%d' % cs} nt{nt} locus {locus} rank {rank} sm {sm} t0 {int(r[2])} t1 {int(r[3])} iters {it}")
q.finalize_tpm(q.fpkm_sum())
q.download()
st = q.stats()
print({k: st[k] for k in ("n_loci", "nnz", "solve_ms", "em_ms", "grid_em_ms", "em_iters_total", "alg_bytes", "grid_alg_bytes")})
if st["grid_em_ms"] > 0:
    print("grid GB/s", st["grid_alg_bytes"] / st["grid_em_ms"] / 1e6)
for r in q.launch_stats():
    print(r)
