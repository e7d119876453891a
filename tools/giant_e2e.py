"""BASELINE configs[3] shape end to end on one GPU at a reduced locus count: host arrays -> results.
   python tools/giant_e2e.py [n_loci] [rows_per_locus]
Prints the stages of the first run of an upload (H2D, one-off layout pass + EM, D2H) and of a second solve of the resident batch."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strawberry_b200 import api, synth  # noqa: E402

n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
b = synth.giant(n_loci=n_loci, rows_per_locus=rows, seed=4)
pinned = api.pinned_batch(b)
q = api.Quantifier()
for attempt in range(2):   # the first run of a process also pays module loading and the first allocations
    q.clear()
    t0 = time.perf_counter()
    q.submit_flat(pinned)
    q.run(b["total_mapped_reads"])
    res = q.results()
    t1 = time.perf_counter()
    st1 = q.stats()
    if attempt == 0:
        print(f"first run of the process: wall {1e3 * (t1 - t0):.1f} ms (upload {st1['upload_ms']:.1f}, solve {st1['solve_ms']:.1f})")
q.solve(b["total_mapped_reads"])
st2 = q.stats()
nnz = st1["nnz"]
print(f"{n_loci} loci x {rows} rows, {nnz} nnz ({nnz * 12 / 1e9:.2f} GB CSR), iterations {res['iters'].tolist()}, statuses {res['status'].tolist()}")
print(f"second upload: wall {1e3 * (t1 - t0):.1f} ms = upload {st1['upload_ms']:.1f} + solve {st1['solve_ms']:.1f} (layout pass included) + download {st1['download_ms']:.2f} ms (+ host)")
print(f"second solve of the resident batch: {st2['solve_ms']:.1f} ms -> layout pass {st1['solve_ms'] - st2['solve_ms']:.1f} ms; "
      f"{st2['grid_alg_bytes'] / st2['grid_em_ms'] / 1e6:.0f} GB/s algorithmic, {st2['frag_iters'] / st2['solve_ms'] / 1e-3:.3e} fragment-iters/s")
for r in q.launch_stats():
    print(r)
