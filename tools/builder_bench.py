"""Class-table build cost per locus: reference LocusContext constructor (compiled reference, oracle/_ref) vs the
libsbq host builder vs host builder with deferred weights + GPU weights_kernel. Run on the GPU box."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import locusgen  # noqa: E402
import oracle  # noqa: E402
from strawberry_b200 import api, builder  # noqa: E402

n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 200
model = builder.Model.normal(230.0, 45.0)
cases = []
for seed in range(7000, 7000 + 4 * n_loci):
    isoforms, hits, rl = locusgen.random_locus(seed, max_frag=3000, max_iso=10, max_exon=14)
    if rl == 75:
        cases.append((isoforms, hits))
    if len(cases) == n_loci:
        break
tfes = [[locusgen.transcript_features(ex) for ex in iso] for iso, _ in cases]
feats = [[(m, builder.pair_features(l, r)) for m, l, r in hits] for _, hits in cases]
n_hits = sum(len(h) for _, h in cases)

t_ref = 0.0
if oracle.have_ref():
    for (iso, hits), tfe in zip(cases, tfes):
        t_ref += float(oracle.ref_locus_context(tfe, hits, read_len=75, mean=230.0, sd=45.0)["ctor_seconds"])
t0 = time.perf_counter()
full = [builder.build_locus(tfe, f, read_len=75, model=model) for tfe, f in zip(tfes, feats)]
t_host = time.perf_counter() - t0
t0 = time.perf_counter()
deferred = [builder.build_locus(tfe, f, read_len=75, model=model, defer_weights=True) for tfe, f in zip(tfes, feats)]
t_defer = time.perf_counter() - t0
nnz = sum(len(t["col"]) for t in full)
q = api.Quantifier()
q.set_insert_model(model, 75)
q.submit_deferred([t["table"] for t in deferred])
q.upload()
w_ms = q.stats()["weights_ms"]
a = q.fetch_alpha()
ah = np.concatenate([t["alpha"] for t in full])
print(dict(loci=len(cases), collapsed_hits=n_hits, classes=sum(len(t["classes"]) for t in full), nnz=nnz,
           reference_ctor_s=round(t_ref, 4), host_builder_s=round(t_host, 4), host_builder_deferred_s=round(t_defer, 4),
           gpu_weights_ms=round(w_ms, 3), alpha_max_rel_err=float(np.max(np.abs(a - ah) / np.maximum(np.abs(ah), 1e-300)))))
print("note: host_builder_* include the Python ctypes marshalling of every locus; reference_ctor_s is C++ only")
