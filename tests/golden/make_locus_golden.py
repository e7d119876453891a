"""Generate tests/golden/locus_golden.json.gz from the UNMODIFIED reference's LocusContext (wide seam of
oracle/_ref/libsbref.so). Run where the reference checkout is present:
    make -C oracle ref && python tests/golden/make_locus_golden.py
Each case stores the inputs (isoform exons, mate CIGARs, masses, read length, insert model) and the reference's
dump: hit features, segments, isoform segments/lengths, classes (coords, float mass, int count, set size, alpha),
iso->class map, theta, FPKM, frac."""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
import locusgen  # noqa: E402
from test_builder import reference_table, specs_for  # noqa: E402

cases = []
for seed in list(range(1000, 1030)):
    isoforms, hits, rl = locusgen.random_locus(seed, max_frag=120)
    spec = specs_for(seed, hits)
    long_read = seed % 17 == 0
    ref = reference_table(oracle, isoforms, hits, rl, spec, long_read)
    for c in ref["classes"]:
        c.pop("frag_lens", None)
    cases.append(dict(seed=seed, isoforms=isoforms, hits=hits, read_len=rl, spec=list(spec), long_read=long_read, ref=ref))
out = os.path.join(HERE, "locus_golden.json.gz")
with gzip.open(out, "wt") as f:
    json.dump(cases, f, separators=(",", ":"))
print("wrote", out, os.path.getsize(out), "bytes;", sum(len(c["ref"]["classes"]) for c in cases), "classes")
