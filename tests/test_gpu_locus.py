"""GPU, end to end through both C ABIs: builder (host) -> sbq_submit -> CUDA EM + epilogue, against what the
UNMODIFIED reference's LocusContext::estimate_abundances produced for the same locus (committed goldens, and
live against oracle/_ref/libsbref.so when it is present)."""
import gzip
import json

import numpy as np
import pytest

import locusgen
from test_builder import GOLD, product_table, reference_table, specs_for

pytestmark = pytest.mark.gpu


def _num(x):
    return float("nan") if x == "nan" else float(x)


def check_case(q, isoforms, hits, read_len, spec, long_read, ref, what):
    _, tab = product_table(isoforms, hits, read_len, spec, long_read)
    q.clear()
    q.submit([(tab["n_iso"], tab["row_ptr"], tab["col"], tab["alpha"], tab["count"], tab["iso_len"])])
    q.run(100000)
    res = q.results()
    assert bool(res["status"][0] != 3) == ref["success"] == ref["em_init"], what
    assert bool(res["status"][0] in (0, 1)) == ref["em_run"], what
    th_ref = np.array([_num(x) for x in ref["theta"]])
    tot = max(1.0, float(np.sum(tab["count"])))
    scale = np.maximum(np.abs(th_ref), 1e-9 * tot)
    assert (np.abs(res["theta"] - th_ref) / scale).max() <= 1e-6, (what, res["theta"], th_ref)
    if ref["success"]:
        # with kMinIsoformFrac = 0 nothing is erased: every isoform is reported, in order
        assert [iso["id"] for iso in ref["isoforms"]] == list(range(tab["n_iso"])), what
        fpkm_ref = np.array([_num(iso["fpkm"]) for iso in ref["isoforms"]])
        frac_ref = np.array([_num(iso["frac"]) for iso in ref["isoforms"]])
        for got, want in ((res["fpkm"], fpkm_ref), (res["frac"], frac_ref)):
            fin = np.isfinite(want)
            assert np.array_equal(np.isfinite(got), fin), what
            sc = np.maximum(np.abs(want[fin]), 1e-9 * max(1e-300, float(np.max(np.abs(want[fin]))) if fin.any() else 1.0))
            assert (np.abs(got[fin] - want[fin]) / sc).max() <= 1e-6 if fin.any() else True, what
        assert (res["keep"] != 0).all()
    else:
        assert (res["keep"] == 0).all()


def test_locus_end_to_end_vs_reference_goldens(sbq_lib_path):
    from strawberry_b200 import api
    q = api.Quantifier()
    cases = json.load(gzip.open(GOLD, "rt"))
    for case in cases:
        hits = [(m, (l[0], [tuple(x) for x in l[1]]) if l else None, (r[0], [tuple(x) for x in r[1]]) if r else None)
                for m, l, r in case["hits"]]
        isoforms = [[tuple(e) for e in iso] for iso in case["isoforms"]]
        check_case(q, isoforms, hits, case["read_len"], tuple(case["spec"]), case["long_read"], case["ref"], f"golden seed {case['seed']}")
    q.close()


def test_locus_end_to_end_vs_live_reference(sbq_lib_path, oracle_mod):
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref/libsbref.so not present")
    from strawberry_b200 import api
    q = api.Quantifier()
    for seed in range(300, 360):
        isoforms, hits, rl = locusgen.random_locus(seed)
        spec = specs_for(seed, hits)
        ref = reference_table(oracle_mod, isoforms, hits, rl, spec, seed % 17 == 0)
        check_case(q, isoforms, hits, rl, spec, seed % 17 == 0, ref, f"seed {seed}")
    q.close()
