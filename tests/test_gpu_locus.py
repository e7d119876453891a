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


def test_gpu_weights_match_host_builder(sbq_lib_path):
    """SURVEY 8f.1 (weight half of the class-table build on the device): alpha computed by weights_kernel from the
    builder's per-entry descriptors equals the host builder's alpha, and the EM on top of it gives the same answer."""
    from strawberry_b200 import api, builder
    import numpy as np
    for kind in ("normal", "emp"):
        rng = np.random.default_rng(5)
        model = builder.Model.normal(230.0, 45.0) if kind == "normal" else builder.Model.empirical(
            [int(x) for x in np.clip(rng.normal(230, 45, 400), 60, 600).astype(int)])
        host_alpha, tables, host_loci = [], [], []
        n_multi = 0
        for seed in range(400, 460):
            isoforms, hits, rl = locusgen.random_locus(seed)
            if rl != 50:
                continue                      # one read length per context
            tfe = [locusgen.transcript_features(ex) for ex in isoforms]
            feats = [(m, builder.pair_features(l, r)) for m, l, r in hits]
            ht = builder.build_locus(tfe, feats, read_len=rl, model=model)
            dt = builder.build_locus(tfe, feats, read_len=rl, model=model, defer_weights=True)
            assert np.array_equal(ht["col"], dt["col"]) and np.array_equal(ht["count"], dt["count"])
            host_alpha.append(ht["alpha"])
            host_loci.append((ht["n_iso"], ht["row_ptr"], ht["col"], ht["alpha"], ht["count"], ht["iso_len"]))
            tables.append(dt["table"])
            n_multi += sum(len(c["coords"]) > 4 for c in ht["classes"])
        assert len(tables) >= 10 and n_multi > 0
        q = api.Quantifier()
        q.set_insert_model(model, 50)
        q.submit_deferred(tables)
        q.run(100000)
        gpu = q.results()
        alpha_gpu = q.fetch_alpha()
        alpha_host = np.concatenate(host_alpha)
        rel = np.abs(alpha_gpu - alpha_host) / np.maximum(np.abs(alpha_host), 1e-300)
        assert rel[alpha_host != 0].max() < 1e-12 and (alpha_gpu[alpha_host == 0] == 0).all(), rel.max()
        assert q.stats()["weights_ms"] > 0
        q2 = api.Quantifier()
        q2.submit(host_loci)
        q2.run(100000)
        host = q2.results()
        assert np.array_equal(gpu["status"], host["status"]) and np.array_equal(gpu["iters"], host["iters"])
        assert np.allclose(gpu["theta"], host["theta"], rtol=1e-9, atol=1e-9)
        q.close(), q2.close()
