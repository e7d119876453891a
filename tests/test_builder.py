"""CPU: the host class-table builder (libsbq.so, include/sbq_builder.h) against the UNMODIFIED reference's
LocusContext (oracle/_ref/libsbref.so wide seam) and against committed goldens generated from it."""
import gzip
import json
import os

import numpy as np
import pytest

import locusgen
from util import GOLDEN_DIR

GOLD = os.path.join(GOLDEN_DIR, "locus_golden.json.gz")


def product_table(isoforms, hits, read_len, model_spec, long_read=False):
    from strawberry_b200 import builder
    tfe = [locusgen.transcript_features(ex) for ex in isoforms]
    feats = [builder.pair_features(l, r) for _, l, r in hits]
    if model_spec[0] == "normal":
        model = builder.Model.normal(model_spec[1], model_spec[2])
    else:
        model = builder.Model.empirical(model_spec[1])
    tab = builder.build_locus(tfe, [(m, f) for (m, _, _), f in zip(hits, feats)], read_len=read_len, model=model, long_read=long_read)
    return feats, tab


def reference_table(oracle_mod, isoforms, hits, read_len, model_spec, long_read=False):
    tfe = [locusgen.transcript_features(ex) for ex in isoforms]
    kw = dict(mean=model_spec[1], sd=model_spec[2]) if model_spec[0] == "normal" else dict(frag_lens=model_spec[1])
    return oracle_mod.ref_locus_context(tfe, hits, read_len=read_len, long_read=long_read, total_mapped_reads=100000, **kw)


def compare(feats, tab, ref, what):
    # a11: Contig(PairedHit) features
    for h, (f, rh) in enumerate(zip(feats, ref["hits"])):
        assert [list(x) for x in f] == rh["feats"], f"{what}: hit {h} features"
    # a12: disjoint segments; a8: isoform segments and lengths
    assert [list(s) for s in tab["segs"]] == ref["segs"], f"{what}: segments"
    assert [[list(tab["segs"][s]) for s in segs] for segs in tab["iso_segs"]] == ref["iso_segs"], f"{what}: isoform segments"
    assert list(tab["iso_len"]) == ref["iso_len"], f"{what}: isoform lengths"
    # a3 / a7: classes in first-seen order, coordinates, set sizes, float mass and truncated count -- bit exact
    assert len(tab["classes"]) == len(ref["classes"]), f"{what}: class count"
    for c, (pc, rc) in enumerate(zip(tab["classes"], ref["classes"])):
        assert [list(tab["segs"][s]) for s in pc["coords"]] == rc["coords"], f"{what}: class {c} coords"
        assert pc["nfrag"] == rc["nfrags"] and pc["count"] == rc["count"], f"{what}: class {c} count {pc} vs {rc['count']}"
        assert np.float32(pc["mass"]) == np.float32(rc["count_f"]), f"{what}: class {c} float mass"
    # iso -> classes map and a4/a5 weights
    iso2 = {}
    for c in range(len(tab["classes"])):
        for k in range(tab["row_ptr"][c], tab["row_ptr"][c + 1]):
            iso2.setdefault(int(tab["col"][k]), []).append(c)
            w_ref = ref["classes"][c]["weights"][str(int(tab["col"][k]))]
            w = float(tab["alpha"][k])
            assert abs(w - w_ref) <= 1e-12 * max(abs(w_ref), 1e-300), f"{what}: alpha class {c} iso {tab['col'][k]}: {w} vs {w_ref}"
        assert len(ref["classes"][c]["weights"]) == tab["row_ptr"][c + 1] - tab["row_ptr"][c]
    assert {str(k): v for k, v in iso2.items()} == {k: v for k, v in ref["iso2bins"].items()}, f"{what}: iso2bins"


def specs_for(seed, hits):
    rng = np.random.default_rng(seed + 7)
    kind = seed % 3
    if kind == 0:
        return ("normal", float(rng.choice([180, 220, 300])), float(rng.choice([20, 40, 80])))
    if kind == 1:
        return ("emp", [int(x) for x in np.clip(rng.normal(230, 45, 300), 60, 600).astype(int)], None)
    return ("normal", 200.0, 80.0)


def test_builder_matches_compiled_reference(oracle_mod, sbq_lib_path):
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref/libsbref.so not built (needs the reference checkout)")
    n_multi = 0
    for seed in range(160):
        isoforms, hits, rl = locusgen.random_locus(seed)
        spec = specs_for(seed, hits)
        long_read = seed % 17 == 0
        ref = reference_table(oracle_mod, isoforms, hits, rl, spec, long_read)
        feats, tab = product_table(isoforms, hits, rl, spec, long_read)
        compare(feats, tab, ref, f"seed {seed}")
        n_multi += sum(len(c["coords"]) > 4 for c in tab["classes"])
    assert n_multi > 20, "the sweep must exercise the > 4 segment effective-length branch"


def test_builder_matches_committed_goldens(sbq_lib_path):
    cases = json.load(gzip.open(GOLD, "rt"))
    assert len(cases) >= 20
    for case in cases:
        hits = [(m, tuple(l) if l else None, tuple(r) if r else None) for m, l, r in case["hits"]]
        hits = [(m, (l[0], [tuple(x) for x in l[1]]) if l else None, (r[0], [tuple(x) for x in r[1]]) if r else None) for m, l, r in hits]
        isoforms = [[tuple(e) for e in iso] for iso in case["isoforms"]]
        spec = tuple(case["spec"])
        feats, tab = product_table(isoforms, hits, case["read_len"], spec, case["long_read"])
        compare(feats, tab, case["ref"], f"golden seed {case['seed']}")


def test_dedup_under_code_blind_comparator(oracle_mod, sbq_lib_path):
    """a 5S50M variant of an existing 50M pair collapses into the same _frags element and
    its mass is dropped from the class count."""
    isoforms = [[(1001, 1200), (1501, 1700), (2001, 2300)], [(1001, 1200), (2001, 2300)]]
    hits = [(2.0, (1050, [(0, 50)]), (1120, [(0, 50)])), (3.0, (1050, [(4, 5), (0, 50)]), (1120, [(0, 50)]))]
    spec = ("normal", 200.0, 20.0)
    feats, tab = product_table(isoforms, hits, 50, spec)
    assert tab["classes"][0]["nfrag"] == 1 and tab["classes"][0]["count"] == 2
    if oracle_mod.have_ref():
        compare(feats, tab, reference_table(oracle_mod, isoforms, hits, 50, spec), "dedup")
    # quirk: mates that abut exactly (gap 0) take the overlap-merge path, fail its "f.right < next.left" test and are
    # dropped as ref_id -1 (include/contig.h:111-138)
    feats2, tab2 = product_table(isoforms, [(1.0, (1100, [(0, 50)]), (1150, [(0, 50)]))], 50, spec)
    assert feats2 == [[]] and tab2["n_dropped"] == 1 and not tab2["classes"]


def test_toy_locus_alpha_known_answers(sbq_lib_path):
    """The four alpha values SURVEY Appendix A.2 quotes from the oracle for the 3-exon / 2-isoform toy locus."""
    isoforms = [[(1001, 1200), (1501, 1700), (2001, 2300)], [(1001, 1200), (2001, 2300)]]
    hits = [(1.0, (1180, [(0, 21), (3, 300), (0, 29)]), (1600, [(0, 50)])), (3.0, (1150, [(0, 50)]), (2050, [(0, 50)])),
            (1.0, (1181, [(0, 20), (3, 800), (0, 30)]), (2100, [(0, 50)]))]
    _, tab = product_table(isoforms, hits, 50, ("normal", 200.0, 20.0))
    got = sorted(float(a) for a in tab["alpha"])
    for want in (0.36791757561521959, 0.64130904079814721):
        assert any(abs(g - want) < 1e-15 for g in got), (want, got)


def test_effective_len_branches(sbq_lib_path):
    from strawberry_b200 import builder
    assert builder.effective_len([300], [], 200, 50) == 101                       # one segment: len - fl + 1
    assert builder.effective_len([100, 100], [], 150, 50) == 51                   # two segments, no_gap_ef
    assert builder.effective_len([100, 100], [], 250, 50) == 0                    # longer than both
    # three segments: hitting all three + skipping the middle one partition the two-end count
    full = builder.effective_len([100, 30, 100], [], 180, 50)
    skip = builder.effective_len([100, 30, 100], [1], 180, 50)
    assert full >= 0 and skip >= 0 and full + skip > 0
