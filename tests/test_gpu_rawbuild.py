"""GPU: fragment-class assignment ON THE DEVICE (sbq_submit_raw, strawberry_b200/csrc/sbq_rawbuild.cuh; SURVEY 8f.1, rows a3 /
a7 / a9) against the host builder - itself bit-exact against the compiled reference's LocusContext (tests/test_builder.py) - and,
where oracle/_ref/libsbref.so is present, against that reference directly: class coordinates and first-seen ids, the class of
every hit, set sizes under the code-blind comparator, float masses and integer counts, the class x isoform CSR (all bit-exact)
and alpha (<= 1e-12 relative: the same weights_kernel that serves the deferred host tables)."""
import numpy as np
import pytest

import locusgen
from test_builder import reference_table, specs_for

pytestmark = pytest.mark.gpu


def inputs(seed, **kw):
    from strawberry_b200 import builder
    isoforms, hits, rl = locusgen.random_locus(seed, **kw)
    tfe = [locusgen.transcript_features(ex) for ex in isoforms]
    feats = [builder.pair_features(l, r) for _, l, r in hits]
    return isoforms, hits, rl, tfe, [(m, f) for (m, _, _), f in zip(hits, feats)]


def device_tables(loci, model, read_len, long_read=False):
    """loci: list of (tfe, hit list). One raw batch -> per-locus dicts in the shape builder.build_locus returns."""
    from strawberry_b200 import api, builder
    q = api.Quantifier()
    q.set_insert_model(model, read_len)
    for tfe, hl in loci:
        builder.submit_raw(q, tfe, hl, read_len=read_len, long_read=long_read)
    q.upload()
    n_hit = sum(len(hl) for _, hl in loci)
    cls = builder.fetch_raw_classes(q, n_hit)
    b = q.fetch_batch()
    out, h0 = [], 0
    for l, (tfe, hl) in enumerate(loci):
        r0, r1 = int(b["loc_row_off"][l]), int(b["loc_row_off"][l + 1])
        k0 = int(b["row_ptr"][r0])
        classes = []
        for c in range(r0, r1):
            rep = int(cls["class_rep"][c])
            classes.append(dict(coords=[int(x) for x in cls["coords"][rep][:cls["ncoord"][rep]]], count=int(b["count"][c]),
                                mass=float(cls["class_mass"][c]), nfrag=int(cls["class_nfrag"][c])))
        out.append(dict(classes=classes, row_ptr=b["row_ptr"][r0:r1 + 1] - k0, col=b["col"][k0:int(b["row_ptr"][r1])],
                        alpha=b["alpha"][k0:int(b["row_ptr"][r1])], hit_class=cls["hit_class"][h0:h0 + len(hl)],
                        iso_len=b["iso_len"][int(b["loc_iso_off"][l]):int(b["loc_iso_off"][l + 1])]))
        h0 += len(hl)
    return q, out


def assert_same_table(dev, host, what):
    assert len(dev["classes"]) == len(host["classes"]), f"{what}: class count {len(dev['classes'])} vs {len(host['classes'])}"
    for c, (d, h) in enumerate(zip(dev["classes"], host["classes"])):
        assert d["coords"] == h["coords"], f"{what}: class {c} coordinates"
        assert d["nfrag"] == h["nfrag"] and d["count"] == h["count"], f"{what}: class {c} count {d} vs {h}"
        assert np.float32(d["mass"]) == np.float32(h["mass"]), f"{what}: class {c} float mass"
    assert np.array_equal(dev["hit_class"], host["hit_class"]), f"{what}: class of every hit"
    assert np.array_equal(dev["row_ptr"], host["row_ptr"]) and np.array_equal(dev["col"], host["col"]), f"{what}: CSR structure"
    assert np.array_equal(dev["iso_len"], host["iso_len"])
    rel = np.abs(dev["alpha"] - host["alpha"]) / np.maximum(np.abs(host["alpha"]), 1e-300)
    assert len(rel) == 0 or rel.max() <= 1e-12, f"{what}: alpha rel err {rel.max():.2e}"


@pytest.mark.parametrize("group", [0, 1, 2])
def test_device_class_tables_match_host_builder_and_reference(oracle_mod, sbq_lib_path, group):
    """The 160-locus sweep of tests/test_builder.py, batched by (read length, insert model): one raw upload per batch."""
    from strawberry_b200 import builder
    seeds = [s for s in range(160) if s % 3 == group and s % 17 != 0]
    by_key = {}
    for s in seeds:
        isoforms, hits, rl, tfe, hl = inputs(s)
        spec = specs_for(s, hits)
        key = (rl,) + ((spec[0], spec[1], spec[2]) if spec[0] == "normal" else ("emp", s))
        by_key.setdefault(key, []).append((s, isoforms, hits, rl, tfe, hl, spec))
    n_cls = 0
    for key, items in by_key.items():
        spec, rl = items[0][6], items[0][3]
        model = builder.Model.normal(spec[1], spec[2]) if spec[0] == "normal" else builder.Model.empirical(spec[1])
        q, dev = device_tables([(it[4], it[5]) for it in items], model, rl)
        q.close()
        for it, d in zip(items, dev):
            host = builder.build_locus(it[4], it[5], read_len=rl, model=model)
            assert_same_table(d, host, f"seed {it[0]}")
            n_cls += len(host["classes"])
            if oracle_mod.have_ref() and it[0] % 5 == 0:      # and against the compiled reference itself
                ref = reference_table(oracle_mod, it[1], it[2], rl, spec)
                assert len(ref["classes"]) == len(d["classes"])
                for c, rc in enumerate(ref["classes"]):
                    assert d["classes"][c]["nfrag"] == rc["nfrags"] and d["classes"][c]["count"] == rc["count"], (it[0], c)
                    assert np.float32(d["classes"][c]["mass"]) == np.float32(rc["count_f"])
    assert n_cls > 200


def test_device_class_tables_long_read_and_code_blind_dedup(sbq_lib_path):
    from strawberry_b200 import builder
    model = builder.Model.normal(200.0, 20.0)
    # the 5S50M variant of an existing 50M pair is the same _frags element: its mass is dropped;
    # mates that abut exactly are dropped as ref_id -1
    isoforms = [[(1001, 1200), (1501, 1700), (2001, 2300)], [(1001, 1200), (2001, 2300)]]
    tfe = [locusgen.transcript_features(ex) for ex in isoforms]
    raw = [(2.0, (1050, [(0, 50)]), (1120, [(0, 50)])), (3.0, (1050, [(4, 5), (0, 50)]), (1120, [(0, 50)])), (1.0, (1100, [(0, 50)]), (1150, [(0, 50)]))]
    hl = [(m, builder.pair_features(l, r)) for m, l, r in raw]
    for long_read in (False, True):
        q, dev = device_tables([(tfe, hl)], model, 50, long_read=long_read)
        q.close()
        host = builder.build_locus(tfe, hl, read_len=50, model=model, long_read=long_read)
        assert_same_table(dev[0], host, f"dedup long_read={long_read}")
        assert dev[0]["classes"][0]["nfrag"] == 1 and dev[0]["classes"][0]["count"] == 2 and list(dev[0]["hit_class"]) == [0, 0, -1]


def test_device_class_tables_big_loci_and_em(oracle_mod, sbq_lib_path):
    """Loci with >= 10 000 collapsed hits: device table == host table, and the EM on the device-built batch == the oracle on
    the host-built CSR (status, iterations, theta / FPKM / frac within 1e-6)."""
    from strawberry_b200 import builder, synth
    from util import assert_matches_oracle
    model = builder.Model.normal(220.0, 40.0)
    rl, loci = 50, []
    for seed in (1001, 1002, 1003):
        isoforms, _, _, tfe, _ = inputs(seed, max_exon=14, max_iso=9, max_frag=50)
        hits = locusgen.make_hits(np.random.default_rng(seed), isoforms, 60000, rl, frag_mean=220, frag_sd=40, noise=True)
        loci.append((tfe, [(m, builder.pair_features(l, r)) for m, l, r in hits]))
    assert max(len(h) for _, h in loci) >= 10000
    q, dev = device_tables(loci, model, rl)
    parts = []
    for (tfe, hl), d in zip(loci, dev):
        host = builder.build_locus(tfe, hl, read_len=rl, model=model)
        assert_same_table(d, host, f"big locus with {len(hl)} hits")
        parts.append(dict(loc_row_off=np.array([0, len(host["count"])]), loc_iso_off=np.array([0, host["n_iso"]]), row_ptr=host["row_ptr"], col=host["col"],
                          alpha=host["alpha"], count=host["count"], iso_len=host["iso_len"], total_mapped_reads=1_000_000))
    b = synth.concat(parts)
    q.solve(1_000_000)
    q.finalize_tpm(q.fpkm_sum())
    q.download()
    res = q.results()
    q.close()
    ora = oracle_mod.quantify_batch(b, 1_000_000)
    assert_matches_oracle(res, ora, b, "EM on the device-built class table")


def test_fractional_masses_are_refused(sbq_lib_path):
    """Masses that are not multiples of 1/2 (--allow-multimapped-hits) would make the float class mass depend on the order of
    the reference's std::set: the device builder refuses such a batch instead of guessing."""
    from strawberry_b200 import api, builder
    isoforms, hits, rl, tfe, hl = inputs(5)
    hl = [(m / 3.0, f) for m, f in hl]
    q = api.Quantifier()
    q.set_insert_model(builder.Model.normal(200.0, 40.0), rl)
    builder.submit_raw(q, tfe, hl, read_len=rl)
    with pytest.raises(api.SbqError) as e:
        q.upload()
    assert e.value.code == api.SBQ_ERR_UNSUPPORTED
    q.close()
