"""CPU: the plain-C restatement (oracle/sbq_oracle.c) against the reference's golden outputs and,
when the compiled reference is present (oracle/_ref/libsbref.so), against the reference itself."""
import numpy as np
import pytest

from strawberry_b200 import synth
from util import load_golden


def test_oracle_matches_reference_golden(oracle_mod):
    b, theta_ref, rc_ref, iters, status = load_golden()
    res = oracle_mod.quantify_batch(b, b["total_mapped_reads"])
    assert np.array_equal(res["status"], status)
    assert np.array_equal(res["iters"], iters)
    # rc bits of the reference: init ok (1) / run ok (2)
    exp_rc = np.array([3, 3, 1, 0])[res["status"]]
    assert np.array_equal(exp_rc, rc_ref)
    rel = np.abs(res["theta"] - theta_ref) / np.maximum(np.abs(theta_ref), 1e-300)
    assert rel.max() < 1e-11, rel.max()
    assert set(np.unique(status)) == {0, 1, 2, 3}, "golden must exercise every locus outcome"


def test_dense_and_csr_restatements_agree_bitwise(oracle_mod):
    b = synth.human_shaped(n_loci=200, total_fragments=300_000, seed=7, max_rows=300)
    for l in range(200):
        T, rp, col, al, cnt, _ = synth.locus_slice(b, l)
        st_s, th_s, it_s = oracle_mod.em_csr(T, rp, col, al, cnt)
        st_d, th_d, it_d = oracle_mod.em_dense(cnt, synth.densify(T, rp, col, al))
        assert (st_s, it_s) == (st_d, it_d)
        assert np.array_equal(th_s, th_d)


def test_restatement_vs_compiled_reference(oracle_mod):
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref/libsbref.so not built (needs the reference checkout)")
    b = synth.human_shaped(n_loci=400, total_fragments=500_000, seed=23, max_rows=800)
    worst = 0.0
    for l in range(400):
        T, rp, col, al, cnt, _ = synth.locus_slice(b, l)
        rc, th_ref = oracle_mod.ref_em(cnt, synth.densify(T, rp, col, al))
        st, th, _ = oracle_mod.em_csr(T, rp, col, al, cnt)
        assert rc == {0: 3, 1: 3, 2: 1, 3: 0}[st]
        worst = max(worst, float(np.max(np.abs(th - th_ref) / np.maximum(np.abs(th_ref), 1e-300))))
    assert worst < 1e-11, worst


def test_batch_driver_threads_and_tpm(oracle_mod):
    b = synth.human_shaped(n_loci=300, total_fragments=200_000, seed=5, max_rows=200)
    r1 = oracle_mod.quantify_batch(b, b["total_mapped_reads"], min_iso_frac=0.01, n_threads=1)
    r4 = oracle_mod.quantify_batch(b, b["total_mapped_reads"], min_iso_frac=0.01, n_threads=4)
    for k in ("theta", "fpkm", "frac", "tpm", "keep", "iters", "status"):
        assert np.array_equal(r1[k], r4[k], equal_nan=True), k
    kept = r1["keep"] != 0
    assert abs(np.nansum(r1["tpm"][kept]) - 1e6) < 1e-3          # src/alignments.cpp:1821-1829
    assert (r1["frac"][kept] >= 0.01).all() and (~kept).any()     # src/estimate.cpp:346-355


def test_mass_conservation_property(oracle_mod):
    """Every EM update conserves the kept mass: sum_j theta_j == sum of kept-row counts (OK loci, iters > 1)."""
    b = synth.human_shaped(n_loci=300, total_fragments=200_000, seed=9, max_rows=200)
    r = oracle_mod.quantify_batch(b, b["total_mapped_reads"])
    for l in np.nonzero((r["status"] == 0) & (r["iters"] > 1))[0]:
        T, rp, col, al, cnt, _ = synth.locus_slice(b, l)
        kept = np.array([(al[rp[i]:rp[i + 1]] > 1e-5).any() for i in range(len(cnt))])
        t0 = b["loc_iso_off"][l]
        assert abs(r["theta"][t0:t0 + T].sum() - cnt[kept].sum()) <= 1e-9 * max(1, cnt.sum())
