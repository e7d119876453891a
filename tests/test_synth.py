import numpy as np

from strawberry_b200 import synth


def test_human_shaped_is_seeded_and_well_formed():
    a = synth.human_shaped(n_loci=500, total_fragments=100_000, seed=3)
    b = synth.human_shaped(n_loci=500, total_fragments=100_000, seed=3)
    for k in ("loc_row_off", "loc_iso_off", "row_ptr", "col", "alpha", "count", "iso_len"):
        assert np.array_equal(a[k], b[k])
    assert a["count"].sum() == 100_000 == a["total_mapped_reads"]
    T = np.diff(a["loc_iso_off"])
    assert T.min() >= 1 and T.max() <= 200
    for l in range(500):
        t, rp, col, al, cnt, il = synth.locus_slice(a, l)
        assert col.min() >= 0 and col.max() < t
        for i in range(len(cnt)):
            c = col[rp[i]:rp[i + 1]]
            assert len(c) >= 1 and (np.diff(c) > 0).all()
    assert (a["alpha"] == 5e-6).any()


def test_giant_columns_are_distinct_and_sorted():
    g = synth.giant(n_loci=2, rows_per_locus=2000, seed=1)
    for l in range(2):
        t, rp, col, al, cnt, il = synth.locus_slice(g, l)
        assert 500 <= t <= 800 and (cnt == 1).all()
        for i in range(0, 2000, 97):
            c = col[rp[i]:rp[i + 1]]
            assert (np.diff(c) > 0).all() and c.max() < t
    nnz_per_row = np.diff(g["row_ptr"]).mean()
    assert 40 < nnz_per_row < 56
