"""GPU: the on-device giant-locus generator (sbq_synth_giant, BASELINE configs[3] input) against its numpy restatement
(strawberry_b200.synth.giant_device) - bit-equal arrays - and the solve of a device-generated batch against the CPU oracle."""
import numpy as np
import pytest

from strawberry_b200 import synth
from util import FLAT_KEYS, assert_matches_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def q(sbq_lib_path):
    from strawberry_b200 import api
    qq = api.Quantifier()
    yield qq
    qq.close()


@pytest.mark.parametrize("ids,rows,kw", [([0, 5, 199], 3000, {}), ([7], 20_000, dict(iso_lo=40, iso_hi=60, mean_extra=70.0)),
                                          ([3, 2, 1, 0], 4097, dict(seed=11, mean_extra=3.0))])
def test_device_generator_equals_numpy_restatement(q, ids, rows, kw):
    q.clear()
    q.synth_giant(ids, rows, **kw)
    got = q.fetch_batch()
    ref = synth.giant_device(ids, rows, **kw)
    for k in FLAT_KEYS:
        assert got[k].dtype == ref[k].dtype and np.array_equal(got[k], ref[k]), k     # weights too: every step is exact in fp64
    st = q.stats()
    assert st["n_loci"] == len(ids) and st["nnz"] == int(ref["row_ptr"][-1]) and st["h2d_bytes"] == 0


def test_a_locus_does_not_depend_on_the_partition(q):
    """(seed, global id) alone defines a locus: generating id 5 on its own gives the same arrays as inside a larger call."""
    q.clear()
    q.synth_giant([5], 2500)
    alone = q.fetch_batch()
    q.clear()
    q.synth_giant([9, 5, 1], 2500)
    b = q.fetch_batch()
    T, rp, col, alpha, count, iso_len = synth.locus_slice(b, 1)
    T0, rp0, col0, alpha0, count0, iso_len0 = synth.locus_slice(alone, 0)
    assert T == T0 and all(np.array_equal(x, y) for x, y in ((rp, rp0), (col, col0), (alpha, alpha0), (count, count0), (iso_len, iso_len0)))


def test_device_generated_batch_solves_like_the_oracle(q, oracle_mod):
    ids, rows = [0, 1, 2], 40_000          # 1.9 M non-zeros each: the giant-locus (grid) tier
    q.clear()
    q.synth_giant(ids, rows)
    q.solve(rows * len(ids))
    q.finalize_tpm(q.fpkm_sum())
    q.download()
    res = q.results()
    assert q.stats()["loci_grid"] == len(ids)
    ref = synth.giant_device(ids, rows)
    ora = oracle_mod.quantify_batch(ref, ref["total_mapped_reads"], n_threads=3)
    assert_matches_oracle(res, ora, ref, "device-generated giant loci")
    # a re-upload of a device-only batch is refused (nothing to upload from), clear + submit works again
    from strawberry_b200 import api
    with pytest.raises(api.SbqError) as e:
        q.upload()
    assert e.value.code == api.SBQ_ERR_STATE
