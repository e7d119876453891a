"""Seeded synthetic locus batches at the CSR level.

The layout produced here is the flat batch ``sbq_submit_flat`` takes (include/sbq.h):

    loc_row_off int64[L+1]   first row of each locus
    loc_iso_off int64[L+1]   first isoform of each locus
    row_ptr     int64[R+1]   global non-zero offsets, one CSR for the whole batch
    col         int32[nnz]   isoform index LOCAL to the locus
    alpha       f64[nnz]     class weight alpha_ij   (LocusContext::set_theory_bin_weight output)
    count       int32[R]     class counts n_i         ((int)ExonBin::read_count())
    iso_len     int32[T]     isoform lengths L_j      (Contig::exonic_length())

GENERATOR_VERSION goes into every report next to the seed.
"""
import numpy as np

GENERATOR_VERSION = "sbq-synth-1"


def _largest_remainder(weights, total):
    """Round ``weights * total / sum(weights)`` to non-negative ints that sum to ``total``."""
    w = np.asarray(weights, dtype=np.float64)
    x = w * (total / w.sum())
    f = np.floor(x).astype(np.int64)
    rem = int(total - f.sum())
    if rem > 0:
        order = np.argsort(-(x - f), kind="stable")[:rem]
        f[order] += 1
    return f


def human_shaped(n_loci=20000, total_fragments=10_000_000, seed=2, max_iso=200, max_rows=5000,
                 tiny_alpha_rate=0.02):
    """BASELINE config 2 (and config 5 with n_loci=60000, total_fragments=10**8).

    T_l = min(ceil(Pareto(1.3)), max_iso); R_l = clamp(ceil(T_l * LogNormal(1.6, 0.8)), 1, max_rows);
    row i is compatible with k ~ 1 + Binomial(T_l - 1, p_l) isoforms, p_l ~ Beta(2, 5);
    alpha ~ 10^U(-4, -1.5), a ``tiny_alpha_rate`` share of rows gets one 5e-6 entry (row filter);
    locus totals ~ Pareto(1.1) normalised to ``total_fragments``, split over rows by Dirichlet(0.3),
    integer with the sum preserved; isoform lengths ~ U{400..8000}.
    """
    rng = np.random.default_rng(seed)
    T = np.minimum(np.ceil(rng.pareto(1.3, n_loci) + 1.0), max_iso).astype(np.int64)
    R = np.clip(np.ceil(T * rng.lognormal(1.6, 0.8, n_loci)), 1, max_rows).astype(np.int64)
    p = rng.beta(2.0, 5.0, n_loci)
    N = _largest_remainder(rng.pareto(1.1, n_loci) + 1.0, total_fragments)

    loc_row_off = np.zeros(n_loci + 1, np.int64)
    loc_iso_off = np.zeros(n_loci + 1, np.int64)
    np.cumsum(R, out=loc_row_off[1:])
    np.cumsum(T, out=loc_iso_off[1:])
    total_rows = int(loc_row_off[-1])

    row_nnz = np.empty(total_rows, np.int64)
    cols, counts = [], np.empty(total_rows, np.int32)
    for l in range(n_loci):
        t, r = int(T[l]), int(R[l])
        r0 = int(loc_row_off[l])
        k = 1 + rng.binomial(t - 1, p[l], r) if t > 1 else np.ones(r, np.int64)
        row_nnz[r0:r0 + r] = k
        if t == 1:
            cols.append(np.zeros(r, np.int32))
        else:
            # k smallest random keys per row -> k distinct isoforms, emitted in ascending order
            keys = rng.random((r, t))
            kth = np.sort(keys, axis=1)[np.arange(r), k - 1]
            mask = keys <= kth[:, None]
            cols.append(np.nonzero(mask)[1].astype(np.int32))
        counts[r0:r0 + r] = _largest_remainder(rng.dirichlet(np.full(r, 0.3)) + 1e-300, int(N[l]))
    col = np.concatenate(cols)
    row_ptr = np.zeros(total_rows + 1, np.int64)
    np.cumsum(row_nnz, out=row_ptr[1:])
    nnz = int(row_ptr[-1])
    assert nnz == len(col)
    alpha = 10.0 ** rng.uniform(-4.0, -1.5, nnz)
    tiny_rows = np.nonzero(rng.random(total_rows) < tiny_alpha_rate)[0]
    alpha[row_ptr[tiny_rows]] = 5e-6
    iso_len = rng.integers(400, 8001, int(loc_iso_off[-1])).astype(np.int32)
    return dict(loc_row_off=loc_row_off, loc_iso_off=loc_iso_off, row_ptr=row_ptr, col=col, alpha=alpha,
                count=counts, iso_len=iso_len, total_mapped_reads=int(N.sum()),
                meta=dict(generator=GENERATOR_VERSION, kind="human_shaped", seed=seed, n_loci=n_loci,
                          total_fragments=int(total_fragments)))


def giant(n_loci=2, rows_per_locus=200_000, seed=4, iso_lo=500, iso_hi=800, mean_extra=47.0):
    """BASELINE config 4 shape at a caller-chosen row count: T ~ U{iso_lo..iso_hi}; every row is one
    fragment (n_i = 1) compatible with k ~ 1 + Poisson(mean_extra) isoforms. Columns are drawn by
    stratified sampling (one per stratum of width T/k) so they are distinct and ascending without
    an R x T key matrix."""
    rng = np.random.default_rng(seed)
    T = rng.integers(iso_lo, iso_hi + 1, n_loci).astype(np.int64)
    loc_row_off = np.arange(n_loci + 1, dtype=np.int64) * rows_per_locus
    loc_iso_off = np.zeros(n_loci + 1, np.int64)
    np.cumsum(T, out=loc_iso_off[1:])
    total_rows = n_loci * rows_per_locus
    k = 1 + rng.poisson(mean_extra, total_rows).astype(np.int64)
    Trow = np.repeat(T, rows_per_locus)
    k = np.minimum(k, Trow)
    row_ptr = np.zeros(total_rows + 1, np.int64)
    np.cumsum(k, out=row_ptr[1:])
    nnz = int(row_ptr[-1])
    row_of = np.repeat(np.arange(total_rows, dtype=np.int64), k)
    m = np.arange(nnz, dtype=np.int64) - row_ptr[row_of]
    kk, tt = k[row_of], Trow[row_of]
    lo = (m * tt) // kk
    hi = ((m + 1) * tt) // kk
    col = (lo + np.floor(rng.random(nnz) * (hi - lo)).astype(np.int64)).astype(np.int32)
    alpha = 10.0 ** rng.uniform(-4.0, -1.5, nnz)
    count = np.ones(total_rows, np.int32)
    iso_len = rng.integers(400, 8001, int(loc_iso_off[-1])).astype(np.int32)
    return dict(loc_row_off=loc_row_off, loc_iso_off=loc_iso_off, row_ptr=row_ptr, col=col, alpha=alpha,
                count=count, iso_len=iso_len, total_mapped_reads=int(total_rows),
                meta=dict(generator=GENERATOR_VERSION, kind="giant", seed=seed, n_loci=n_loci,
                          rows_per_locus=rows_per_locus))


def collapsed_giant(n_loci=1, rows=20000, n_iso=500, density=0.05, seed=4):
    """Config 4's 'collapsed-faithful' variant the dense reference can still run."""
    rng = np.random.default_rng(seed)
    loc_row_off = np.arange(n_loci + 1, dtype=np.int64) * rows
    loc_iso_off = np.arange(n_loci + 1, dtype=np.int64) * n_iso
    cols, nnz_row = [], []
    for _ in range(n_loci * rows):
        k = max(1, rng.binomial(n_iso, density))
        cols.append(np.sort(rng.choice(n_iso, k, replace=False)).astype(np.int32))
        nnz_row.append(k)
    col = np.concatenate(cols)
    row_ptr = np.zeros(n_loci * rows + 1, np.int64)
    np.cumsum(nnz_row, out=row_ptr[1:])
    alpha = 10.0 ** rng.uniform(-4.0, -1.5, len(col))
    count = rng.integers(0, 100, n_loci * rows).astype(np.int32)
    iso_len = rng.integers(400, 8001, n_loci * n_iso).astype(np.int32)
    return dict(loc_row_off=loc_row_off, loc_iso_off=loc_iso_off, row_ptr=row_ptr, col=col, alpha=alpha,
                count=count, iso_len=iso_len, total_mapped_reads=int(count.sum()),
                meta=dict(generator=GENERATOR_VERSION, kind="collapsed_giant", seed=seed, n_loci=n_loci))


def locus_slice(batch, l):
    """One locus of a flat batch as (T, row_ptr_local, col, alpha, count, iso_len)."""
    r0, r1 = int(batch["loc_row_off"][l]), int(batch["loc_row_off"][l + 1])
    t0, t1 = int(batch["loc_iso_off"][l]), int(batch["loc_iso_off"][l + 1])
    rp = batch["row_ptr"][r0:r1 + 1]
    k0, k1 = int(rp[0]), int(rp[-1])
    return (t1 - t0, rp - k0, batch["col"][k0:k1], batch["alpha"][k0:k1], batch["count"][r0:r1],
            batch["iso_len"][t0:t1])


def densify(T, row_ptr, col, alpha):
    R = len(row_ptr) - 1
    a = np.zeros((R, T))
    rows = np.repeat(np.arange(R), np.diff(row_ptr))
    a[rows, col] = alpha
    return a


def concat(batches):
    """Concatenate flat batches (used to mix shapes in tests)."""
    out = {}
    lro, lio, rp = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)], [np.zeros(1, np.int64)]
    n_row = n_iso = nnz = 0
    for b in batches:
        lro.append(np.asarray(b["loc_row_off"][1:], np.int64) + n_row)
        lio.append(np.asarray(b["loc_iso_off"][1:], np.int64) + n_iso)
        rp.append(np.asarray(b["row_ptr"][1:], np.int64) + nnz)
        n_row += int(b["loc_row_off"][-1])
        n_iso += int(b["loc_iso_off"][-1])
        nnz += int(b["row_ptr"][-1])
    out["loc_row_off"], out["loc_iso_off"], out["row_ptr"] = map(np.concatenate, (lro, lio, rp))
    for k, dt in (("col", np.int32), ("alpha", np.float64), ("count", np.int32), ("iso_len", np.int32)):
        out[k] = np.concatenate([np.asarray(b[k], dtype=dt) for b in batches])
    out["total_mapped_reads"] = int(sum(b["total_mapped_reads"] for b in batches))
    out["meta"] = dict(generator=GENERATOR_VERSION, kind="concat")
    return out
