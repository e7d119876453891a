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


def covariates(batch, seed=3):
    """BASELINE config 3: per-row covariates for the bias-in-EM mode, R x 5 =
    (gc, gc^2, gc^3, log(bin_len) / 5, mean_frag_len / 1000) with gc ~ Beta(8, 8), bin_len ~ U{50..400}, mean_frag_len ~ N(250, 30)."""
    rng = np.random.default_rng(seed)
    R = int(batch["loc_row_off"][-1])
    gc = rng.beta(8.0, 8.0, R)
    return np.stack([gc, gc ** 2, gc ** 3, np.log(rng.integers(50, 401, R)) / 5.0, rng.normal(250.0, 30.0, R) / 1000.0], axis=1)


DEVICE_GENERATOR_VERSION = "sbq-synth-dev-1"
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _sm64(x):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic), as strawberry_b200/csrc/sbq_synth.cuh::sm64."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def _synth_key(seed, ids, stream):
    with np.errstate(over="ignore"):
        return _sm64(_sm64(np.uint64(seed) ^ np.uint64(0x5851F42D4C957F2D)) + np.uint64(4) * np.asarray(ids, np.uint64) + np.uint64(stream))


def _synth_u(h):
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def giant_device_iso(seed, ids, iso_lo=500, iso_hi=800):
    """T of the device-generated giant loci `ids` (host-side hash only)."""
    return (iso_lo + (_sm64(_synth_key(seed, ids, 0)) % np.uint64(iso_hi - iso_lo + 1)).astype(np.int64)).astype(np.int64)


def giant_device(locus_ids, rows_per_locus, seed=4, iso_lo=500, iso_hi=800, mean_extra=47.0):
    """numpy restatement of the ON-DEVICE giant-locus generator (sbq_synth_giant, csrc/sbq_synth.cuh): bit-equal row
    pointers, columns, weights, counts and lengths. A locus is a pure function of (seed, global locus id): T ~ U{iso_lo..iso_hi};
    every row is one fragment with k ~ 1 + Poisson(mean_extra) (CDF inversion, capped at min(T, 255)) compatible isoforms drawn
    by stratified sampling; alpha = (1 + u) 2^-(6 + e), e ~ U{0..7} (piecewise log-uniform on [1.2e-4, 3.1e-2), exact in fp64)."""
    import math
    ids = np.asarray(locus_ids, dtype=np.int64)
    L, rows = len(ids), int(rows_per_locus)
    T = giant_device_iso(seed, ids, iso_lo, iso_hi)
    cdf = np.empty(255)
    pk = math.exp(-mean_extra)
    cdf[0] = pk
    for k in range(1, 255):
        pk = pk * (mean_extra / float(k))
        cdf[k] = cdf[k - 1] + pk
    i = np.arange(rows, dtype=np.uint64)
    degs, cols, alphas, lens = [], [], [], []
    with np.errstate(over="ignore"):
        for l in range(L):
            t = int(T[l])
            u = _synth_u(_sm64(_synth_key(seed, ids[l], 1) + i))
            k = 1 + np.minimum(np.searchsorted(cdf, u, side="right"), 254)
            k = np.minimum(k, min(t, 255)).astype(np.int64)
            row_of = np.repeat(np.arange(rows, dtype=np.int64), k)
            start = np.concatenate([[0], np.cumsum(k)])
            m = np.arange(int(start[-1]), dtype=np.int64) - start[row_of]
            kk = k[row_of]
            v = _sm64(_synth_key(seed, ids[l], 2) + (np.uint64(256) * row_of.astype(np.uint64) + m.astype(np.uint64)))
            lo, hi = (m * t) // kk, ((m + 1) * t) // kk
            cols.append((lo + (_synth_u(v) * (hi - lo).astype(np.float64)).astype(np.int64)).astype(np.int32))
            w = _sm64(v)
            alphas.append(np.ldexp(1.0 + _synth_u(w), -(6 + (w & np.uint64(7)).astype(np.int64)).astype(np.int32)))
            degs.append(k)
            lens.append((400 + (_sm64(_synth_key(seed, ids[l], 3) + np.arange(t, dtype=np.uint64)) % np.uint64(7601)).astype(np.int64)).astype(np.int32))
    k_all = np.concatenate(degs) if L else np.zeros(0, np.int64)
    row_ptr = np.zeros(L * rows + 1, np.int64)
    np.cumsum(k_all, out=row_ptr[1:])
    loc_iso_off = np.zeros(L + 1, np.int64)
    np.cumsum(T, out=loc_iso_off[1:])
    return dict(loc_row_off=np.arange(L + 1, dtype=np.int64) * rows, loc_iso_off=loc_iso_off, row_ptr=row_ptr,
                col=np.concatenate(cols), alpha=np.concatenate(alphas), count=np.ones(L * rows, np.int32),
                iso_len=np.concatenate(lens), total_mapped_reads=int(L * rows),
                meta=dict(generator=DEVICE_GENERATOR_VERSION, kind="giant_device", seed=seed, locus_ids=ids.tolist(),
                          rows_per_locus=rows))


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
