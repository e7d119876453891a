"""Seeded synthetic loci at the alignment level (SURVEY Appendix C style): a gene with alternative
isoforms (exon skipping, alternative 5'/3' ends -> partially overlapping exons) and paired / single
reads with M/N/S/I/D CIGARs, collapsed with a multiplicity. Used by the builder parity tests."""
import numpy as np

M, I, D, N, S = 0, 1, 2, 3, 4


def make_gene(rng, n_exon, n_iso, exon_len=(30, 400), intron_len=(60, 900), start=1000):
    base, pos = [], start
    for _ in range(n_exon):
        l = int(rng.integers(exon_len[0], exon_len[1] + 1))
        base.append((pos, pos + l - 1))
        pos += l + int(rng.integers(intron_len[0], intron_len[1] + 1))
    isoforms = []
    for _ in range(n_iso):
        keep = [e for e in base if rng.random() < 0.75]
        if not keep:
            keep = [base[int(rng.integers(0, n_exon))]]
        exons = []
        for (l, r) in keep:
            if rng.random() < 0.25 and r - l > 24:      # alternative splice site: shift one boundary inwards
                if rng.random() < 0.5:
                    l += int(rng.integers(1, (r - l) // 2))
                else:
                    r -= int(rng.integers(1, (r - l) // 2))
            exons.append((l, r))
        if exons not in isoforms:
            isoforms.append(exons)
    return isoforms


def transcript_features(exons):
    feats = []
    for k, (l, r) in enumerate(exons):
        if k:
            pl, pr = exons[k - 1]
            feats.append((1, pr + 1, l - pr - 1))
        feats.append((0, l, r - l + 1))
    return feats


def _blocks(exons, t0, t1):
    """genomic blocks of transcript interval [t0, t1)"""
    out, acc = [], 0
    for (l, r) in exons:
        ln = r - l + 1
        a, b = max(t0, acc), min(t1, acc + ln)
        if a < b:
            out.append((l + a - acc, l + b - acc - 1))
        acc += ln
    return out


def _cigar(blocks, rng, noise):
    ops = []
    for k, (l, r) in enumerate(blocks):
        if k:
            ops.append((N, l - blocks[k - 1][1] - 1))
        ln = r - l + 1
        if noise and ln > 12 and rng.random() < 0.08:       # insertion inside a match block
            a = int(rng.integers(3, ln - 3))
            ops += [(M, a), (I, int(rng.integers(1, 4))), (M, ln - a)]
        elif noise and ln > 12 and rng.random() < 0.08:     # deletion: reference span stays, read bases shrink
            a = int(rng.integers(3, ln - 6))
            d = int(rng.integers(1, 4))
            ops += [(M, a), (D, d), (M, ln - a - d)]
        else:
            ops.append((M, ln))
    if noise and rng.random() < 0.1:
        ops = [(S, int(rng.integers(1, 6)))] + ops
    if noise and rng.random() < 0.1:
        ops = ops + [(S, int(rng.integers(1, 6)))]
    return (blocks[0][0], ops)


def make_hits(rng, isoforms, n_frag, read_len, frag_mean=220, frag_sd=40, noise=True, single_rate=0.08):
    """-> list of (mass, left_mate, right_mate) sorted by (left, right) like HitCluster::uniq_hits()."""
    lens = [sum(r - l + 1 for l, r in ex) for ex in isoforms]
    frags = {}
    for _ in range(n_frag):
        t = int(rng.integers(0, len(isoforms)))
        L = lens[t]
        if L < read_len + 2:
            continue
        fl = int(np.clip(rng.normal(frag_mean, frag_sd), read_len, L))
        s = int(rng.integers(0, L - fl + 1))
        lb = _blocks(isoforms[t], s, s + read_len)
        rb = _blocks(isoforms[t], s + fl - read_len, s + fl)
        left, right = _cigar(lb, rng, noise), _cigar(rb, rng, noise)
        u = rng.random()
        if u < single_rate / 2:
            right = None
        elif u < single_rate:
            left = None
        key = (left[0] if left else -1, tuple(left[1]) if left else (), right[0] if right else -1, tuple(right[1]) if right else ())
        frags[key] = frags.get(key, 0) + 1
    hits = []
    for (lp, lo, rp, ro), mult in frags.items():
        left = (lp, list(lo)) if lp >= 0 else None
        right = (rp, list(ro)) if rp >= 0 else None
        hits.append((float(mult), left, right))

    def span(h):
        ms = [m for m in (h[1], h[2]) if m is not None]
        lo = min(m[0] for m in ms)
        hi = max(m[0] + sum(l for o, l in m[1] if o in (M, N, D)) - 1 for m in ms)
        return (lo, hi)
    hits.sort(key=span)
    return hits


def random_locus(seed, **kw):
    rng = np.random.default_rng(seed)
    n_exon = int(rng.integers(1, kw.get("max_exon", 10)))
    n_iso = int(rng.integers(1, kw.get("max_iso", 6)))
    small = rng.random() < 0.35                           # many short exons -> classes spanning > 4 segments
    isoforms = make_gene(rng, n_exon + (6 if small else 0), n_iso, exon_len=(18, 70) if small else (30, 400),
                         intron_len=(30, 200) if small else (60, 900))
    read_len = int(rng.choice([36, 50, 75]))
    hits = make_hits(rng, isoforms, int(rng.integers(5, kw.get("max_frag", 400))), read_len,
                     frag_mean=float(rng.choice([160, 220, 300])), frag_sd=float(rng.choice([15, 40, 80])),
                     noise=kw.get("noise", True))
    return isoforms, hits, read_len
