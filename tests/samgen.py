"""Seeded SAM + GTF generator for the end-to-end integration test: genes laid out on one
chromosome, alternative isoforms, paired-end M/N reads sampled from the isoforms, coordinate-sorted SAM with
NH:i:1 and XS:A:+ on spliced mates, GTF with transcript and exon lines."""
import numpy as np

import locusgen


def _cigar_str(ops):
    return "".join(f"{l}{'MIDNS'[o]}" for o, l in ops)


def write_dataset(sam_path, gtf_path, n_genes=120, seed=1, read_len=75, frags_per_gene=(20, 600)):
    rng = np.random.default_rng(seed)
    genes, pos = [], 10_000
    for g in range(n_genes):
        n_exon = int(rng.integers(2, 12))
        n_iso = int(min(12, np.ceil(rng.pareto(1.2) + 1)))
        isoforms = locusgen.make_gene(rng, n_exon, n_iso, exon_len=(90, 400), intron_len=(200, 1500), start=pos)
        isoforms = [ex for ex in isoforms if sum(r - l + 1 for l, r in ex) >= 2 * read_len + 20]
        if not isoforms:
            continue
        genes.append(isoforms)
        pos = max(r for ex in isoforms for _, r in ex) + 5_000
    chrom_len = pos + 10_000
    recs, rid = [], 0
    with open(gtf_path, "w") as gtf:
        for g, isoforms in enumerate(genes):
            expr = rng.pareto(1.1, len(isoforms)) + 0.05
            expr /= expr.sum()
            n_frag = int(rng.integers(*frags_per_gene))
            for t, exons in enumerate(isoforms):
                l, r = exons[0][0], exons[-1][1]
                attr = f'gene_id "G{g}"; transcript_id "G{g}.T{t}";'
                gtf.write(f"chr1\tsynth\ttranscript\t{l}\t{r}\t.\t+\t.\t{attr}\n")
                for (el, er) in exons:
                    gtf.write(f"chr1\tsynth\texon\t{el}\t{er}\t.\t+\t.\t{attr}\n")
            lens = [sum(r - l + 1 for l, r in ex) for ex in isoforms]
            for _ in range(n_frag):
                t = int(rng.choice(len(isoforms), p=expr))
                L = lens[t]
                fl = int(np.clip(rng.normal(250, 30), read_len + 1, L))
                s = int(rng.integers(0, L - fl + 1))
                lb = locusgen._blocks(isoforms[t], s, s + read_len)
                rb = locusgen._blocks(isoforms[t], s + fl - read_len, s + fl)
                if lb[0][0] == rb[0][0]:
                    continue                      # mates starting at the same position are discarded by the reference
                lpos, lops = locusgen._cigar(lb, rng, noise=False)
                rpos, rops = locusgen._cigar(rb, rng, noise=False)
                tlen = rb[-1][1] - lpos + 1
                name = f"r{rid}"
                rid += 1
                for pos_, ops, flag, mpos, tl in ((lpos, lops, 99, rpos, tlen), (rpos, rops, 147, lpos, -tlen)):
                    tags = "NH:i:1" + ("\tXS:A:+" if any(o == 3 for o, _ in ops) else "")
                    recs.append((pos_, f"{name}\t{flag}\tchr1\t{pos_}\t255\t{_cigar_str(ops)}\t=\t{mpos}\t{tl}\t{'A' * read_len}\t{'I' * read_len}\t{tags}\n"))
    recs.sort(key=lambda x: x[0])
    with open(sam_path, "w") as sam:
        sam.write("@HD\tVN:1.0\tSO:coordinate\n")
        sam.write(f"@SQ\tSN:chr1\tLN:{chrom_len}\n")
        for _, line in recs:
            sam.write(line)
    return dict(n_genes=len(genes), n_isoforms=sum(len(g) for g in genes), n_fragments=len(recs) // 2)


def _gene_records(args):
    """fragments of one gene (own seeded stream): sorted (pos, SAM line) records"""
    seed, g, isoforms, n_frag, read_len = args
    rng = np.random.default_rng([seed, g])
    expr = rng.pareto(1.1, len(isoforms)) + 0.05
    expr /= expr.sum()
    lens = [sum(r - l + 1 for l, r in ex) for ex in isoforms]
    which = rng.choice(len(isoforms), size=n_frag, p=expr)
    recs = []
    for i in range(n_frag):
        t = int(which[i])
        L = lens[t]
        fl = int(np.clip(rng.normal(250, 30), read_len + 1, L))
        s = int(rng.integers(0, L - fl + 1))
        lb = locusgen._blocks(isoforms[t], s, s + read_len)
        rb = locusgen._blocks(isoforms[t], s + fl - read_len, s + fl)
        if lb[0][0] == rb[0][0]:
            continue
        lpos, lops = locusgen._cigar(lb, rng, noise=False)
        rpos, rops = locusgen._cigar(rb, rng, noise=False)
        tlen = rb[-1][1] - lpos + 1
        name = f"g{g}r{i}"
        for pos_, ops, flag, mpos, tl in ((lpos, lops, 99, rpos, tlen), (rpos, rops, 147, lpos, -tlen)):
            tags = "NH:i:1" + ("\tXS:A:+" if any(o == 3 for o, _ in ops) else "")
            recs.append((pos_, f"{name}\t{flag}\tchr1\t{pos_}\t255\t{_cigar_str(ops)}\t=\t{mpos}\t{tl}\t{'A' * read_len}\t{'I' * read_len}\t{tags}\n"))
    recs.sort(key=lambda x: x[0])
    return recs


def write_dataset_parallel(sam_path, gtf_path, n_genes=3000, seed=1, read_len=75, frags_per_gene=(20, 700), workers=None):
    """Same kind of data as write_dataset, sized for >= 1 M fragments: the genes are laid out by one seeded stream, the
    fragments of gene g come from their own stream (seed, g) in a process pool. Deterministic for a given seed."""
    import concurrent.futures
    import os
    rng = np.random.default_rng(seed)
    genes, pos = [], 10_000
    for g in range(n_genes):
        n_exon = int(rng.integers(2, 12))
        n_iso = int(min(12, np.ceil(rng.pareto(1.2) + 1)))
        isoforms = locusgen.make_gene(rng, n_exon, n_iso, exon_len=(90, 400), intron_len=(200, 1500), start=pos)
        isoforms = [ex for ex in isoforms if sum(r - l + 1 for l, r in ex) >= 2 * read_len + 20]
        if not isoforms:
            continue
        genes.append(isoforms)
        pos = max(r for ex in isoforms for _, r in ex) + 5_000
    chrom_len = pos + 10_000
    n_frags = rng.integers(frags_per_gene[0], frags_per_gene[1], len(genes))
    with open(gtf_path, "w") as gtf:
        for g, isoforms in enumerate(genes):
            for t, exons in enumerate(isoforms):
                attr = f'gene_id "G{g}"; transcript_id "G{g}.T{t}";'
                gtf.write(f"chr1\tsynth\ttranscript\t{exons[0][0]}\t{exons[-1][1]}\t.\t+\t.\t{attr}\n")
                for (el, er) in exons:
                    gtf.write(f"chr1\tsynth\texon\t{el}\t{er}\t.\t+\t.\t{attr}\n")
    jobs = [(seed, g, isoforms, int(n_frags[g]), read_len) for g, isoforms in enumerate(genes)]
    n_rec = 0
    with open(sam_path, "w") as sam, concurrent.futures.ProcessPoolExecutor(workers or os.cpu_count() or 1) as pool:
        sam.write("@HD\tVN:1.0\tSO:coordinate\n")
        sam.write(f"@SQ\tSN:chr1\tLN:{chrom_len}\n")
        for recs in pool.map(_gene_records, jobs, chunksize=16):      # genes do not overlap: concatenation is coordinate-sorted
            n_rec += len(recs)
            sam.writelines(line for _, line in recs)
    return dict(n_genes=len(genes), n_isoforms=sum(len(g) for g in genes), n_fragments=n_rec // 2)
