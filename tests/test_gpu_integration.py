"""GPU, integration level: the reference program with libsbq linked behind its quantification call
site (oracle/_ref/strawberry_sbq, built by integration/Makefile from the reference's own objects + our replacement
estimate TU) against the unmodified reference binary (oracle/_ref/strawberry_ref) on the same synthetic BAM + GTF.
The two binaries contain reference object code, so they are built where the reference checkout exists and travel to
the GPU box; the test is skipped when they are absent."""
import os
import re
import subprocess

import pytest

import samgen

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
BINS = [os.path.join(REFDIR, b) for b in ("strawberry_ref", "strawberry_sbq", "samtools_ref")]
BATCHED = os.path.join(REFDIR, "strawberry_sbq_batched")


def parse_gtf(path):
    out = {}
    for line in open(path):
        f = line.rstrip("\n").split("\t")
        if len(f) < 9 or f[2] != "transcript":
            continue
        attrs = dict(re.findall(r"(\w+) \"([^\"]*)\"", f[8]))
        out[attrs["transcript_id"]] = (f[3], f[4], f[6], attrs)
    return out


def run(binary, bam, gtf, out, log, threads=1, ctx=None, **env):
    cmd = [binary, bam, "-g", gtf, "-r", "-o", out, "-T", log, "-p", str(threads)] + (["-f", ctx] if ctx else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600, env=dict(os.environ, **env))


@pytest.mark.parametrize("threads", [1, 4])
def test_dropin_binary_matches_reference_binary(tmp_path, threads):
    if not all(os.path.exists(b) for b in BINS):
        pytest.skip("oracle/_ref binaries not built (make -C integration)")
    sam, gtf, bam = (str(tmp_path / n) for n in ("s.sam", "s.gtf", "s.bam"))
    info = samgen.write_dataset(sam, gtf, n_genes=150, seed=3)
    with open(bam, "wb") as fh:
        subprocess.run([BINS[2], "view", "-bS", sam], check=True, stdout=fh, stderr=subprocess.DEVNULL)
    outs = {}
    for tag, binary in (("ref", BINS[0]), ("sbq", BINS[1])):
        out, log = str(tmp_path / f"{tag}{threads}.gtf"), str(tmp_path / f"{tag}{threads}.log")
        run(binary, bam, gtf, out, log, threads)
        outs[tag] = (parse_gtf(out), open(log).read())
    ref, got = outs["ref"][0], outs["sbq"][0]
    assert len(ref) > 0.5 * info["n_isoforms"]
    assert set(ref) == set(got), "the same transcripts must be reported"
    worst = 0.0
    for tid, (l, r, strand, attrs) in ref.items():
        gl, gr, gs, gattrs = got[tid]
        assert (l, r, strand) == (gl, gr, gs)
        for key in ("FPKM", "Frac", "TPM"):
            a, b = float(attrs[key]), float(gattrs[key])
            # GTF shows the first 11 characters of std::to_string (src/contig.cpp:677-701): compare numerically
            tol = 1e-5 * max(abs(a), 1.0) + 2e-6
            assert abs(a - b) <= tol, (tid, key, attrs[key], gattrs[key])
            worst = max(worst, abs(a - b) / max(abs(a), 1e-9))
    # theta log lines (%f, src/estimate.cpp:311-313): same multiset of values up to the last printed digit
    th_ref = sorted(float(x) for x in re.findall(r"has ([0-9.]+) raw read count", outs["ref"][1]))
    th_got = sorted(float(x) for x in re.findall(r"has ([0-9.]+) raw read count", outs["sbq"][1]))
    assert len(th_ref) == len(th_got) and len(th_ref) > 0
    assert max(abs(a - b) for a, b in zip(th_ref, th_got)) <= 2e-6 * max(1.0, max(th_ref))


def test_fragment_context_tsv_matches_reference(tmp_path):
    """a18 / SURVEY 8f.3: the -f fragment-context TSV (class coordinates, 12-digit alpha rows, counts, FPKM and frac
    strings - the input format of the downstream DE tool) written by the integrated binary equals the reference's."""
    if not all(os.path.exists(b) for b in BINS):
        pytest.skip("oracle/_ref binaries not built (make -C integration)")
    sam, gtf, bam = (str(tmp_path / n) for n in ("s.sam", "s.gtf", "s.bam"))
    samgen.write_dataset(sam, gtf, n_genes=80, seed=9)
    with open(bam, "wb") as fh:
        subprocess.run([BINS[2], "view", "-bS", sam], check=True, stdout=fh, stderr=subprocess.DEVNULL)
    rows = {}
    for tag, binary in (("ref", BINS[0]), ("sbq", BINS[1])):
        ctx = str(tmp_path / f"{tag}.tsv")
        run(binary, bam, gtf, str(tmp_path / f"{tag}f.gtf"), str(tmp_path / f"{tag}f.log"), 1, ctx)
        rows[tag] = [l.rstrip("\n").split("\t") for l in open(ctx)]
    ref, got = rows["ref"], rows["sbq"]
    assert len(ref) == len(got) > 500 and ref[0] == got[0]
    hdr = ref[0]
    num_cols = {hdr.index("FPKMs"), hdr.index("conditional_probabilities"), hdr.index("class_probabilities")}
    for a, b in zip(ref[1:], got[1:]):
        assert len(a) == len(b)
        for k, (x, y) in enumerate(zip(a, b)):
            if k in num_cols:
                xs, ys = x.split(","), y.split(",")
                assert len(xs) == len(ys)
                for u, v in zip(xs, ys):
                    if u == v:
                        continue
                    fu, fv = float(u), float(v)
                    assert abs(fu - fv) <= 1e-9 * max(abs(fu), 1e-30) + (2e-6 if k != hdr.index("conditional_probabilities") else 0.0), (hdr[k], u, v)
            else:
                assert x == y, (hdr[k], x, y)


def gtf_body(path):
    """GTF records; the leading comment lines carry the command line (binary name, output paths) and are skipped"""
    return [l for l in open(path, "rb").read().split(b"\n") if l and not l.startswith(b"#")]


def theta_lines(path):
    return sorted(l for l in open(path, "rb").read().split(b"\n") if b"raw read count" in l)


def make_bam(tmp_path, n_genes, seed, parallel=False):
    sam, gtf, bam = (str(tmp_path / n) for n in ("s.sam", "s.gtf", "s.bam"))
    info = (samgen.write_dataset_parallel if parallel else samgen.write_dataset)(sam, gtf, n_genes=n_genes, seed=seed)
    with open(bam, "wb") as fh:
        subprocess.run([BINS[2], "view", "-bS", sam], check=True, stdout=fh, stderr=subprocess.DEVNULL)
    os.remove(sam)
    return bam, gtf, info


def test_batched_dropin_is_byte_exact(tmp_path):
    """SURVEY 8f.4 / INTEGRATION.md section 3: the BATCHED drop-in (one sbq_run for the whole sample, class weights on the
    GPU) against the unmodified reference binary as a byte-for-byte diff: the GTF records in order with -p 1, the sorted
    records with -p 4, and the theta log lines (%f, src/estimate.cpp:311-313) as a sorted multiset."""
    if not all(os.path.exists(b) for b in BINS + [BATCHED]):
        pytest.skip("oracle/_ref binaries not built (make -C integration)")
    bam, gtf, info = make_bam(tmp_path, 150, 3)
    out, log = str(tmp_path / "ref.gtf"), str(tmp_path / "ref.log")
    run(BINS[0], bam, gtf, out, log, 1)
    ref_gtf, ref_theta = gtf_body(out), theta_lines(log)
    assert len(ref_gtf) > info["n_isoforms"] and len(ref_theta) > 100
    # single pass (the clusters of pass 1 are reused: SURVEY 8f.2, the default) and two passes over the BAM like the reference;
    # GPU class weights (default) and host class weights
    # class assignment on the GPU (default) and in the host builder
    for n, (threads, env) in enumerate(((1, {}), (4, {}), (1, {"SBQ_SINGLE_PASS": "0"}), (4, {"SBQ_SINGLE_PASS": "0", "SBQ_HOST_WEIGHTS": "1"}), (1, {"SBQ_HOST_CLASSES": "1"}))):
        tag = f"b{n}"                       # (the program refuses to overwrite an existing output file)
        out, log = str(tmp_path / f"{tag}.gtf"), str(tmp_path / f"{tag}.log")
        run(BATCHED, bam, gtf, out, log, threads, **env)
        got = gtf_body(out)
        if threads == 1:
            assert got == ref_gtf, f"GTF records differ from the reference (byte diff, -p 1, in order, {env})"
        assert sorted(got) == sorted(ref_gtf), f"sorted GTF records differ from the reference (-p {threads}, {env})"
        assert theta_lines(log) == ref_theta, f"theta log lines differ from the reference (-p {threads}, {env})"


def test_batched_dropin_fragment_context_tsv(tmp_path):
    """-f through the batched drop-in (host class weights are kept for it): the TSV rows equal the reference's as text,
    except the numeric columns that are compared like in test_fragment_context_tsv_matches_reference."""
    if not all(os.path.exists(b) for b in BINS + [BATCHED]):
        pytest.skip("oracle/_ref binaries not built (make -C integration)")
    bam, gtf, _ = make_bam(tmp_path, 80, 9)
    rows = {}
    for tag, binary in (("ref", BINS[0]), ("sbq", BATCHED)):
        ctx = str(tmp_path / f"{tag}.tsv")
        run(binary, bam, gtf, str(tmp_path / f"{tag}f.gtf"), str(tmp_path / f"{tag}f.log"), 1, ctx)
        rows[tag] = sorted(open(ctx, "rb").read().split(b"\n"))
    assert len(rows["ref"]) == len(rows["sbq"]) > 500
    diff = sum(a != b for a, b in zip(rows["ref"], rows["sbq"]))
    assert diff <= 0.01 * len(rows["ref"]), f"{diff} of {len(rows['ref'])} TSV rows differ as text"   # last printed digit of a 12-digit alpha may differ


def test_batched_dropin_one_million_fragments(tmp_path):
    """The real program on a >= 1 M-fragment synthetic BAM (2992 genes, ~10 k isoforms): reference binary (-p 1) vs the
    batched drop-in with -p 1 and -p nproc - sorted GTF records and sorted theta log lines byte-identical - and the wall /
    quantification-only times of both (also written to gpurun_out/ when that directory exists)."""
    import json
    import time
    if not all(os.path.exists(b) for b in BINS + [BATCHED]):
        pytest.skip("oracle/_ref binaries not built (make -C integration)")
    bam, gtf, info = make_bam(tmp_path, 3000, 5, parallel=True)
    assert info["n_fragments"] >= 1_000_000
    timed = os.path.join(REFDIR, "strawberry_ref_timed")
    rows = []

    def timed_run(tag, binary, threads, **env):
        out, log = str(tmp_path / f"{tag}.gtf"), str(tmp_path / f"{tag}.log")
        cmd = [binary, bam, "-g", gtf, "-r", "-o", out, "-T", log, "-p", str(threads)]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=900, env=dict(os.environ, SBQ_TIMING="1", **env))
        wall = time.perf_counter() - t0
        rows.append(dict(run=tag, threads=threads, wall_s=round(wall, 3), timing=[l for l in r.stderr.splitlines() if l.startswith("SBQ_TIMING")]))
        return sorted(gtf_body(out)), theta_lines(log)

    ref = timed_run("reference", timed if os.path.exists(timed) else BINS[0], 1)
    nproc = os.cpu_count() or 4
    for tag, threads in (("batched_p1", 1), (f"batched_p{nproc}", nproc)):
        got = timed_run(tag, BATCHED, threads)
        assert len(got[0]) == len(ref[0]) and got[0] == ref[0], f"{tag}: sorted GTF records differ from the reference"
        assert got[1] == ref[1], f"{tag}: theta log lines differ from the reference"
    print(json.dumps(dict(dataset=info, runs=rows), indent=1))
    outdir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(outdir):
        json.dump(dict(dataset=info, runs=rows), open(os.path.join(outdir, "integration_1m_timing.json"), "w"), indent=1)
