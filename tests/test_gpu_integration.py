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


def parse_gtf(path):
    out = {}
    for line in open(path):
        f = line.rstrip("\n").split("\t")
        if len(f) < 9 or f[2] != "transcript":
            continue
        attrs = dict(re.findall(r"(\w+) \"([^\"]*)\"", f[8]))
        out[attrs["transcript_id"]] = (f[3], f[4], f[6], attrs)
    return out


def run(binary, bam, gtf, out, log, threads=1, ctx=None):
    cmd = [binary, bam, "-g", gtf, "-r", "-o", out, "-T", log, "-p", str(threads)] + (["-f", ctx] if ctx else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)


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
