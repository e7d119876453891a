"""Wall time and quantification-only time of the real program on a >= 1 M-fragment synthetic BAM: the unmodified reference
(-p 1 / -p nproc), the per-locus drop-in, and the batched drop-in (-p 1 / -p nproc, 1 / N GPUs). Run on the GPU box:

    python tools/integration_bench.py [n_genes] > gpurun_out/integration_bench.md
"""
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import samgen  # noqa: E402

R = os.path.join(ROOT, "oracle", "_ref")
n_genes = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
tmp = tempfile.mkdtemp(prefix="sbq_integ_")
sam, gtf, bam = (os.path.join(tmp, n) for n in ("s.sam", "s.gtf", "s.bam"))
info = samgen.write_dataset_parallel(sam, gtf, n_genes=n_genes, seed=5)
subprocess.run([os.path.join(R, "samtools_ref"), "view", "-bS", sam], check=True, stdout=open(bam, "wb"), stderr=subprocess.DEVNULL)
os.remove(sam)
nproc = os.cpu_count() or 4
try:
    import torch
    n_gpu = torch.cuda.device_count()
except Exception:
    n_gpu = 1


def body(path):
    return sorted(l for l in open(path, "rb").read().split(b"\n") if l and not l.startswith(b"#"))


def go(tag, binary, threads, **env):
    out, log = os.path.join(tmp, tag + ".gtf"), os.path.join(tmp, tag + ".log")
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(R, binary), bam, "-g", gtf, "-r", "-o", out, "-T", log, "-p", str(threads)], stdout=subprocess.DEVNULL,
                       stderr=subprocess.PIPE, text=True, env=dict(os.environ, SBQ_TIMING="1", **env))
    wall = time.perf_counter() - t0
    tm = " ".join(l for l in r.stderr.splitlines() if l.startswith("SBQ_TIMING"))
    num = lambda k: float(re.search(k + r" ([0-9.]+)", tm).group(1)) if re.search(k + r" ([0-9.]+)", tm) else None
    return dict(tag=tag, rc=r.returncode, wall=wall, gtf=body(out) if r.returncode == 0 else None, table=num("class_table_ms"), est=num("estimate_abundances_ms"),
                stage=num(r" stage_ms"), run=num("run_ms"), solve=num("solve"), upload=num("upload"), finish=num("finish_ms"), walk=num(r"walk\+stage_ms"))


runs = [go("reference -p 1", "strawberry_ref_timed", 1), go(f"reference -p {nproc}", "strawberry_ref_timed", nproc),
        go("per-locus drop-in -p 1", "strawberry_sbq", 1), go("batched drop-in -p 1", "strawberry_sbq_batched", 1),
        go(f"batched drop-in -p {nproc}", "strawberry_sbq_batched", nproc),
        go("batched drop-in -p 1, host class assignment", "strawberry_sbq_batched", 1, SBQ_HOST_CLASSES="1"),
        go("batched drop-in -p 1, two passes over the BAM", "strawberry_sbq_batched", 1, SBQ_SINGLE_PASS="0"),
        go("batched drop-in -p 1, two passes, host weights", "strawberry_sbq_batched", 1, SBQ_SINGLE_PASS="0", SBQ_HOST_WEIGHTS="1")]
if n_gpu > 1:
    runs.append(go(f"batched drop-in -p {nproc}, {n_gpu} GPUs", "strawberry_sbq_batched", nproc, SBQ_N_GPUS=str(n_gpu)))
ref = runs[0]["gtf"]
print(f"Dataset: {info['n_fragments']} paired fragments, {info['n_genes']} genes, {info['n_isoforms']} isoforms (tests/samgen.py write_dataset_parallel, seed 5); "
      f"host: {nproc} cores; `-g s.gtf -r`. Quantification-only = time inside the quantification entry points, summed over loci "
      f"(reference: LocusContext::assign_exon_bin + set_theory_bin_weight + estimate_abundances; drop-in: class-table build + staging + sbq_run + per-locus tail).\n")
print("| run | wall s | quantification-only s | of which class table / weights | EM + epilogue | sorted GTF == reference |")
print("|---|---|---|---|---|---|")
for r in runs:
    if r["rc"] != 0:
        print(f"| {r['tag']} | failed rc={r['rc']} | | | | |")
        continue
    if r["est"] is not None:
        q, tb, em = (r["table"] + r["est"]) / 1e3, r["table"] / 1e3, r["est"] / 1e3
    else:
        q = ((r["table"] or 0) + (r["stage"] or 0) + (r["run"] or 0) + (r["finish"] or 0)) / 1e3
        tb, em = (r["table"] or 0) / 1e3, ((r["run"] or 0) if r["run"] else (r["stage"] or 0)) / 1e3
    print(f"| {r['tag']} | {r['wall']:.2f} | {q:.3f} | {tb:.3f} | {em:.3f} | {'yes' if r['gtf'] == ref else 'NO'} |")
