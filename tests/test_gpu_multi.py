"""GPU (needs >= 2 devices, skipped otherwise): loci partitioned over 2 ranks with NCCL give the same answer as
one GPU - bitwise for theta/FPKM/frac (a locus' plan depends only on the locus), TPM up to the order of one sum."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from strawberry_b200 import partition, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    b = synth.human_shaped(n_loci=3000, total_fragments=1_500_000, seed=77)
    full = partition.quantify_distributed(b, b["total_mapped_reads"], partition.GpuLocal(device=rank))
    if rank == 0:
        np.savez(out, **full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus_match_one(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from strawberry_b200 import api, synth
    out = str(tmp_path / "multi.npz")
    mp.spawn(_worker, args=(2, 29600 + os.getpid() % 2000, out), nprocs=2, join=True)
    got = np.load(out)
    b = synth.human_shaped(n_loci=3000, total_fragments=1_500_000, seed=77)
    q = api.Quantifier(device=0)
    q.submit_flat(b)
    q.run(b["total_mapped_reads"])
    one = q.results()
    for k in ("theta", "fpkm", "frac", "keep", "iters", "status"):
        assert np.array_equal(got[k], one[k], equal_nan=True), k
    assert np.allclose(got["tpm"], one["tpm"], rtol=1e-12, equal_nan=True)


def _multi_vs_single(n_gpus, batch, **cfg):
    from strawberry_b200 import api
    one = api.Quantifier(device=0, **cfg)
    one.submit_flat(batch)
    one.run(batch["total_mapped_reads"])
    ref = one.results()
    one.close()
    q = api.Quantifier(device=0, n_gpus=n_gpus, **cfg)
    q.submit_flat(batch)
    q.run(batch["total_mapped_reads"])
    got, st, dev = q.results(), q.stats(), q.locus_devices()
    q.close()
    return ref, got, st, dev


@pytest.mark.parametrize("n_gpus", [2, 4, 8])
def test_c_abi_multi_gpu_context_matches_one_gpu(n_gpus):
    """sbq_config.n_gpus = N inside ONE process (no torch.distributed): LPT partition in sbq_upload, one ncclAllReduce of the
    FPKM sums, results back in submit order. theta / FPKM / frac / keep / iters / status bitwise equal to one GPU, TPM up to
    the order of one sum; every device gets loci."""
    import torch
    if torch.cuda.device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    from strawberry_b200 import synth
    b = synth.concat([synth.human_shaped(n_loci=3000, total_fragments=1_500_000, seed=77), synth.giant(n_loci=2, rows_per_locus=20_000, seed=4)])
    ref, got, st, dev = _multi_vs_single(n_gpus, b, min_iso_frac=0.01)
    for k in ("theta", "fpkm", "frac", "keep", "iters", "status"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    assert np.allclose(got["tpm"], ref["tpm"], rtol=1e-12, equal_nan=True)
    assert sorted(set(dev.tolist())) == list(range(n_gpus))
    assert st["n_loci"] == 3002 and st["nnz"] == int(b["row_ptr"][-1])


def test_c_abi_multi_gpu_fewer_loci_than_devices():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from strawberry_b200 import synth
    b = synth.human_shaped(n_loci=1, total_fragments=5_000, seed=3)
    ref, got, st, dev = _multi_vs_single(2, b)
    for k in ("theta", "fpkm", "frac", "tpm", "keep", "iters", "status"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k


def test_c_abi_n_gpus_beyond_the_box_is_refused():
    import torch
    from strawberry_b200 import api
    with pytest.raises(api.SbqError) as e:
        api.Quantifier(device=0, n_gpus=torch.cuda.device_count() + 1)
    assert e.value.code == api.SBQ_ERR_NO_DEVICE


def test_integrated_binary_on_n_gpus_gives_the_same_gtf(tmp_path):
    """The batched drop-in binary with SBQ_N_GPUS = all devices of the box (loci LPT-partitioned inside libsbq, one
    ncclAllReduce for the TPM denominator) writes the same GTF records as on one GPU."""
    import subprocess
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import samgen
    refdir = os.path.join(ROOT, "oracle", "_ref")
    binary, samtools = os.path.join(refdir, "strawberry_sbq_batched"), os.path.join(refdir, "samtools_ref")
    if not (os.path.exists(binary) and os.path.exists(samtools)):
        pytest.skip("oracle/_ref binaries not built (make -C integration)")
    sam, gtf, bam = (str(tmp_path / x) for x in ("s.sam", "s.gtf", "s.bam"))
    samgen.write_dataset(sam, gtf, n_genes=200, seed=12)
    with open(bam, "wb") as fh:
        subprocess.run([samtools, "view", "-bS", sam], check=True, stdout=fh, stderr=subprocess.DEVNULL)
    outs = []
    for tag, env in (("one", {}), ("many", {"SBQ_N_GPUS": str(n)})):
        out = str(tmp_path / f"{tag}.gtf")
        subprocess.run([binary, bam, "-g", gtf, "-r", "-o", out, "-T", str(tmp_path / f"{tag}.log"), "-p", "2"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600, env=dict(os.environ, **env))
        outs.append(sorted(l for l in open(out, "rb").read().split(b"\n") if l and not l.startswith(b"#")))
    assert len(outs[0]) > 500 and outs[0] == outs[1]


def test_raw_loci_on_two_gpus_match_one():
    """Device class assignment (sbq_submit_raw) in a 2-GPU context: every raw locus is dealt to a device at submit time, the class
    tables are built on both devices concurrently, results come back in submit order - bitwise equal to the single-device
    context (theta / FPKM / frac / keep / iters / status; TPM up to the order of one sum)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import locusgen
    from strawberry_b200 import api, builder
    loci = []
    for seed in range(40, 64):
        isoforms, hits, _ = locusgen.random_locus(seed)
        loci.append(([locusgen.transcript_features(ex) for ex in isoforms], [(m, builder.pair_features(l, r)) for m, l, r in hits]))
    out = []
    for n in (1, 2):
        q = api.Quantifier(device=0, n_gpus=n, min_iso_frac=0.01)
        q.set_insert_model(builder.Model.normal(200.0, 40.0), 50)
        for tfe, hl in loci:
            builder.submit_raw(q, tfe, hl, read_len=50)
        q.run(250_000)
        out.append((q.results(), q.locus_devices(), q.stats()))
        q.close()
    (ref, _, st1), (got, dev, st2) = out
    for k in ("theta", "fpkm", "frac", "keep", "iters", "status"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    assert np.allclose(got["tpm"], ref["tpm"], rtol=1e-12, equal_nan=True)
    assert sorted(set(dev.tolist())) == [0, 1]
    assert st2["n_loci"] == st1["n_loci"] == len(loci) and st2["nnz"] == st1["nnz"] and st2["n_row"] == st1["n_row"]
