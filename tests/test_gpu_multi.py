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
