"""Multi-GPU host logic: loci are independent, so they are partitioned across ranks (one process per GPU).

The only exchange step of the path is the TPM denominator, the sum of FPKM over all surviving isoforms of all
loci (reference src/alignments.cpp:1821-1829): one scalar all-reduce. Everything here is host-side plumbing;
the per-rank solve is the CUDA engine (`GpuLocal`), there is no CPU path in the product.
"""
import numpy as np

FLAT = (("col", np.int32), ("alpha", np.float64))


def locus_cost(batch):
    """Planning cost per locus: non-zeros (iteration counts are unknown before the solve)."""
    rp, lro = np.asarray(batch["row_ptr"]), np.asarray(batch["loc_row_off"])
    return (rp[lro[1:]] - rp[lro[:-1]]).astype(np.int64) + np.diff(lro) + np.diff(np.asarray(batch["loc_iso_off"]))


def lpt_partition(cost, n_parts):
    """Greedy longest-processing-time partition, deterministic on every rank: loci by descending cost (ties by
    index) onto the currently lightest part (ties by part index). Returns one ascending index array per part."""
    cost = np.asarray(cost, dtype=np.int64)
    order = np.lexsort((np.arange(len(cost)), -cost))
    load = np.zeros(n_parts, np.int64)
    owner = np.empty(len(cost), np.int32)
    for l in order:
        p = int(np.argmin(load))
        owner[l] = p
        load[p] += cost[l]
    return [np.nonzero(owner == p)[0] for p in range(n_parts)]


def take(batch, idx):
    """Sub-batch with the loci `idx` (ascending), in the flat layout of include/sbq.h."""
    idx = np.asarray(idx, dtype=np.int64)
    lro, lio, rp = (np.asarray(batch[k]) for k in ("loc_row_off", "loc_iso_off", "row_ptr"))
    R, T = lro[idx + 1] - lro[idx], lio[idx + 1] - lio[idx]
    new_lro = np.concatenate([[0], np.cumsum(R)]).astype(np.int64)
    new_lio = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
    rows = np.concatenate([np.arange(lro[l], lro[l + 1]) for l in idx]) if len(idx) else np.zeros(0, np.int64)
    isos = np.concatenate([np.arange(lio[l], lio[l + 1]) for l in idx]) if len(idx) else np.zeros(0, np.int64)
    nnz_row = rp[rows + 1] - rp[rows] if len(rows) else np.zeros(0, np.int64)
    new_rp = np.concatenate([[0], np.cumsum(nnz_row)]).astype(np.int64)
    if len(rows):
        ent = np.repeat(rp[rows] - new_rp[:-1], nnz_row) + np.arange(new_rp[-1])
    else:
        ent = np.zeros(0, np.int64)
    out = dict(loc_row_off=new_lro, loc_iso_off=new_lio, row_ptr=new_rp,
               col=np.asarray(batch["col"])[ent], alpha=np.asarray(batch["alpha"])[ent],
               count=np.asarray(batch["count"])[rows], iso_len=np.asarray(batch["iso_len"])[isos],
               total_mapped_reads=batch["total_mapped_reads"])
    return out, isos


class GpuLocal:
    """Per-rank solver on the CUDA engine."""

    def __init__(self, device=None, **config):
        from . import api
        self.q = api.Quantifier(device=-1 if device is None else device, **config)

    def solve(self, batch, total_mapped_reads):
        self.q.clear()
        self.q.submit_flat(batch)
        self.q.upload()
        self.q.solve(total_mapped_reads)
        return self.q.fpkm_sum()

    def finalize(self, global_fpkm_sum):
        self.q.finalize_tpm(global_fpkm_sum)
        self.q.download()
        return self.q.results()


def quantify_distributed(batch, total_mapped_reads, local, group=None, gather=True):
    """Partition `batch` over the ranks of `group` (torch.distributed), solve the local share with `local`
    (GpuLocal in production), all-reduce the TPM denominator and, if `gather`, reassemble the full result on
    every rank in the original locus / isoform order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    parts = lpt_partition(locus_cost(batch), world)
    sub, isos = take(batch, parts[rank])
    n_local = len(parts[rank])
    local_sum = local.solve(sub, total_mapped_reads) if n_local else 0.0
    t = torch.tensor([local_sum], dtype=torch.float64)
    if world > 1:
        backend = dist.get_backend(group)
        if backend == "nccl":
            t = t.cuda()
        dist.all_reduce(t, group=group)          # the path's only collective
    global_sum = float(t.item())
    res = local.finalize(global_sum) if n_local else dict(theta=np.zeros(0), fpkm=np.zeros(0), frac=np.zeros(0), tpm=np.zeros(0),
                                                          keep=np.zeros(0, np.int32), iters=np.zeros(0, np.int32), status=np.zeros(0, np.int32))
    res = dict(res, fpkm_sum=global_sum, loci=parts[rank], isoforms=isos)
    if not gather or world == 1:
        if world == 1:
            return res
        return res
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: res[k] for k in ("theta", "fpkm", "frac", "tpm", "keep", "iters", "status", "loci", "isoforms")}, group=group)
    n_iso, n_loc = int(np.asarray(batch["loc_iso_off"])[-1]), len(batch["loc_row_off"]) - 1
    full = dict(theta=np.zeros(n_iso), fpkm=np.zeros(n_iso), frac=np.zeros(n_iso), tpm=np.zeros(n_iso), keep=np.zeros(n_iso, np.int32),
                iters=np.zeros(n_loc, np.int32), status=np.zeros(n_loc, np.int32), fpkm_sum=global_sum)
    for g in gathered:
        for k in ("theta", "fpkm", "frac", "tpm", "keep"):
            full[k][g["isoforms"]] = g[k]
        for k in ("iters", "status"):
            full[k][g["loci"]] = g[k]
    return full
