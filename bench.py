#!/usr/bin/env python
"""bench.py - the reference's headline metric on B200: fragments*EM-iters/sec (and wall time to
converge) of the per-locus Latent-Class-Model EM on a synthetic 10M-fragment, ~20k-locus human-shaped
batch (BASELINE.json configs[1]; generator strawberry_b200.synth.human_shaped, seed 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (EM to convergence + FPKM/frac/filter epilogue + TPM denominator) over
the rank's batch. `value` is timed with the batch resident in HBM (CUDA events inside libsbq on the
stream the kernels run on); `e2e` is the same work through the public host-buffer call sbq_run from
page-locked host arrays, H2D and D2H inside the timed region. N > 1: weak scaling - every rank owns a
full copy of the batch (seed 2); loci are independent so there is no data-path collective, only the scalar
TPM-denominator all-reduce (NCCL) per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fragments*EM-iters/sec"
UNIT = "fragment-iters/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_workload(batch):
    """The whole workload as the CPU sample (the dense reference finishes it in seconds)."""
    import oracle
    ora = oracle.quantify_batch(batch, batch["total_mapped_reads"], n_threads=os.cpu_count() or 1)
    frag_iters = int((np.add.reduceat(batch["count"].astype(np.int64), batch["loc_row_off"][:-1]) * ora["iters"]).sum())
    return oracle, frag_iters


def cpu_baseline_sample(batch):
    """CPU baseline, one thread, on the full 20k-locus workload: the reference's own EmSolver
    (oracle/_ref/libsbref.so, kind "reference") when it was built, else our C port (kind "port")."""
    oracle, frag_iters = cpu_workload(batch)
    if oracle.have_ref():
        secs, kind = oracle.ref_em_batch(batch, n_threads=1)["seconds"], "reference"
    else:
        secs, kind = oracle.quantify_batch(batch, batch["total_mapped_reads"], n_threads=1)["seconds"], "port"
    return dict(value=frag_iters / secs, unit=UNIT, cores=1, kind=kind,
                sample=f"the full workload, one pass: 20000 loci, {frag_iters} fragment-iters, {secs:.2f} s on one core "
                       f"({'reference EmSolver::init/run, dense Eigen' if kind == 'reference' else 'C port of the reference EM'})")


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores (all of them)."""
    if rank != 0:
        return
    from strawberry_b200 import synth
    batch = synth.human_shaped(seed=2)
    oracle, frag_iters = cpu_workload(batch)
    cores = os.cpu_count() or 1
    use_ref = oracle.have_ref()
    kind = "reference" if use_ref else "port"

    def step():
        if use_ref:
            return oracle.ref_em_batch(batch, n_threads=cores)["seconds"]
        return oracle.quantify_batch(batch, batch["total_mapped_reads"], n_threads=cores)["seconds"]

    for _ in range(args.warmup):
        step()
    secs = [step() for _ in range(args.steps)]
    ms = 1e3 * float(np.mean(secs))
    value = frag_iters / (ms / 1e3)
    sample = (f"the full workload per step: 20000 loci, {frag_iters} fragment-iters; "
              f"{'reference EmSolver::init/run (dense Eigen, src/estimate.cpp:366-488), one std::thread per core over loci' if use_ref else 'C port of the reference EM, pthreads'}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: synthetic 10M-fragment paired-end human-shaped loci (20000 loci), quantification only",
                       "generator": synth.GENERATOR_VERSION, "seed": 2, "l2": "n/a (CPU)"},
            "wall_ms_to_converge": ms,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry exactly one JSON line. Libraries write there behind Python's back (NCCL prints its version banner to
    fd 1 when NCCL_DEBUG is set), so fd 1 is pointed at stderr for the run and the line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-giant", action="store_true", help="skip the giant-locus roofline leg")
    ap.add_argument("--giant-rows", type=int, default=1_000_000)
    ap.add_argument("--giant-loci", type=int, default=2)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from strawberry_b200 import api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- workload: BASELINE configs[1], one full batch per rank (weak scaling)
    batch = synth.human_shaped(seed=2)   # the same 10M-fragment batch on every rank: identical work per GPU
    total_reads = batch["total_mapped_reads"]
    pinned = api.pinned_batch(batch)
    q = api.Quantifier(device=local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    gsum = torch.zeros(1, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tpm_exchange():
        """the path's only exchange step: sum of FPKM over all ranks (src/alignments.cpp:1821-1824)"""
        if world > 1:
            q.fpkm_sum_to_device(gsum.data_ptr())
            dist.all_reduce(gsum)
            q.finalize_tpm(float(gsum.item()))
        else:
            q.finalize_tpm(q.fpkm_sum())

    def resident_step():
        flush.fill_(1)               # L2 flush between timed iterations (not timed: events live inside sbq_solve)
        torch.cuda.synchronize()
        q.solve(total_reads)
        tpm_exchange()
        return q.stats()["solve_ms"]

    def e2e_step():
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        q.clear()
        q.submit_flat(pinned)        # page-locked arrays are used in place
        q.upload()
        q.solve(total_reads)
        tpm_exchange()
        q.download()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    # ---- device-resident leg ("value")
    q.submit_flat(pinned)
    q.upload()
    for _ in range(args.warmup):
        resident_step()
    barrier()
    clk = ClockSampler(local_rank)
    clk.__enter__()                      # sampled over both timed legs (resident + end-to-end)
    t_wall0 = time.perf_counter()
    step_ms = [resident_step() for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    q.download()
    st = q.stats()
    launches = q.launch_stats()
    res = q.results()
    frag_iters = st["frag_iters"]
    ms_local = float(np.mean(step_ms))

    # ---- end-to-end leg (host buffers through the public call)
    for _ in range(2):
        e2e_step()
    barrier()
    e2e_s = [e2e_step() for _ in range(args.steps)]
    barrier()
    clk.__exit__(None, None, None)
    st_e2e = q.stats()
    e2e_local = float(np.mean(e2e_s))

    # ---- max over ranks, whole-job aggregate
    agg = torch.tensor([ms_local, e2e_local], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(frag_iters)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_per_step, e2e_per_step = float(agg[0]), float(agg[1])
    total_frag_iters = float(tot[0])

    # ---- roofline. The EM kernels of a step (one warp-tier launch + one launch per cluster size) run concurrently on
    # separate streams, so the phase is timed as a whole with CUDA events on the context's main stream (em_ms) and
    # every launch also carries its own event pair on its own stream (launches[]). `kernel` names the launch with
    # the most algorithmic bytes.
    peak, peak_src = load_peaks()
    dom = max(launches, key=lambda r: r["alg_bytes"])
    achieved = st["alg_bytes"] / (st["em_ms"] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": f"{dom['kernel']} (cluster_size={dom['cluster_size']}) + {len(launches) - 1} concurrent EM launches",
                "kernel_ms": st["em_ms"], "alg_bytes_per_launch": st["alg_bytes"], "peak_source": peak_src,
                "note": "the 58 MB CSR of the batch is read from HBM once and then lives in shared memory (cluster tier) / L1 (warp tier) for up "
                        "to 1000 sequential EM iterations, so this step is bound by per-iteration latency of its longest loci, not by HBM; "
                        "roofline_giant is the HBM-bound kernel of this path",
                "launches": [{k: r[k] for k in ("kernel", "cluster_size", "n_loci", "nnz", "ms", "alg_bytes", "max_iters")} for r in launches]}

    line = {"metric": METRIC, "value": total_frag_iters / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: synthetic 10M-fragment paired-end human-shaped loci (20000 loci), quantification only; one such batch per GPU",
                       "generator": synth.GENERATOR_VERSION, "seed": 2, "loci_per_gpu": st["n_loci"], "rows_per_gpu": st["n_row"],
                       "isoforms_per_gpu": st["n_iso"], "nnz_per_gpu": st["nnz"], "fragments_per_gpu": int(batch["count"].sum()),
                       "em_iters_total_per_gpu": st["em_iters_total"], "max_iter": 1000, "theta_tol": 1e-2,
                       "tiers": {"warp": st["loci_warp"], "cluster": st["loci_cta"], "grid": st["loci_grid"]},
                       "l2": "flushed between timed iterations (256 MiB fill)", "parallelism": f"loci x{world}"},
            "wall_ms_to_converge": ms_per_step,
            "e2e": {"value": total_frag_iters / e2e_per_step, "unit": UNIT, "ms_per_step": e2e_per_step * 1e3,
                    "h2d_bytes_per_step": st_e2e["h2d_bytes"], "d2h_bytes_per_step": st_e2e["d2h_bytes"],
                    "stages_ms": {"upload": st_e2e["upload_ms"], "solve": st_e2e["solve_ms"], "download": st_e2e["download_ms"]}},
            "gpu_launches": int((st["kernel_launches"]) * args.steps),
            "timed_region_wall_s": t_wall,
            "roofline": roofline,
            "clocks": clk.summary(),
            "statuses": {k: int(v) for k, v in zip(("ok", "iter_cap", "zero_denom", "no_rows"), np.bincount(res["status"], minlength=4))}}

    if rank == 0:
        # ---- giant-locus leg: the HBM-bound multi-CTA kernel (BASELINE configs[3] shape, scaled to fit a short run)
        if not args.no_giant:
            try:
                gb = synth.giant(n_loci=args.giant_loci, rows_per_locus=args.giant_rows, seed=4)
                qg = api.Quantifier(device=local_rank)
                qg.submit_flat(gb)
                qg.upload()
                g_ms, g_bytes = [], 0
                for i in range(1 + 3):
                    flush.fill_(1)
                    torch.cuda.synchronize()
                    qg.solve(gb["total_mapped_reads"])
                    if i:
                        g_ms.append(qg.stats()["grid_em_ms"])
                qg.finalize_tpm(qg.fpkm_sum())
                qg.download()
                gst = qg.stats()
                g_kernel = next((r["kernel"] for r in qg.launch_stats() if r["kernel"].startswith("em_grid")), "em_grid_kernel")
                g_ach = gst["grid_alg_bytes"] / (float(np.mean(g_ms)) * 1e-3) / 1e9
                # DRAM bytes per non-zero and pass measured by ncu --set full on these kernels (profiles/r01_grid_dual_ncu_summary.txt:
                # dram__bytes_read.sum + write = 4.59 GB for 9 passes over 48.0 M non-zeros = 10.6 B; r01_grid_tma_ncu_summary.txt:
                # 10.3 B): u16 columns are streamed, so the kernels move LESS than the 12 B/nnz the roofline counts as algorithmic
                passes = gst["em_iters_total"] + args.giant_loci
                g_traffic = (10.63 if g_kernel == "em_grid_dual_kernel" else 10.34) * (gst["nnz"] / args.giant_loci) * passes
                line["roofline_giant"] = {"bound": "hbm", "achieved": g_ach, "peak": peak, "unit": "GB/s", "frac": g_ach / peak,
                                          "traffic": g_traffic, "traffic_source": "scaled from the ncu --set full capture of the same kernel in profiles/ (bytes per non-zero and pass)",
                                          "kernel": g_kernel, "kernel_ms": float(np.mean(g_ms)),
                                          "alg_bytes_per_launch": gst["grid_alg_bytes"], "peak_source": peak_src,
                                          "workload": f"configs[3] shape: {args.giant_loci} loci x {args.giant_rows} rows, n_i=1, k~1+Poisson(47), T~U{{500..800}} "
                                                      f"({gst['nnz']} nnz, {gst['em_iters_total']} EM iterations in total); CSR {gst['nnz'] * 12 / 1e9:.2f} GB > L2",
                                          "value": gst["frag_iters"] / (float(np.mean(g_ms)) * 1e-3), "unit_value": UNIT}
                qg.close()
            except Exception as e:   # the headline line must still be printed
                line["roofline_giant"] = {"error": repr(e)}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_sample(batch)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
