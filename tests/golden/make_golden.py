"""Generate tests/golden/em_golden.npz from the UNMODIFIED reference (oracle/_ref/libsbref.so).

Run in the build container (needs /root/reference to build the reference seam library):
    make -C oracle ref && python tests/golden/make_golden.py
The fixture holds a flat batch of loci (strawberry_b200.synth layout) plus, per locus, the theta the
reference's EmSolver::init/run returned (theta_ref), its return-code bits (rc_ref: bit0 init, bit1 run)
and the iteration count / status of our restatement on the same input (the reference does not expose
its iteration count).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from strawberry_b200 import synth  # noqa: E402


def locus(T, rows, counts, iso_len=None):
    """rows: list of {col: alpha} dicts."""
    rp, col, al = [0], [], []
    for r in rows:
        for c in sorted(r):
            col.append(c), al.append(r[c])
        rp.append(len(col))
    return dict(loc_row_off=np.array([0, len(rows)], np.int64), loc_iso_off=np.array([0, T], np.int64),
                row_ptr=np.array(rp, np.int64), col=np.array(col, np.int32), alpha=np.array(al, np.float64),
                count=np.array(counts, np.int32),
                iso_len=np.array(iso_len if iso_len is not None else [1000 + 100 * j for j in range(T)], np.int32),
                total_mapped_reads=int(sum(counts)))


def edge_cases():
    out = []
    out.append(locus(1, [{0: 0.01}, {0: 0.002}], [10, 5]))                       # T = 1
    out.append(locus(1, [{0: 0.01}, {0: 5e-6}], [10, 7]))                        # T = 1, filtered row carries counts
    out.append(locus(2, [{0: 5e-6}, {1: 1e-5}], [3, 4]))                         # every row filtered -> NO_ROWS
    out.append(locus(3, [], []))                                                 # no rows at all -> NO_ROWS
    out.append(locus(2, [{0: 0.1}, {1: 0.2}], [0, 0]))                           # total = 0 -> zero denominator at it 0
    out.append(locus(2, [{0: 0.1, 1: 0.1}, {1: 0.2}, {0: 0.3}], [5, 0, 7]))      # zero-count row
    out.append(locus(3, [{0: 0.1}, {0: 0.05, 1: 0.02}, {1: 0.03}], [4, 0, 0]))   # theta_1 -> 0 then d == 0 later
    out.append(locus(3, [{0: 0.1, 1: 0.0}, {0: 0.05, 1: 0.02}], [4, 9]))         # explicit zero, empty column 2
    out.append(locus(2, [{0: 1e-5, 1: 2e-5}, {0: 0.3}], [6, 2]))                 # entry exactly at the threshold
    # slow convergence: two almost collinear isoforms and a large count -> iteration cap
    rows = [{0: 0.010, 1: 0.0100001}, {0: 0.020, 1: 0.0200003}, {0: 0.015, 1: 0.0149998}]
    out.append(locus(2, rows, [40000000, 30000000, 50000000]))
    rows = [{0: 0.010, 1: 0.0101, 2: 0.0099}, {0: 0.020, 1: 0.0203, 2: 0.02}, {1: 0.015, 2: 0.0149}, {0: 0.01, 2: 0.0101}]
    out.append(locus(3, rows, [4000000, 3000000, 5000000, 100000]))
    return out


def main():
    assert oracle.have_ref(), "build oracle/_ref/libsbref.so first (make -C oracle ref)"
    parts = edge_cases()
    parts.append(synth.human_shaped(n_loci=120, total_fragments=400_000, seed=101, max_rows=600))
    parts.append(synth.human_shaped(n_loci=4, total_fragments=2_000_000, seed=102, max_iso=60, max_rows=1500))
    parts.append(synth.collapsed_giant(n_loci=1, rows=1500, n_iso=120, density=0.05, seed=103))
    b = synth.concat(parts)
    L = len(b["loc_row_off"]) - 1
    theta_ref = np.zeros(int(b["loc_iso_off"][-1]))
    rc_ref = np.zeros(L, np.int32)
    iters = np.zeros(L, np.int32)
    status = np.zeros(L, np.int32)
    worst = 0.0
    for l in range(L):
        T, rp, col, al, cnt, il = synth.locus_slice(b, l)
        t0 = int(b["loc_iso_off"][l])
        rc, th = oracle.ref_em(cnt, synth.densify(T, rp, col, al).reshape(len(cnt), T))
        st, th_o, it = oracle.em_csr(T, rp, col, al, cnt)
        theta_ref[t0:t0 + T], rc_ref[l], iters[l], status[l] = th, rc, it, st
        assert rc == {0: 3, 1: 3, 2: 1, 3: 0}[st], (l, rc, st)
        worst = max(worst, float(np.max(np.abs(th_o - th) / np.maximum(np.abs(th), 1e-300))))
    print(f"{L} loci, statuses {np.bincount(status, minlength=4)}, max iters {iters.max()}, "
          f"restatement vs reference max rel err {worst:.3e}")
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "em_golden.npz")
    np.savez_compressed(out, theta_ref=theta_ref, rc_ref=rc_ref, iters=iters, status=status,
                        **{k: b[k] for k in ("loc_row_off", "loc_iso_off", "row_ptr", "col", "alpha", "count", "iso_len")},
                        total_mapped_reads=np.int64(b["total_mapped_reads"]))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
