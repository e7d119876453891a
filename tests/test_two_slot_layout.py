"""CPU: properties of the bank-aligned two-slot layout of the giant-locus kernel (strawberry_b200/csrc/sbq_grid_dual.cuh),
restated in numpy: the slot maps are bijections onto disjoint ranges, the second slot of a column lies in another 8-byte
bank, and the two-choice allocation of dual_prepare_kernel brings the fullest of the 16 banks of a BASELINE configs[3] row
(k ~ 1 + Poisson(47) of T ~ U{500..800} isoforms) from ~7 entries down to ~4 - the step count the kernel's register path
(at most 6 steps) is sized for. The CUDA implementation itself is checked against the oracle in tests/test_gpu_em.py."""
import numpy as np


def tp(T):
    return (T + 15) & ~15


def slot_b(j, Tp):
    return Tp + (j & ~15) + ((j + (j >> 4)) & 15)


def two_choice_loads(cols):
    """g6 prepare, phase B: greedy least-loaded bank, then two improvement sweeps. Returns the 16 bank loads."""
    ba, bb = cols & 15, (cols + (cols >> 4)) & 15
    load = np.zeros(16, int)
    pick = np.zeros(len(cols), bool)
    for e in range(len(cols)):
        pick[e] = load[bb[e]] < load[ba[e]]
        load[bb[e] if pick[e] else ba[e]] += 1
    for _ in range(2):
        for e in range(len(cols)):
            cur, alt = (bb[e], ba[e]) if pick[e] else (ba[e], bb[e])
            if load[cur] > load[alt] + 1:
                load[cur] -= 1
                load[alt] += 1
                pick[e] = not pick[e]
    return load


def test_slot_maps_are_disjoint_bijections():
    for T in (1, 15, 16, 17, 90, 718, 783, 800, 4048):
        Tp = tp(T)
        j = np.arange(T)
        b = slot_b(j, Tp)
        assert len(np.unique(b)) == T and b.min() >= Tp and b.max() < 2 * Tp       # slot A = j < Tp <= slot B < 2 Tp
        assert (2 * Tp + 64) * 8 <= 65535                                          # slot * 8 fits the u16 stream
        other_bank = (b & 15) != (j & 15)
        assert np.array_equal(other_bank, ((j >> 4) & 15) != 0)                    # same bank only in every 16th 16-block


def test_two_choice_allocation_flattens_the_banks():
    rng = np.random.default_rng(7)
    one, two = [], []
    for _ in range(1500):
        T = int(rng.integers(500, 801))
        k = min(T, 1 + int(rng.poisson(47)))
        m = np.arange(k)
        lo, hi = (m * T) // k, ((m + 1) * T) // k                                  # synth.giant: one column per stratum
        cols = lo + np.floor(rng.random(k) * (hi - lo)).astype(int)
        one.append(np.bincount(cols & 15, minlength=16).max())
        load = two_choice_loads(cols)
        assert load.sum() == k
        two.append(load.max())
    one, two = np.array(one), np.array(two)
    assert one.mean() > 6.0                  # a single slot per column: the fullest bank holds ~7 of ~48 entries
    assert two.mean() < 4.1                  # two slots: ~3.85
    assert (two > 6).mean() < 0.002          # rows the register path (6 steps) cannot take are rare (they are walked instead)
