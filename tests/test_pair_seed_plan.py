"""Host side of the pair-of-pieces seed tables: which sets a library gets (mirge_pair_seed_plan, no
device call), and the pigeonhole argument the kernels rely on, restated with numpy on random placements."""
import ctypes as C
import itertools

import numpy as np
import pytest

from mirge_b200 import _capi
from mirge_b200.engine import LIB_IDS


def plan(lib, total, threshold=16):
    out = (C.c_int * 16)()
    n = _capi.lib.mirge_pair_seed_plan(LIB_IDS[lib] if isinstance(lib, str) else lib, total, threshold, out)
    return [tuple(out[4 * i:4 * i + 4]) for i in range(n)]


def test_plan_for_the_bench_libraries():
    # miRNA library (57 k bases), isomiR round -v 2 -> four pieces: 3-base pieces for 12..15-base regions (plain 4/5-base
    # seeds: 668 / 167 expected candidates), 4-base pieces up to 20 (plain 6-base seeds: 42), not beyond (7-base seeds: 10)
    assert plan('mirna', 57_000) == [(4, 3, 12, 15), (4, 4, 16, 20)]
    # ncRNA library (3 Mbases), -n 1 -> three pieces; region 15 has no plain table at all and rides along
    assert plan('ncrna_others', 3_000_000) == [(3, 5, 15, 17), (3, 6, 18, 19)]
    # small libraries keep the plain seeds, and so do rounds without mismatches (mRNA, -n 0)
    assert plan('hairpin', 160_000) == []
    assert plan('snorna', 130_000) == []
    assert plan('mrna', 40_000_000) == []


def test_plan_thresholds_and_limits():
    assert plan('mirna', 57_000, 0) == []                                  # switched off
    assert plan('mirna', 57_000, 1) == [(4, 3, 12, 15), (4, 4, 16, 23)]    # wherever a set can serve
    assert plan('hairpin', 160_000, 1) == [(3, 5, 15, 17), (3, 6, 18, 23)]
    assert plan('mirna', 11, 1) == []                                      # shorter than the pieces
    assert plan('mirna', 13, 1) == [(4, 3, 12, 15)]
    assert plan(8, 10 ** 8, 1) == []                                       # the genome slot has its own index
    assert plan(99, 10 ** 6, 1) == [] and plan(-1, 10 ** 6, 1) == []
    for total in (10 ** 4, 10 ** 5, 10 ** 6, 10 ** 7, 10 ** 8):            # rows are well formed whatever the size
        for lib in LIB_IDS:
            for thr in (1, 4, 16, 64):
                for m, pl, lo, hi in plan(lib, total, thr):
                    assert m in (3, 4) and lo >= m * pl and lo <= hi <= 23


@pytest.mark.parametrize('e,pl', [(1, 5), (1, 6), (2, 3), (2, 4)])
def test_two_pieces_survive(e, pl):
    """e mismatches (read N included) inside the first m * pl bases leave two of the m = e + 2 pieces intact, so the
    placement sits in the bucket of one of the m(m-1)/2 pair keys -- for every placement and mismatch pattern."""
    rng = np.random.default_rng(100 * e + pl)
    m = e + 2
    pairs = list(itertools.combinations(range(m), 2))
    assert len(pairs) == m * (m - 1) // 2 and pairs == sorted(pairs)       # the order k2_map.cuh:gap_piece enumerates
    ref = rng.integers(0, 4, 5000)
    for _ in range(3000):
        n = int(rng.integers(m * pl, 40))
        g = int(rng.integers(0, len(ref) - n))
        q = ref[g:g + n].copy()
        for p in rng.choice(n, int(rng.integers(0, e + 1)), replace=False):
            q[p] = 4 if rng.random() < 0.2 else (q[p] + int(rng.integers(1, 4))) % 4      # 4 = N: matches nothing
        intact = [bool(np.array_equal(q[i * pl:(i + 1) * pl], ref[g + i * pl:g + (i + 1) * pl])) for i in range(m)]
        assert sum(intact) >= 2
        assert any(intact[i] and intact[j] for i, j in pairs)
