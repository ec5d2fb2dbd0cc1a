"""Pair-of-pieces seed tables (k2_map.cuh:GapSet) only shorten the candidate lists of short queries: with the tables
forced on for every library and region they can serve, forced off, and at the default threshold, the rounds must
report the alignments of the CPU oracle bit for bit."""
import numpy as np
import pytest

from conftest import ILLUMINA
from test_gpu_parity import _compare, _load_libs, _oracle_run

pytestmark = pytest.mark.gpu


def _engine_run(world, fqs, adapter, pair_seeds, max_unique=1 << 20, debug=0, max_read_len=64, **kw):
    from mirge_b200.engine import Engine, LIB_IDS
    e = Engine(len(fqs), False, max_unique, max_read_len)
    e.set_adapter(adapter, **kw)
    if debug:
        e.set_trim_mode(debug)
    e.pair_seeds(pair_seeds)                 # read when the libraries are loaded
    _load_libs(e, world, LIB_IDS)
    for s, fq in enumerate(fqs):
        e.feed_fastq(s, fq)
    e.annotate()
    e.tally()
    return e


def short_fragments(world, rng, n_per_lib=3000, lo=16, hi=27):
    """Short pieces of every library with 0-3 substitutions, now and then an N: the queries whose seeds are short."""
    reads = []
    for t, (names, seqs) in world.libs.items():
        if not seqs:
            continue
        for _ in range(n_per_lib):
            s = seqs[int(rng.integers(0, len(seqs)))]
            n = int(rng.integers(lo, hi))
            if len(s) < n:
                continue
            a = int(rng.integers(0, len(s) - n + 1))
            r = list(s[a:a + n])
            for _ in range(int(rng.integers(0, 4))):
                p = int(rng.integers(0, n))
                r[p] = 'ACGT'[('ACGT'.index(r[p]) + int(rng.integers(1, 4))) % 4] if r[p] in 'ACGT' else 'A'
            if rng.random() < 0.05:
                r[int(rng.integers(0, n))] = 'N'
            reads.append(''.join(r))
    return reads


def _fq(reads):
    return ''.join('@r%d\n%s\n+\n%s\n' % (k, r, 'I' * len(r)) for k, r in enumerate(reads)).encode()


@pytest.mark.parametrize('pair_seeds', [1, 0, 16])
def test_short_queries_with_and_without_pair_tables(small_world, oracle_mod, pair_seeds):
    rng = np.random.default_rng(21)
    fq = _fq(short_fragments(small_world, rng))
    o = _oracle_run(oracle_mod, small_world, [fq], 'none')
    e = _engine_run(small_world, [fq], 'none', pair_seeds)
    _compare(oracle_mod, o, e)
    assert int((e.export_unique()['round_of'] == 9).sum()) > 100        # the isomiR round (-v 2) did find placements


def test_short_queries_four_word_kernels(small_world, oracle_mod):
    """The same with the 128-nt (four-word) instantiation of the round kernels."""
    rng = np.random.default_rng(22)
    fq = _fq(short_fragments(small_world, rng, 1500, 16, 40))
    o = _oracle_run(oracle_mod, small_world, [fq], 'none')
    e = _engine_run(small_world, [fq], 'none', 1, max_read_len=128)
    _compare(oracle_mod, o, e)


def test_whole_workload_with_pair_tables_everywhere(small_world, oracle_mod):
    fq = small_world.fastq(200000, seed=1234).tobytes()
    o = _oracle_run(oracle_mod, small_world, [fq], ILLUMINA)
    e = _engine_run(small_world, [fq], ILLUMINA, 1)
    _compare(oracle_mod, o, e)


def test_list_routing_does_not_change_results(small_world, oracle_mod):
    """Debug bit 0x1000 sends every miss through the poly-T round instead of past it: same tables."""
    fq = small_world.fastq(100000, seed=77).tobytes()
    o = _oracle_run(oracle_mod, small_world, [fq], ILLUMINA)
    e = _engine_run(small_world, [fq], ILLUMINA, 16, debug=0x1000)
    _compare(oracle_mod, o, e)


def test_library_with_ambiguous_bases_under_pair_tables(small_world, oracle_mod):
    """N in a library is packed as A in the seed tables; the verification has to reject those placements."""
    from synth.world import World
    w = World(seed=9, scale=0.01, n_expressed=40)
    rng = np.random.default_rng(2)
    libs = dict(w.libs)
    for t in ('mirna', 'ncrna_others', 'snorna'):
        names, seqs = libs[t]
        seqs = list(seqs)
        for k in range(0, len(seqs), 3):
            p = int(rng.integers(0, len(seqs[k])))
            seqs[k] = seqs[k][:p] + 'N' + seqs[k][p + 1:]
        libs[t] = (names, seqs)
    w.libs = libs
    reads = short_fragments(w, rng, 2000)
    reads = [r.replace('N', 'A') if rng.random() < 0.5 else r for r in reads]
    fq = _fq(reads)
    o = _oracle_run(oracle_mod, w, [fq], 'none')
    e = _engine_run(w, [fq], 'none', 1)
    _compare(oracle_mod, o, e)
