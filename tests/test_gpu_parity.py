"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit-exact."""
import numpy as np
import pytest

from conftest import ILLUMINA

pytestmark = pytest.mark.gpu


def _load_libs(target, world, ids):
    for t, i in ids.items():
        target.load_library(i, *world.libs[t])


def _oracle_run(O, world, fqs, adapter, spike=False):
    o = O.Oracle(len(fqs), spike, adapter)
    _load_libs(o, world, O.LIB_IDS)
    for s, fq in enumerate(fqs):
        o.feed_fastq(s, fq)
    o.annotate()
    return o


def _engine_run(world, fqs, adapter, spike=False, max_read_len=64, max_unique=1 << 20):
    from mirge_b200.engine import Engine, LIB_IDS
    e = Engine(len(fqs), spike, max_unique, max_read_len)
    e.set_adapter(adapter)
    _load_libs(e, world, LIB_IDS)
    for s, fq in enumerate(fqs):
        e.feed_fastq(s, fq)
    e.annotate()
    e.tally()
    return e


def _compare(O, o, e):
    op, ok = o.stats()
    ep, ek, eu = e.stats()
    assert list(op) == list(ep)
    assert list(ok) == list(ek)
    oe = o.export()
    ee = e.export_unique()
    assert eu == len(oe['seqs'])
    od = {s: k for k, s in enumerate(oe['seqs'])}
    eseq = [s.decode() for s in ee['seq']]
    assert set(eseq) == set(od)
    idx = np.array([od[s] for s in eseq])
    assert np.array_equal(ee['len'], np.array([len(s) for s in eseq]))
    assert np.array_equal(ee['counts'], oe['counts'][idx])
    assert np.array_equal(ee['round_of'], oe['round_of'][idx])
    hit = ee['round_of'] > 0
    assert np.array_equal(ee['ref_of'][hit], oe['ref_of'][idx][hit])
    assert np.array_equal(ee['pos_of'][hit], oe['pos_of'][idx][hit])
    assert np.array_equal(ee['mm_of'][hit], oe['mm_of'][idx][hit])
    q, i, t = o.tally()
    et = e.export_tables()
    assert np.array_equal(et['mir_quant'], q)
    assert np.array_equal(et['mir_iscan'], i)
    assert np.array_equal(et['lib_totals'], t)
    lh = o.len_hist(e.max_read_len)
    assert np.array_equal(et['len_hist'], lh)


def test_single_sample_parity(small_world, oracle_mod):
    fq = small_world.fastq(200000, seed=1000).tobytes()
    o = _oracle_run(oracle_mod, small_world, [fq], ILLUMINA)
    e = _engine_run(small_world, [fq], ILLUMINA)
    _compare(oracle_mod, o, e)


def test_two_samples_spikein_parity(small_world, oracle_mod):
    fqs = [small_world.fastq(50000, seed=1000 + s).tobytes() for s in range(2)]
    o = _oracle_run(oracle_mod, small_world, fqs, ILLUMINA, spike=True)
    e = _engine_run(small_world, fqs, ILLUMINA, spike=True)
    _compare(oracle_mod, o, e)
