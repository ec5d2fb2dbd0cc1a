"""GPU: BASELINE configs at their own sizes.
C1 (1M reads, human-scale synthetic libraries): bit-exact tables against the CPU oracle.
C2 (25M reads, one B200): bit-exact against the CPU oracle on the same 25 M reads (unique set, counts, round /
reference / offset / mismatches per sequence, tallies, CSV tables), plus size-independent properties -- count
conservation, a checksum of checksums that is independent of chunking, and doubling under a second feed."""
import numpy as np
import pytest

from conftest import ILLUMINA

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def full_world():
    from synth.world import World
    return World(seed=20180626, scale=1.0)


def _engine(world, max_unique):
    from mirge_b200.engine import Engine, LIB_IDS
    e = Engine(1, False, max_unique, 64)
    e.set_adapter(ILLUMINA)
    for t, i in LIB_IDS.items():
        e.load_library(i, *world.libs[t])
    return e


def test_config_c1_one_million_reads_matches_oracle(full_world, oracle_mod):
    O = oracle_mod
    fq = full_world.fastq(1_000_000, seed=1000).tobytes()
    o = O.Oracle(1, False, ILLUMINA)
    for t, i in O.LIB_IDS.items():
        o.load_library(i, *full_world.libs[t])
    o.feed_fastq(0, fq)
    o.annotate()
    want = O.run_tables(o, ['s1.fastq'], full_world.libs['mirna'][0], full_world.merges)
    e = _engine(full_world, 1 << 20)
    e.feed_fastq(0, fq)
    e.annotate()
    e.tally()
    from mirge_b200 import tables as T
    p, k, u = e.stats()
    t = e.export_tables()
    log = T.build_logdic(['s1.fastq'], p, k, t['lib_totals'], False)
    mir = T.build_mirdic(full_world.libs['mirna'][0], t['mir_quant'], t['mir_iscan'])
    T.merge_mirnas(mir, full_world.merges, 1)
    T.filter_mirnas(mir, ['s1.fastq'], log, '0.1')
    assert T.annotation_report(['s1.fastq'], log, False) == want['report']
    assert T.mir_counts_csv(['s1.fastq'], log, mir) == want['counts']
    assert T.mir_rpm_csv(['s1.fastq'], log, mir) == want['rpm']
    oe, ee = o.export(), e.export_unique()
    od = {s: (int(c[0]), int(r)) for s, c, r in zip(oe['seqs'], oe['counts'], oe['round_of'])}
    gd = {s.decode(): (int(c[0]), int(r)) for s, c, r in zip(ee['seq'], ee['counts'], ee['round_of'])}
    assert gd == od


def _checksum(u):
    """Order-independent checksum of (sequence, count, round, ref, pos) over all uniques."""
    h = np.zeros(len(u['seq']), np.uint64)
    raw = u['seq'].view(np.uint8).reshape(len(u['seq']), -1).astype(np.uint64)
    mult = np.uint64(1099511628211)
    for col in range(raw.shape[1]):
        h = (h ^ raw[:, col]) * mult
    h ^= u['counts'][:, 0].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    h ^= (u['round_of'].astype(np.int64).astype(np.uint64) << np.uint64(40)) ^ (u['ref_of'].astype(np.int64).astype(np.uint64) << np.uint64(8))
    return int(np.bitwise_xor.reduce(h)), int(h.sum(dtype=np.uint64))


def _assert_equals_oracle(O, world, fq_bytes, e, u, t, p, k):
    """The CUDA path's exports against the oracle run on the same reads, sequence by sequence."""
    from mirge_b200 import tables as T
    o = O.Oracle(1, False, ILLUMINA)
    for tn, i in O.LIB_IDS.items():
        o.load_library(i, *world.libs[tn])
    o.feed_fastq(0, fq_bytes, 33)
    o.annotate()
    oe = o.export()
    op, ok = o.stats()
    assert (int(p[0]), int(k[0])) == (int(op[0]), int(ok[0]))
    oseq = np.array([s.encode() for s in oe['seqs']], dtype=u['seq'].dtype)
    assert len(oseq) == len(u['seq'])
    io, ig = np.argsort(oseq, kind='stable'), np.argsort(u['seq'], kind='stable')
    assert np.array_equal(oseq[io], u['seq'][ig])                              # the same unique set
    assert np.array_equal(oe['counts'][io], u['counts'][ig])
    oro, gro = oe['round_of'][io], u['round_of'][ig].astype(np.int32)
    assert np.array_equal(oro, gro)
    hit = oro > 0
    for key in ('ref_of', 'pos_of', 'mm_of'):
        assert np.array_equal(oe[key][io][hit], u[key][ig].astype(np.int32)[hit]), key
    q, isc, tot = o.tally()
    assert np.array_equal(t['mir_quant'], q) and np.array_equal(t['mir_iscan'], isc) and np.array_equal(t['lib_totals'], tot)
    assert np.array_equal(t['len_hist'][:, 0], o.len_hist(len(t['len_hist']) - 1)[:, 0])
    names = world.libs['mirna'][0]
    want = O.run_tables(o, ['s1.fastq'], names, world.merges)
    log = T.build_logdic(['s1.fastq'], p, k, t['lib_totals'], False)
    mir = T.build_mirdic(names, t['mir_quant'], t['mir_iscan'])
    T.merge_mirnas(mir, world.merges, 1)
    T.filter_mirnas(mir, ['s1.fastq'], log, '0.1')
    assert T.annotation_report(['s1.fastq'], log, False) == want['report']
    assert T.mir_counts_csv(['s1.fastq'], log, mir) == want['counts']
    assert T.mir_rpm_csv(['s1.fastq'], log, mir) == want['rpm']


def test_config_c2_properties_at_25m_reads(full_world, oracle_mod):
    import torch
    n = 25_000_000
    rec = full_world.record_bytes(50)
    host = torch.empty(n * rec, dtype=torch.uint8, pin_memory=True)
    full_world.fastq(n, seed=1000, out=host.numpy())
    e = _engine(full_world, 1 << 24)
    e.feed_fastq(0, host.numpy(), 33)
    e.annotate()
    e.tally()
    p, k, U = e.stats()
    t = e.export_tables()
    u = e.export_unique()
    kept = int(k[0])
    assert int(p[0]) == n and 0.9 * n < kept < n
    assert int(u['counts'].sum()) == kept                               # collapse conserves reads
    assert int(t['len_hist'].sum()) == kept
    tot = t['lib_totals'][0]
    assert int(tot[1:11].sum()) == kept and int(tot[0]) == U            # every kept read lands in exactly one class
    assert int(t['mir_quant'].sum()) == int(tot[1]) and (t['mir_iscan'] <= t['mir_quant']).all()
    assert int(u['counts'][(u['round_of'] == 1) | (u['round_of'] == 9)].sum()) == int(tot[1])
    assert len(set(u['seq'].tolist())) == U                             # no duplicate keys in the table
    assert (u['len'] >= 16).all() and int(u['len'].max()) <= 50
    _assert_equals_oracle(oracle_mod, full_world, host.numpy().tobytes(), e, u, t, p, k)   # bit-exact at full size
    c1 = _checksum(u)
    # same reads fed as two device-resident halves: identical table (chunking independence)
    e.reset()
    dev = host.cuda()
    half = (n // 2) * rec
    e.feed_fastq_device(0, dev.data_ptr(), half, 33)
    e.feed_fastq_device(0, dev.data_ptr() + half, n * rec - half, 33)
    e.annotate()
    e.tally()
    u2 = e.export_unique()
    assert _checksum(u2) == c1
    assert np.array_equal(e.export_tables()['lib_totals'], t['lib_totals'])
    # feeding everything once more doubles every count and changes nothing else
    e.feed_fastq_device(0, dev.data_ptr(), n * rec, 33)
    e.annotate()
    e.tally()
    p3, k3, U3 = e.stats()
    t3 = e.export_tables()
    assert (int(p3[0]), int(k3[0]), U3) == (2 * n, 2 * kept, U)
    assert np.array_equal(t3['mir_quant'], 2 * t['mir_quant']) and np.array_equal(t3['lib_totals'][0][1:], 2 * tot[1:])
