"""GPU: whole `miRge2.0 annotate` runs through the CLI against the oracle's tables,
chunked / device-resident feeding, the exchange records round trip and error paths."""
import glob
import gzip
import os

import numpy as np
import pytest

from conftest import ILLUMINA

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def world_dir(tmp_path_factory, small_world):
    root = tmp_path_factory.mktemp('libs')
    small_world.write(str(root))
    return str(root)


def _oracle_tables(O, world, fq_list, names, spike):
    o = O.Oracle(len(fq_list), spike, ILLUMINA)
    for t, i in O.LIB_IDS.items():
        o.load_library(i, *world.libs[t])
    for s, fq in enumerate(fq_list):
        o.feed_fastq(s, fq)
    o.annotate()
    return o, O.run_tables(o, names, world.libs['mirna'][0], world.merges)


@pytest.mark.parametrize('spike', [False, True])
def test_cli_matches_oracle_tables(tmp_path, small_world, world_dir, oracle_mod, spike):
    from mirge_b200.__main__ import main
    fqs = [small_world.fastq(40000, seed=1000 + s).tobytes() for s in range(2)]
    p1 = tmp_path / 'sampleA.fastq'
    p1.write_bytes(fqs[0])
    p2 = tmp_path / 'sampleB.fastq.gz'
    with gzip.open(p2, 'wb') as f:
        f.write(fqs[1])
    argv = ['annotate', '-s', str(p1), str(p2), '-o', str(tmp_path), '-d', 'mirbase', '-pb', '/nonexistent', '-lib', world_dir,
            '-sp', 'synth', '-ad', 'illumina', '-cpu', '4', '-tcf', '-di'] + (['-spikeIn'] if spike else [])
    out = main(argv)
    assert os.path.basename(out).startswith('miRge.')
    names = ['sampleA.fastq', 'sampleB.fastq']          # .gz is stripped from the column name (__main__.py:339-345)
    o, tabs = _oracle_tables(oracle_mod, small_world, fqs, names, spike)
    assert open(os.path.join(out, 'annotation.report.csv')).read() == tabs['report']
    assert open(os.path.join(out, 'miR.Counts.csv')).read() == tabs['counts']
    assert open(os.path.join(out, 'miR.RPM.csv')).read() == tabs['rpm']
    mapped, unmapped = oracle_mod.mapped_unmapped_rows(tabs['seqDic'], spike)
    hdr = 'uniqueSequence,annotFlag,exact miRNA,hairpin miRNA,mature tRNA,primary tRNA,snoRNA,rRNA,ncrna others,mRNA,isomiR miRNA' \
          + (',spike-in' if spike else '') + ',sampleA.fastq,sampleB.fastq'
    for fn, want in (('mapped.csv', mapped), ('unmapped.csv', unmapped)):
        lines = open(os.path.join(out, fn)).read().strip().split('\n')
        assert lines[0] == hdr
        assert set(lines[1:]) == want and len(lines) - 1 == len(want)     # row order is unspecified in the reference
    assert len(glob.glob(os.path.join(out, '*.trim.collapse.fa'))) == 2
    fa = open(os.path.join(out, 'sampleA.trim.collapse.fa')).read().split('\n')
    top = max(tabs['seqDic'].items(), key=lambda kv: (kv[1]['quant'][0], kv[0]))
    assert fa[0] == '>seq1_%d' % top[1]['quant'][0] and fa[1] == top[0]
    want1, want2 = oracle_mod.isomir_tables(tabs['seqDic'], names, tabs['logDic'])       # -di tables, row order unspecified upstream
    for fn, want in (('isomirs.csv', want1), ('isomirs.samples.csv', want2)):
        got_lines = open(os.path.join(out, fn)).read().strip().split('\n')
        want_lines = want.strip().split('\n')
        assert got_lines[0] == want_lines[0] and sorted(got_lines[1:]) == sorted(want_lines[1:]) and len(got_lines) > 10


def test_cli_aborts_when_a_sample_has_no_mirna(tmp_path, world_dir):
    from mirge_b200.__main__ import main
    rng = np.random.default_rng(0)
    reads = [''.join('ACGT'[x] for x in rng.integers(0, 4, 40)) for _ in range(200)]
    p = tmp_path / 'noise.fastq'
    p.write_text(''.join('@r%d\n%s\n+\n%s\n' % (k, r, 'I' * 40) for k, r in enumerate(reads)))
    with pytest.raises(SystemExit) as ei:
        main(['annotate', '-s', str(p), '-o', str(tmp_path), '-lib', world_dir, '-sp', 'synth', '-ad', 'illumina'])
    assert ei.value.code == 1                                              # filter.py:25-32


def test_chunked_and_device_resident_feeds_agree(small_world):
    import torch
    from mirge_b200.engine import Engine, LIB_IDS
    fq = small_world.fastq(30000, seed=5)
    rec = small_world.record_bytes(50)
    tables = []
    for mode in ('one', 'chunks', 'device', 'no_trailing_newline'):
        e = Engine(1, False, 1 << 18, 64)
        e.set_adapter(ILLUMINA)
        for t, i in LIB_IDS.items():
            e.load_library(i, *small_world.libs[t])
        if mode == 'one':
            e.feed_fastq(0, fq, 33)
        elif mode == 'chunks':
            for a in range(0, 30000, 7000):
                e.feed_fastq(0, fq[a * rec:min(a + 7000, 30000) * rec], 33)
                e.sync()
        elif mode == 'device':
            d = torch.from_numpy(fq).cuda()
            e.feed_fastq_device(0, d.data_ptr(), d.numel(), 33)
        else:
            e.feed_fastq(0, fq[:-1], 33)
        e.annotate()
        e.tally()
        u = e.export_unique()
        t = e.export_tables()
        tables.append((sorted(zip(u['seq'].tolist(), u['counts'][:, 0].tolist(), u['round_of'].tolist())),
                       t['mir_quant'].tolist(), t['lib_totals'].tolist(), [int(x) for x in e.stats()[0]]))
    assert tables[0] == tables[1] == tables[2] == tables[3]


def test_exchange_records_round_trip(small_world):
    """Export records for 3 owners, import them all again: same table (multi-GPU exchange on one device)."""
    import torch
    from mirge_b200.engine import Engine, LIB_IDS, owner_of
    fqs = [small_world.fastq(20000, seed=50 + s) for s in range(2)]
    e = Engine(2, False, 1 << 18, 64)
    e.set_adapter(ILLUMINA)
    for t, i in LIB_IDS.items():
        e.load_library(i, *small_world.libs[t])
    for s in range(2):
        e.feed_fastq(s, fqs[s], 33)
    u0 = e.export_unique()
    before = dict(zip(u0['seq'].tolist(), map(tuple, u0['counts'].tolist())))
    U, rb = len(u0['seq']), e.record_bytes()
    n_rec = int((u0['counts'] > 0).sum())                         # one record per (sequence, sample) with reads
    assert rb == 32
    buf = torch.empty(n_rec * rb, dtype=torch.uint8, device='cuda')
    pc = e.export_records_device(3, buf.data_ptr(), n_rec)
    assert int(pc.sum()) == n_rec and n_rec > U
    own = np.zeros(3, np.int64)
    for s, c in zip(u0['seq'].tolist(), u0['counts']):
        own[owner_of(s, 3)] += int((c > 0).sum())
    assert own.tolist() == [int(x) for x in pc]                 # host hash == device hash
    e.reset_unique()
    off = 0
    for part in range(3):                                         # import part by part, duplicated once to test count merging
        e.import_records_device(buf.data_ptr() + off * rb, int(pc[part]))
        off += int(pc[part])
    e.import_records_device(buf.data_ptr(), int(pc[0]))
    e.sync()
    u1 = e.export_unique()
    after = dict(zip(u1['seq'].tolist(), map(tuple, u1['counts'].tolist())))
    first = set(s for s in u0['seq'].tolist() if owner_of(s, 3) == 0)
    assert set(after) == set(before)
    for s, c in before.items():
        assert after[s] == (tuple(2 * x for x in c) if s in first else c)


def test_error_paths():
    from mirge_b200.engine import Engine, MirgeError
    e = Engine(1, False, 1 << 12, 64)
    e.set_adapter(ILLUMINA)
    with pytest.raises(MirgeError) as ei:
        e.feed_fastq(0, b'@r\nACGT\n+\nIIII\n@r2\nACGT\n+\n', 33)     # 7 lines
        e.sync()                                                        # lines are counted on the device: surfaces at the next sync
    assert ei.value.code == -3 and 'whole number' in str(ei.value)
    e.reset()
    e.feed_fastq(0, b'@r\nACGT\n+\nIII\n', 33)                         # quality shorter than sequence
    with pytest.raises(MirgeError) as ei:
        e.sync()
    assert ei.value.code == -3
    e2 = Engine(1, False, 1 << 12, 64)
    e2.set_adapter('none')
    e2.feed_fastq(0, ('@r\n%s\n+\n%s\n' % ('A' * 70, 'I' * 70)).encode(), 33)
    with pytest.raises(MirgeError) as ei:
        e2.sync()
    assert ei.value.code == -4
    e3 = Engine(1, False, 64, 64)
    e3.set_adapter('none')
    rng = np.random.default_rng(0)
    fq = ''.join('@r\n%s\n+\n%s\n' % (''.join('ACGT'[x] for x in rng.integers(0, 4, 30)), 'I' * 30) for _ in range(500)).encode()
    e3.feed_fastq(0, fq, 33)
    with pytest.raises(MirgeError) as ei:
        e3.sync()
    assert ei.value.code == -5
    for bad in ('ACGTNACGT', '1234', 'A' * 65):
        with pytest.raises(MirgeError):
            e3.set_adapter(bad)
    with pytest.raises(MirgeError):
        Engine(1, False, 1 << 12, 300)
