"""Host tail (merge / filter / writers): product code in mirge_b200/tables.py against
the oracle's literal restatement of the reference functions, plus hand-made quirk cases."""
import numpy as np
import pytest

from mirge_b200 import tables as T
from oracle import oracle as O


def test_py2_float_str():
    assert T.py2_float_str(5.0) == '5.0'
    assert T.py2_float_str(1000000.0 * 3 / 7) == '428571.428571'
    assert T.py2_float_str(0.1 + 0.2) == '0.3'
    assert T.py2_float_str(1e16) == '1e+16'
    assert T.py2_float_str(123456789012.0) == '123456789012.0'
    for x in (0.0, 1.5, 2.0 / 3, 123456.789e3, 7e-7):
        assert T.py2_float_str(x) == O.py2_str_float(x)


def _dics(names, quant, iscan):
    return ({n: {'quant': list(map(int, quant[k])), 'iscan': list(map(int, iscan[k]))} for k, n in enumerate(names)},
            {n: {'quant': list(map(int, quant[k])), 'iscan': list(map(int, iscan[k]))} for k, n in enumerate(names)})


def test_merge_creation_quirk_and_deletion():
    names = ['a', 'b', 'c', 'd', 'e']
    quant = np.array([[5, 0], [0, 0], [0, 0], [7, 1], [0, 3]])
    iscan = np.array([[4, 0], [0, 0], [0, 0], [2, 1], [0, 3]])
    lines = ['a/b,a,b\n', 'c/x,c,missing\n', 'd/e,d,e\n']
    mine, ref = _dics(names, quant, iscan)
    T.merge_mirnas(mine, lines, 2)
    O.miRNAmerge(lines, ['s1', 's2'], ref)
    assert mine == ref
    assert 'a/b' in mine and mine['a/b']['quant'] == [5, 0]
    assert 'c/x' not in mine and 'c' not in mine          # all-zero members: merged row never created, member removed
    assert mine['d/e'] == {'quant': [7, 4], 'iscan': [2, 4]}


def test_merge_filter_random_against_restatement():
    rng = np.random.default_rng(0)
    for trial in range(50):
        n, S = 30, int(rng.integers(1, 4))
        names = ['m%d' % i for i in range(n)]
        quant = rng.integers(0, 40, (n, S)) * (rng.random((n, S)) < 0.6)
        iscan = (quant * rng.random((n, S))).astype(int)
        lines = []
        perm = rng.permutation(n)
        for g in range(0, 12, 3):
            lines.append(','.join(['merged%d' % g] + [names[i] for i in perm[g:g + 3]]) + '\n')
        mine, ref = _dics(names, quant, iscan)
        samples = ['s%d.fastq' % i for i in range(S)]
        log1 = {'quantStats': [{} for _ in range(S)]}
        log2 = {'quantStats': [{} for _ in range(S)]}
        T.merge_mirnas(mine, lines, S)
        O.miRNAmerge(lines, samples, ref)
        assert mine == ref
        err1 = err2 = None
        try:
            T.filter_mirnas(mine, samples, log1, '0.1')
        except T.ZeroMiRNAError as e:
            err1 = e
        try:
            O.filter_(ref, samples, log2, '0.1')
        except O.ZeroMiRNA as e:
            err2 = e
        assert (err1 is None) == (err2 is None)
        assert mine == ref
        if err1 is None:
            assert log1 == log2
            assert T.mir_counts_csv(samples, log1, mine) == O.mir_counts_csv(samples, log2, ref)
            assert T.mir_rpm_csv(samples, log1, mine) == O.mir_rpm_csv(samples, log2, ref)


def test_filter_rules():
    d = {'keep': {'quant': [10], 'iscan': [2]}, 'iscan1': {'quant': [10], 'iscan': [1]},
         'ratio': {'quant': [100], 'iscan': [9]}, 'zero': {'quant': [0], 'iscan': [0]}}
    log = {'quantStats': [{}]}
    T.filter_mirnas(d, ['s'], log, '0.1')
    assert d['keep']['quant'] == [10] and d['iscan1']['quant'] == [0] and d['ratio']['quant'] == [0]
    assert log['quantStats'][0] == {'mirnaReadsFiltered': 10, 'mirnaUniqFiltered': 1}
    with pytest.raises(T.ZeroMiRNAError):
        T.filter_mirnas({'x': {'quant': [3], 'iscan': [1]}}, ['s'], {'quantStats': [{}]}, '0.1')


def test_report_row_and_spikein_header_quirk():
    log = T.build_logdic(['a.fastq'], [100], [90], np.arange(12).reshape(1, 12), True)
    log['quantStats'][0].update(mirnaReadsFiltered=1, mirnaUniqFiltered=1)
    rep = T.annotation_report(['a.fastq'], log, True)
    hdr, row = rep.strip().split('\n')
    assert len(row.split(',')) == len(hdr.split(',')) + 1      # the header never gains the spike-in column
    assert rep == O.annotation_report_csv(['a.fastq'], log, True)


def test_arrow_and_pandas_writers_of_mapped_csv_agree(tmp_path):
    """mapped.csv / unmapped.csv: the Arrow writer used for millions of rows gives the bytes pandas gives."""
    import os
    import numpy as np
    from mirge_b200 import tables
    rng = np.random.default_rng(1)
    U = 5000
    lens = rng.integers(16, 64, U).astype(np.uint8)
    raw = np.zeros((U, 64), np.uint8)
    for i in range(U):
        raw[i, :lens[i]] = rng.choice(np.frombuffer(b'ACGTN', np.uint8), lens[i])
    uniq = dict(seq=raw.view('S64').reshape(U), len=lens, counts=rng.integers(0, 1000, (U, 2)).astype(np.uint32),
                round_of=rng.integers(0, 11, U).astype(np.int8), ref_of=rng.integers(0, 50, U).astype(np.int32))
    lib_names = {k: ['n%d_%d' % (k, j) for j in range(50)] for k in range(10)}
    for spike in (False, True):
        names = ['exact miRNA', 'hairpin miRNA', 'mature tRNA', 'primary tRNA', 'snoRNA', 'rRNA', 'ncrna others', 'mRNA', 'isomiR miRNA'] \
            + (['spike-in'] if spike else [])
        u = dict(uniq)
        if not spike:
            u['round_of'] = np.minimum(uniq['round_of'], 9).astype(np.int8)
        a, b = tmp_path / ('a%d' % spike), tmp_path / ('b%d' % spike)
        a.mkdir()
        b.mkdir()
        tables._write_mapped_unmapped_arrow(str(a), u, lib_names, names, ['s1.fastq', 's2.fastq'], spike)
        tables._write_mapped_unmapped_pandas(str(b), u, lib_names, names, ['s1.fastq', 's2.fastq'], spike)
        for fn in ('mapped.csv', 'unmapped.csv'):
            assert open(os.path.join(str(a), fn)).read() == open(os.path.join(str(b), fn)).read(), fn
