"""Committed golden fixture (tests/golden/case1.json.gz, made by make_golden.py):
the oracle must keep reproducing it (CPU), and the CUDA path must match it (GPU)."""
import gzip
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def case():
    with gzip.open(os.path.join(HERE, 'golden', 'case1.json.gz'), 'rt') as f:
        return json.load(f)


def test_oracle_reproduces_golden(case, oracle_mod):
    O = oracle_mod
    o = O.Oracle(2, case['spike_in'], case['adapter'])
    for t, i in O.LIB_IDS.items():
        o.load_library(i, *case['libs'][t])
    for s, fq in enumerate(case['fastq']):
        st, sp = o.trim_fastq(fq.encode(), 33)
        assert [st.tolist(), sp.tolist()] == case['trims'][s]
        o.feed_fastq(s, fq.encode(), 33)
    o.annotate()
    e = o.export()
    got = {s: [[int(x) for x in c], int(r), int(rf), int(p), int(m)]
           for s, c, r, rf, p, m in zip(e['seqs'], e['counts'], e['round_of'], e['ref_of'], e['pos_of'], e['mm_of'])}
    assert got == case['unique']
    tabs = O.run_tables(o, case['samples'], case['libs']['mirna'][0], case['merges'])
    assert tabs['report'] == case['report'] and tabs['counts'] == case['counts'] and tabs['rpm'] == case['rpm']


@pytest.mark.gpu
def test_cuda_path_matches_golden(case):
    from mirge_b200.engine import Engine, LIB_IDS
    from mirge_b200 import tables as T
    e = Engine(2, case['spike_in'], 1 << 16, 64)
    e.set_adapter(case['adapter'])
    for t, i in LIB_IDS.items():
        e.load_library(i, *case['libs'][t])
    for s, fq in enumerate(case['fastq']):
        e.feed_fastq(s, fq.encode(), 33)
    e.annotate()
    e.tally()
    u = e.export_unique()
    got = {}
    for k, s in enumerate(u['seq']):
        hit = u['round_of'][k] > 0
        got[s.decode()] = [[int(x) for x in u['counts'][k]], int(u['round_of'][k]), int(u['ref_of'][k]) if hit else 0,
                           int(u['pos_of'][k]) if hit else 0, int(u['mm_of'][k]) if hit else 0]
    assert got == case['unique']
    t = e.export_tables()
    assert t['mir_quant'].tolist() == case['mir_quant'] and t['mir_iscan'].tolist() == case['mir_iscan']
    assert t['lib_totals'].tolist() == case['lib_totals']
    p, k, _ = e.stats()
    log = T.build_logdic(case['samples'], p, k, t['lib_totals'], case['spike_in'])
    mir = T.build_mirdic(case['libs']['mirna'][0], t['mir_quant'], t['mir_iscan'])
    T.merge_mirnas(mir, case['merges'], 2)
    T.filter_mirnas(mir, case['samples'], log, '0.1')
    assert T.annotation_report(case['samples'], log, case['spike_in']) == case['report']
    assert T.mir_counts_csv(case['samples'], log, mir) == case['counts']
    assert T.mir_rpm_csv(case['samples'], log, mir) == case['rpm']
