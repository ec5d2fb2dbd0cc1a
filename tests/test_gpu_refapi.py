"""GPU: the reference-shaped stage functions (mirge_b200.refapi), called in the order of
src/mirge/__main__.py:346-389, must leave the four dicts as the oracle's restatement does."""
import os

import numpy as np
import pytest

from conftest import ILLUMINA

pytestmark = pytest.mark.gpu


def test_reference_call_sequence(tmp_path, small_world, oracle_mod):
    from mirge_b200 import refapi, tables
    O = oracle_mod
    root = str(tmp_path / 'libs')
    small_world.write(root)
    idx = os.path.join(root, 'synth', 'index.Libs')
    files = {k: os.path.join(idx, 'synth_' + v) for k, v in dict(
        mirna='mirna_miRBase', hairpin='hairpin_miRBase', mature_trna='mature_trna', pre_trna='pre_trna', snorna='snorna',
        rrna='rrna', ncrna_others='ncrna_others', mrna='mrna', spike='spike-in').items()}
    fqs = [small_world.fastq(30000, seed=77 + s).tobytes() for s in range(2)]
    paths = []
    for s, fq in enumerate(fqs):
        p = tmp_path / ('smp%d.fastq' % s)
        p.write_bytes(fq)
        paths.append(str(p))
    sampleList = ['smp0.fastq', 'smp1.fastq']
    spikeIn = True
    annotNameList = ['exact miRNA', 'hairpin miRNA', 'mature tRNA', 'primary tRNA', 'snoRNA', 'rRNA', 'ncrna others', 'mRNA',
                     'isomiR miRNA', 'spike-in']
    # --- the reference's main(), lines 318-389, with the GPU-backed functions ---
    seqDic, mirDic, readLengthDic = {}, {}, {}
    logDic = {'quantStats': [], 'annotStats': []}
    refapi.start_session(len(sampleList), spikeIn, 1 << 18, 64)
    for i, sampleFile in enumerate(paths):
        cleaned = os.path.join(str(tmp_path), os.path.splitext(sampleList[i])[0] + '.trim.fastq')
        phred, processed, kept = refapi.trim_file(sampleFile, ILLUMINA, cleaned, 4)
        logDic['quantStats'].append({'filename': sampleList[i], 'totalReads': processed, 'trimmedReads': kept})
        refapi.quantReads(cleaned, seqDic, readLengthDic, len(sampleList), i, sampleList, True, spikeIn)
    refapi.runAnnotationPipeline('/nonexistent', seqDic, '4', False, annotNameList, str(tmp_path), logDic, files['mirna'],
                                 files['hairpin'], files['mature_trna'], files['pre_trna'], files['snorna'], files['rrna'],
                                 files['ncrna_others'], files['mrna'], spikeIn, files['spike'], sampleList=sampleList)
    refapi.summarize(seqDic, sampleList, logDic, mirDic, files['mirna'], str(tmp_path), spikeIn, '/nonexistent')
    # --- oracle ---
    o = O.Oracle(2, spikeIn, ILLUMINA)
    for t, i in O.LIB_IDS.items():
        o.load_library(i, *small_world.libs[t])
    for s, fq in enumerate(fqs):
        o.feed_fastq(s, fq)
    rp, ra = o.annotate()
    want_seq = o.seqDic()
    assert seqDic == want_seq
    proc, kept = o.stats()
    wlog = {'quantStats': [{'filename': s, 'totalReads': int(proc[i]), 'trimmedReads': int(kept[i])} for i, s in enumerate(sampleList)],
            'annotStats': []}
    wmir = {}
    O.summarize(want_seq, sampleList, wlog, wmir, small_world.libs['mirna'][0], spikeIn)
    assert mirDic == wmir
    assert logDic['quantStats'] == wlog['quantStats']
    assert [d['readsProcessed'] for d in logDic['annotStats']] == [int(x) for x in rp]
    assert [d['readsAligned'] for d in logDic['annotStats']] == [int(x) for x in ra]
    lh = o.len_hist(64)
    want_len = {L: [int(x) for x in lh[L]] for L in range(65) if lh[L].any()}
    assert readLengthDic == want_len
    assert os.path.isfile(str(tmp_path / 'smp0.trim.collapse.fa')) and os.path.isfile(str(tmp_path / 'smp1.trim.collapse.fa'))
    # merge + filter on both sides, then the tables
    tables.merge_mirnas(mirDic, small_world.merges, 2)
    tables.filter_mirnas(mirDic, sampleList, logDic, '0.1')
    O.miRNAmerge(small_world.merges, sampleList, wmir)
    O.filter_(wmir, sampleList, wlog, '0.1')
    assert tables.mir_counts_csv(sampleList, logDic, mirDic) == O.mir_counts_csv(sampleList, wlog, wmir)
    assert tables.annotation_report(sampleList, logDic, spikeIn) == O.annotation_report_csv(sampleList, wlog, spikeIn)
