"""-gff output (<sample>_isomiRs.gff): the product's interval arithmetic
(mirge_b200/gff.py) against the oracle's literal dash-padding restatement (oracle/oracle.py)."""
import os

import numpy as np
import pytest

from mirge_b200 import gff


def test_known_answers(oracle_mod):
    O = oracle_mod
    assert O.NT2CODE == gff.TRIPLET_CODE and len(gff.TRIPLET_CODE) == 64
    assert (gff.TRIPLET_CODE['AAA'], gff.TRIPLET_CODE['ACG'], gff.TRIPLET_CODE['ACT'], gff.TRIPLET_CODE['TTT']) == ('@', 'd', 's', '%')
    # ACG TAC GTA C+AA -> d N W v, two A's of padding
    assert gff.read_uid('ACGTACGTAC') == O.make_id('ACGTACGTAC') == 'dNWv2'
    assert gff.read_uid('ACGTAC') == O.make_id('ACGTAC') == 'dN'
    assert gff.read_uid('ACGNAC') == O.make_id('ACGNAC') == '.'
    assert gff.read_uid('ACGTACGN') == O.make_id('ACGTACGN') == '.'
    assert gff.cigar_against('ACGTA', 'ACGTA') == O.make_cigar('ACGTA', 'ACGTA') == '5M'
    assert gff.cigar_against('ACGTA', 'ACCTA') == O.make_cigar('ACGTA', 'ACCTA') == '2MG2M'
    assert gff.cigar_against('ACGTA', 'ACGT-') == O.make_cigar('ACGTA', 'ACGT-') == '4MI'
    assert gff.cigar_against('GACGT', '-ACGA') == O.make_cigar('GACGT', '-ACGA') == 'I3MT'
    # hairpin 0..29, mature at 10..31 (22 nt), entry = hairpin[8:38]
    pre = 'GGCTAGCTAG' + 'TGAGGTAGTAGGTTGTATAGTT' + 'TTAGGGTCACACCCACCACTGG'
    lib = pre[8:38]
    mir = lib[2:-6]
    assert gff.describe_read(pre, lib, mir, 3, False) == ('ref_miRNA', 'NA', 11, 32, '22M')
    assert gff.describe_read(pre, lib, mir[:-1], 3, False) == ('isomiR', 'iso_3p:-1', 11, 31, '21M')
    assert gff.describe_read(pre, lib, mir[1:] + 'T', 4, False) == ('isomiR', 'iso_5p:-1,iso_3p:+1', 12, 33, '22M')
    # isomiR round: bowtie saw read[1:-2] at POS 3 + 1 -> the read starts at the mature; non-templated AA at the 3' end
    read = mir + 'AA'
    assert gff.describe_read(pre, lib, read, 4, True) == ('isomiR', 'iso_add:+2', 11, 34, '22MA' + 'A')
    # a substitution at padded column 12 (third base of the mature)
    read = mir[:2] + 'C' + mir[3:]
    assert gff.describe_read(pre, lib, read, 4, True) == ('isomiR', 'iso_snpcentral_supp', 11, 32, '2MC19M')
    assert gff.describe_read('ACGT' * 10, lib, mir, 3, False) is None


def _rand(rng, n):
    return ''.join('ACGT'[x] for x in rng.integers(0, 4, n))


def test_describe_read_matches_literal_restatement(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(5)
    kinds = set()
    for trial in range(30000):
        M = int(rng.integers(18, 26))
        left = int(rng.integers(0, 12))            # includes matures closer than 2 nt to the hairpin start
        right = int(rng.integers(0, 12))           # ... and closer than 6 nt to its end
        mir = _rand(rng, M)
        pre = _rand(rng, left) + mir + _rand(rng, right)
        at = pre.find(mir)
        flank5 = ('XX' + pre)[at:at + 2].replace('X', _rand(rng, 1))      # entry = 2 nt + mature + 6 nt, cut from the hairpin where it reaches
        tail = pre[at + M:at + M + 6]
        lib = flank5 + mir + tail + _rand(rng, 6 - len(tail))
        isomir = bool(rng.integers(0, 2))
        # where bowtie placed the (trimmed) read inside the entry
        n = int(rng.integers(16, 30))
        if isomir:
            pos = int(rng.integers(1, 8))
            core = lib[pos - 1:pos - 1 + n - 3]
            if len(core) < n - 3:                  # bowtie never reports a placement that runs off the entry
                continue
            core = list(core)
            for _ in range(int(rng.integers(0, 3))):
                core[int(rng.integers(0, len(core)))] = 'ACGT'[rng.integers(0, 4)]
            read = _rand(rng, 1) + ''.join(core) + _rand(rng, 2)
            if rng.random() < 0.5:                 # templated ends
                x = at - 2 + pos - 2
                read = (gff._hairpin_slice(pre, x, x + 1).replace('-', 'A') + read[1:-2]
                        + gff._hairpin_slice(pre, x + n - 2, x + n).replace('-', 'C'))
        else:
            pos = int(rng.integers(1, 8))
            read = lib[pos - 1:pos - 1 + n]
            if len(read) < 16:
                continue
        want = O.fill_terminal(pre, lib, lib[2:-6], read, pos, 8 if isomir else 0)
        got = gff.describe_read(pre, lib, read, pos, isomir)
        if not want[4]:
            assert got is None
            continue
        lit = O.analyze_alignment(*want[:4])
        assert got == lit, (pre, lib, read, pos, isomir, got, lit)
        kinds.update(lit[1].split(','))
    assert {'NA', 'iso_snp', 'iso_snp_seed', 'iso_5p:-1', 'iso_3p:+1', 'iso_add:+1', 'iso_5p:+1'} <= kinds


def test_stem_loop_names(oracle_mod, tmp_path):
    O = oracle_mod
    mirbase = ('##gff-version 3\n'
               'chr1\t.\tmiRNA_primary_transcript\t1\t90\t.\t+\t.\tID=MI01;Alias=MI01;Name=hsa-mir-9-2\n'
               'chr2\t.\tmiRNA_primary_transcript\t1\t90\t.\t+\t.\tID=MI02;Alias=MI02;Name=hsa-mir-9-1\n'
               'chr3\t.\tmiRNA_primary_transcript\t1\t90\t.\t+\t.\tID=MI03;Alias=MI03;Name=hsa-mir-7b\n'
               'chr3\t.\tmiRNA_primary_transcript\t1\t90\t.\t+\t.\tID=MI04;Alias=MI04;Name=hsa-mir-7a\n'
               'chr1\t.\tmiRNA\t5\t26\t.\t+\t.\tID=MIMAT1;Alias=MIMAT1;Name=hsa-miR-9-5p;Derives_from=MI01\n'
               'chr2\t.\tmiRNA\t5\t26\t.\t+\t.\tID=MIMAT1_1;Alias=MIMAT1;Name=hsa-miR-9-5p;Derives_from=MI02\n'
               'chr3\t.\tmiRNA\t5\t26\t.\t+\t.\tID=MIMAT2;Alias=MIMAT2;Name=hsa-miR-7-5p;Derives_from=MI03\n'
               'chr3\t.\tmiRNA\t5\t26\t.\t+\t.\tID=MIMAT2_1;Alias=MIMAT2;Name=hsa-miR-7-5p;Derives_from=MI04\n'
               'chr3\t.\tmiRNA\t60\t81\t.\t+\t.\tID=MIMAT3;Alias=MIMAT3;Name=hsa-miR-7b-3p;Derives_from=MI03\n')
    p = tmp_path / 'x.gff3'
    p.write_text(mirbase)
    got = gff.stem_loop_names(str(p), 'miRBase')
    assert got == O.extract_pre_mir_name(mirbase, 'miRBase')
    assert got == {'hsa-miR-9-5p': 'hsa-mir-9-1', 'hsa-miR-7-5p': 'hsa-mir-7a', 'hsa-miR-7b-3p': 'hsa-mir-7b'}
    for name in ('hsa-miR-9-5p', 'hsa-miR-9-3p', 'hsa-miR-7b-5p', 'hsa-miR-1234-5p', 'hsa-let-7a-5p', 'hsa-miR-9'):
        assert gff.stem_loop_of(name, got, 'miRBase') == O.infer_pre_mir_name(name, got, 'miRBase')
    assert gff.stem_loop_of('hsa-miR-1234-5p', got, 'miRBase') == 'hsa-mir-1234'
    mgdb = 'chr1\t.\tmiRNA\t5\t26\t.\t+\t.\tID=Hsa-Mir-9-P1_5p;Alias=x\nchr1\t.\tpre_miRNA\t1\t90\t.\t+\t.\tID=Hsa-Mir-9-P1_pre\n'
    p.write_text(mgdb)
    got = gff.stem_loop_names(str(p), 'MirGeneDB')
    assert got == O.extract_pre_mir_name(mgdb, 'MirGeneDB') == {'Hsa-Mir-9-P1_5p': 'Hsa-Mir-9-P1_pre'}
    for name in ('Hsa-Mir-9-P1_5p', 'Hsa-Mir-9-P1_3p*', 'Hsa-Mir-1_3p'):
        assert gff.stem_loop_of(name, got, 'MirGeneDB') == O.infer_pre_mir_name(name, got, 'MirGeneDB')


def test_gff_files_match_literal_restatement_on_the_synthetic_world(oracle_mod, tmp_path):
    """Whole files: the oracle's collapse + mapping rounds feed both implementations (the CUDA path exports the same
    arrays, tests/test_gpu_pipeline.py)."""
    from synth.world import World
    O = oracle_mod
    w = World(seed=20180626, scale=0.05, snpc_names=True)
    samples = ['s1.fastq', 's2.fastq']
    o = O.Oracle(2, False, 'TGGAATTCTCGGGTGCCAAGGAACTCCAG')
    for t, i in O.LIB_IDS.items():
        o.load_library(i, *w.libs[t])
    for k in range(2):
        o.feed_fastq(k, bytes(w.fastq(60000, seed=1000 + k)), 33)
    o.annotate()
    e = o.export()
    mirna_names, mirna_seqs = w.libs['mirna']
    hp_names, hp_seqs = w.libs['hairpin']
    gff3 = tmp_path / 'synth_miRBase.gff3'
    gff3.write_text(w.gff3_text())
    pre_of = gff.stem_loop_names(str(gff3), 'miRBase')
    assert pre_of == O.extract_pre_mir_name(w.gff3_text(), 'miRBase')
    # literal side
    seqDic = o.seqDic()
    hits = {}
    for k, s in enumerate(e['seqs']):
        ro = int(e['round_of'][k])
        if ro in (1, 9):
            hits[s] = (ro - 1, mirna_names[int(e['ref_of'][k])], int(e['pos_of'][k]) + 1)
    want = O.gff_literal(seqDic, hits, samples, dict(zip(mirna_names, mirna_seqs)), dict(zip(hp_names, hp_seqs)), pre_of, 'miRBase')
    # product side, from export_unique-shaped arrays
    uniq = dict(seq=np.array([s.encode() for s in e['seqs']]), counts=np.asarray(e['counts']), round_of=np.asarray(e['round_of']),
                ref_of=np.asarray(e['ref_of']), pos_of=np.asarray(e['pos_of']))
    recs = gff.gff_rows(uniq, mirna_names, mirna_seqs, hp_names, hp_seqs, pre_of, 'miRBase')
    paths = gff.write_gff(str(tmp_path), recs, samples, 'miRBase')
    assert [os.path.basename(p) for p in paths] == ['s1_isomiRs.gff', 's2_isomiRs.gff']
    n_iso = 0
    for sample, path in zip(samples, paths):
        text = open(path).read()
        header, rows = want[sample]
        assert text.startswith(header)
        got_rows = text[len(header):].splitlines(True)
        assert sorted(got_rows) == sorted(rows) and len(rows) > 500
        n_iso += sum('\tisomiR\t' in r for r in rows)
        assert any('.SNP1\t' in r for r in rows) or sample != 's1.fastq'
    assert n_iso > 200


def test_missing_snpc_entry_is_reported():
    uniq = dict(seq=np.array([b'ACGTACGTACGTACGTACGT']), counts=np.array([[1]]), round_of=np.array([1]), ref_of=np.array([0]), pos_of=np.array([2]))
    with pytest.raises(gff.GffError):
        gff.gff_rows(uniq, ['x-miR-1-5p.SNP1'], ['TTACGTACGTACGTACGTACGTAAAAAA'], ['x-mir-1'], ['GGTTACGTACGTACGTACGTACGTAAAAAAGG'], {'x-miR-1-5p': 'x-mir-1'}, 'miRBase')


def test_product_reproduces_committed_gff_fixture(tmp_path):
    """tests/golden/gff_case.json.gz (made by make_golden_gff.py from the oracle's literal restatement)."""
    import gzip
    import json
    with gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'gff_case.json.gz'), 'rt') as f:
        case = json.load(f)
    gff3 = tmp_path / 'x.gff3'
    gff3.write_text(case['gff3'])
    pre_of = gff.stem_loop_names(str(gff3), 'miRBase')
    u = case['uniq']
    uniq = dict(seq=np.array([r[0].encode() for r in u]), counts=np.array([r[1] for r in u], np.uint32), round_of=np.array([r[2] for r in u], np.int8),
                ref_of=np.array([r[3] for r in u], np.int32), pos_of=np.array([r[4] for r in u], np.int32))
    recs = gff.gff_rows(uniq, case['mirna'][0], case['mirna'][1], case['hairpin'][0], case['hairpin'][1], pre_of, 'miRBase')
    for sample, path in zip(case['samples'], gff.write_gff(str(tmp_path), recs, case['samples'], 'miRBase')):
        text = open(path).read()
        assert text.startswith(case['header'][sample])
        assert sorted(text[len(case['header'][sample]):].splitlines(True)) == case['rows'][sample]
