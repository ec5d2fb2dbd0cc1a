"""GPU: `miRge2.0 annotate ... -gff` through the CLI; <sample>_isomiRs.gff must equal the oracle's literal
restatement of the reference's -gff path fed by the oracle's own collapse and mapping rounds.
(Runs last: the field arithmetic itself is covered on the CPU by tests/test_gff.py.)"""
import os

import pytest

from conftest import ILLUMINA

pytestmark = pytest.mark.gpu


def test_cli_gff_matches_oracle(tmp_path, oracle_mod):
    from synth.world import World
    from mirge_b200 import gff
    from mirge_b200.__main__ import main
    O = oracle_mod
    w = World(seed=20180626, scale=0.05, snpc_names=True)
    libs = tmp_path / 'libs'
    w.write(str(libs))
    fqs = [w.fastq(40000, seed=1000 + s).tobytes() for s in range(2)]
    names = ['sampleA.fastq', 'sampleB.fastq']
    for n, fq in zip(names, fqs):
        (tmp_path / n).write_bytes(fq)
    out = main(['annotate', '-s', str(tmp_path / names[0]), str(tmp_path / names[1]), '-o', str(tmp_path), '-lib', str(libs),
                '-sp', 'synth', '-ad', 'illumina', '-gff'])
    o = O.Oracle(2, False, ILLUMINA)
    for t, i in O.LIB_IDS.items():
        o.load_library(i, *w.libs[t])
    for s, fq in enumerate(fqs):
        o.feed_fastq(s, fq)
    o.annotate()
    e = o.export()
    mirna_names, mirna_seqs = w.libs['mirna']
    hp_names, hp_seqs = w.libs['hairpin']
    pre_of = O.extract_pre_mir_name(w.gff3_text(), 'miRBase')
    hits = {}
    for k, s in enumerate(e['seqs']):
        ro = int(e['round_of'][k])
        if ro in (1, 9):
            hits[s] = (ro - 1, mirna_names[int(e['ref_of'][k])], int(e['pos_of'][k]) + 1)
    want = O.gff_literal(o.seqDic(), hits, names, dict(zip(mirna_names, mirna_seqs)), dict(zip(hp_names, hp_seqs)), pre_of, 'miRBase')
    for n in names:
        text = open(os.path.join(out, os.path.splitext(n)[0] + '_isomiRs.gff')).read()
        header, rows = want[n]
        assert text.startswith(header)
        assert sorted(text[len(header):].splitlines(True)) == sorted(rows) and len(rows) > 300
    assert gff.TRIPLET_CODE == O.NT2CODE
