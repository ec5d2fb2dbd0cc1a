"""Writes tests/golden/gff_case.json.gz: the <sample>_isomiRs.gff rows the ORACLE's literal restatement
(oracle/oracle.py:gff_literal -- fillTerminal / analyzeAlignment padded-string logic) produces for a seeded
synthetic world whose SNP entries carry .SNPC companions.  The product's interval arithmetic
(mirge_b200/gff.py) must keep reproducing these rows (tests/test_gff.py).

    python tests/golden/make_golden_gff.py
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as O          # noqa: E402
from synth.world import World           # noqa: E402

ADAPTER = 'TGGAATTCTCGGGTGCCAAGGAACTCCAG'


def main():
    w = World(seed=20180626, scale=0.03, snpc_names=True)
    samples = ['g1.fastq', 'g2.fastq']
    o = O.Oracle(2, False, ADAPTER)
    for t, i in O.LIB_IDS.items():
        o.load_library(i, *w.libs[t])
    for k in range(2):
        o.feed_fastq(k, bytes(w.fastq(15000, seed=77 + k)), 33)
    o.annotate()
    e = o.export()
    mirna_names, mirna_seqs = w.libs['mirna']
    hp_names, hp_seqs = w.libs['hairpin']
    pre_of = O.extract_pre_mir_name(w.gff3_text(), 'miRBase')
    keep = [k for k in range(len(e['seqs'])) if int(e['round_of'][k]) in (1, 9)]
    hits = {e['seqs'][k]: (int(e['round_of'][k]) - 1, mirna_names[int(e['ref_of'][k])], int(e['pos_of'][k]) + 1) for k in keep}
    want = O.gff_literal(o.seqDic(), hits, samples, dict(zip(mirna_names, mirna_seqs)), dict(zip(hp_names, hp_seqs)), pre_of, 'miRBase')
    case = dict(samples=samples, gff3=w.gff3_text(), mirna=[mirna_names, mirna_seqs], hairpin=[hp_names, hp_seqs],
                uniq=[[e['seqs'][k], [int(x) for x in e['counts'][k]], int(e['round_of'][k]), int(e['ref_of'][k]), int(e['pos_of'][k])] for k in keep],
                header={s: want[s][0] for s in samples}, rows={s: sorted(want[s][1]) for s in samples})
    with gzip.open(os.path.join(HERE, 'gff_case.json.gz'), 'wt') as f:
        json.dump(case, f)
    print('rows', {s: len(case['rows'][s]) for s in samples}, 'isomiR rows', sum('\tisomiR\t' in r for s in samples for r in case['rows'][s]))


if __name__ == '__main__':
    main()
