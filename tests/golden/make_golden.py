"""Regenerates tests/golden/case1.json.gz.

The reference cannot run offline (Python 2 + cutadapt + bowtie are absent, and
it ships no golden vectors), so these fixtures are OUTPUTS OF THE CPU ORACLE on a
small seeded world: they pin the oracle against regressions and give the GPU
tests a committed target that does not depend on the oracle being rebuilt.
    python tests/golden/make_golden.py
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from synth.world import World          # noqa: E402
from oracle import oracle as O         # noqa: E402

ADAPTER = 'TGGAATTCTCGGGTGCCAAGGAACTCCAG'


def main():
    w = World(seed=20180626, scale=0.004, n_expressed=40)
    fqs = [w.fastq(3000, seed=1000 + s).tobytes().decode() for s in range(2)]
    o = O.Oracle(2, True, ADAPTER)
    for t, i in O.LIB_IDS.items():
        o.load_library(i, *w.libs[t])
    trims = []
    for s, fq in enumerate(fqs):
        st, sp = o.trim_fastq(fq.encode(), 33)
        trims.append([st.tolist(), sp.tolist()])
        o.feed_fastq(s, fq.encode(), 33)
    o.annotate()
    e = o.export()
    samples = ['s1.fastq', 's2.fastq']
    tabs = O.run_tables(o, samples, w.libs['mirna'][0], w.merges)
    q, i, t = o.tally()
    case = {
        'adapter': ADAPTER, 'spike_in': True, 'samples': samples,
        'libs': {k: [v[0], v[1]] for k, v in w.libs.items()}, 'merges': w.merges, 'fastq': fqs, 'trims': trims,
        'unique': {s: [[int(x) for x in c], int(r), int(rf), int(p), int(m)]
                   for s, c, r, rf, p, m in zip(e['seqs'], e['counts'], e['round_of'], e['ref_of'], e['pos_of'], e['mm_of'])},
        'mir_quant': q.tolist(), 'mir_iscan': i.tolist(), 'lib_totals': t.tolist(),
        'report': tabs['report'], 'counts': tabs['counts'], 'rpm': tabs['rpm'],
    }
    with gzip.open(os.path.join(HERE, 'case1.json.gz'), 'wt') as f:
        json.dump(case, f)
    print('unique', len(case['unique']), 'bytes', os.path.getsize(os.path.join(HERE, 'case1.json.gz')))


if __name__ == '__main__':
    main()
