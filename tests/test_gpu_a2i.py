"""`annotate -ai` through the CLI on the GPU: the A-to-I report must be byte-identical to the one the same host code
writes when the CPU oracle supplies the annotation and the two genome passes (writeDataToCSV.py:1260, :1488)."""
import os

import pytest

from conftest import ILLUMINA

pytestmark = pytest.mark.gpu
GENOME_SLOT = 8


def test_cli_ai_report_equals_oracle_driven_report(tmp_path, oracle_mod):
    from synth.world import World
    from mirge_b200 import a2i
    from mirge_b200.__main__ import main
    from mirge_b200.libs import LibraryDir
    O = oracle_mod
    w = World(seed=20180626, scale=0.05)
    root = tmp_path / 'libs'
    w.write(str(root))
    gnames, chroms = w.write_genome(str(root), total=36000)
    fqs = [w.fastq(60000, seed=500 + s).tobytes() for s in range(2)]
    paths = []
    for s, fq in enumerate(fqs):
        p = tmp_path / ('s%d.fastq' % (s + 1))
        p.write_bytes(fq)
        paths.append(str(p))
    out = main(['annotate', '-s'] + paths + ['-o', str(tmp_path), '-lib', str(root), '-sp', 'synth', '-ad', 'illumina', '-ai'])
    got = open(os.path.join(out, 'a2IEditing.report.csv')).read()
    assert len(got.strip().split('\n')) > 5

    samples = ['s1.fastq', 's2.fastq']
    o = O.Oracle(2, False, ILLUMINA)
    for t, i in O.LIB_IDS.items():
        o.load_library(i, *w.libs[t])
    o.load_library(GENOME_SLOT, gnames, chroms)
    for s in range(2):
        o.feed_fastq(s, fqs[s])
    o.annotate()
    tabs = O.run_tables(o, samples, w.libs['mirna'][0], w.merges)
    # same sequence order as the GPU run (Python-2 dict order upstream: only the detail file depends on it)
    import csv
    order = [r[0] for r in csv.reader(open(os.path.join(out, 'mapped.csv')))][1:]
    sd = tabs['seqDic']
    items = [(s, sd[s]['annot'][1], sd[s]['annot'][9], list(sd[s]['quant'])) for s in order if s in sd]
    ld = LibraryDir(str(root), 'synth', 'miRBase')
    ref_dir = tmp_path / 'ref'
    ref_dir.mkdir()
    a2i.write_a2i(str(ref_dir), items, samples, tabs['logDic'], tabs['mirDic'], ld.mature_sequences(), ld.merged_names(),
                  ld.repetitive_mirnas(), lambda seqs, n: [o.genome_hits(GENOME_SLOT, s, 0, n, 28, 2) for s in seqs], log=lambda *a: None)
    assert got == open(os.path.join(str(ref_dir), 'a2IEditing.report.csv')).read()
    assert open(os.path.join(out, 'a2IEditing.report.newform.csv')).read() == open(os.path.join(str(ref_dir), 'a2IEditing.report.newform.csv')).read()
