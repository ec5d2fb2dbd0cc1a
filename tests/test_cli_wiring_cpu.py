"""CPU: the whole `miRge2.0 annotate` command line (argument handling, library directory, stage order,
every table writer incl. -gff / -di / -tcf) with the device Engine replaced by an oracle-backed stand-in.
The GPU tests run the same command line on the real Engine; this one keeps the host wiring covered
where no GPU is present.  (Test infrastructure only: the product never falls back to the oracle.)"""
import os

import numpy as np
import pytest

from conftest import ILLUMINA


class OracleEngine:
    """The slice of mirge_b200.engine.Engine that pipeline.AnnotateRun uses, answered by the CPU oracle."""

    def __init__(self, n_samples=1, spike_in=False, max_unique=1 << 22, max_read_len=64, device=0):
        from oracle import oracle as O
        self.O, self.S, self.spike_in = O, int(n_samples), bool(spike_in)
        self.max_read_len = 64 if max_read_len <= 64 else 128
        self.o = None
        self.lib_names = {}
        self._libs = []

    def set_adapter(self, spec, *a, **k):
        self.o = self.O.Oracle(self.S, self.spike_in, spec)
        for args in self._libs:
            self.o.load_library(*args)

    def load_library(self, lib_id, names, seqs):
        self.lib_names[lib_id] = list(names)
        self._libs.append((lib_id, names, seqs))
        if self.o is not None:
            self.o.load_library(lib_id, names, seqs)

    @staticmethod
    def detect_phred(buf):
        from oracle import oracle as O
        return O.Oracle.detect_phred(bytes(buf))

    def feed_fastq(self, sample, buf, phred_base=None):
        return self.o.feed_fastq(sample, bytes(buf), phred_base)

    def sync(self):
        pass

    def close(self):
        pass

    def stats(self):
        p, k = self.o.stats()
        return p, k, len(self.o.export()['seqs'])

    def annotate(self):
        self._rounds = self.o.annotate()

    def tally(self):
        self._tally = self.o.tally()

    def export_unique(self, with_seq=True):
        e = self.o.export()
        return dict(seq=np.array([s.encode() for s in e['seqs']], dtype='S%d' % self.max_read_len), len=np.array([len(s) for s in e['seqs']], np.uint8),
                    counts=np.asarray(e['counts'], np.uint32), round_of=np.asarray(e['round_of'], np.int8),
                    ref_of=np.asarray(e['ref_of'], np.int32), pos_of=np.asarray(e['pos_of'], np.int32), mm_of=np.asarray(e['mm_of'], np.int8))

    def export_tables(self):
        q, i, t = self._tally
        rp, ra = self._rounds
        return dict(mir_quant=q, mir_iscan=i, lib_totals=t, len_hist=self.o.len_hist(self.max_read_len + 1)[:self.max_read_len + 1],
                    round_processed=rp, round_aligned=ra)


@pytest.fixture()
def cli(monkeypatch):
    from mirge_b200 import pipeline
    monkeypatch.setattr(pipeline, 'Engine', OracleEngine)
    from mirge_b200.__main__ import main
    return main


def test_cli_writes_every_table_with_gff(cli, tmp_path, oracle_mod):
    from synth.world import World
    O = oracle_mod
    w = World(seed=20180626, scale=0.05, snpc_names=True)
    libs = tmp_path / 'libs'
    w.write(str(libs))
    fqs = [w.fastq(30000, seed=1000 + s).tobytes() for s in range(2)]
    names = ['sampleA.fastq', 'sampleB.fastq']
    for n, fq in zip(names, fqs):
        (tmp_path / n).write_bytes(fq)
    out = cli(['annotate', '-s', str(tmp_path / names[0]), str(tmp_path / names[1]), '-o', str(tmp_path), '-lib', str(libs),
               '-sp', 'synth', '-ad', 'illumina', '-gff', '-di', '-tcf', '-pb', '/none', '-cpu', '3'])
    o = O.Oracle(2, False, ILLUMINA)
    for t, i in O.LIB_IDS.items():
        o.load_library(i, *w.libs[t])
    for s, fq in enumerate(fqs):
        o.feed_fastq(s, fq)
    o.annotate()
    tabs = O.run_tables(o, names, w.libs['mirna'][0], w.merges)
    assert open(os.path.join(out, 'annotation.report.csv')).read() == tabs['report']
    assert open(os.path.join(out, 'miR.Counts.csv')).read() == tabs['counts']
    assert open(os.path.join(out, 'miR.RPM.csv')).read() == tabs['rpm']
    for fn in ('mapped.csv', 'unmapped.csv', 'isomirs.csv', 'isomirs.samples.csv', 'sampleA.trim.collapse.fa', 'sampleB.trim.collapse.fa',
               'sampleA_isomiRs.gff', 'sampleB_isomiRs.gff'):
        assert os.path.getsize(os.path.join(out, fn)) > 0, fn
    # the GFF against the literal restatement
    e = o.export()
    mirna_names, mirna_seqs = w.libs['mirna']
    hp_names, hp_seqs = w.libs['hairpin']
    pre_of = O.extract_pre_mir_name(w.gff3_text(), 'miRBase')
    hits = {s: (int(e['round_of'][k]) - 1, mirna_names[int(e['ref_of'][k])], int(e['pos_of'][k]) + 1)
            for k, s in enumerate(e['seqs']) if int(e['round_of'][k]) in (1, 9)}
    want = O.gff_literal(o.seqDic(), hits, names, dict(zip(mirna_names, mirna_seqs)), dict(zip(hp_names, hp_seqs)), pre_of, 'miRBase')
    for n in names:
        text = open(os.path.join(out, os.path.splitext(n)[0] + '_isomiRs.gff')).read()
        header, rows = want[n]
        assert text.startswith(header) and sorted(text[len(header):].splitlines(True)) == sorted(rows) and len(rows) > 200


def test_cli_gff_needs_the_coordinate_file(cli, tmp_path, small_world, capsys):
    libs = tmp_path / 'libs'
    small_world.write(str(libs))
    os.remove(os.path.join(str(libs), 'synth', 'annotation.Libs', 'synth_miRBase.gff3'))
    p = tmp_path / 's.fastq'
    p.write_bytes(small_world.fastq(200, seed=1).tobytes())
    with pytest.raises(SystemExit) as ex:
        cli(['annotate', '-s', str(p), '-o', str(tmp_path), '-lib', str(libs), '-sp', 'synth', '-ad', 'illumina', '-gff'])
    assert ex.value.code == 1 and 'gff3' in capsys.readouterr().out
