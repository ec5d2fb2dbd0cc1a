"""Annotate-mode orchestration (src/mirge/__main__.py:318-392) on one GPU.

Same order as the reference: per sample trim + collapse, then ONE annotation
pass over the union of unique sequences, then summarize -> merge -> filter ->
report writers.  The reference-shaped entry points (trim_file, quantReads,
runAnnotationPipeline, summarize) keep their names and argument meaning; state
that the reference keeps in Python dicts lives in an Engine (device memory).
"""
import gzip
import os
import time

import numpy as np

from . import gff as gff_mod
from . import tables
from .engine import Engine, LIB_IDS, MirgeError
from .libs import LibraryDir

CHUNK_BYTES = 256 << 20


def iter_fastq_chunks(path, chunk_bytes=CHUNK_BYTES):
    """Yield record-aligned chunks (whole 4-line records) of a .fastq / .fastq.gz file."""
    opener = gzip.open if path.endswith('.gz') else open
    carry = b''
    with opener(path, 'rb') as f:
        while True:
            data = f.read(chunk_bytes)
            if not data:
                break
            buf = carry + data
            n_nl = buf.count(b'\n')
            drop = n_nl % 4
            cut = len(buf)
            if not buf.endswith(b'\n'):
                cut = buf.rfind(b'\n') + 1
            for _ in range(drop):
                cut = buf.rfind(b'\n', 0, cut - 1) + 1
            if cut > 0:
                yield buf[:cut]
            carry = buf[cut:]
    if carry:
        yield carry     # last record without a trailing newline


def peek_max_read_len(paths, n_records=2000):
    mx = 0
    for p in paths:
        opener = gzip.open if p.endswith('.gz') else open
        with opener(p, 'rb') as f:
            for k, line in enumerate(f):
                if k >= 4 * n_records:
                    break
                if k % 4 == 1:
                    mx = max(mx, len(line.rstrip(b'\r\n')))
    return mx


class AnnotateRun:
    """State of one `miRge2.0 annotate` run: the Engine plus the host-side names."""

    def __init__(self, sample_paths, library_path, species, db='miRBase', adapter='none', spike_in=False,
                 max_unique=None, max_read_len=None, device=0, log=print):
        self.sample_paths = list(sample_paths)
        self.sample_names = []
        for p in self.sample_paths:     # __main__.py:339-345
            b = os.path.basename(p)
            self.sample_names.append(b[:-3] if b.endswith('.gz') else b)
        self.spike_in = bool(spike_in)
        self.adapter = adapter
        self.log = log
        self.libdir = LibraryDir(library_path, species, db, spike_in)
        self.libdir.check()
        if max_read_len is None:
            max_read_len = peek_max_read_len(self.sample_paths)
        if max_read_len > 128:
            raise MirgeError(-4, 'reads longer than 128 nt are not supported (found %d)' % max_read_len)
        if max_unique is None:
            total = sum(os.path.getsize(p) * (4 if p.endswith('.gz') else 1) for p in self.sample_paths)
            max_unique = int(min(max(1 << 20, total // 100), 192 << 20))
        self.max_unique, self.max_read_len, self.device = int(max_unique), max_read_len, device
        self.engine = None
        self.annot_name_list = ['exact miRNA', 'hairpin miRNA', 'mature tRNA', 'primary tRNA', 'snoRNA', 'rRNA', 'ncrna others',
                                'mRNA', 'isomiR miRNA'] + (['spike-in'] if spike_in else [])     # __main__.py:333-335
        self.timers = {}

    def _new_engine(self):
        if self.engine is not None:
            self.engine.close()
        e = Engine(len(self.sample_paths), self.spike_in, self.max_unique, max(self.max_read_len, 1), self.device)
        e.set_adapter(self.adapter)
        for t in self.libdir.types:
            names, seqs = self.libdir.load(t)
            e.load_library(LIB_IDS[t], names, seqs)
        self.engine = e

    # -- reference-shaped stages ------------------------------------------------
    def trim_file(self, infile, sample_index):
        """trim_file + quantReads for one sample (trim_file.py:89, quantReads.py:3):
        returns (phred, processed, kept) like trim_file."""
        phred = None
        for chunk in iter_fastq_chunks(infile):
            if phred is None:
                phred = Engine.detect_phred(chunk)        # decided by read #0, trim_file.py:104-109
            self.engine.feed_fastq(sample_index, chunk, phred)
            self.engine.sync()
        p, k, _ = self.engine.stats()
        return (phred or 33), int(p[sample_index]), int(k[sample_index])

    def run(self):
        """Trim+collapse every sample, annotate, tally.  Grows the table and restarts when it fills up."""
        while True:
            try:
                self._new_engine()
                t0 = time.time()
                for i, path in enumerate(self.sample_paths):
                    self.log('Performing quantitation analysis of %s...' % self.sample_names[i])
                    t1 = time.time()
                    self.trim_file(path, i)
                    self.log('It takes: %.2fs' % (time.time() - t1))
                self.log('\nPerforming annotation for all of the collasped sequences...')
                t2 = time.time()
                self.engine.annotate()
                self.engine.sync()
                self.log('All annotation cycles completed (%.2f sec).\n' % (time.time() - t2))
                self.engine.tally()
                self.engine.sync()
                self.timers = {'quant': t2 - t0, 'annotate': time.time() - t2}
                return self
            except MirgeError as e:
                if e.code == -4 and self.max_read_len <= 64:
                    # a read beyond the first 2000 records is longer than 64 nt (peek_max_read_len only samples):
                    # rerun with the 128-nt record layout; longer reads remain the one hard limit
                    self.max_read_len = 128
                    self.log('a read is longer than 64 nt; retrying with max_read_len=128')
                    continue
                if e.code != -5 or self.max_unique >= (1 << 31):
                    raise
                self.max_unique *= 2
                self.log('unique-sequence table full; retrying with max_unique=%d' % self.max_unique)

    # -- outputs ------------------------------------------------------------------
    def write_outputs(self, outdir, cano_ratio='0.1', tcf=False, isomir_diff=False, gff=False):
        """summarize -> miRNAmerge -> filter -> generateReport / writeDataToCSV (__main__.py:381-389)."""
        e = self.engine
        p, k, _ = e.stats()
        t = e.export_tables()
        names = e.lib_names[0]
        logDic = tables.build_logdic(self.sample_names, p, k, t['lib_totals'], self.spike_in)
        rp, ra = t['round_processed'], t['round_aligned']
        logDic['annotStats'] = [{'readsProcessed': int(rp[r]), 'readsAligned': int(ra[r])} for r in range(len(self.annot_name_list))]
        mirDic = tables.build_mirdic(names, t['mir_quant'], t['mir_iscan'])
        merge_lines = self.libdir.merge_lines()
        if merge_lines is None:
            self.log('Cannot find merges file, skipping merge step.\n')
        tables.merge_mirnas(mirDic, merge_lines, len(self.sample_names))
        tables.filter_mirnas(mirDic, self.sample_names, logDic, cano_ratio)
        with open(os.path.join(outdir, 'annotation.report.csv'), 'w') as f:
            f.write(tables.annotation_report(self.sample_names, logDic, self.spike_in))
        uniq = e.export_unique()
        tables.write_mapped_unmapped(outdir, uniq, e.lib_names, self.annot_name_list, self.sample_names, self.spike_in)
        if isomir_diff:
            tables.write_isomirs(outdir, uniq, e.lib_names, self.sample_names, logDic)
        with open(os.path.join(outdir, 'miR.Counts.csv'), 'w') as f:
            f.write(tables.mir_counts_csv(self.sample_names, logDic, mirDic))
        with open(os.path.join(outdir, 'miR.RPM.csv'), 'w') as f:
            f.write(tables.mir_rpm_csv(self.sample_names, logDic, mirDic))
        if tcf:
            tables.write_trim_collapse_fa(outdir, uniq, self.sample_names)
        if gff:     # <sample>_isomiRs.gff (writeDataToCSV.py:621-646)
            pre_of = gff_mod.stem_loop_names(self.libdir.gff3, self.libdir.db)
            mir_names, mir_seqs = self.libdir.load('mirna')
            hp_names, hp_seqs = self.libdir.load('hairpin')
            recs = gff_mod.gff_rows(uniq, mir_names, mir_seqs, hp_names, hp_seqs, pre_of, self.libdir.db)
            gff_mod.write_gff(outdir, recs, self.sample_names, self.libdir.db)
        return dict(logDic=logDic, mirDic=mirDic, len_hist=t['len_hist'])
