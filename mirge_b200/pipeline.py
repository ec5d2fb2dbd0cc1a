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

from . import a2i as a2i_mod
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


def find_record_start(f, offset, size, window=1 << 18):
    """First record boundary at or after byte `offset` of an uncompressed FASTQ file: the start of a line that begins
    with '@', whose line after next begins with '+', whose sequence and quality lines are equally long -- and, when
    visible, whose successor looks the same (a quality line may begin with '@'; a sequence line never begins with '+')."""
    if offset <= 0:
        return 0
    if offset >= size:
        return size
    f.seek(offset - 1)
    buf = f.read(window + 1)
    nl = buf.find(b'\n')                                  # buf[0] is the byte before `offset`
    if nl < 0:
        return size
    starts = [nl + 1]
    while len(starts) < 16:
        nl = buf.find(b'\n', starts[-1])
        if nl < 0:
            break
        starts.append(nl + 1)

    def line(i):
        return buf[starts[i]:starts[i + 1] - 1].rstrip(b'\r')

    def looks_like_record(i):
        if i + 4 >= len(starts):
            return None
        return buf[starts[i]:starts[i] + 1] == b'@' and buf[starts[i + 2]:starts[i + 2] + 1] == b'+' and len(line(i + 1)) == len(line(i + 3))

    for i in range(4):
        a = looks_like_record(i)
        if a is None:
            break
        b = looks_like_record(i + 4)
        if a and b is not False:
            return offset - 1 + starts[i]
    raise MirgeError(-3, 'cannot find a FASTQ record boundary near byte %d (records longer than %d bytes?)' % (offset, window // 8))


def iter_fastq_chunks_shard(path, rank, world, chunk_bytes=CHUNK_BYTES):
    """Record-aligned chunks of this rank's share of one sample: an uncompressed file is cut into `world` byte ranges
    whose ends are moved to record boundaries; a .gz stream cannot be entered in the middle, so every rank inflates it
    and keeps every world-th chunk."""
    if world <= 1:
        yield from iter_fastq_chunks(path, chunk_bytes)
        return
    if path.endswith('.gz'):
        per = max(1 << 20, chunk_bytes // world)
        for k, chunk in enumerate(iter_fastq_chunks(path, per)):
            if k % world == rank:
                yield chunk
        return
    size = os.path.getsize(path)
    with open(path, 'rb') as f:
        lo = find_record_start(f, size * rank // world, size)
        hi = find_record_start(f, size * (rank + 1) // world, size)
        f.seek(lo)
        carry = b''
        left = hi - lo
        while left > 0:
            data = f.read(min(chunk_bytes, left))
            if not data:
                break
            left -= len(data)
            buf = carry + data
            if left == 0:
                yield buf
                carry = b''
                break
            n_nl = buf.count(b'\n')
            cut = buf.rfind(b'\n') + 1
            for _ in range(n_nl % 4):
                cut = buf.rfind(b'\n', 0, cut - 1) + 1
            if cut > 0:
                yield buf[:cut]
            carry = buf[cut:]
        if carry:
            yield carry


def record_aligned_ranges(path, lo, hi, chunk_bytes=CHUNK_BYTES):
    """[(a, b)] byte ranges of an uncompressed FASTQ file covering [lo, hi) in whole records, about chunk_bytes each."""
    size = os.path.getsize(path)
    out = []
    with open(path, 'rb') as f:
        a = lo
        while a < hi:
            b = hi if a + chunk_bytes >= hi else min(hi, find_record_start(f, a + chunk_bytes, size))
            if b <= a:
                b = hi
            out.append((a, b))
            a = b
    return out


class PinnedChunkReader:
    """Streams record-aligned byte ranges of an uncompressed FASTQ file through two pinned staging buffers: a small
    thread pool preads the next range (the GIL is released inside the system call) while the GPU works on the current
    one, so the copies to the device are true asynchronous DMA instead of staged pageable copies."""

    def __init__(self, path, ranges, threads=8):
        import concurrent.futures
        self.path, self.ranges, self.threads = path, ranges, max(1, threads)
        cap = max((b - a for a, b in ranges), default=0) + 64
        self.bufs = [self._alloc(cap), self._alloc(cap)]
        self.pool = concurrent.futures.ThreadPoolExecutor(self.threads)
        self.fd = os.open(path, os.O_RDONLY)

    @staticmethod
    def _alloc(n):
        try:
            import torch
            t = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            return (t, t.numpy())            # the tensor owns the pinned allocation
        except Exception:
            return (None, np.empty(n, np.uint8))

    def _fill(self, k):
        a, b = self.ranges[k]
        view = memoryview(self.bufs[k % 2][1])[:b - a]
        step = -(-(b - a) // self.threads)
        step = (step + 4095) & ~4095

        def part(x):
            got = x
            end = min(b - a, x + step)
            while got < end:
                n = os.preadv(self.fd, [view[got:end]], a + got)
                if n <= 0:
                    raise IOError('short read on %s' % self.path)
                got += n
        futs = [self.pool.submit(part, x) for x in range(0, b - a, step)]
        return futs

    def __iter__(self):
        pending = self._fill(0) if self.ranges else []
        for k in range(len(self.ranges)):
            for f in pending:
                f.result()
            a, b = self.ranges[k]
            pending = self._fill(k + 1) if k + 1 < len(self.ranges) else []
            yield self.bufs[k % 2][1][:b - a]       # the consumer must be done with it before the next-but-one range is read

    def close(self):
        self.pool.shutdown(wait=True)
        os.close(self.fd)


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
                 max_unique=None, max_read_len=None, device=0, log=print, rank=0, world=1):
        self.rank, self.world = int(rank), int(world)
        self._stream = None
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
        if self.world > 1:          # the exchange runs torch.distributed collectives: everything on one torch stream
            import torch
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=torch.device('cuda', self.device))
            e.set_stream(self._stream.cuda_stream)
        e.set_adapter(self.adapter)
        for t in self.libdir.types:
            names, seqs = self.libdir.load(t)
            e.load_library(LIB_IDS[t], names, seqs)
        if self.world == 1 and hasattr(e, 'annotate_while_feeding'):
            e.annotate_while_feeding(True)      # no exchange follows: the rounds hide behind the copies
        self.engine = e

    # -- reference-shaped stages ------------------------------------------------
    def trim_file(self, infile, sample_index):
        """trim_file + quantReads for one sample (trim_file.py:89, quantReads.py:3):
        returns (phred, processed, kept) like trim_file."""
        phred = None
        if self.world > 1:                                # decided by read #0 of the FILE on every rank, trim_file.py:104-109
            opener = gzip.open if infile.endswith('.gz') else open
            with opener(infile, 'rb') as f:
                phred = Engine.detect_phred(f.read(1 << 16))
        if not infile.endswith('.gz'):
            # plain text: this rank's byte range, cut at record boundaries, streamed through pinned double buffers
            size = os.path.getsize(infile)
            with open(infile, 'rb') as f:
                lo = find_record_start(f, size * self.rank // self.world, size) if self.world > 1 else 0
                hi = find_record_start(f, size * (self.rank + 1) // self.world, size) if self.world > 1 else size
            reader = PinnedChunkReader(infile, record_aligned_ranges(infile, lo, hi))
            try:
                for chunk in reader:
                    if phred is None:
                        phred = Engine.detect_phred(chunk)    # decided by read #0, trim_file.py:104-109
                    self.engine.feed_fastq(sample_index, chunk, phred)
                    self.engine.sync()                        # the next range is already being read in the background
            finally:
                reader.close()
        else:
            for chunk in iter_fastq_chunks_shard(infile, self.rank, self.world):
                if phred is None:
                    phred = Engine.detect_phred(chunk)        # decided by read #0, trim_file.py:104-109
                self.engine.feed_fastq(sample_index, chunk, phred)
                self.engine.sync()
        p, k, _ = self.engine.stats()
        return (phred or 33), int(p[sample_index]), int(k[sample_index])

    def _agree(self, code):
        """Distributed runs: every rank learns the most severe error code of the phase (0 = fine)."""
        if self.world <= 1:
            return code
        import torch
        import torch.distributed as dist
        t = torch.tensor([code], dtype=torch.int64, device=torch.device('cuda', self.device))
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())

    def run(self):
        """Trim+collapse every sample (this rank's file chunks), exchange (X1), annotate, tally, reduce (X2).  Grows the
        table and restarts when it fills up -- on every rank alike."""
        import contextlib
        while True:
            code, err = 0, None
            self._new_engine()
            scope = contextlib.nullcontext()
            if self.world > 1:
                import torch
                scope = torch.cuda.stream(self._stream)
            with scope:
                try:
                    t0 = time.time()
                    for i, path in enumerate(self.sample_paths):
                        self.log('Performing quantitation analysis of %s...' % self.sample_names[i])
                        t1 = time.time()
                        self.trim_file(path, i)
                        self.log('It takes: %.2fs' % (time.time() - t1))
                except MirgeError as e:
                    code, err = e.code, e
                code = self._agree(code)
                if code == 0:
                    try:
                        if self.world > 1:
                            from . import dist as mdist
                            mdist.hash_partition(self.engine)
                        self.log('\nPerforming annotation for all of the collasped sequences...')
                        t2 = time.time()
                        self.engine.annotate()
                        self.engine.sync()
                        self.log('All annotation cycles completed (%.2f sec).\n' % (time.time() - t2))
                        self.engine.tally()
                        if self.world > 1:
                            mdist.reduce_engine_tables(self.engine)
                        self.engine.sync()
                        self.timers = {'quant': t2 - t0, 'annotate': time.time() - t2}
                    except MirgeError as e:
                        code, err = e.code, e
                    code = self._agree(code)
                    if code == 0:
                        return self
            if code == -4 and self.max_read_len <= 64:
                # a read beyond the first 2000 records is longer than 64 nt (peek_max_read_len only samples):
                # rerun with the 128-nt record layout; longer reads remain the one hard limit
                self.max_read_len = 128
                self.log('a read is longer than 64 nt; retrying with max_read_len=128')
                continue
            if code != -5 or self.max_unique >= (1 << 31):
                raise err if err is not None else MirgeError(code, 'another rank failed')
            self.max_unique *= 2
            self.log('unique-sequence table full; retrying with max_unique=%d' % self.max_unique)

    def gather_unique(self):
        """export_unique() of the whole job on rank 0 (None elsewhere): every sequence is owned by exactly one rank."""
        uniq = self.engine.export_unique()
        if self.world <= 1:
            return uniq
        import torch.distributed as dist
        parts = [None] * self.world if self.rank == 0 else None
        dist.gather_object(uniq, parts, dst=0)
        if self.rank != 0:
            return None
        return {k: np.concatenate([p[k] for p in parts]) for k in uniq}

    # -- outputs ------------------------------------------------------------------
    def write_outputs(self, outdir, cano_ratio='0.1', tcf=False, isomir_diff=False, gff=False, a_to_i=False):
        """summarize -> miRNAmerge -> filter -> generateReport / writeDataToCSV (__main__.py:381-389)."""
        e = self.engine
        uniq = self.gather_unique()
        if self.rank != 0:
            return None
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
        if a_to_i:  # writeDataToCSV.py:1222-1615
            self.write_a2i(outdir, uniq, logDic, mirDic)
        return dict(logDic=logDic, mirDic=mirDic, len_hist=t['len_hist'])

    def write_a2i(self, outdir, uniq, logDic, mirDic):
        """-ai: A-to-I editing report; the two genome bowtie runs become GPU genome passes (csrc/k2_genome.cuh)."""
        e = self.engine
        if 8 not in e.lib_names:
            e.load_genome(*self.libdir.load_genome())
        names = e.lib_names[0]
        items = []
        for s, c, ro, rf in zip(uniq['seq'], uniq['counts'], uniq['round_of'], uniq['ref_of']):
            if ro == 1 or ro == 9:
                items.append((s.decode(), names[rf] if ro == 1 else '', names[rf] if ro == 9 else '', [int(x) for x in c]))

        def genome_pass(seqs, n):
            return e.genome_pass(seqs, 0, n, 28, 2)            # bowtie genome -n <n> -f -a -3 2

        return a2i_mod.write_a2i(outdir, items, self.sample_names, logDic, mirDic, self.libdir.mature_sequences(),
                                 self.libdir.merged_names(), self.libdir.repetitive_mirnas(), genome_pass, self.log)
