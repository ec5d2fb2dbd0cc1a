"""The reference's stage functions, same names and argument meaning, backed by the GPU.

`src/mirge/__main__.py:346-389` calls, in this order,
    trim_file(infile, adapter, outfile, threads, phred)            utils/trim_file.py:89
    quantReads(inputfile, seqDic, readLengthDic, sampleCount, sampleIndex, sampleList, tcf, spikeIn)   utils/quantReads.py:3
    runAnnotationPipeline(bowtieBinary, seqDic, numCPU, phred64, annotNameList, outputdir, logDic, file_mirna, ...)  utils/runAnnotationPipeline.py:566
    summarize(seqDic, sampleList, logDic, mirDic, file_mirna, outputdir, spikeIn, bowtieBinary)        utils/summarize.py:3
and owns four dicts (seqDic, mirDic, readLengthDic, logDic) that these functions
mutate in place.  The functions below keep that contract so a maintainer can swap
the imports; the arithmetic happens on the device and the dicts are filled from
its exports.  A `Session` stands in for the state the reference keeps in files
between the calls (the trimmed FASTQ, SeqToAnnot.fasta/.sam).

Differences a caller can observe: `outfile` is not written (the reference deletes
it again in quantReads.py:46), `threads` / `bowtieBinary` / `numCPU` are ignored,
and index paths (`file_mirna`, ...) are resolved to the `<index>.fa` next to them.
"""
import os

from . import gff as gff_mod
from . import tables
from .engine import Engine, LIB_IDS, ROUND_LIB
from .libs import read_fasta
from .pipeline import iter_fastq_chunks, peek_max_read_len

_SESSION = None


class Session:
    def __init__(self, sample_count, spike_in=False, max_unique=1 << 22, max_read_len=64, device=0):
        self.engine = Engine(sample_count, spike_in, max_unique, max_read_len, device)
        self.spike_in = bool(spike_in)
        self.adapter = None
        self.pending = {}          # trimmed "file" name -> sample slot it was fed into
        self.next_slot = 0
        self.uniq = None


def start_session(sample_count, spike_in=False, max_unique=1 << 22, max_read_len=64, device=0):
    """Replaces the dict initialisation at __main__.py:318-330."""
    global _SESSION
    if _SESSION is not None:
        _SESSION.engine.close()
    _SESSION = Session(sample_count, spike_in, max_unique, max_read_len, device)
    return _SESSION


def _session():
    if _SESSION is None:
        raise RuntimeError('mirge_b200.refapi.start_session(sample_count, spike_in) must be called first')
    return _SESSION


def trim_file(infile, adapter, outfile, threads=1, phred=33):
    """-> (phred, processed, kept), utils/trim_file.py:89-134."""
    s = _session()
    if s.adapter != adapter:
        s.engine.set_adapter(adapter)
        s.adapter = adapter
    slot = s.next_slot
    s.next_slot += 1
    base = None
    for chunk in iter_fastq_chunks(infile):
        if base is None:
            base = Engine.detect_phred(chunk)
        s.engine.feed_fastq(slot, chunk, base)
        s.engine.sync()
    p, k, _ = s.engine.stats()
    s.pending[outfile] = slot
    return (base if base is not None else phred), int(p[slot]), int(k[slot])


def quantReads(inputfile, seqDic, readLengthDic, sampleCount, sampleIndex, sampleList, trimmed_collapsed_fa, spikeIn):
    """Adds sample `sampleIndex` to seqDic / readLengthDic, utils/quantReads.py:3-46."""
    s = _session()
    slot = s.pending.pop(inputfile)
    if slot != sampleIndex:
        raise RuntimeError('samples must be quantified in the order they were trimmed')
    u = s.engine.export_unique()
    n_annot = 11 if spikeIn else 10
    for k, seq in enumerate(u['seq']):
        c = int(u['counts'][k, sampleIndex])
        if c == 0:
            continue
        key = seq.decode()
        e = seqDic.get(key)
        if e is None:
            e = {'quant': [0] * sampleCount, 'annot': [0] + [''] * (n_annot - 1), 'length': len(key)}
            seqDic[key] = e
        e['quant'][sampleIndex] = c
    lh = s.engine.export_tables_prelim()['len_hist']
    for L in range(lh.shape[0]):
        if lh[L, sampleIndex]:
            readLengthDic.setdefault(L, [0] * sampleCount)[sampleIndex] = int(lh[L, sampleIndex])
    if trimmed_collapsed_fa:
        tables.write_trim_collapse_fa(os.path.dirname(inputfile), {'seq': u['seq'], 'counts': u['counts']},
                                      [sampleList[i] if i == sampleIndex else None for i in range(sampleCount)])


def runAnnotationPipeline(bowtieBinary, seqDic, numCPU, phred64, annotNameList, outputdir, logDic, file_mirna, file_hairpin,
                          file_mature_tRNA, file_pre_tRNA, file_snoRNA, file_rRNA, file_ncrna_others, file_mrna, spikeIn,
                          file_spikeIn, gff_output=False, miRNamePreNameDic=None, isomiRContentDic=None, miRNA_database=None,
                          trf_output=False, trnaStruDic=None, trfContentDic=None, sampleList=None):
    """Fills seqDic[seq]['annot'] and logDic['annotStats'], utils/runAnnotationPipeline.py:566-705."""
    s = _session()
    files = {'mirna': file_mirna, 'hairpin': file_hairpin, 'mature_trna': file_mature_tRNA, 'pre_trna': file_pre_tRNA,
             'snorna': file_snoRNA, 'rrna': file_rRNA, 'ncrna_others': file_ncrna_others, 'mrna': file_mrna}
    if spikeIn:
        files['spike-in'] = file_spikeIn
    for t, base in files.items():
        names, seqs = read_fasta(base + '.fa')
        s.engine.load_library(LIB_IDS[t], names, seqs)
    s.engine.annotate()
    u = s.engine.export_unique()
    s.uniq = u
    for k, seq in enumerate(u['seq']):
        ro = int(u['round_of'][k])
        if ro:
            a = seqDic[seq.decode()]['annot']
            a[0] = 1
            a[ro] = s.engine.lib_names[ROUND_LIB[ro - 1]][int(u['ref_of'][k])]
    if gff_output and isomiRContentDic is not None:
        # updateIsomiRDic / updateIsomiRDic2 (runAnnotationPipeline.py:382-446): same keys, values as strings
        mir_names, mir_seqs = read_fasta(file_mirna + '.fa')
        hp_names, hp_seqs = read_fasta(file_hairpin + '.fa')
        recs = gff_mod.gff_rows(u, mir_names, mir_seqs, hp_names, hp_seqs, miRNamePreNameDic or {}, miRNA_database or 'miRBase')
        for seq, r in recs.items():
            isomiRContentDic[seq] = {'miRName': r['miRName'], 'preMiRName': r['preMiRName'], 'type': r['type'], 'variant': r['variant'],
                                     'pre_start': str(r['pre_start']), 'pre_end': str(r['pre_end']), 'cigar': r['cigar'], 'strand': '+',
                                     'uid': r['uid'], 'start': str(r['start']), 'annot': 1, 'filter': 'Pass'}
    t = s.engine.export_tables_prelim()
    for r in range(len(annotNameList)):
        logDic['annotStats'].append({'readsProcessed': int(t['round_processed'][r]), 'readsAligned': int(t['round_aligned'][r]),
                                     'cpuTime': 0.0})


def summarize(seqDic, sampleList, logDic, mirDic, file_mirna, outputdir, spikeIn, bowtieBinary):
    """Fills mirDic and logDic['quantStats'][i], utils/summarize.py:3-66."""
    s = _session()
    s.engine.tally()
    t = s.engine.export_tables()
    names = s.engine.lib_names[0]
    mirDic.update(tables.build_mirdic(names, t['mir_quant'], t['mir_iscan']))
    for i in range(len(sampleList)):
        for c, key in enumerate(tables.TOTALS_KEYS):
            if key == 'spikeInReads' and not spikeIn:
                continue
            logDic['quantStats'][i][key] = int(t['lib_totals'][i, c])
