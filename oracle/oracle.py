"""CPU ORACLE (Python side) for the miRge2.0 annotate-mode read pipeline.

TEST INFRASTRUCTURE ONLY -- only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  Nothing under
mirge_b200/ does.

PARITY UNPINNED: the reference ships no golden vectors and its engines
(cutadapt 1.16, bowtie 1.1.2) are not installable here; see mirge_oracle.c.

Two layers:
  * ctypes bindings to libmirge_oracle.so (trim / collapse / mapping rounds /
    integer tallies), and
  * literal Python-3 restatements of the reference's dict-based host stages
    (quantReads -> seqDic, summarize, miRNAmerge, filter and the CSV writers),
    each citing the reference file:line it follows, so whole-table parity can
    be checked end to end.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LIB_IDS = {'mirna': 0, 'hairpin': 1, 'mature_trna': 2, 'pre_trna': 3, 'snorna': 4,
           'rrna': 5, 'ncrna_others': 6, 'mrna': 7, 'spike-in': 9}


def build():
    """Compile libmirge_oracle.so (building the checker is not using it)."""
    subprocess.check_call(['make', '-s', '-C', _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libmirge_oracle.so')
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_new.restype = C.c_void_p
        L.orc_new.argtypes = [C.c_int, C.c_int]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_set_adapter.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_detect_phred.argtypes = [C.c_char_p, C.c_uint64]
        L.orc_feed_fastq.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_uint64, C.c_int, C.c_int]
        L.orc_trim_fastq.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_reset.argtypes = [C.c_void_p]
        L.orc_set_tie_policy.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_trim_params.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
        L.orc_n_unique.restype = C.c_uint64
        L.orc_n_unique.argtypes = [C.c_void_p]
        L.orc_seq_bytes.restype = C.c_uint64
        L.orc_seq_bytes.argtypes = [C.c_void_p]
        L.orc_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_load_lib.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int]
        L.orc_annotate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_export.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        L.orc_len_hist.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_tally.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_quality_trim_index.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_locate.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_void_p]
        L.orc_locate_unbanded.argtypes = L.orc_locate.argtypes
        L.orc_match_back.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_map_one.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_map_bruteforce.argtypes = L.orc_map_one.argtypes
        L.orc_map_genome_all.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------
# scalar entry points (known-answer tests)
# --------------------------------------------------------------------------
def quality_trim_index(qual, cutoff_front, cutoff_back, base=33):
    out = np.zeros(2, np.int32)
    q = qual.encode() if isinstance(qual, str) else qual
    lib().orc_quality_trim_index(q, len(q), cutoff_front, cutoff_back, base, _p(out[0:1]), _p(out[1:2]))
    return int(out[0]), int(out[1])


def locate(adapter, read, max_error_rate=0.12, min_overlap=3, banded=True):
    """Aligner.locate for a BACK adapter -> (astart, astop, rstart, rstop, matches, errors) or None."""
    out = np.zeros(6, np.int32)
    a, r = adapter.encode(), read.encode()
    f = lib().orc_locate if banded else lib().orc_locate_unbanded
    ok = f(a, len(a), r, len(r), max_error_rate, min_overlap, _p(out))
    return tuple(int(x) for x in out) if ok else None


def match_back(adapter, read, max_error_rate=0.12, min_overlap=3):
    """Adapter.match_to for a 3' adapter: (rstart, matches) or None (exact str.find first, else locate)."""
    out = np.zeros(2, np.int32)
    a, r = adapter.encode(), read.encode()
    ok = lib().orc_match_back(a, len(a), r, len(r), max_error_rate, min_overlap, _p(out[0:1]), _p(out[1:2]))
    return (int(out[0]), int(out[1])) if ok else None


def py_locate(adapter, read, max_error_rate=0.12, min_overlap=3):
    """Pure-Python twin of mirge_oracle.c:orc_locate (cutadapt _align.pyx locate, BACK)."""
    m, n = len(adapter), len(read)
    k = int(max_error_rate * m)
    col = [[i, 0, 0] for i in range(m + 1)]  # cost, origin, matches
    best = dict(ref_stop=m, query_stop=n, cost=m + n, origin=0, matches=0)
    last = min(m, k + 1)
    for j in range(1, n + 1):
        tmp = list(col[0])
        col[0][1] = j
        for i in range(1, last + 1):
            if adapter[i - 1] == read[j - 1]:
                cost, origin, matches = tmp[0], tmp[1], tmp[2] + 1
            else:
                cd, cdel, cins = tmp[0] + 1, col[i][0] + 1, col[i - 1][0] + 1
                if cd <= cdel and cd <= cins:
                    cost, origin, matches = cd, tmp[1], tmp[2]
                elif cins <= cdel:
                    cost, origin, matches = cins, col[i - 1][1], col[i - 1][2]
                else:
                    cost, origin, matches = cdel, col[i][1], col[i][2]
            tmp = list(col[i])
            col[i] = [cost, origin, matches]
        while last >= 0 and col[last][0] > k:
            last -= 1
        if last < m:
            last += 1
        else:
            length = m + min(col[m][1], 0)
            cost, matches = col[m][0], col[m][2]
            if length >= min_overlap and cost <= length * max_error_rate and (
                    matches > best['matches'] or (matches == best['matches'] and cost < best['cost'])):
                best.update(matches=matches, cost=cost, origin=col[m][1], ref_stop=m, query_stop=j)
                if cost == 0 and matches == m:
                    break
    for i in range(m, -1, -1):
        length = i + min(col[i][1], 0)
        cost, matches = col[i][0], col[i][2]
        if length >= min_overlap and cost <= length * max_error_rate and (
                matches > best['matches'] or (matches == best['matches'] and cost < best['cost'])):
            best.update(matches=matches, cost=cost, origin=col[i][1], ref_stop=i, query_stop=n)
    if best['cost'] == m + n:
        return None
    return (0, best['ref_stop'], best['origin'], best['query_stop'], best['matches'], best['cost'])


# --------------------------------------------------------------------------
# pipeline object
# --------------------------------------------------------------------------
class Oracle:
    def __init__(self, n_samples=1, spike_in=False, adapter='none', threads=0):
        self.S = n_samples
        self.spike_in = bool(spike_in)
        self.threads = threads or (os.cpu_count() or 1)
        self.h = lib().orc_new(n_samples, int(self.spike_in))
        self.lib_names = {}
        self.set_adapter(adapter)

    def close(self):
        if self.h:
            lib().orc_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(lib().orc_last_error(self.h).decode())

    def reset(self):
        lib().orc_reset(self.h)

    def set_trim_params(self, error_rate=0.12, min_overlap=3, min_length=16, qual_cutoff=10):
        lib().orc_set_trim_params(self.h, error_rate, min_overlap, min_length, qual_cutoff)

    def set_adapter(self, spec):
        self._check(lib().orc_set_adapter(self.h, spec.encode()))

    def set_tie_policy(self, flip):
        """Test knob: flip=True inverts every round's (ref, pos) preference among equally good hits."""
        lib().orc_set_tie_policy(self.h, int(bool(flip)))

    @staticmethod
    def detect_phred(buf):
        return lib().orc_detect_phred(buf, len(buf))

    def feed_fastq(self, sample, buf, phred_base=None):
        if phred_base is None:
            phred_base = self.detect_phred(buf)
        self._check(lib().orc_feed_fastq(self.h, sample, buf, len(buf), phred_base, self.threads))
        return phred_base

    def trim_fastq(self, buf, phred_base=33, max_rec=None):
        max_rec = max_rec or (buf.count(b'\n') // 4 + 1)
        st = np.zeros(max_rec, np.int32)
        sp = np.zeros(max_rec, np.int32)
        n = lib().orc_trim_fastq(self.h, buf, len(buf), phred_base, _p(st), _p(sp), max_rec)
        return st[:n], sp[:n]

    def stats(self):
        a = np.zeros(self.S, np.uint64)
        b = np.zeros(self.S, np.uint64)
        lib().orc_stats(self.h, _p(a), _p(b))
        return a, b

    def load_library(self, lib_id, names, seqs):
        off = np.zeros(len(seqs) + 1, np.uint64)
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
        cat = ''.join(seqs).encode()
        self._check(lib().orc_load_lib(self.h, lib_id, cat, _p(off), len(seqs)))
        self.lib_names[lib_id] = list(names)

    def annotate(self):
        rp = np.zeros(10, np.uint64)
        ra = np.zeros(10, np.uint64)
        self._check(lib().orc_annotate(self.h, self.threads, _p(rp), _p(ra)))
        return rp, ra

    def map_one(self, lib_id, read, mode, param, highest=False, brute=False):
        r = np.zeros(3, np.int32)
        f = lib().orc_map_bruteforce if brute else lib().orc_map_one
        ok = f(self.h, lib_id, read.encode(), len(read), mode, param, int(highest), _p(r[0:1]), _p(r[1:2]), _p(r[2:3]))
        return (int(r[0]), int(r[1]), int(r[2])) if ok else None

    def genome_hits(self, lib_id, read, mode, param, seedlen=28, trim3=0, norc=False):
        """Alignments of read (minus trim3 3' bases) on both strands of library lib_id, per mismatch count 0..3
        (bowtie -a)."""
        cnt = np.zeros(4, np.uint32)
        lib().orc_map_genome_all(self.h, lib_id, read.encode(), len(read), mode, param, seedlen, trim3, int(norc), _p(cnt))
        return [int(x) for x in cnt]

    def a2i_genome_filter(self, lib_id, reads):
        """writeDataToCSV.py:1260-1281: `bowtie genome -n 1 -a -3 2`; a read is retained when it aligns and its
        lowest occupied mismatch stratum holds exactly one alignment."""
        kept = {}
        for r in reads:
            c = self.genome_hits(lib_id, r, 0, 1, 28, 2)
            occupied = [x for x in c if x]
            if occupied and occupied[0] == 1:
                kept[r] = True
        return kept

    def export(self):
        U = int(lib().orc_n_unique(self.h))
        nb = int(lib().orc_seq_bytes(self.h))
        seqs = np.zeros(max(nb, 1), np.uint8)
        off = np.zeros(U, np.uint64)
        ln = np.zeros(U, np.uint32)
        cnt = np.zeros((U, self.S), np.uint32)
        ro = np.zeros(U, np.int32)
        rf = np.zeros(U, np.int32)
        ps = np.zeros(U, np.int32)
        mm = np.zeros(U, np.int32)
        lib().orc_export(self.h, _p(seqs), _p(off), _p(ln), _p(cnt), _p(ro), _p(rf), _p(ps), _p(mm))
        raw = seqs.tobytes()
        strs = [raw[int(o):int(o) + int(l)].decode() for o, l in zip(off, ln)]
        return dict(seqs=strs, counts=cnt, round_of=ro, ref_of=rf, pos_of=ps, mm_of=mm)

    def len_hist(self, max_len=256):
        out = np.zeros((max_len + 1, self.S), np.uint64)
        lib().orc_len_hist(self.h, _p(out), max_len)
        return out

    def tally(self):
        nmir = len(self.lib_names.get(0, []))
        q = np.zeros((nmir, self.S), np.int64)
        i = np.zeros((nmir, self.S), np.int64)
        t = np.zeros((self.S, 12), np.int64)
        lib().orc_tally(self.h, _p(q), _p(i), _p(t))
        return q, i, t

    # -- reference-shaped dicts ------------------------------------------------
    def seqDic(self):
        """seqDic as built by quantReads.py:9-16 and filled by updateAnnotDic
        (runAnnotationPipeline.py:341-352): seq -> {'quant','annot','length'}."""
        e = self.export()
        nslots = 11 if self.spike_in else 10
        round_lib = [0, 1, 2, 3, 4, 5, 6, 7, 0, 9]
        d = {}
        for k, s in enumerate(e['seqs']):
            annot = [0] + [''] * (nslots - 1)
            ro = int(e['round_of'][k])
            if ro:
                annot[0] = 1
                annot[ro] = self.lib_names[round_lib[ro - 1]][int(e['ref_of'][k])]
            d[s] = {'quant': [int(x) for x in e['counts'][k]], 'annot': annot, 'length': len(s)}
        return d


# --------------------------------------------------------------------------
# literal restatements of the reference's dict-based host stages
# --------------------------------------------------------------------------
def py2_str_float(x):
    """Python 2 str(float): repr with 12 significant digits, '.0' kept."""
    s = '%.12g' % x
    if '.' not in s and 'e' not in s and 'n' not in s and 'i' not in s:
        s += '.0'
    return s


def summarize(seqDic, sampleList, logDic, mirDic, mirList, spikeIn):
    """summarize.py:3-66; mirList stands in for `bowtie-inspect -n` (:6-9)."""
    for name in mirList:
        mirDic[name] = {'quant': [0] * len(sampleList), 'iscan': [0] * len(sampleList)}
    keys = ['trimmedUniq', 'mirnaReads', 'hairpinReads', 'maturetrnaReads', 'pretrnaReads', 'snornaReads',
            'rrnaReads', 'ncrnaOthersReads', 'mrnaReads', 'remReads']
    for i in range(len(sampleList)):
        for k in keys:
            logDic['quantStats'][i][k] = 0
        if spikeIn:
            logDic['quantStats'][i]['spikeInReads'] = 0
    for seqKey in seqDic:
        e = seqDic[seqKey]
        for i in range(len(sampleList)):
            q = e['quant'][i]
            if q != 0:
                st = logDic['quantStats'][i]
                st['trimmedUniq'] += 1
                a = e['annot']
                if a[1] != '' or a[9] != '':
                    st['mirnaReads'] += q
                    if a[1] != '':
                        mirDic[a[1]]['quant'][i] += q
                        mirDic[a[1]]['iscan'][i] += q
                    else:
                        mirDic[a[9]]['quant'][i] += q
                elif a[2] != '':
                    st['hairpinReads'] += q
                elif a[3] != '':
                    st['maturetrnaReads'] += q
                elif a[4] != '':
                    st['pretrnaReads'] += q
                elif a[5] != '':
                    st['snornaReads'] += q
                elif a[6] != '':
                    st['rrnaReads'] += q
                elif a[7] != '':
                    st['ncrnaOthersReads'] += q
                elif a[8] != '':
                    st['mrnaReads'] += q
                else:
                    if spikeIn and a[10] != '':
                        st['spikeInReads'] += q
                    else:
                        st['remReads'] += q


def miRNAmerge(merge_lines, sampleList, mirDic):
    """miRNAmerge.py:13-41 (merge_lines = rows of <sp>_merges_<DB>.csv)."""
    delList = []
    for line in merge_lines:
        c = line.strip().split(',')
        for i in range(1, len(c)):
            for j in range(len(sampleList)):
                try:
                    if mirDic[c[i]]['quant'][j] > 0:
                        try:
                            mirDic[c[0]]['quant'][j] = mirDic[c[0]]['quant'][j] + mirDic[c[i]]['quant'][j]
                            mirDic[c[0]]['iscan'][j] = mirDic[c[0]]['iscan'][j] + mirDic[c[i]]['iscan'][j]
                        except KeyError:
                            mirDic[c[0]] = {'quant': [0] * len(sampleList), 'iscan': [0] * len(sampleList)}
                            mirDic[c[0]]['quant'][j] = mirDic[c[i]]['quant'][j]
                            mirDic[c[0]]['iscan'][j] = mirDic[c[i]]['iscan'][j]
                    delList.append(c[i])
                except KeyError:
                    pass
    for item in set(delList):
        mirDic.pop(item, None)


class ZeroMiRNA(Exception):
    """filter.py:25-32 aborts (sys.exit(1)) when a sample has no miRNA reads."""


def filter_(mirDic, sampleList, logDic, canoRatioTmp):
    """filter.py:3-32."""
    thr = float(canoRatioTmp)
    for k in mirDic:
        for i in range(len(sampleList)):
            try:
                v = float(mirDic[k]['iscan'][i]) / mirDic[k]['quant'][i]
            except ZeroDivisionError:
                v = 1.1
            if mirDic[k]['iscan'][i] < 2 or v < thr:
                mirDic[k]['quant'][i] = 0
    for k in mirDic:
        for i in range(len(sampleList)):
            if mirDic[k]['quant'][i] > 0:
                st = logDic['quantStats'][i]
                st['mirnaReadsFiltered'] = st.get('mirnaReadsFiltered', 0) + mirDic[k]['quant'][i]
                st['mirnaUniqFiltered'] = st.get('mirnaUniqFiltered', 0) + 1
    for i in range(len(sampleList)):
        if logDic['quantStats'][i].get('mirnaReadsFiltered', 0) == 0:
            raise ZeroMiRNA(sampleList[i])


def annotation_report_csv(sampleList, logDic, spikeIn):
    """generateReport.py:106-108,191-201 (CSV only)."""
    out = ['File name(s),Total Input Reads,Trimmed Reads(all),Trimmed Reads(unique),All miRNA Reads,Filtered miRNA Reads,'
           'Unique miRNAs,Hairpin miRNAs,mature tRNA Reads,primary tRNA Reads,snoRNA Reads,rRNA Reads,ncRNA others,'
           'mRNA Reads,Remaining Reads\n']
    for i, s in enumerate(sampleList):
        st = logDic['quantStats'][i]
        keys = ['totalReads', 'trimmedReads', 'trimmedUniq', 'mirnaReads', 'mirnaReadsFiltered', 'mirnaUniqFiltered',
                'hairpinReads', 'maturetrnaReads', 'pretrnaReads', 'snornaReads', 'rrnaReads', 'ncrnaOthersReads', 'mrnaReads']
        if spikeIn:
            keys.append('spikeInReads')
        keys.append('remReads')
        out.append(s + ',' + ','.join(str(st[k]) for k in keys) + '\n')
    return ''.join(out)


def mir_counts_csv(sampleList, logDic, mirDic):
    """writeDataToCSV.py:1190-1205."""
    out = ['miRNA' + ''.join(',' + s for s in sampleList) + '\n']
    out.append('miRNAtotal' + ''.join(',' + str(logDic['quantStats'][i]['mirnaReadsFiltered']) for i in range(len(sampleList))) + '\n')
    for k in sorted(mirDic.keys()):
        out.append(k + ''.join(',' + str(mirDic[k]['quant'][i]) for i in range(len(sampleList))) + '\n')
    return ''.join(out)


def mir_rpm_csv(sampleList, logDic, mirDic):
    """writeDataToCSV.py:1207-1219."""
    out = ['miRNA' + ''.join(',' + s for s in sampleList) + '\n']
    for k in sorted(mirDic.keys()):
        row = k
        for i in range(len(sampleList)):
            t = logDic['quantStats'][i]['mirnaReadsFiltered']
            row += (',' + py2_str_float(1000000.0 * mirDic[k]['quant'][i] / t)) if t > 0 else ',0'
        out.append(row + '\n')
    return ''.join(out)


def mapped_unmapped_rows(seqDic, spikeIn):
    """writeDataToCSV.py:582-619,1172-1188 -> (mapped row set, unmapped row set)."""
    n = 11 if spikeIn else 10
    mapped, unmapped = set(), set()
    for s, e in seqDic.items():
        row = ','.join([s] + [str(e['annot'][i]) for i in range(n)] + [str(q) for q in e['quant']])
        (mapped if e['annot'][0] > 0 else unmapped).add(row)
    return mapped, unmapped


def run_tables(o, sampleList, mirna_names, merge_lines, canoRatio='0.1', drop=()):
    """__main__.py:375-389 after the GPU-replaceable part: summarize -> merge -> filter -> tables.
    drop: sequences left out of the tables (tests/test_tie_policy.py isolates the tie-dependent reads with it)."""
    seqDic = o.seqDic()
    for s in drop:
        del seqDic[s]
    proc, kept = o.stats()
    logDic = {'quantStats': [{'filename': s, 'totalReads': int(proc[i]), 'trimmedReads': int(kept[i])}
                             for i, s in enumerate(sampleList)], 'annotStats': []}
    mirDic = {}
    summarize(seqDic, sampleList, logDic, mirDic, mirna_names, o.spike_in)
    miRNAmerge(merge_lines, sampleList, mirDic)
    filter_(mirDic, sampleList, logDic, canoRatio)
    return dict(seqDic=seqDic, logDic=logDic, mirDic=mirDic,
                report=annotation_report_csv(sampleList, logDic, o.spike_in),
                counts=mir_counts_csv(sampleList, logDic, mirDic),
                rpm=mir_rpm_csv(sampleList, logDic, mirDic))


def isomir_tables(seqDic, sampleList, logDic):
    """writeDataToCSV.py:584-607 (isomirDic) and :1090-1170 (isomirs.csv, isomirs.samples.csv), literal.
    Returns (isomirs.csv text, isomirs.samples.csv text); row order is Python-2 dict order upstream."""
    import math

    def calcEntropy(inputList):                      # writeDataToCSV.py:344-351
        sum1 = sum(inputList)
        entropy = 0
        for i in range(len(inputList)):
            if inputList[i] > 1:
                freq = float(inputList[i]) / sum1
                entropy = entropy + -1 * freq * math.log(freq, 2)
        return entropy

    isomirDic = {}
    for seqKey in seqDic.keys():
        if seqDic[seqKey]['annot'][0] > 0:
            isomir = seqDic[seqKey]['annot'][9]
            mirna = seqDic[seqKey]['annot'][1]
            if isomir != '' or mirna != '':
                if isomir != '':
                    key, key2 = isomir, 'isomirs'
                else:
                    key, key2 = mirna, 'mirnas'
                if '.SNP' in key:
                    key = key.split('.SNP')[0]
                if key not in isomirDic:
                    isomirDic[key] = {'mirnas': {}, 'isomirs': {}}
                isomirDic[key][key2][seqKey] = [seqDic[seqKey]['quant'][i] for i in range(len(sampleList))]
    out1 = ['miRNA,sequence' + ''.join(',' + s for s in sampleList) + ',Entropy\n']
    out2 = ['miRNA' + ''.join(',' + s + ' isomir+miRNA Entropy' + ',' + s + ' Canonical Sequence' + ',' + s + ' Canonical RPM'
                              + ',' + s + ' Top Isomir RPM' for s in sampleList) + '\n']
    for miRNA in isomirDic.keys():
        sampleIsomirs = {i: [] for i in range(len(sampleList))}
        samplemiRNAs = [0] * len(sampleList)
        for miRNASeq in isomirDic[miRNA]['mirnas'].keys():
            sampleArray = isomirDic[miRNA]['mirnas'][miRNASeq]
            for i in range(len(sampleArray)):
                samplemiRNAs[i] = samplemiRNAs[i] + sampleArray[i]
        for miRNASeq in isomirDic[miRNA]['isomirs'].keys():
            entry = [miRNA, miRNASeq]
            sampleArray = isomirDic[miRNA]['isomirs'][miRNASeq]
            for i in range(len(sampleArray)):
                sampleIsomirs[i].append(sampleArray[i])
            entropy = calcEntropy(sampleArray)
            maxEntropy = math.log(len(sampleArray), 2)
            entropy = 'NA' if maxEntropy == 0 else py2_str_float(entropy / maxEntropy)
            for i in range(len(sampleArray)):
                sampleArray[i] = py2_str_float(sampleArray[i] * 1000000.0 / logDic['quantStats'][i]['mirnaReadsFiltered'])
            out1.append(','.join(entry + sampleArray + [str(entropy)]) + '\n')
        isomirOut = [miRNA]
        for sampleLane in range(len(sampleList)):
            rpmFactor = 1000000.0 / logDic['quantStats'][sampleLane]['mirnaReadsFiltered']
            if len(sampleIsomirs[sampleLane]) > 0:
                topIsomir = max(sampleIsomirs[sampleLane]) * rpmFactor
                isomirSum = sum(sampleIsomirs[sampleLane]) * rpmFactor
                sampleIsomirs[sampleLane].append(samplemiRNAs[sampleLane])
                sampleEntropyWithmiRNA = calcEntropy(sampleIsomirs[sampleLane])
                miRNARPM = samplemiRNAs[sampleLane] * rpmFactor
                maxEntropy = len(sampleIsomirs[sampleLane])
                if maxEntropy > 1:
                    sampleEntropyWithmiRNA = py2_str_float(sampleEntropyWithmiRNA / (math.log(maxEntropy, 2)))
                else:
                    sampleEntropyWithmiRNA = 'NA'
                isomirOut.append(sampleEntropyWithmiRNA)
                combined = miRNARPM + isomirSum
                isomirOut.append(py2_str_float(100.0 * miRNARPM / combined) if combined > 0 else 'NA')
                isomirOut.append(py2_str_float(miRNARPM))
                isomirOut.append(py2_str_float(topIsomir))
                out2.append(','.join(isomirOut) + '\n')
    return ''.join(out1), ''.join(out2)


# --------------------------------------------------------------------------
# -gff: <sample>_isomiRs.gff, literal restatement
# --------------------------------------------------------------------------
NT2CODE = dict(zip(                                           # runAnnotationPipeline.py:613-620
    [a + b + c for a in 'ACGT' for b in 'ACGT' for c in 'ACGT'],
    '@fcoladsmkhwgebpvtDnx#yiCEGSrjqHT84FVXZ6KM$AWY35LNJzU9P07IuBQOR%'))


def extract_pre_mir_name(gff3_text, miRNA_database):
    """extractPreMiRName.py:9-62: miRNA name -> stem-loop name from <species>_<DB>.gff3."""
    def pickOptimal(names):                                   # :9-17
        tails = [n.split('-')[-1] for n in names]
        if all(t.isdigit() for t in tails):
            cand = [[int(tails[i]), names[i]] for i in range(len(names))]
        else:
            cand = [[names[i]] for i in range(len(names))]
        cand.sort()
        return cand[0][-1]

    out = {}
    if miRNA_database == 'miRBase':
        cor, order, loopName = {}, [], {}
        for line in gff3_text.splitlines(True):
            if line[0] != '#':
                c = line.strip().split('\t')
                if c[2] == 'miRNA':
                    name = c[-1].split('Name=')[1].split(';')[0].strip()
                    loop = c[-1].split('Derives_from=')[1].split(';')[0].strip()
                    if name not in order:
                        order.append(name)
                        cor[name] = {loop: (c[0], c[6], int(c[3]) - 1, int(c[4]) - 1)}
                    else:
                        cor[name][loop] = (c[0], c[6], int(c[3]) - 1, int(c[4]) - 1)
                if c[2] == 'miRNA_primary_transcript':
                    alias = c[-1].split('Alias=')[1].split(';')[0].strip()
                    loopName[alias] = c[-1].split('Name=')[1].split(';')[0].strip()
        for name in order:
            if len(cor[name]) > 1:
                out[name] = pickOptimal([loopName[k] for k in cor[name].keys()])
            else:
                out[name] = loopName[list(cor[name].keys())[0]]
    else:
        for line in gff3_text.splitlines(True):
            if line[0] != '#':
                c = line.strip().split('\t')
                if c[2] == 'miRNA':
                    name = c[-1].split('ID=')[1].split(';')[0].strip()
                    out[name] = '_'.join(name.split('_')[:-1]) + '_pre'
    return out


def infer_pre_mir_name(canon, pre_of, miRNA_database):
    """runAnnotationPipeline.py:353-380."""
    if canon in pre_of:
        return pre_of[canon]
    if miRNA_database == 'miRBase':
        stem = '-'.join(canon.split('-')[:-1])
        for k in (stem, stem + '-5p', stem + '-3p'):
            if k in pre_of:
                return pre_of[k]
        return canon.replace('-5p', '').replace('-3p', '').replace('miR', 'mir')
    return canon.replace('_5p*', '').replace('_3p*', '').replace('_5p', '').replace('_3p', '') + '_pre'


def _ldash(s):
    return len(s) - len(s.lstrip('-'))


def _rdash(s):
    return len(s) - len(s.rstrip('-'))


def fill_terminal(pre, lib, mir, read, start, indexValue):
    """fillTerminal, runAnnotationPipeline.py:86-172: pad the hairpin, the library entry, the mature and the read
    with '-' so that they line up column by column."""
    if mir not in pre:                                        # .index() raising ValueError (:159-164)
        return '', '', '', '', False
    at = pre.index(mir)
    if at + len(mir) + 6 > len(pre):
        pre = pre + '-' * (at + len(mir) + 6 - len(pre))
    else:
        lib = lib + '-' * (len(pre) - at - len(mir) - 6)
    if at < 2:
        pre = '-' * (2 - at) + pre
    else:
        lib = '-' * (at - 2) + lib
    at2 = pre.index(mir)
    mir = '-' * at2 + mir + '-' * (len(pre) - at2 - len(mir))
    if indexValue == 0:                                       # exact miRNA hit: the read sits at 1-based `start` of the entry
        head, i = '', 0
        for ch in lib:
            if ch == '-':
                head += '-'
            else:
                i += 1
                if i == start:
                    break
                head += '-'
        read = head + read
        read = read + '-' * (len(lib) - len(read))
    else:                                                     # isomiR round: bowtie saw read[1:-2]
        room = len(lib) - _ldash(lib) - start + 1
        if len(read) - 1 > room:
            extra = len(read) - 1 - room
            lib, pre, mir = lib + '-' * extra, pre + '-' * extra, mir + '-' * extra
        else:
            read = read + '-' * (room - len(read) + 1)
        if start - 1 + _ldash(lib) < 1:
            pad = 1 - start + 1 - _ldash(lib)
            pre, mir, lib = '-' * pad + pre, '-' * pad + mir, '-' * pad + lib
        else:
            read = '-' * (start - 1 + _ldash(lib) - 1) + read
    return pre, lib, mir, read, True


def make_cigar(seq, mature):
    """make_cigar, runAnnotationPipeline.py:180-212 (runs of M are counted, everything else stays one character)."""
    raw = ''
    for a, b in zip(seq, mature):
        if a == b:
            raw += 'M'
        elif a != '-' and b != '-':
            raw += a
        elif a == '-':
            raw += 'D'
        else:
            raw += 'I'
    short, cache, counter = '', '', 1
    for c in raw:
        if c != cache and cache != '' and cache == 'M':
            short += 'M' if counter == 1 else str(counter) + 'M'
            counter = 1
            cache = c
        if c != 'M':
            short += c
        if c == cache and c == 'M':
            counter += 1
        cache = c
    if cache == 'M':
        short += 'M' if counter == 1 else str(counter) + 'M'
    return short


def make_id(seq):
    """make_id, runAnnotationPipeline.py:214-235."""
    start, uid, i = 0, '', 0
    for i in range(0, len(seq) + 1, 3):
        if i == 0:
            continue
        tri = seq[start:i]
        if tri not in NT2CODE:
            uid = '.'
            break
        uid += NT2CODE[tri]
        start = i
    if len(seq) > i and uid != '.':
        dummy = 'A' * (3 - (len(seq) - i))
        tri = seq[i:len(seq)] + dummy
        if tri in NT2CODE:
            uid += NT2CODE[tri] + str(len(dummy))
        else:
            uid = '.'
    return uid


def analyze_alignment(pre, lib, mir, read):
    """analyzeAlignment, runAnnotationPipeline.py:237-339 -> (type, variant, pre_start, pre_end, cigar)."""
    assert len({len(pre), len(lib), len(mir), len(read)}) == 1
    variants = []
    if mir == read:
        kind = 'ref_miRNA'
        variants.append('NA')
    else:
        kind = 'isomiR'
        s1, e1 = _ldash(mir), len(mir) - _rdash(mir)
        snp, snp_pos = False, ''
        for i in range(s1, e1):
            if read[i] != '-' and read[i] != mir[i]:
                snp = True
                if 1 <= i <= 6:
                    snp_pos = '_seed'
                elif i == 7:
                    snp_pos = '_central_offset'
                elif 8 <= i <= 11:
                    snp_pos = '_central'
                elif 12 <= i <= 16:
                    snp_pos = 'central_supp'
                else:
                    snp_pos = ''
                break
        p5 = add = p3 = False
        p5_pos = add_pos = p3_pos = ''
        if _ldash(read) > _ldash(mir):
            p5, p5_pos = True, str(_ldash(mir) - _ldash(read))
        elif _ldash(read) < _ldash(mir):
            a = _ldash(read)
            if read[a:s1] != pre[a:s1]:
                snp, snp_pos = True, ''
            p5, p5_pos = True, '+' + str(_ldash(mir) - _ldash(read))
        if _rdash(read) > _rdash(mir):
            p3, p3_pos = True, str(_rdash(mir) - _rdash(read))
        elif _rdash(read) < _rdash(mir):
            z = len(read) - _rdash(read) - 1
            if read[e1:z + 1] != pre[e1:z + 1]:
                add, add_pos = True, '+' + str(_rdash(mir) - _rdash(read))
            else:
                p3, p3_pos = True, '+' + str(_rdash(mir) - _rdash(read))
        if snp:
            variants.append('iso_snp' + snp_pos)
        if add:
            variants.append('iso_add:' + add_pos)
        if p5:
            variants.append('iso_5p:' + p5_pos)
        if p3:
            variants.append('iso_3p:' + p3_pos)
    if _ldash(pre) > 0:
        if _ldash(read) >= _ldash(pre):
            pre_start = _ldash(read) - _ldash(pre) + 1
        else:
            pre_start = 1 - (_ldash(pre) - _ldash(read))
    else:
        pre_start = _ldash(read) + 1
    pre_end = pre_start + len(read) - _ldash(read) - _rdash(read) - 1
    a, z = _ldash(read), len(read) - _rdash(read) - 1
    return kind, ','.join(variants), pre_start, pre_end, make_cigar(read[a:z + 1], pre[a:z + 1])


def gff_literal(seqDic, hits, sampleList, mirna_lib, hairpin_lib, pre_of, miRNA_database):
    """What runAnnotationPipeline.py:643-655 (updateIsomiRDic / updateIsomiRDic2 after rounds 0 and 8) and
    writeDataToCSV.py:621-646 leave in <sample>_isomiRs.gff.  hits: seq -> (round index 0 or 8, miRNA entry name,
    1-based POS as bowtie reports it); mirna_lib / hairpin_lib: name -> sequence (`bowtie-inspect` output, :603-611).
    Returns {sample: (header text, [row, ...])}; row order is Python-2 dict order upstream."""
    content = {}
    for rnd in (0, 8):
        for seqKey, (r, name, pos) in hits.items():
            if r != rnd:
                continue
            canon = name.split('.')[0] if '.' in name else name
            content[seqKey] = {'miRName': name, 'preMiRName': infer_pre_mir_name(canon, pre_of, miRNA_database),
                               'start': str(pos), 'cigar': (str(len(seqKey)) + 'M'), 'annot': 0, 'filter': 'Pass',
                               'uid': make_id(seqKey)}
        for seqKey in list(content.keys()):
            e = content[seqKey]
            if e['annot'] != 0:
                continue
            lib = mirna_lib[e['miRName']]
            mir = lib[2:-6]
            pre = hairpin_lib[e['preMiRName']]
            if '.SNP' in e['miRName'] and '.SNPC' not in e['miRName']:
                canon_lib = mirna_lib[e['miRName'].split('.')[0] + '.SNPC']
                canon_mir = canon_lib[2:-6]
                canon_pre = hairpin_lib[infer_pre_mir_name(e['miRName'].split('.')[0], pre_of, miRNA_database)]
                at = canon_pre.index(canon_mir)               # ValueError = the reference's 'Error happens.' exit
                pre = canon_pre[:at] + mir + canon_pre[at + len(mir):]
            preO, libO, mirO, readO, ok = fill_terminal(pre, lib, mir, seqKey, int(e['start']), rnd)
            if ok:
                kind, variant, ps, pe, cigar = analyze_alignment(preO, libO, mirO, readO)
                e.update(type=kind, pre_start=str(ps), pre_end=str(pe), variant=str(variant), strand='+', annot=1, cigar=cigar)
            else:
                del content[seqKey]
    source = 'miRBase22' if miRNA_database == 'miRBase' else miRNA_database + '2.0'
    out = {}
    for i, sample in enumerate(sampleList):
        sampleName = os.path.splitext(sample)[0]
        header = '# GFF3 adapted for miRNA sequencing data\n## VERSION 0.0.1\n## source-ontology: %s\n## COLDATA: %s\n' % (source, sampleName)
        rows = []
        for seqKey, e in content.items():
            n = seqDic[seqKey]['quant'][i]
            if n >= 1:
                rows.append('\t'.join([e['miRName'], source, e['type'], e['pre_start'], e['pre_end'], '.', e['strand'], '.']) + '\t' +
                            ';'.join(['Read ' + seqKey, ' UID ' + e['uid'], ' Name ' + e['miRName'], ' Parent ' + e['preMiRName'],
                                      ' Variant ' + e['variant'], ' Cigar ' + e['cigar'], ' Expression ' + str(n), ' Filter ' + e['filter']]) + '\n')
        out[sample] = (header, rows)
    return out
