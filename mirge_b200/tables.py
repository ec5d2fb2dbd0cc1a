"""Host tail of the annotate pipeline: what remains after the GPU tallies.

O(#miRNA x samples) bookkeeping and the output tables, restated from
    utils/summarize.py:12-32       mirDic / logDic initialisation
    utils/miRNAmerge.py:13-41      merged-miRNA folding (with its creation quirk)
    utils/filter.py:3-32           iscan / canonical-ratio filter, zero-miRNA abort
    utils/generateReport.py:106-108,191-201   annotation.report.csv
    utils/writeDataToCSV.py:582-619,1172-1219 mapped / unmapped / miR.Counts / miR.RPM
    utils/writeDataToCSV.py:584-607,1090-1170 isomirs.csv / isomirs.samples.csv (-di)
    utils/quantReads.py:26-44      <sample>.trim.collapse.fa (-tcf)
Integer inputs come from the device (Engine.export_tables / export_unique).
"""
import math
import os

import numpy as np

from .engine import TOTALS_KEYS, ROUND_LIB


class ZeroMiRNAError(Exception):
    """filter.py:25-32: a sample without any filtered miRNA read aborts the run."""


def py2_float_str(x):
    """str(float) as Python 2 prints it (12 significant digits, '.0' kept)."""
    s = '%.12g' % x
    if not any(ch in s for ch in '.einf'):
        s += '.0'
    return s


def build_logdic(sample_names, processed, kept, lib_totals, spike_in):
    """logDic['quantStats'] rows (__main__.py:350-368, summarize.py:20-32)."""
    stats = []
    for i, name in enumerate(sample_names):
        d = {'filename': name, 'totalReads': int(processed[i]), 'trimmedReads': int(kept[i])}
        for c, key in enumerate(TOTALS_KEYS):
            if key == 'spikeInReads' and not spike_in:
                continue
            d[key] = int(lib_totals[i, c])
        stats.append(d)
    return {'quantStats': stats, 'annotStats': []}


def build_mirdic(mir_names, mir_quant, mir_iscan):
    """mirDic as summarize leaves it (summarize.py:12-15,41-44)."""
    return {n: {'quant': [int(x) for x in mir_quant[k]], 'iscan': [int(x) for x in mir_iscan[k]]}
            for k, n in enumerate(mir_names)}


def merge_mirnas(mirDic, merge_lines, n_samples):
    """miRNAmerge.py:13-41.  A merged row only comes into being when some member
    has a non-zero count in some sample; members present in mirDic are removed."""
    if merge_lines is None:
        return mirDic
    dele = set()
    for line in merge_lines:
        c = line.strip().split(',')
        merged = c[0]
        for member in c[1:]:
            if member not in mirDic:
                continue
            for j in range(n_samples):
                q = mirDic[member]['quant'][j]
                if q > 0:
                    if merged not in mirDic:
                        mirDic[merged] = {'quant': [0] * n_samples, 'iscan': [0] * n_samples}
                        mirDic[merged]['quant'][j] = q
                        mirDic[merged]['iscan'][j] = mirDic[member]['iscan'][j]
                    else:
                        mirDic[merged]['quant'][j] += q
                        mirDic[merged]['iscan'][j] += mirDic[member]['iscan'][j]
            dele.add(member)
    for m in dele:
        mirDic.pop(m, None)
    return mirDic


def filter_mirnas(mirDic, sample_names, logDic, cano_ratio):
    """filter.py:3-32."""
    thr = float(cano_ratio)
    n = len(sample_names)
    for e in mirDic.values():
        for i in range(n):
            q, c = e['quant'][i], e['iscan'][i]
            ratio = (float(c) / q) if q != 0 else 1.1
            if c < 2 or ratio < thr:
                e['quant'][i] = 0
    for i in range(n):
        vals = [e['quant'][i] for e in mirDic.values() if e['quant'][i] > 0]
        if vals:
            logDic['quantStats'][i]['mirnaReadsFiltered'] = int(sum(vals))
            logDic['quantStats'][i]['mirnaUniqFiltered'] = len(vals)
    for i in range(n):
        if logDic['quantStats'][i].get('mirnaReadsFiltered', 0) == 0:
            raise ZeroMiRNAError('No miRNA reads were found in sample %s. Please check your files and provided arguments.\n'
                                 % sample_names[i])


REPORT_HEADER = ('File name(s),Total Input Reads,Trimmed Reads(all),Trimmed Reads(unique),All miRNA Reads,Filtered miRNA Reads,'
                 'Unique miRNAs,Hairpin miRNAs,mature tRNA Reads,primary tRNA Reads,snoRNA Reads,rRNA Reads,ncRNA others,'
                 'mRNA Reads,Remaining Reads\n')


def annotation_report(sample_names, logDic, spike_in):
    """generateReport.py:106-108,191-201; with -spikeIn the row gains a value the header never names (reference quirk)."""
    keys = ['totalReads', 'trimmedReads', 'trimmedUniq', 'mirnaReads', 'mirnaReadsFiltered', 'mirnaUniqFiltered', 'hairpinReads',
            'maturetrnaReads', 'pretrnaReads', 'snornaReads', 'rrnaReads', 'ncrnaOthersReads', 'mrnaReads']
    keys += (['spikeInReads'] if spike_in else []) + ['remReads']
    rows = [REPORT_HEADER]
    for i, s in enumerate(sample_names):
        st = logDic['quantStats'][i]
        rows.append(s + ',' + ','.join(str(st[k]) for k in keys) + '\n')
    return ''.join(rows)


def mir_counts_csv(sample_names, logDic, mirDic):
    """writeDataToCSV.py:1190-1205 (byte-wise sorted miRNA names)."""
    n = len(sample_names)
    rows = ['miRNA' + ''.join(',' + s for s in sample_names) + '\n',
            'miRNAtotal' + ''.join(',' + str(logDic['quantStats'][i]['mirnaReadsFiltered']) for i in range(n)) + '\n']
    for k in sorted(mirDic):
        rows.append(k + ''.join(',' + str(v) for v in mirDic[k]['quant']) + '\n')
    return ''.join(rows)


def mir_rpm_csv(sample_names, logDic, mirDic):
    """writeDataToCSV.py:1207-1219."""
    n = len(sample_names)
    tot = [logDic['quantStats'][i]['mirnaReadsFiltered'] for i in range(n)]
    rows = ['miRNA' + ''.join(',' + s for s in sample_names) + '\n']
    for k in sorted(mirDic):
        q = mirDic[k]['quant']
        rows.append(k + ''.join((',' + py2_float_str(1000000.0 * q[i] / tot[i])) if tot[i] > 0 else ',0' for i in range(n)) + '\n')
    return ''.join(rows)


def annot_columns(uniq, lib_names, spike_in):
    """annot[1..9(10)] name columns of seqDic (runAnnotationPipeline.py:345) as numpy object arrays."""
    n_slots = 10 if spike_in else 9
    U = len(uniq['round_of'])
    cols = []
    for slot in range(1, n_slots + 1):
        col = np.full(U, '', dtype=object)
        sel = np.flatnonzero(uniq['round_of'] == slot)
        if len(sel):
            names = np.asarray(lib_names[ROUND_LIB[slot - 1]], dtype=object)
            col[sel] = names[uniq['ref_of'][sel]]
        cols.append(col)
    return cols


def write_mapped_unmapped(outdir, uniq, lib_names, annot_name_list, sample_names, spike_in):
    """mapped.csv / unmapped.csv (writeDataToCSV.py:582-619,1172-1188).  Row order is
    Python-2 dict order in the reference, i.e. unspecified; rows here follow unique id.
    Millions of rows: the columns are assembled as Arrow arrays straight from the exported buffers (no per-row Python
    objects) and written by Arrow's multi-threaded CSV writer; pandas is the fallback for names that need quoting."""
    try:
        _write_mapped_unmapped_arrow(outdir, uniq, lib_names, annot_name_list, sample_names, spike_in)
    except Exception:
        _write_mapped_unmapped_pandas(outdir, uniq, lib_names, annot_name_list, sample_names, spike_in)


def _write_mapped_unmapped_arrow(outdir, uniq, lib_names, annot_name_list, sample_names, spike_in):
    import pyarrow as pa
    import pyarrow.csv as pc
    U = len(uniq['round_of'])
    lens = uniq['len'].astype(np.int64)
    raw = uniq['seq'].view(np.uint8).reshape(U, -1) if U else np.zeros((0, 1), np.uint8)
    offsets = np.zeros(U + 1, np.int32)
    np.cumsum(lens, out=offsets[1:])
    data = raw[np.arange(raw.shape[1])[None, :] < lens[:, None]]
    cols = {'uniqueSequence': pa.Array.from_buffers(pa.string(), U, [None, pa.py_buffer(offsets), pa.py_buffer(np.ascontiguousarray(data))]),
            'annotFlag': pa.array((uniq['round_of'] > 0).astype(np.int64))}
    n_slots = 10 if spike_in else 9
    for slot, name in zip(range(1, n_slots + 1), annot_name_list):
        hit = uniq['round_of'] == slot
        idx = np.where(hit, uniq['ref_of'], 0).astype(np.int32)
        names = pa.array([str(x) for x in lib_names.get(ROUND_LIB[slot - 1], [])] or [''], type=pa.string())
        cols[name] = pa.DictionaryArray.from_arrays(pa.array(idx, mask=~hit), names).cast(pa.string())
    for i, s_ in enumerate(sample_names):
        cols[s_] = pa.array(np.ascontiguousarray(uniq['counts'][:, i]))
    tab = pa.table(cols)
    header = (','.join(['uniqueSequence', 'annotFlag'] + list(annot_name_list) + list(sample_names)) + '\n').encode()
    mapped = pa.array(uniq['round_of'] > 0)
    opts = pc.WriteOptions(include_header=False, quoting_style='none')
    for fn, mask in (('mapped.csv', mapped), ('unmapped.csv', pa.compute.invert(mapped))):
        with open(os.path.join(outdir, fn), 'wb') as f:
            f.write(header)
            pc.write_csv(tab.filter(mask), f, opts)


def _write_mapped_unmapped_pandas(outdir, uniq, lib_names, annot_name_list, sample_names, spike_in):
    import pandas as pd
    cols = annot_columns(uniq, lib_names, spike_in)
    flag = (uniq['round_of'] > 0).astype(np.int64)
    df = pd.DataFrame({'uniqueSequence': np.char.decode(uniq['seq'], 'ascii'), 'annotFlag': flag})
    for name, col in zip(annot_name_list, cols):
        df[name] = col
    for i, s in enumerate(sample_names):
        df[s] = uniq['counts'][:, i]
    df[flag > 0].to_csv(os.path.join(outdir, 'mapped.csv'), index=False)
    df[flag == 0].to_csv(os.path.join(outdir, 'unmapped.csv'), index=False)


def write_trim_collapse_fa(outdir, uniq, sample_names):
    """-tcf: <sample>.trim.collapse.fa, records >seq<rank>_<count> sorted by (count, seq) descending (quantReads.py:26-44)."""
    seqs = np.char.decode(uniq['seq'], 'ascii')
    for i, s in enumerate(sample_names):
        if s is None:
            continue
        c = uniq['counts'][:, i]
        sel = np.flatnonzero(c > 0)
        deco = sorted(zip(c[sel].tolist(), seqs[sel].tolist()), reverse=True)
        with open(os.path.join(outdir, os.path.splitext(s)[0] + '.trim.collapse.fa'), 'w') as f:
            f.write(''.join('>seq%d_%d\n%s\n' % (k + 1, cnt, sq) for k, (cnt, sq) in enumerate(deco)))


def calc_entropy(vals):
    """writeDataToCSV.py:344-351."""
    tot = sum(vals)
    ent = 0
    for v in vals:
        if v > 1:
            f = float(v) / tot
            ent += -1 * f * math.log(f, 2)
    return ent


def write_isomirs(outdir, uniq, lib_names, sample_names, logDic):
    """-di: isomirs.csv and isomirs.samples.csv (writeDataToCSV.py:584-607,1090-1170)."""
    n = len(sample_names)
    names = lib_names[0]
    seqs = np.char.decode(uniq['seq'], 'ascii')
    iso = {}
    for u in np.flatnonzero((uniq['round_of'] == 1) | (uniq['round_of'] == 9)):
        key = names[uniq['ref_of'][u]]
        kind = 'isomirs' if uniq['round_of'][u] == 9 else 'mirnas'
        if '.SNP' in key:
            key = key.split('.SNP')[0]
        iso.setdefault(key, {'mirnas': {}, 'isomirs': {}})[kind][seqs[u]] = [int(x) for x in uniq['counts'][u]]
    tot = [logDic['quantStats'][i]['mirnaReadsFiltered'] for i in range(n)]
    with open(os.path.join(outdir, 'isomirs.csv'), 'w') as f1, open(os.path.join(outdir, 'isomirs.samples.csv'), 'w') as f2:
        f1.write('miRNA,sequence' + ''.join(',' + s for s in sample_names) + ',Entropy\n')
        f2.write('miRNA' + ''.join(',%s isomir+miRNA Entropy,%s Canonical Sequence,%s Canonical RPM,%s Top Isomir RPM' % (s, s, s, s)
                                   for s in sample_names) + '\n')
        for mir, d in iso.items():
            sample_iso = [[] for _ in range(n)]
            sample_mir = [0] * n
            for arr in d['mirnas'].values():
                for i in range(n):
                    sample_mir[i] += arr[i]
            for sq, arr in d['isomirs'].items():
                for i in range(n):
                    sample_iso[i].append(arr[i])
                ent = calc_entropy(arr)
                mx = math.log(len(arr), 2)
                ent_s = 'NA' if mx == 0 else py2_float_str(ent / mx)
                f1.write(','.join([mir, sq] + [py2_float_str(arr[i] * 1000000.0 / tot[i]) for i in range(n)] + [ent_s]) + '\n')
            out = [mir]
            for i in range(n):
                rpm = 1000000.0 / tot[i]
                if len(sample_iso[i]) > 0:
                    top = max(sample_iso[i]) * rpm
                    iso_sum = sum(sample_iso[i]) * rpm
                    sample_iso[i].append(sample_mir[i])
                    ent = calc_entropy(sample_iso[i])
                    mir_rpm = sample_mir[i] * rpm
                    k = len(sample_iso[i])
                    out.append(py2_float_str(ent / math.log(k, 2)) if k > 1 else 'NA')
                    comb = mir_rpm + iso_sum
                    out.append(py2_float_str(100.0 * mir_rpm / comb) if comb > 0 else 'NA')
                    out.append(py2_float_str(mir_rpm))
                    out.append(py2_float_str(top))
                    f2.write(','.join(out) + '\n')
