"""A-to-I editing analysis of `miRge2.0 annotate -ai` (src/mirge/utils/writeDataToCSV.py:1222-1615).

Host-side restatement, function for function, of the reference block; its two `bowtie <genome> ... -a -3 2` runs
(:1260, :1488) are served by the GPU genome pass (csrc/k2_genome.cuh) through Engine.genome_pass.  Outputs:
a2IEditing.detail.txt, a2IEditing.report.csv, a2IEditing.report.newform.csv (the heat-map PDF drawn by an R script
at :1611-1615 is not produced).

Two things the reference leaves to its environment are fixed here and flagged UNVERIFIED-OFFLINE:
  * `pairwise2.align.localms(target, seq, 2, -1, -20, -20)[0]` (:109): with gaps at -20 the reads of one miRNA align to
    its canonical sequence without gaps, so the alignment is the best-scoring DIAGONAL (match 2, mismatch -1, local);
    among equally good end cells the last one in row-major order is taken (Biopython's traceback pops the last start
    first).  Both sequences come back whole, padded with '-' to a common frame, as pairwise2 returns them.
  * the order of Python 2 dict keys (:1225, :1294): rows are produced in first-seen order of the unique-sequence
    table; the final report is sorted (:1451, `sort -t, -k1,1 -k2,2n`, byte order) so only a2IEditing.detail.txt and
    ties of that sort depend on it.
"""
import math
import os

A2I_PERCENTAGE_CUTOFF = 2.0      # writeDataToCSV.py:567


# ---------------------------------------------------------------------------------------------------------------------
# small helpers, writeDataToCSV.py:14-99
# ---------------------------------------------------------------------------------------------------------------------
def start_and_end(seq1, seq2):                      # :14-18
    start = seq2.index(seq1)
    return start, start + len(seq1) - 1


def dash_count(seq):                                # :20-33
    head = len(seq) - len(seq.lstrip('-'))
    tail = len(seq) - len(seq.rstrip('-'))
    if head == len(seq):                            # all dashes: the reference's two loops both run to the end
        tail = len(seq)
    return head, tail


def remove_dash(seq):                               # :68-70
    head, tail = dash_count(seq)
    return seq[head:len(seq) - tail]


def judge_align(target_new, seq_tmp):               # :35-66
    mismatch_limit, head_shift, tail_shift = 1, 1, 3
    state = True
    head_t, tail_t = dash_count(target_new)
    head_s, tail_s = dash_count(seq_tmp)
    len1 = len(target_new) - head_t - tail_t
    match_limit = len1 - tail_shift - mismatch_limit
    start1 = head_t
    end1 = len(target_new) - head_t - 1 - tail_shift            # as written upstream (head, not tail, dashes)
    start2 = head_s
    end2 = len(seq_tmp) - tail_s - 1
    if start2 - start1 > head_shift:
        return False
    mismatch = match = 0
    for pos in range(head_t, min(end1, end2) + 1):
        if seq_tmp[pos] == '-':
            continue
        if target_new[pos] != seq_tmp[pos]:
            mismatch += 1
        else:
            match += 1
    limit = match_limit - head_shift if start2 - start1 == head_shift else match_limit
    if mismatch > mismatch_limit:
        state = False
    if match < limit:
        state = False
    return state


def refine_name(name):                              # :72-83
    if '.fastq' in name:
        k = name.index('.fastq')
        return name[:k] + name[k + len('.fastq'):]
    return name


def local_align_gapless(target, seq, match=2, mismatch=-1):
    """localms(target, seq, 2, -1, -20, -20)[0][:2] restated (module docstring): best local diagonal."""
    best, cell = 0, None
    prev = [0] * (len(seq) + 1)
    for i in range(1, len(target) + 1):
        cur = [0] * (len(seq) + 1)
        ti = target[i - 1]
        for j in range(1, len(seq) + 1):
            v = prev[j - 1] + (match if ti == seq[j - 1] else mismatch)
            if v < 0:
                v = 0
            cur[j] = v
            if v > 0 and v >= best:
                best, cell = v, (i, j)
        prev = cur
    if cell is None:
        return None
    d = cell[0] - cell[1]                           # target offset minus read offset along the diagonal
    a, b = target, seq
    if d > 0:
        b = '-' * d + b
    elif d < 0:
        a = '-' * (-d) + a
    n = max(len(a), len(b))
    return a + '-' * (n - len(a)), b + '-' * (n - len(b))


def align_to_target(target, seq_list):              # align2TargetSeq, :101-143
    aligned, states = [], []
    for seq in seq_list:
        r = local_align_gapless(target, seq)
        if r is None:                               # nothing in common: pairwise2 returns no alignment (IndexError upstream)
            raise ValueError('no local alignment between %s and %s' % (target, seq))
        t_new, s_new = r
        states.append(judge_align(t_new, s_new))
        if not aligned:
            aligned += [t_new, s_new]
        elif t_new == aligned[0]:
            aligned.append(s_new)
        else:
            head1 = t_new.index(target)
            tail1 = len(t_new) - head1 - len(target)
            head2 = aligned[0].index(target)
            tail2 = len(aligned[0]) - head2 - len(target)
            if head1 >= head2 and tail1 >= tail2:
                aligned = [(head1 - head2) * '-' + x + (tail1 - tail2) * '-' for x in aligned]
                aligned.append(s_new)
            elif head1 >= head2 and tail1 < tail2:
                aligned = [(head1 - head2) * '-' + x for x in aligned]
                aligned.append(s_new + (tail2 - tail1) * '-')
            elif head1 < head2 and tail1 >= tail2:
                aligned = [x + (tail1 - tail2) * '-' for x in aligned]
                aligned.append((head2 - head1) * '-' + s_new)
            else:
                aligned.append((head2 - head1) * '-' + s_new + (tail2 - tail1) * '-')
    return aligned, states


def binom_cdf(k, n, p):
    """scipy.stats.binom.cdf(k, n, p) (:221)."""
    from scipy import stats
    return float(stats.binom.cdf(k, n, p))


def a2i_editing(target, seq_list, count_list, mir_name, detail, retained, start_base='A', end_base='G'):
    """A2IEditing, :145-229."""
    p_mismatch, tail_shift = 0.001, 5
    aligned, states = align_to_target(target, seq_list)
    start_label, end_label = start_and_end(target, aligned[0])
    target_aligned = aligned[0]
    head_count = dash_count(target_aligned)[0]
    canonical = count_true = seq_true = 0
    pos_count, positions, kept_new = {}, [], []
    for j, seq in enumerate(aligned[1:]):
        if states[j] and remove_dash(seq) in retained:
            if seq_list[j] in target:
                canonical += count_list[j]
            kept_new.append(seq)
            seq_true += 1
            count_true += count_list[j]
            for i in range(start_label, end_label + 1 - tail_shift):
                if i < len(target_aligned) and i < len(seq) and target_aligned[i] == start_base and seq[i] == end_base:
                    pos = i + 1 - head_count
                    if pos not in positions:
                        positions.append(pos)
                        pos_count[pos] = count_list[j]
                    else:
                        pos_count[pos] += count_list[j]
    detail.write('Canonical_Seq of %s: %s\n' % (mir_name, target))
    detail.write('seqList size is: %d, %d\n' % (len(seq_list), len(count_list)))
    for k, item in enumerate(aligned[1:]):
        detail.write('\t'.join([item, str(count_list[k]), str(states[k])]) + '\n')
    detail.write('****************\n')
    detail.write('retained seqList size is: %d\n' % seq_true)
    for k, item in enumerate(aligned[1:]):
        if states[k]:
            detail.write('\t'.join([item, str(count_list[k]), str(states[k])]) + '\n')
    detail.write('****************\n')
    detail.write('retained sequences after filering are:\n')
    for k, item in enumerate(aligned[1:]):
        if states[k] and remove_dash(item) in retained:
            detail.write('\t'.join([item, str(count_list[k]), str(states[k])]) + '\n')
    ratio, pvalue = {}, {}
    for pos in positions:
        ratio[pos] = float(pos_count[pos]) / count_true if count_true else 0
        if count_true - pos_count[pos] >= 0:
            pvalue[pos] = binom_cdf(count_true - pos_count[pos], count_true, 1 - p_mismatch)
        else:
            pvalue[pos] = 1.0
    return kept_new, positions, pos_count, ratio, pvalue, count_true, seq_true, canonical


# ---------------------------------------------------------------------------------------------------------------------
# genome filters (the two bowtie runs)
# ---------------------------------------------------------------------------------------------------------------------
def unique_best_hits(seqs, counts):
    """:1266-1281: a read is retained when it aligns and its lowest occupied mismatch stratum holds ONE alignment."""
    kept = {}
    for s, c in zip(seqs, counts):
        occupied = [int(x) for x in c if int(x)]
        if occupied and occupied[0] == 1:
            kept[s] = True
    return kept


# ---------------------------------------------------------------------------------------------------------------------
# the -ai block
# ---------------------------------------------------------------------------------------------------------------------
def write_a2i(outdir, seq_items, sample_list, logDic, mirDic, mir_name_seq, mir_merged_name, removed_mirnas, genome_pass, log=print):
    """seq_items: [(sequence, exact-miRNA name or '', isomiR-miRNA name or '', [count per sample])] in table order
    (seqDic with annot[1] / annot[9], :1225-1249).  genome_pass(seqs, n) -> per-sequence stratum counts of
    `bowtie genome -n <n> -f -a -3 2`.  Returns the path of a2IEditing.report.csv."""
    S = len(sample_list)
    mrf = [logDic['quantStats'][i]['mirnaReadsFiltered'] for i in range(S)]
    rpm = lambda c, i: 1000000.0 * c / mrf[i]                                   # noqa: E731
    seq_info = {}
    by_mir = {}
    for seq, exact, iso, quant in seq_items:                                   # :1225-1249
        if exact == '' and iso == '':
            continue
        seq_info[seq] = (exact, quant)
        name_tmp = exact if exact != '' else iso
        name = mir_merged_name.get(name_tmp, name_tmp)
        if exact != '' or any(rpm(quant[i], i) >= 1 for i in range(S)):
            by_mir.setdefault(name, []).append(seq)
    to_map = [s for name in by_mir for s in by_mir[name]]
    retained = unique_best_hits(to_map, genome_pass(to_map, 1)) if to_map else {}     # :1251-1287
    values, order = {}, []                                                     # mirNamePositionValueDict / List
    detail_path = os.path.join(outdir, 'a2IEditing.detail.txt')
    with open(detail_path, 'w') as detail:
        for name, seqs in by_mir.items():                                      # :1294-1322
            target = mir_name_seq[name]
            for i in range(S):
                if rpm(mirDic[name]['quant'][i], i) >= 1:
                    cl, sl = [], []
                    for s in seqs:
                        exact, quant = seq_info[s]
                        if rpm(quant[i], i) >= 1 or (exact != '' and quant[i] > 0):
                            sl.append(s)
                            cl.append(quant[i])
                    if len(cl) > 1 and min(cl) > 0:
                        kept, positions, pc, ratio, pv, count_true, seq_true, canonical = \
                            a2i_editing(target, sl, cl, name, detail, retained)
                        for pos in positions:
                            key = ':'.join([name, str(pos)])
                            if key not in values:
                                order.append(key)
                                values[key] = [[] for _ in range(S)]
                            values[key][i] = values[key][i] + [kept, str(sum(cl)), str(len([x for x in cl if x > 0])), str(count_true),
                                                                 str(seq_true), str(canonical), str(pc[pos]), ratio[pos], pv[pos]]
    # (the per-base-pair mismatch tally of :1324-1351 only feeds a file the reference deletes again at :1537)
    for i in range(S):                                                         # Benjamini-Hochberg, :1353-1364
        plist = sorted([values[k][i][8], k] for k in order if len(values[k][i]) != 0)
        for rank, (p, k) in enumerate(plist):
            values[k][i].append(p * len(plist) / (rank + 1))
    exact_of = {s: info[0] for s, info in seq_info.items()}

    def check_seq_list(kept):                                                  # checkSeqList, :85-99
        for s in kept:
            if exact_of.get(remove_dash(s), '') != '':
                return True
        return False

    header = 'miRNA,A-to-I position in the miRNA,miRNA sequence'
    for i in range(S):
        n = refine_name(sample_list[i])
        header += ',' + ','.join([n + '.readCount', n + '.readCount.canonical', n + '.RPM.canonical', n + '.readCount.mismatch',
                                  n + '.RPM.mismatch', n + '.AtoI.percentage', n + '.AtoI.adjusted.pValue'])
    rows = []
    for key in order:                                                          # :1373-1428
        name, pos = key.split(':')[0].strip(), key.split(':')[1].strip()
        row = ','.join([name, pos, mir_name_seq[name]])
        for i in range(S):
            c = values[key][i]
            row += ','
            if len(c) != 0:
                row += '%s,%s,%.2f,%s,%.2f' % (c[3], c[5], rpm(int(c[5]), i), c[6], rpm(int(c[6]), i))
                if check_seq_list(c[0]):
                    row += ',%.2f%%' % (c[7] * 100)
                    row += (',%.2E' % c[9]) if c[9] <= 0.05 else ',NS'
                else:
                    row += ',NE,NE'
            else:
                row += ','.join(['NE'] * 7)
        rows.append(row)

    def is_float(x):
        try:
            float(x)
            return True
        except ValueError:
            return False

    kept_rows = []
    for row in rows:                                                           # :1429-1449: some sample has a numeric adjusted p
        content = row.strip().split(',')
        portion = ((len(content) - 9) // 7) + 1
        if any(9 + i * 7 < len(content) and is_float(content[9 + i * 7]) for i in range(portion)):
            kept_rows.append(row)
    kept_rows.sort(key=lambda r: (r.split(',')[0].encode(), int(r.split(',')[1])))     # :1451
    judge = {}
    for row in kept_rows:                                                      # :1456-1484
        content = row.strip().split(',')
        mism = ''.join('G' if idx == int(content[1]) - 1 else ch for idx, ch in enumerate(content[2]))
        crpm = [float(content[5 + i * 7]) if is_float(content[5 + i * 7]) else 0 for i in range(((len(content) - 9) // 7) + 1)]
        judge[':'.join(content[:2])] = {'mismathedSeq': mism, 'RPMRemoveState': not any(x >= 1 for x in crpm),
                                        'ReElementRemoveState': content[0] in removed_mirnas, 'SeqRemoveState': False}
    mism_seqs = [v['mismathedSeq'] for v in judge.values()]
    if mism_seqs:                                                              # :1485-1502: any alignment at all removes the site
        hits = genome_pass(mism_seqs, 0)
        aligned = {s for s, c in zip(mism_seqs, hits) if sum(int(x) for x in c) > 0}
        for v in judge.values():
            if v['mismathedSeq'] in aligned:
                v['SeqRemoveState'] = True
    report = os.path.join(outdir, 'a2IEditing.report.csv')
    final_rows = []
    with open(report, 'w') as f:                                               # :1504-1515
        f.write(header + '\n')
        for row in kept_rows:
            j = judge[':'.join(row.split(',')[:2])]
            if not (j['RPMRemoveState'] or j['ReElementRemoveState'] or j['SeqRemoveState']):
                f.write(row + '\n')
                final_rows.append(row)
    write_newform(os.path.join(outdir, 'a2IEditing.report.newform.csv'), header, final_rows, log)
    return report


SAMPLE_ALIASES = {'SRR837842': 'Colon 1', 'SRR837839': 'Colon 2', 'SRR5127219': 'Colon cell', 'SRR1646473': 'Colon cancer 1',
                  'SRR1646493': 'Colon cancer 2', 'SRR1917324': 'DKO1', 'SRR1917336': 'DLD1', 'SRR1917329': 'DKS8',
                  'SRR567638': 'Placenta 2'}                                  # :1541-1550


def write_newform(path, header, rows, log=print):
    """a2IEditing.report.newform.csv, :1552-1606 (input of the heat-map script)."""
    samples = []
    for item in header.strip().split(','):
        if '.AtoI.percentage' in item and item.split('.')[0] not in samples:
            samples.append(item.split('.')[0])
    names, contents = [], {}
    for row in rows:
        content = row.strip().split(',')
        key = ':'.join(content[:2])
        tmp = content[3:]
        lst = []
        for item in [tmp[i:i + 7] for i in range(0, len(tmp), 7)]:
            if item[6] in ('NE', 'NS'):
                lst.append(('NA', 'NA'))
            else:
                lst.append((item[5][:-1], str(math.log(float(item[4]), 2))))
        state = False
        for item in lst:
            try:
                if float(item[0]) >= A2I_PERCENTAGE_CUTOFF:
                    state = True
                    break
            except ValueError:
                pass
        if state:
            names.append(key)
            contents[key] = lst
    ranked = sorted(((sum(1 for x in contents[k] if x[0] != 'NA'), k) for k in names), reverse=True)
    with open(path, 'w') as f:
        f.write('miRNA:position,sample,A-to-I percentage,log2RPM\n')
        for idx, sample in enumerate(samples):
            for _, k in ranked:
                f.write(k + ',' + SAMPLE_ALIASES.get(sample, sample) + ',' + ','.join(contents[k][idx]) + '\n')
    if len(ranked) <= 5:
        log('The number of A-to-I editing sites for is less than 10 so that no heatmap is drawn.')
