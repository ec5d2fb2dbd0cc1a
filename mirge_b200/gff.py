"""-gff: <sample>_isomiRs.gff, the mirGFF3 table of miRNA / isomiR reads.

Replaces, for the sequences the exact-miRNA round (0) and the isomiR round (8) annotated,
    utils/extractPreMiRName.py:9-62          miRNA name -> stem-loop name from <sp>_<DB>.gff3
    utils/runAnnotationPipeline.py:353-380   inferPremiRName
    utils/runAnnotationPipeline.py:382-446   updateIsomiRDic / updateIsomiRDic2
    utils/runAnnotationPipeline.py:86-172    fillTerminal
    utils/runAnnotationPipeline.py:180-339   make_cigar, make_id, analyzeAlignment
    utils/writeDataToCSV.py:621-646          the writer
The reference lines the hairpin, the library entry, the mature and the read up by padding four
strings with '-' and then measures runs of dashes.  Here everything is interval arithmetic on ONE
axis -- hairpin base k sits at x = k -- which yields the same fields:
    hairpin  [0, P)            library entry [at - 2, at + M + 6)     (at = first occurrence of the mature)
    mature   [at, at + M)      read          [r0, r0 + n),  r0 = at - 2 + POS - 1   (round 0)
                                                            r0 = at - 2 + POS - 2   (round 8: bowtie saw read[1:-2])
Inputs come from mirge_export_unique (round_of, ref_of, pos_of) -- no SAM files are parsed.
"""
import os

import numpy as np

_CODES = '@fcoladsmkhwgebpvtDnx#yiCEGSrjqHT84FVXZ6KM$AWY35LNJzU9P07IuBQOR%'      # runAnnotationPipeline.py:613-620
TRIPLET_CODE = {a + b + c: _CODES[16 * i + 4 * j + k] for i, a in enumerate('ACGT') for j, b in enumerate('ACGT')
                for k, c in enumerate('ACGT')}


class GffError(Exception):
    """Situations in which the reference stops (KeyError / 'Error happens.' + sys.exit(1))."""


def read_uid(seq):
    """UID of a read: one character per base triplet; a trailing partial triplet is completed with A's and
    followed by their number; any non-ACGT triplet gives '.' (make_id)."""
    full = len(seq) // 3 * 3
    out = []
    for i in range(0, full, 3):
        c = TRIPLET_CODE.get(seq[i:i + 3])
        if c is None:
            return '.'
        out.append(c)
    # the reference's loop variable stops at the last multiple of 3 <= len(seq); 0 for reads shorter than 3
    rest = len(seq) - full
    if rest:
        c = TRIPLET_CODE.get(seq[full:] + 'A' * (3 - rest))
        if c is None:
            return '.'
        out.append(c + str(3 - rest))
    return ''.join(out)


def stem_loop_names(gff3_path, db):
    """miRNA name -> name of the stem-loop it derives from (extractPreMiRName)."""
    out = {}
    with open(gff3_path) as f:
        rows = [line.strip().split('\t') for line in f if line[0] != '#']
    if db != 'miRBase':
        for c in rows:
            if c[2] == 'miRNA':
                name = c[-1].split('ID=')[1].split(';')[0].strip()
                out[name] = '_'.join(name.split('_')[:-1]) + '_pre'
        return out
    loops_of, loop_name = {}, {}
    for c in rows:
        attr = c[-1]
        if c[2] == 'miRNA':
            name = attr.split('Name=')[1].split(';')[0].strip()
            loops_of.setdefault(name, {})[attr.split('Derives_from=')[1].split(';')[0].strip()] = True
        if c[2] == 'miRNA_primary_transcript':
            loop_name[attr.split('Alias=')[1].split(';')[0].strip()] = attr.split('Name=')[1].split(';')[0].strip()
    for name, loops in loops_of.items():
        cands = [loop_name[a] for a in loops]
        if len(cands) == 1:
            out[name] = cands[0]
            continue
        tails = [c.split('-')[-1] for c in cands]             # several loci: lowest numeric suffix, else lowest name
        if all(t.isdigit() for t in tails):
            out[name] = min(zip((int(t) for t in tails), cands))[1]
        else:
            out[name] = min(cands)
    return out


def stem_loop_of(canonical, pre_of, db):
    """inferPremiRName: look the miRNA up, else derive the stem-loop name from the miRNA name."""
    if canonical in pre_of:
        return pre_of[canonical]
    if db == 'miRBase':
        stem = canonical.rsplit('-', 1)[0] if '-' in canonical else ''
        for k in (stem, stem + '-5p', stem + '-3p'):
            if k in pre_of:
                return pre_of[k]
        return canonical.replace('-5p', '').replace('-3p', '').replace('miR', 'mir')
    return canonical.replace('_5p*', '').replace('_3p*', '').replace('_5p', '').replace('_3p', '') + '_pre'


def _hairpin_slice(pre, a, b):
    """Hairpin bases at x in [a, b); '-' where the hairpin does not reach."""
    P = len(pre)
    return '-' * (min(b, 0) - min(a, 0)) + pre[max(a, 0):max(min(b, P), 0)] + '-' * (max(b, P) - max(a, P))


def describe_read(pre, lib_entry, read, pos, isomir_round):
    """(type, variant, pre_start, pre_end, cigar) of one annotated read, or None when the mature is not part of
    the hairpin (the reference drops such reads).  pos = 1-based offset bowtie reported in the library entry."""
    mir = lib_entry[2:-6]
    M, n, P = len(mir), len(read), len(pre)
    at = pre.find(mir)
    if at < 0:
        return None
    r0 = at - 2 + pos - (2 if isomir_round else 1)
    r1 = r0 + n
    m0, m1 = at, at + M
    x0 = min(0, at - 2, r0) if isomir_round else min(0, at - 2)      # left end of the padded strings
    variants = []
    if r0 == m0 and r1 == m1 and read == mir:
        kind = 'ref_miRNA'
        variants.append('NA')
    else:
        kind = 'isomiR'
        snp = None
        for x in range(max(m0, r0), min(m1, r1)):            # first substitution inside the mature
            if read[x - r0] != mir[x - m0]:
                i = x - x0                                    # the reference classifies by PADDED column, not by mature position
                snp = '_seed' if 1 <= i <= 6 else '_central_offset' if i == 7 else '_central' if 8 <= i <= 11 \
                    else 'central_supp' if 12 <= i <= 16 else ''
                break
        five = three = add = None
        if r0 > m0:
            five = str(m0 - r0)
        elif r0 < m0:
            five = '+' + str(m0 - r0)
            if read[:m0 - r0] != _hairpin_slice(pre, r0, m0):  # non-templated 5' bases also flag a substitution
                snp = ''
        if r1 < m1:
            three = str(r1 - m1)
        elif r1 > m1:
            if read[m1 - r0:] != _hairpin_slice(pre, m1, r1):
                add = '+' + str(r1 - m1)                      # non-templated 3' addition
            else:
                three = '+' + str(r1 - m1)
        if snp is not None:
            variants.append('iso_snp' + snp)
        if add is not None:
            variants.append('iso_add:' + add)
        if five is not None:
            variants.append('iso_5p:' + five)
        if three is not None:
            variants.append('iso_3p:' + three)
    return kind, ','.join(variants), r0 + 1, r1, cigar_against(read, _hairpin_slice(pre, r0, r1))


def cigar_against(read, ref):
    """make_cigar: matches are run-length coded ('M', '5M'), a substituted base is written as the read's base,
    'I' where the hairpin has no base."""
    out, run = [], 0
    for a, b in zip(read, ref):
        if a == b:
            run += 1
            continue
        if run:
            out.append('M' if run == 1 else '%dM' % run)
            run = 0
        out.append('I' if b == '-' else a)
    if run:
        out.append('M' if run == 1 else '%dM' % run)
    return ''.join(out)


def gff_rows(uniq, mirna_names, mirna_seqs, hairpin_names, hairpin_seqs, pre_of, db):
    """One record per unique sequence annotated in round 0 or 8: seq -> dict of GFF fields."""
    mirna = dict(zip(mirna_names, mirna_seqs))
    hairpin = dict(zip(hairpin_names, hairpin_seqs))
    seqs = np.char.decode(uniq['seq'], 'ascii')
    recs = {}
    for u in np.flatnonzero((uniq['round_of'] == 1) | (uniq['round_of'] == 9)):
        read = str(seqs[u])
        name = mirna_names[int(uniq['ref_of'][u])]
        canonical = name.split('.')[0] if '.' in name else name
        parent = stem_loop_of(canonical, pre_of, db)
        if parent not in hairpin:
            raise GffError('hairpin library has no entry %s (stem-loop of %s)' % (parent, name))
        pre = hairpin[parent]
        lib_entry = mirna[name]
        if '.SNP' in name and '.SNPC' not in name:            # SNP entry: the hairpin carries the variant mature
            canon_entry = mirna.get(canonical + '.SNPC')
            if canon_entry is None:
                raise GffError('miRNA library has no entry %s.SNPC next to %s' % (canonical, name))
            canon_mir = canon_entry[2:-6]
            at = pre.find(canon_mir)
            if at < 0:
                raise GffError('Error happens: %s is not part of %s' % (canonical, parent))
            mir = lib_entry[2:-6]
            pre = pre[:at] + mir + pre[at + len(mir):]
        d = describe_read(pre, lib_entry, read, int(uniq['pos_of'][u]) + 1, uniq['round_of'][u] == 9)
        if d is None:
            continue
        recs[read] = dict(miRName=name, preMiRName=parent, start=int(uniq['pos_of'][u]) + 1, type=d[0], variant=d[1], pre_start=d[2], pre_end=d[3], cigar=d[4],
                          uid=read_uid(read), counts=[int(c) for c in uniq['counts'][u]])
    return recs


def write_gff(outdir, recs, sample_names, db):
    """<sample>_isomiRs.gff per sample (writeDataToCSV.py:621-646); rows sorted by read (the reference's order is
    Python-2 dict order)."""
    source = 'miRBase22' if db == 'miRBase' else db + '2.0'
    paths = []
    for i, sample in enumerate(sample_names):
        base = os.path.splitext(sample)[0]
        path = os.path.join(outdir, base + '_isomiRs.gff')
        with open(path, 'w') as f:
            f.write('# GFF3 adapted for miRNA sequencing data\n## VERSION 0.0.1\n## source-ontology: %s\n## COLDATA: %s\n' % (source, base))
            for read in sorted(recs):
                r = recs[read]
                if r['counts'][i] >= 1:
                    f.write('\t'.join([r['miRName'], source, r['type'], str(r['pre_start']), str(r['pre_end']), '.', '+', '.']) + '\t' +
                            ';'.join(['Read ' + read, ' UID ' + r['uid'], ' Name ' + r['miRName'], ' Parent ' + r['preMiRName'],
                                      ' Variant ' + r['variant'], ' Cigar ' + r['cigar'], ' Expression ' + str(r['counts'][i]),
                                      ' Filter Pass']) + '\n')
        paths.append(path)
    return paths
