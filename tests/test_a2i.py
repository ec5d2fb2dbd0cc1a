"""-ai A-to-I report (mirge_b200/a2i.py, restating writeDataToCSV.py:1222-1615): known answers for the alignment
helpers, and the whole block on a synthetic sample with the CPU oracle standing in for the genome mapper -- every
reported site is recomputed independently from the raw counts."""
import os

import numpy as np
import pytest

from mirge_b200 import a2i

GENOME_SLOT = 8


def test_dash_helpers_and_local_alignment_frames():
    assert a2i.dash_count('--ACGT-') == (2, 1) and a2i.dash_count('ACGT') == (0, 0)
    assert a2i.remove_dash('--ACGT-') == 'ACGT'
    assert a2i.start_and_end('CGT', '--ACGT-') == (3, 5)
    t = 'TGAGGTAGTAGGTTGTATAGTT'
    # identical, 3' addition, 5' offset of the read into the flank, 5' truncation, one substitution
    assert a2i.local_align_gapless(t, t) == (t, t)
    assert a2i.local_align_gapless(t, t + 'AA') == (t + '--', t + 'AA')
    assert a2i.local_align_gapless(t, 'CC' + t) == ('--' + t, 'CC' + t)
    assert a2i.local_align_gapless(t, t[2:]) == (t, '--' + t[2:])
    sub = t[:9] + 'G' + t[10:]
    assert a2i.local_align_gapless(t, sub) == (t, sub)
    assert a2i.local_align_gapless('AAAA', 'CCCC') is None


def test_local_alignment_is_the_best_diagonal():
    """Independent check: score every diagonal with match 2 / mismatch -1 (maximum-sum run); the restatement's frame must
    lie on a diagonal of maximal score."""
    rng = np.random.default_rng(5)
    for _ in range(300):
        t = ''.join('ACGT'[x] for x in rng.integers(0, 4, int(rng.integers(18, 26))))
        a = int(rng.integers(0, 4))
        s = list(t[a:a + int(rng.integers(12, 24))])
        for _ in range(int(rng.integers(0, 3))):
            s[int(rng.integers(0, len(s)))] = 'ACGT'[int(rng.integers(0, 4))]
        s = ''.join(s) + ''.join('ACGT'[x] for x in rng.integers(0, 4, int(rng.integers(0, 3))))
        best = {}
        for d in range(-len(s) + 1, len(t)):
            run = top = 0
            for i in range(len(t)):
                j = i - d
                if 0 <= j < len(s):
                    run = max(0, run + (2 if t[i] == s[j] else -1))
                    top = max(top, run)
            best[d] = top
        ta, sa = a2i.local_align_gapless(t, s)
        d = a2i.dash_count(sa)[0] - a2i.dash_count(ta)[0]
        assert best[d] == max(best.values()) and len(ta) == len(sa)
        assert a2i.remove_dash(ta) == t and a2i.remove_dash(sa) == s


def test_judge_align_rules():
    t = 'TGAGGTAGTAGGTTGTATAGTT'                       # 22 nt: match limit 22 - 3 - 1 = 18
    ok = lambda s: a2i.judge_align(*a2i.local_align_gapless(t, s))       # noqa: E731
    assert ok(t) and ok(t + 'A') and ok(t[:20])
    assert ok(t[:9] + 'G' + t[10:])                    # one mismatch
    assert not ok(t[:9] + 'G' + t[10:14] + 'C' + t[15:])   # two mismatches
    assert ok(t[1:])                                   # starts one base late (head shift 1): limit drops by one
    assert not ok(t[2:])                               # starts two bases late
    # upstream quirk (:48): the scored window ends at len - headDashes - 1 - 3, so every 5' flank base of the read costs
    # one scored position: one extra base still passes, two do not
    assert ok('A' + t) and not ok('AA' + t)
    assert not ok(t[:15])                              # too few matches


def test_alignment_frame_grows_both_ways():
    t = 'TGAGGTAGTAGGTTGTATAGTT'
    aligned, states = a2i.align_to_target(t, [t, 'CC' + t, t + 'GG', t[1:]])
    assert len(set(len(x) for x in aligned)) == 1 and aligned[0] == '--' + t + '--'
    assert aligned[1:] == ['--' + t + '--', 'CC' + t + '--', '--' + t + 'GG', '---' + t[1:] + '--']
    assert states == [True, False, True, True]      # 'CC' + t: the two-base 5' flank fails the upstream window quirk


def test_binomial_and_site_statistics():
    import math
    tail = sum(math.comb(1000, k) * 0.001 ** k * 0.999 ** (1000 - k) for k in range(10, 60))     # P(at least 10 of 1000 at p = 0.001)
    assert abs(a2i.binom_cdf(990, 1000, 0.999) / tail - 1) < 1e-6
    t = 'TGAGGTAGTAGGTTGTATAGTT'
    edited = t[:6] + 'G' + t[7:]                      # A at position 7 -> G
    import io
    d = io.StringIO()
    kept, pos, pc, ratio, pv, total, nseq, canon = a2i.a2i_editing(t, [t, edited, t + 'A'], [900, 80, 20], 'm', d,
                                                                   {t: True, edited: True, t + 'A': True})
    assert pos == [7] and pc == {7: 80} and total == 1000 and nseq == 3 and canon == 900
    assert abs(ratio[7] - 0.08) < 1e-12 and pv[7] < 1e-50
    # a read the genome filter dropped takes no part; positions inside the last five bases are not scored
    late = t[:18] + 'G' + t[19:]
    kept, pos, pc, ratio, pv, total, nseq, canon = a2i.a2i_editing(t, [t, edited, late], [900, 80, 50], 'm', d, {t: True, late: True})
    assert pos == [] and total == 950 and nseq == 2
    assert 'retained seqList size is: 2' in d.getvalue()


@pytest.fixture(scope='module')
def a2i_world(tmp_path_factory):
    from synth.world import World
    w = World(seed=20180626, scale=0.05)
    root = tmp_path_factory.mktemp('a2ilibs')
    w.write(str(root))
    names, chroms = w.write_genome(str(root), total=36000)      # small: the oracle scans every offset
    return w, str(root), names, chroms


def test_whole_block_with_the_oracle_as_genome_mapper(a2i_world, oracle_mod, tmp_path):
    from conftest import ILLUMINA
    from mirge_b200.libs import LibraryDir
    w, root, gnames, chroms = a2i_world
    O = oracle_mod
    samples = ['s1.fastq', 's2.fastq']
    o = O.Oracle(2, False, ILLUMINA)
    for t, i in O.LIB_IDS.items():
        o.load_library(i, *w.libs[t])
    o.load_library(GENOME_SLOT, gnames, chroms)
    for s in range(2):
        o.feed_fastq(s, w.fastq(60000, seed=500 + s).tobytes())
    o.annotate()
    tabs = O.run_tables(o, samples, w.libs['mirna'][0], w.merges)
    items = [(s, v['annot'][1], v['annot'][9], list(v['quant'])) for s, v in tabs['seqDic'].items()]
    ld = LibraryDir(root, 'synth', 'miRBase')

    cache = {}

    def hits(s, n):
        if (s, n) not in cache:
            cache[(s, n)] = o.genome_hits(GENOME_SLOT, s, 0, n, 28, 2)
        return cache[(s, n)]

    def genome_pass(seqs, n):
        return [hits(s, n) for s in seqs]

    rep = a2i.write_a2i(str(tmp_path), items, samples, tabs['logDic'], tabs['mirDic'], ld.mature_sequences(), ld.merged_names(),
                        ld.repetitive_mirnas(), genome_pass, log=lambda *a: None)
    lines = open(rep).read().strip().split('\n')
    assert lines[0].startswith('miRNA,A-to-I position in the miRNA,miRNA sequence,s1.readCount,s1.readCount.canonical')
    assert len(lines) > 5                                                  # planted A>G substitutions of the expressed miRNAs
    rows = [l.split(',') for l in lines[1:]]
    assert rows == sorted(rows, key=lambda r: (r[0].encode(), int(r[1])))
    mature = ld.mature_sequences()
    mrf = [tabs['logDic']['quantStats'][i]['mirnaReadsFiltered'] for i in range(2)]
    merged = ld.merged_names()
    n_checked = 0
    for r in rows:
        name, pos, target = r[0], int(r[1]), r[2]
        assert mature[name] == target and target[pos - 1] == 'A' and pos <= len(target) - 5
        for i in range(2):
            f = r[3 + 7 * i:10 + 7 * i]
            if f[0] == 'NE' or f[5] == 'NE':
                continue
            # independent recount: reads of this miRNA whose base aligned to `pos` is G, over reads kept by the rules
            edited = total = 0
            for s, exact, iso, q in items:
                nm = exact or iso
                if not nm or merged.get(nm, nm) != name or q[i] == 0:
                    continue
                if not (1e6 * q[i] / mrf[i] >= 1 or exact):
                    continue
                ta, sa = a2i.local_align_gapless(target, s)
                if not a2i.judge_align(ta, sa):
                    continue
                c = hits(s, 1)
                occ = [x for x in c if x]
                if not occ or occ[0] != 1:
                    continue
                total += q[i]
                k = a2i.dash_count(ta)[0] + pos - 1
                if k < len(sa) and sa[k] == 'G':
                    edited += q[i]
            assert int(f[0]) == total and int(f[3]) == edited, (name, pos, i)
            assert f[5] == '%.2f%%' % (100.0 * edited / total)
            n_checked += 1
    assert n_checked > 5
    assert os.path.isfile(os.path.join(str(tmp_path), 'a2IEditing.detail.txt'))
    assert open(os.path.join(str(tmp_path), 'a2IEditing.report.newform.csv')).readline() == 'miRNA:position,sample,A-to-I percentage,log2RPM\n'
