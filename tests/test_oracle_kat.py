"""Known-answer tests pinning the CPU oracle (hand-derived from the cutadapt 1.16 /
bowtie 1.1.2 rules restated in SURVEY.md Appendix A/B; the reference ships none)."""
import numpy as np
import pytest

from conftest import ILLUMINA

X20 = 'ACACACACACACACACACAC'


@pytest.mark.parametrize('qual,base,expect', [
    ('IIIII', 33, (0, 5)),          # all high quality: nothing trimmed
    ('IIII#', 33, (0, 4)),          # one bad base at the 3' end
    ('II###', 33, (0, 2)),
    ('####', 33, (0, 0)),           # start >= stop -> empty read
    ('I#I##', 33, (0, 3)),          # BWA rule: partial sum goes negative at the good base
    ('#5#', 33, (0, 2)),            # 8, then 8 + 10 - 20 < 0
    ('#+#', 33, (0, 0)),            # q == cutoff keeps the running sum, argmax moves to 0
    ('<<hhh', 64, (2, 5)),          # negative phred64 qualities trim the 5' end even with cutoff 0
])
def test_quality_trim_index(oracle_mod, qual, base, expect):
    assert oracle_mod.quality_trim_index(qual, 0, 10, base) == expect


def _trim(O, adapter, read, qual=None):
    o = O.Oracle(1, False, adapter)
    fq = ('@r\n%s\n+\n%s\n' % (read, qual or 'I' * len(read))).encode()
    st, sp = o.trim_fastq(fq, 33)
    return int(st[0]), int(sp[0])


def test_full_adapter_at_each_offset(oracle_mod):
    for p in (0, 1, 15, 16, 20, 30):
        ins = (X20 * 2)[:p]
        st, sp = _trim(oracle_mod, ILLUMINA, ins + ILLUMINA + 'ACGT')
        assert (st, sp) == ((0, p) if p >= 16 else (0, -1))       # shorter than 16 nt -> dropped


def test_partial_adapter_min_overlap(oracle_mod):
    assert _trim(oracle_mod, ILLUMINA, X20 + 'TGG') == (0, 20)      # 3-nt overlap is removed
    assert _trim(oracle_mod, ILLUMINA, X20 + 'TG') == (0, 22)       # 2-nt overlap is not


def test_partial_adapter_error_budget(oracle_mod):
    # allowed errors: overlap 3-8 -> 0, 9-16 -> 1, 17-24 -> 2, 25-29 -> 3
    assert _trim(oracle_mod, ILLUMINA, X20 + 'TGGATTTCT') == (0, 20)   # 9 nt, one substitution: accepted
    assert _trim(oracle_mod, ILLUMINA, X20 + 'TGGATTTC') == (0, 28)    # 8 nt, one substitution: rejected
    m = oracle_mod.locate(ILLUMINA, X20 + 'TGGATTTCT')
    assert m == (0, 9, 20, 29, 8, 1)


def test_indel_inside_adapter(oracle_mod):
    a_del = ILLUMINA[:10] + ILLUMINA[11:]            # adapter with one base missing
    assert _trim(oracle_mod, ILLUMINA, X20 + a_del) == (0, 20)
    a_ins = ILLUMINA[:10] + 'A' + ILLUMINA[10:]      # adapter with one extra base
    assert _trim(oracle_mod, ILLUMINA, X20 + a_ins) == (0, 20)


def test_length_filter_and_n(oracle_mod):
    assert _trim(oracle_mod, ILLUMINA, 'ACGTACGTACGTACG' + ILLUMINA)[1] == -1        # 15 nt: dropped
    assert _trim(oracle_mod, ILLUMINA, 'ACGTACGTACGTACGT' + ILLUMINA) == (0, 16)     # 16 nt: kept
    # N never matches an adapter base: the 29-mer with one N is found with one error instead
    a_n = ILLUMINA[:5] + 'N' + ILLUMINA[6:]
    assert _trim(oracle_mod, ILLUMINA, X20 + a_n) == (0, 20)
    assert _trim(oracle_mod, 'none', X20 + ILLUMINA) == (0, 49)                      # -ad none: untouched
    assert _trim(oracle_mod, '+3', X20 + 'ACGT') == (3, 24)                          # UnconditionalCutter(3)


def test_quality_then_adapter(oracle_mod):
    read = X20 + ILLUMINA[:12] + 'GGGG'
    qual = 'I' * (20 + 12) + '####'
    assert _trim(oracle_mod, ILLUMINA, read, qual) == (0, 20)     # tail removed first, then the 12-nt overlap


def test_c_python_and_unbanded_locate_agree(oracle_mod):
    rng = np.random.default_rng(0)
    adapters = [ILLUMINA, 'AAAAAAAAAAAA', 'ACGTACGTAC', 'TGGAATTC']
    for _ in range(3000):
        a = adapters[int(rng.integers(0, len(adapters)))]
        n = int(rng.integers(0, 60))
        ins = ''.join('ACGTN'[x] for x in rng.integers(0, 5, n))
        t = list(a[:int(rng.integers(0, len(a) + 1))])
        for _ in range(int(rng.integers(0, 3))):
            if t:
                t[int(rng.integers(0, len(t)))] = 'ACGT'[int(rng.integers(0, 4))]
        read = ins[:int(rng.integers(0, n + 1))] + ''.join(t) + ins[:int(rng.integers(0, 4))]
        c = oracle_mod.locate(a, read)
        assert c == oracle_mod.py_locate(a, read)
        assert c == oracle_mod.locate(a, read, banded=False)   # the Ukkonen band is only an optimisation


def test_cutadapt_user_guide_examples(oracle_mod):
    """The 3' adapter table of the cutadapt user guide ("Adapter types", -a ADAPTER), recalled from the
    published documentation (the docs are not in this image): full adapter anywhere, full adapter at the
    3' end and a partial adapter at the 3' end are removed with everything after them; with the default
    10 % error rate a 7-letter adapter tolerates no error."""
    O = oracle_mod

    def trimmed(read, rate=0.1, overlap=3):
        m = O.match_back('ADAPTER', read, rate, overlap)
        return read if m is None else read[:m[0]]
    assert trimmed('MYSEQUENCEADAPTERSOMETHINGELSE') == 'MYSEQUENCE'
    assert trimmed('MYSEQUENCEADAPTER') == 'MYSEQUENCE'
    assert trimmed('MYSEQUENCEADAP') == 'MYSEQUENCE'
    assert trimmed('ADAPTERSOMETHING') == ''
    assert trimmed('MYSEQUENCEAD') == 'MYSEQUENCEAD'            # overlap 2 < minimum overlap 3
    assert trimmed('MYSEQUENCEAD', overlap=2) == 'MYSEQUENCE'
    assert trimmed('MYSEQUENCEADABTER') == 'MYSEQUENCEADABTER'  # one error in 7 letters: 1 > 0.1 * 7
    assert trimmed('MYSEQUENCEADABTER', rate=0.2) == 'MYSEQUENCE'
    assert trimmed('MYSEQUENCE') == 'MYSEQUENCE'


# -- bowtie acceptance rules ---------------------------------------------------
def _lib_oracle(O, seqs, lib_id=0):
    o = O.Oracle(1, False, 'none')
    o.load_library(lib_id, ['r%d' % i for i in range(len(seqs))], seqs)
    return o


def _rand(rng, n):
    return ''.join('ACGT'[x] for x in rng.integers(0, 4, n))


def _sub(s, p):
    return s[:p] + ('A' if s[p] != 'A' else 'C') + s[p + 1:]


def test_n_mode_seed_and_quality_budget(oracle_mod):
    rng = np.random.default_rng(1)
    ref = _rand(rng, 200)
    o = _lib_oracle(oracle_mod, [_rand(rng, 50), ref])
    read = ref[30:70]                                   # 40 nt
    assert o.map_one(0, read, 0, 0) == (1, 30, 0)
    assert o.map_one(0, _sub(_sub(read, 30), 35), 0, 0) == (1, 30, 2)       # -n 0: two mismatches beyond the 28-nt seed
    assert o.map_one(0, _sub(read, 27), 0, 0) is None                       # one mismatch inside the seed
    assert o.map_one(0, _sub(_sub(_sub(read, 30), 35), 38), 0, 0) is None   # three mismatches: 90 > -e 70
    assert o.map_one(0, _sub(read, 5), 0, 1) == (1, 30, 1)                  # -n 1
    assert o.map_one(0, _sub(_sub(read, 5), 20), 0, 1) is None              # two seed mismatches
    assert o.map_one(0, _sub(_sub(read, 5), 33), 0, 1) == (1, 30, 2)
    short = ref[100:118]                                # the whole read is the seed
    assert o.map_one(0, _sub(short, 17), 0, 0) is None
    assert o.map_one(0, _sub(short, 17), 0, 1) == (1, 100, 1)


def test_v_mode_and_entry_bounds(oracle_mod):
    rng = np.random.default_rng(2)
    a, b = _rand(rng, 60), _rand(rng, 60)
    o = _lib_oracle(oracle_mod, [a, b])
    assert o.map_one(0, a[50:] + b[:10], 1, 2) is None                      # a placement may not span two entries
    r = b[5:30]
    assert o.map_one(0, _sub(_sub(r, 3), 20), 1, 2) == (1, 5, 2)
    assert o.map_one(0, _sub(_sub(_sub(r, 3), 12), 20), 1, 2) is None
    assert o.map_one(0, r[:10] + 'N' + r[11:], 1, 1) == (1, 5, 1)           # N in the read is a mismatch
    o2 = _lib_oracle(oracle_mod, [a[:20] + 'N' + a[21:]])
    assert o2.map_one(0, a[10:30], 1, 2) is None                            # ambiguous reference base: unindexed
    assert o2.map_one(0, a[21:45], 1, 0) == (0, 21, 0)


def test_tie_policy(oracle_mod):
    rng = np.random.default_rng(3)
    core = _rand(rng, 22)
    seqs = ['GG' + core + 'TTTTTT', _rand(rng, 40), 'CC' + core + 'AAAAAA', 'CC' + _sub(core, 10) + 'AAAAAA']
    o = _lib_oracle(oracle_mod, seqs)
    assert o.map_one(0, core, 0, 0) == (0, 2, 0)                            # lowest (ref, pos)
    assert o.map_one(0, core, 1, 1, highest=True) == (2, 2, 0)              # best stratum only, highest (ref, pos)
    assert o.map_one(0, _sub(core, 10), 1, 2) == (3, 2, 0)                  # --best: fewest mismatches first


def test_indexed_search_equals_bruteforce(oracle_mod):
    rng = np.random.default_rng(4)
    seqs = [_rand(rng, int(rng.integers(30, 120))) for _ in range(40)]
    seqs[7] = seqs[7][:15] + 'NN' + seqs[7][17:]
    seqs.append(seqs[3][5:60])                                               # duplicated region: ties
    o = _lib_oracle(oracle_mod, seqs)
    for _ in range(1500):
        s = seqs[int(rng.integers(0, len(seqs)))]
        n = int(rng.integers(11, min(50, len(s)) + 1))
        p = int(rng.integers(0, len(s) - n + 1))
        r = s[p:p + n]
        for _ in range(int(rng.integers(0, 4))):
            r = _sub(r, int(rng.integers(0, n)))
        for mode, param in ((0, 0), (0, 1), (1, 0), (1, 1), (1, 2)):
            for hi in (False, True):
                assert o.map_one(0, r, mode, param, hi) == o.map_one(0, r, mode, param, hi, brute=True)


def test_round_order_and_polyT(oracle_mod):
    """runAnnotationPipeline.py:636-705 on a hand-built world."""
    rng = np.random.default_rng(5)
    mature = _rand(rng, 22)
    mir = 'AC' + mature + 'GATTAC'
    hp = _rand(rng, 30) + mir + _rand(rng, 30)
    pre = _rand(rng, 80)
    o = oracle_mod.Oracle(1, False, 'none')
    libs = {'mirna': [mir], 'hairpin': [hp], 'mature_trna': [_rand(rng, 70)], 'pre_trna': [pre], 'snorna': [_rand(rng, 90)],
            'rrna': [_rand(rng, 90)], 'ncrna_others': [_rand(rng, 90)], 'mrna': [_rand(rng, 300)]}
    for t, s in libs.items():
        o.load_library(oracle_mod.LIB_IDS[t], [t], s)
    reads = {
        'exact': mature,                               # round 1 (exact miRNA)
        'long': hp[28:28 + 27],                        # 27 nt > 25: skips round 1, hairpin round 2
        'isomir': 'T' + _sub(mature, 9)[1:] + 'AA',    # mismatching ends are cut by -5 1 -3 2, one mismatch inside: round 9
        'polyt': pre[20:33] + 'TTTT',                  # 13-nt remainder + poly-T: round 4
        'polyt_short': pre[20:30] + 'TTTTTT',          # 10-nt remainder: not eligible
        'noise': _rand(rng, 30),
    }
    fq = ''.join('@%s\n%s\n+\n%s\n' % (k, s, 'I' * len(s)) for k, s in reads.items()).encode()
    o.feed_fastq(0, fq, 33)
    o.annotate()
    e = o.export()
    got = {s: int(r) for s, r in zip(e['seqs'], e['round_of'])}
    assert got[reads['exact']] == 1
    assert got[reads['long']] == 2
    assert got[reads['isomir']] == 9
    assert got[reads['polyt']] == 4
    assert got[reads['polyt_short']] == 0
    assert got[reads['noise']] == 0
