"""GPU trimming parity on adversarial reads: the bit-vector fast path, the literal
DP slow path and the CPU oracle must produce the same collapsed table."""
import numpy as np
import pytest

from conftest import ILLUMINA

pytestmark = pytest.mark.gpu


def _fastq(reads, quals=None):
    out = []
    for k, r in enumerate(reads):
        q = quals[k] if quals is not None else 'I' * len(r)
        out.append('@r%d\n%s\n+\n%s\n' % (k, r, q))
    return ''.join(out).encode()


def _mutate(rng, s, nsub, nins, ndel):
    s = list(s)
    for _ in range(nsub):
        if s:
            p = rng.integers(0, len(s)); s[p] = 'ACGT'[rng.integers(0, 4)]
    for _ in range(nins):
        p = rng.integers(0, len(s) + 1); s.insert(p, 'ACGT'[rng.integers(0, 4)])
    for _ in range(ndel):
        if s:
            del s[rng.integers(0, len(s))]
    return ''.join(s)


def adversarial_reads(rng, adapter, n, maxlen=64, alphabet='ACGT'):
    reads = []
    m = len(adapter)
    for _ in range(n):
        kind = rng.integers(0, 6)
        ins_len = int(rng.integers(0, maxlen))
        if kind == 5:      # low-complexity / adapter-like insert
            unit = adapter[int(rng.integers(0, m)):][:int(rng.integers(1, 6))] or 'A'
            insert = (unit * maxlen)[:ins_len]
        else:
            insert = ''.join(alphabet[x] for x in rng.integers(0, len(alphabet), ins_len))
        if kind == 0:
            tail = ''
        else:
            tail = _mutate(rng, adapter, int(rng.integers(0, 4)), int(rng.integers(0, 3)), int(rng.integers(0, 3)))
            if kind == 2:  # second partial copy
                tail = tail[:int(rng.integers(0, m))] + adapter
        r = (insert + tail)[:maxlen]
        r = r[:int(rng.integers(max(1, len(r) - 12), len(r) + 1))] if len(r) > 1 and kind == 3 else r
        if rng.random() < 0.05 and r:
            p = int(rng.integers(0, len(r))); r = r[:p] + 'N' + r[p + 1:]
        reads.append(r if r else 'A')
    return reads


def _tables(O, reads_fq, adapter, max_read_len=64, force_slow=False, literal=False, mode=None, **kw):
    from mirge_b200.engine import Engine
    o = O.Oracle(1, False, adapter)
    if kw:
        o.set_trim_params(**kw)
    o.feed_fastq(0, reads_fq, 33)
    oe = o.export()
    want = {s: int(c[0]) for s, c in zip(oe['seqs'], oe['counts'])}
    e = Engine(1, False, 1 << 20, max_read_len)
    e.set_adapter(adapter, **kw)
    if force_slow:
        e.force_slow_path(True, literal)
    if mode is not None:
        e.set_trim_mode(mode)
    e.feed_fastq(0, reads_fq, 33)
    ee = e.export_unique()
    got = {s.decode(): int(c[0]) for s, c in zip(ee['seq'], ee['counts'])}
    p, k, u = e.stats()
    op, ok = o.stats()
    assert int(k[0]) == int(ok[0])
    return want, got, e.slow_reads(), int(p[0])


@pytest.mark.parametrize('adapter', [ILLUMINA, 'TGGAATTCTCGGGTGCCAAGG', 'AAAAAAAAAAAAAAAA', 'ACACACACACACACACACAC',
                                     'TGGAATTC', 'AGATCGGAAGAGCACACGTCTGAACTCCAGTC', 'AGATCGGAAGAGC,TGGAATTCTCGGGTGCCAAGG'])
def test_adversarial_reads_fast_and_slow(oracle_mod, adapter):
    rng = np.random.default_rng(7)
    reads = adversarial_reads(rng, adapter.split(',')[0], 60000)
    fq = _fastq(reads)
    want, got, slow, n = _tables(oracle_mod, fq, adapter)
    assert got == want
    assert slow < n          # the fast path is actually taken
    want2, got2, slow2, n2 = _tables(oracle_mod, fq, adapter, force_slow=True)               # all-column bit-vector path
    n_long = sum(len(r) >= 16 for r in reads)            # K0 does not emit reads the length filter must drop anyway
    assert got2 == want2 and slow2 == n_long
    want3, got3, slow3, n3 = _tables(oracle_mod, fq, adapter, force_slow=True, literal=True)  # literal cutadapt DP
    assert got3 == want3 and slow3 == n_long
    want4, got4, _, _ = _tables(oracle_mod, fq, adapter, mode=4)                             # exact-match prefilter off
    assert got4 == want4


def prefilter_reads(rng, adapter, n, maxlen=64):
    """Reads that end in an EXACT adapter prefix (or hold the whole adapter) behind inserts built from
    random bases, adapter pieces and edited adapter copies: the shapes the prefilter settles without DP,
    and the ones that must make it back off."""
    m = len(adapter)
    reads = []
    for _ in range(n):
        ilen = int(rng.integers(0, maxlen + 1))
        ins = []
        kind = int(rng.integers(0, 4))
        while len(ins) < ilen:
            mode = 0 if kind == 0 else int(rng.integers(0, 4))
            if mode == 0:
                ins.append('ACGT'[rng.integers(0, 4)])
            elif mode == 1:
                a = int(rng.integers(0, m)); ins.extend(adapter[a:a + int(rng.integers(1, m + 1))])
            elif mode == 2:
                ins.extend(_mutate(rng, adapter[:int(rng.integers(1, m + 1))], int(rng.integers(0, 3)), int(rng.integers(0, 2)), int(rng.integers(0, 2))))
            else:
                ins.append(adapter[rng.integers(0, m)])
        ins = ''.join(ins[:ilen])
        tail = adapter[:int(rng.integers(0, m + 1))]
        if rng.random() < 0.15:
            tail = adapter + ''.join('ACGT'[x] for x in rng.integers(0, 4, int(rng.integers(0, 8))))
        r = (ins + tail)[:maxlen]
        if rng.random() < 0.03 and r:
            q = int(rng.integers(0, len(r))); r = r[:q] + 'N' + r[q + 1:]
        reads.append(r if r else 'A')
    return reads


@pytest.mark.parametrize('adapter', [ILLUMINA, 'TGGAATTCTCGGGTGCCAAGG', 'AAAAAAAAAAAAAAAA', 'ACACACACACACACACACAC', 'TGGAATTC',
                                     'AGATCGGAAGAGCACACGTCTGAACTCCAGTC', 'TTTTGGAATTTTGGAATTGGAATT'])
@pytest.mark.parametrize('maxlen', [50, 64, 128])
def test_prefilter_classes_match_literal_dp(oracle_mod, adapter, maxlen):
    rng = np.random.default_rng(len(adapter) * 131 + maxlen)
    reads = prefilter_reads(rng, adapter, 80000, maxlen)
    fq = _fastq(reads)
    mrl = 64 if maxlen <= 64 else 128
    from mirge_b200.engine import Engine
    o = oracle_mod.Oracle(1, False, adapter)
    o.feed_fastq(0, fq, 33)
    oe = o.export()
    want = {s: int(c[0]) for s, c in zip(oe['seqs'], oe['counts'])}
    e = Engine(1, False, 1 << 20, mrl)
    e.set_adapter(adapter)
    e.feed_fastq(0, fq, 33)
    ee = e.export_unique()
    got = {s.decode(): int(c[0]) for s, c in zip(ee['seq'], ee['counts'])}
    assert got == want
    slow, c1, c2, _ = e.trim_classes()
    assert c2 < len(reads)               # some reads are settled by exact comparisons alone (c2 includes the class-1 reads the piece filter passes on)
    if adapter == ILLUMINA:
        assert c1 > 0


def test_long_adapter_and_long_reads(oracle_mod):
    rng = np.random.default_rng(11)
    adapter = 'AGATCGGAAGAGCACACGTCTGAACTCCAGTCACATCACGATCTCGTAT'   # 49 nt: literal DP only
    reads = adversarial_reads(rng, adapter, 20000, maxlen=120)
    want, got, slow, n = _tables(oracle_mod, _fastq(reads), adapter, max_read_len=128)
    assert got == want and slow == sum(len(r) >= 16 for r in reads)
    reads = adversarial_reads(rng, ILLUMINA, 20000, maxlen=128)
    want, got, slow, n = _tables(oracle_mod, _fastq(reads), ILLUMINA, max_read_len=128)
    assert got == want


def test_quality_trim_and_modes(oracle_mod):
    rng = np.random.default_rng(3)
    reads = adversarial_reads(rng, ILLUMINA, 20000, maxlen=60)
    quals = []
    for r in reads:
        q = rng.choice(list('#,:FI5!'), len(r), p=[.15, .1, .15, .3, .2, .05, .05])
        quals.append(''.join(q))
    fq = _fastq(reads, quals)
    for adapter in (ILLUMINA, 'none', '+3', '+-4'):
        want, got, _, _ = _tables(oracle_mod, fq, adapter)
        assert got == want, adapter


def test_phred64_base(oracle_mod):
    rng = np.random.default_rng(5)
    reads = adversarial_reads(rng, ILLUMINA, 5000, maxlen=50)
    quals = [''.join(rng.choice(list('BCDfghZ;'), len(r))) for r in reads]   # ';' < 64: negative phred64 quality
    fq = _fastq(reads, quals)
    from mirge_b200.engine import Engine
    assert Engine.detect_phred(fq) == oracle_mod.Oracle.detect_phred(fq)
    o = oracle_mod.Oracle(1, False, ILLUMINA)
    base = o.feed_fastq(0, fq)
    e = Engine(1, False, 1 << 16, 64)
    e.set_adapter(ILLUMINA)
    assert e.feed_fastq(0, fq) == base
    oe, ee = o.export(), e.export_unique()
    assert {s: int(c[0]) for s, c in zip(oe['seqs'], oe['counts'])} == {s.decode(): int(c[0]) for s, c in zip(ee['seq'], ee['counts'])}


@pytest.mark.parametrize('kw', [dict(error_rate=0.0), dict(error_rate=0.2, min_overlap=1), dict(error_rate=0.34, min_overlap=5),
                                dict(error_rate=0.12, min_overlap=3, min_length=1, qual_cutoff=20)])
def test_non_default_trim_parameters(oracle_mod, kw):
    """error rate / min overlap / min length / quality cutoff are constants in the reference; the kernels take them as parameters."""
    rng = np.random.default_rng(13)
    reads = adversarial_reads(rng, ILLUMINA, 30000)
    quals = [''.join(rng.choice(list('#,5I'), len(r), p=[.1, .1, .2, .6])) for r in reads]
    fq = _fastq(reads, quals)
    for mode in (dict(), dict(force_slow=True), dict(force_slow=True, literal=True)):
        want, got, _, _ = _tables(oracle_mod, fq, ILLUMINA, **mode, **kw)
        assert got == want, (kw, mode)
