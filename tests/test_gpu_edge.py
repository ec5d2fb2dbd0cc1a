"""GPU edge cases: long reads (W=4 path) through every round, ragged / empty / CRLF /
lower-case input, reads made of N, libraries holding ambiguous bases."""
import numpy as np
import pytest

from conftest import ILLUMINA

pytestmark = pytest.mark.gpu


def _both(O, world, fq, adapter, spike=False, max_read_len=64, libs=None):
    from mirge_b200.engine import Engine, LIB_IDS
    libs = libs or world.libs
    o = O.Oracle(1, spike, adapter)
    e = Engine(1, spike, 1 << 18, max_read_len)
    e.set_adapter(adapter)
    for t, i in LIB_IDS.items():
        if t in libs:
            o.load_library(O.LIB_IDS[t], *libs[t])
            e.load_library(i, *libs[t])
    if len(fq):
        o.feed_fastq(0, fq, 33)
    e.feed_fastq(0, fq, 33)
    o.annotate()
    e.annotate()
    e.tally()
    oe, ee = o.export(), e.export_unique()
    want = {s: (int(c[0]), int(r), int(rf) if r else 0, int(p) if r else 0, int(m) if r else 0)
            for s, c, r, rf, p, m in zip(oe['seqs'], oe['counts'], oe['round_of'], oe['ref_of'], oe['pos_of'], oe['mm_of'])}
    got = {s.decode(): (int(c[0]), int(r), int(rf) if r else 0, int(p) if r else 0, int(m) if r else 0)
           for s, c, r, rf, p, m in zip(ee['seq'], ee['counts'], ee['round_of'], ee['ref_of'], ee['pos_of'], ee['mm_of'])}
    assert got == want
    q, i, t = o.tally()
    et = e.export_tables()
    assert np.array_equal(et['mir_quant'], q) and np.array_equal(et['mir_iscan'], i) and np.array_equal(et['lib_totals'], t)
    op, ok = o.stats()
    ep, ek, _ = e.stats()
    assert (int(ep[0]), int(ek[0])) == (int(op[0]), int(ok[0]))
    return got


def test_long_reads_all_rounds(small_world, oracle_mod):
    """100-nt reads (W=4 kernels): long inserts reach the hairpin / mRNA / ncRNA rounds untrimmed."""
    fq = small_world.fastq(40000, seed=9, L=100).tobytes()
    got = _both(oracle_mod, small_world, fq, ILLUMINA, spike=True, max_read_len=128)
    rounds = {v[1] for v in got.values()}
    assert {1, 2, 3, 5, 6, 7, 8, 9}.issubset(rounds)
    # collapse keys and mapped reads longer than 64 nt: 70-100 nt mRNA / ncRNA fragments, 0-2 mismatches past the seed
    rng = np.random.default_rng(1)
    reads = []
    for _ in range(3000):
        lib = small_world.libs['mrna' if rng.random() < 0.5 else 'ncrna_others'][1]
        s = lib[int(rng.integers(0, len(lib)))]
        n = int(rng.integers(70, 101))
        if len(s) < n:
            continue
        a = int(rng.integers(0, len(s) - n + 1))
        r = list(s[a:a + n])
        for _ in range(int(rng.integers(0, 3))):
            p = int(rng.integers(28, n))
            r[p] = 'ACGT'[(('ACGT'.index(r[p])) + 1) % 4]
        reads.append((''.join(r) + ILLUMINA)[:120])
    fq2 = ''.join('@r%d\n%s\n+\n%s\n' % (k, r, 'I' * len(r)) for k, r in enumerate(reads)).encode()
    got2 = _both(oracle_mod, small_world, fq2, ILLUMINA, spike=True, max_read_len=128)
    assert max(len(s) for s in got2) > 90
    assert {7, 8}.issubset({v[1] for v in got2.values()})


def test_ragged_and_degenerate_input(small_world, oracle_mod):
    rng = np.random.default_rng(2)
    mir = small_world.libs['mirna'][1][0][2:24]
    reads = ['', 'A', 'N' * 30, 'ACGT' * 16, mir + ILLUMINA[:20], mir[:10] + 'N' + mir[11:] + ILLUMINA,
             'T' * 40, mir + 'TTTT', ILLUMINA, ILLUMINA[3:] + 'ACGTACGTACGTACGTACGT']
    reads += [''.join('ACGTN'[x] for x in rng.integers(0, 5, int(rng.integers(0, 64)))) for _ in range(3000)]
    for eol in ('\n', '\r\n'):
        fq = ''.join('@r%d%s%s%s+%s%s%s' % (k, eol, r, eol, eol, 'I' * len(r), eol) for k, r in enumerate(reads)).encode()
        if eol == '\r\n':
            # cutadapt strips the line ending; the oracle and the kernel both drop a trailing CR
            pass
        got = _both(oracle_mod, small_world, fq, ILLUMINA)
        assert (mir.upper() in got)


def test_lower_case_reads_are_canonicalised():
    """Documented deviation: the reference keeps the original bytes of a read in the
    collapse key; the packed representation upper-cases and turns every non-ACGT character into N."""
    from mirge_b200.engine import Engine
    e = Engine(1, False, 1 << 12, 64)
    e.set_adapter('none')
    r = 'ACGTACGTACGTACGTACGT'
    fq = ''.join('@r\n%s\n+\n%s\n' % (x, 'I' * 20) for x in (r, r.lower(), r[:5] + '.' + r[6:], r[:5] + 'n' + r[6:])).encode()
    e.feed_fastq(0, fq, 33)
    u = e.export_unique()
    got = {s.decode(): int(c[0]) for s, c in zip(u['seq'], u['counts'])}
    assert got == {r: 2, r[:5] + 'N' + r[6:]: 2}


def test_empty_chunk_and_no_kept_reads(small_world, oracle_mod):
    from mirge_b200.engine import Engine
    e = Engine(1, False, 1 << 12, 64)
    e.set_adapter(ILLUMINA)
    e.feed_fastq(0, b'', 33)
    e.feed_fastq(0, b'@r\nACGT\n+\nIIII\n', 33)        # too short: dropped
    e.annotate()
    e.tally()
    p, k, u = e.stats()
    assert (int(p[0]), int(k[0]), u) == (1, 0, 0)
    t = e.export_tables()
    assert int(t['lib_totals'].sum()) == 0 and int(t['len_hist'].sum()) == 0


def test_library_with_ambiguous_bases_and_tiny_entries(oracle_mod, small_world):
    rng = np.random.default_rng(4)

    def rnd(n):
        return ''.join('ACGT'[x] for x in rng.integers(0, 4, n))
    hp = [rnd(90) for _ in range(30)]
    hp[3] = hp[3][:40] + 'NNNN' + hp[3][44:]
    hp[7] = 'acgu' * 20                                  # lower case and U in a library
    hp.append('ACGT')                                    # entry shorter than any seed
    hp.append('')                                        # empty entry
    libs = dict(small_world.libs)
    libs['hairpin'] = (['h%d' % i for i in range(len(hp))], hp)
    reads = [hp[3][30:62], hp[3][44:80], hp[7][2:40].upper().replace('U', 'T'), hp[1][10:50], hp[29][50:90], hp[28][60:] + hp[29][:10]]
    fq = ''.join('@r%d\n%s\n+\n%s\n' % (k, r, 'I' * len(r)) for k, r in enumerate(reads)).encode()
    got = _both(oracle_mod, small_world, fq, 'none', libs=libs)
    assert got[reads[0]][1] == 0            # overlaps the ambiguous stretch
    assert got[reads[1]][1] == 2
    assert got[reads[2]][1] == 2
    assert got[reads[5]][1] == 0            # spans two entries


def test_short_reads_use_the_scan_fallback(oracle_mod):
    """min_length below 16 (not reachable from the reference CLI): reads too short for any seed
    table fall back to scanning every library offset; results must still equal the oracle."""
    from mirge_b200.engine import Engine, LIB_IDS
    from synth.world import World
    w = World(seed=5, scale=0.002, n_expressed=30)
    rng = np.random.default_rng(8)
    reads = []
    for t in ('mirna', 'hairpin', 'mature_trna', 'snorna', 'mrna'):
        for s in w.libs[t][1][:20]:
            for _ in range(6):
                n = int(rng.integers(8, 20))
                a = int(rng.integers(0, len(s) - n + 1))
                r = s[a:a + n]
                if rng.random() < 0.3:
                    p = int(rng.integers(0, n)); r = r[:p] + 'ACGT'[('ACGT'.index(r[p]) + 1) % 4] + r[p + 1:]
                reads.append(r)
    fq = ''.join('@r%d\n%s\n+\n%s\n' % (k, r, 'I' * len(r)) for k, r in enumerate(reads)).encode()
    o = oracle_mod.Oracle(1, False, 'none')
    oracle_mod.lib().orc_set_adapter(o.h, b'none')
    # the oracle's min_length lives in its trim config: rebuild it through the C struct is not exposed, so drop short reads by hand
    e = Engine(1, False, 1 << 14, 64)
    e.set_adapter('none', min_length=8)
    for t, i in LIB_IDS.items():
        o.load_library(oracle_mod.LIB_IDS[t], *w.libs[t])
        e.load_library(i, *w.libs[t])
    e.feed_fastq(0, fq, 33)
    e.annotate()
    u = e.export_unique()
    got = {s.decode(): (int(r), int(rf), int(p), int(m)) for s, r, rf, p, m in zip(u['seq'], u['round_of'], u['ref_of'], u['pos_of'], u['mm_of'])}
    assert min(len(s) for s in got) == 8
    # oracle, read by read, through the round table (runAnnotationPipeline.py:574-599)
    rounds = [(0, -26, 0, 0, 0, 0, 0, 0), (1, 25, 0, 1, 0, 0, 0, 0), (2, 0, 1, 1, 0, 0, 1, 0), (3, 0, 1, 0, 0, 0, 1, 1), (4, 0, 0, 1, 0, 0, 0, 0),
              (5, 0, 0, 1, 0, 0, 0, 0), (6, 0, 0, 1, 0, 0, 0, 0), (7, 0, 0, 0, 0, 0, 0, 0), (0, 0, 1, 2, 1, 2, 0, 0)]
    for s, g in got.items():
        want = (0, 0, 0, 0)
        for ri, (lib, lf, mode, param, t5, t3, hi, polyt) in enumerate(rounds):
            if lf < 0 and not len(s) < -lf:
                continue
            if lf > 0 and not len(s) > lf:
                continue
            q = s
            if polyt:
                t = len(s) - len(s.rstrip('T'))
                if t < 3 or len(s) - t < 11:
                    continue
                q = s[:len(s) - t]
            else:
                q = s[t5:len(s) - t3]
                if len(q) <= 0:
                    continue
            hit = o.map_one(lib, q, mode, param, bool(hi), brute=True)
            if hit:
                want = (ri + 1,) + hit
                break
        assert g == want, (s, g, want)
