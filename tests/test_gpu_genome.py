"""Genome passes on the GPU (csrc/k2_genome.cuh) against the oracle's exhaustive scan (orc_map_genome_all): both strands,
all alignments per mismatch stratum, -3 trimming, -l seed length, --norc, the A-to-I unique-best-hit filter and the -m 3
suppression -- for the three bowtie command lines of the reference (writeDataToCSV.py:1260, :1488, __main__.py:538)."""
import numpy as np
import pytest

from conftest import ILLUMINA

pytestmark = pytest.mark.gpu
GENOME_SLOT = 8


def _rc(s):
    return s[::-1].translate(str.maketrans('ACGTN', 'TGCAN'))


def _genome(rng, n_chrom=4, size=30000):
    """Random chromosomes with planted repeats (exact and 1-2 mismatch copies, both strands) and runs of N."""
    chroms = []
    units = [''.join('ACGT'[x] for x in rng.integers(0, 4, 40)) for _ in range(12)]
    for c in range(n_chrom):
        s = list(''.join('ACGT'[x] for x in rng.integers(0, 4, size + int(rng.integers(0, 999)))))
        for _ in range(60):
            u = list(units[int(rng.integers(0, len(units)))])
            for _ in range(int(rng.integers(0, 3))):
                p = int(rng.integers(0, len(u)))
                u[p] = 'ACGT'[('ACGT'.index(u[p]) + 1 + int(rng.integers(0, 3))) % 4]
            u = ''.join(u)
            if rng.random() < 0.5:
                u = _rc(u)
            p = int(rng.integers(0, len(s) - 40))
            s[p:p + 40] = list(u)
        for _ in range(3):
            p = int(rng.integers(0, len(s) - 30))
            s[p:p + int(rng.integers(1, 25))] = 'N' * int(rng.integers(1, 25))
        chroms.append(''.join(s))
    return chroms, units


def _reads(rng, chroms, units, n):
    out = []
    for _ in range(n):
        kind = rng.random()
        L = int(rng.integers(16, 41))
        if kind < 0.35:
            u = units[int(rng.integers(0, len(units)))]
            a = int(rng.integers(0, 40 - min(L, 40) + 1))
            r = u[a:a + min(L, 40)]
        elif kind < 0.85:
            c = chroms[int(rng.integers(0, len(chroms)))]
            a = int(rng.integers(0, len(c) - L))
            r = c[a:a + L]
        else:
            r = ''.join('ACGT'[x] for x in rng.integers(0, 4, L))
        r = list(r)
        for _ in range(int(rng.integers(0, 4))):
            p = int(rng.integers(0, len(r)))
            r[p] = 'ACGTN'[int(rng.integers(0, 5))]
        r = ''.join(r)
        if rng.random() < 0.5:
            r = _rc(r)
        if 'N' * len(r) == r or len(r) == 0:
            continue
        out.append(r)
    # chromosome ends and boundary-crossing windows
    out += [chroms[0][:20], chroms[0][-20:], chroms[0][-10:] + chroms[1][:10], _rc(chroms[1][:18]), chroms[2][-16:]]
    return out


def _both(O, chroms):
    from mirge_b200.engine import Engine
    o = O.Oracle(1, False, 'none')
    names = ['chr%d' % (i + 1) for i in range(len(chroms))]
    o.load_library(GENOME_SLOT, names, chroms)
    e = Engine(1, False, 1 << 16, 64)
    e.set_adapter('none')
    e.load_genome(names, chroms)
    return o, e


@pytest.mark.parametrize('mode,param,seedlen,trim3,norc', [
    (0, 1, 28, 2, False),       # writeDataToCSV.py:1260  -n 1 -a -3 2
    (0, 0, 28, 2, False),       # writeDataToCSV.py:1488  -n 0 -a -3 2
    (0, 0, 25, 0, False),       # __main__.py:538         -n 0 -l 25 -a
    (0, 1, 20, 0, True),        # shorter seed, forward strand only
    (1, 2, 28, 1, False),       # -v 2
    (1, 3, 28, 0, False),       # -v 3
])
def test_all_hits_per_stratum(oracle_mod, mode, param, seedlen, trim3, norc):
    rng = np.random.default_rng(100 + 10 * mode + param + seedlen)
    chroms, units = _genome(rng)
    o, e = _both(oracle_mod, chroms)
    reads = _reads(rng, chroms, units, 600)
    got = e.genome_pass(reads, mode, param, seedlen, trim3, norc)
    n_hit = 0
    for r, g in zip(reads, got):
        want = o.genome_hits(GENOME_SLOT, r, mode, param, seedlen, trim3, norc)
        assert [int(x) for x in g] == want, (r, mode, param, seedlen, trim3, norc)
        n_hit += sum(want) > 0
    assert n_hit > 100


def test_a2i_filter_and_m3_policy_on_table_ids(oracle_mod):
    """Queries taken from the unique table by id; the host policies of the two call sites on top of the counts."""
    rng = np.random.default_rng(7)
    chroms, units = _genome(rng, n_chrom=3, size=20000)
    o, e = _both(oracle_mod, chroms)
    reads = [r for r in _reads(rng, chroms, units, 400) if 'N' not in r and 16 <= len(r) <= 40]
    reads = sorted(set(reads))
    fq = ''.join('@r%d\n%s\n+\n%s\n' % (k, r, 'I' * len(r)) for k, r in enumerate(reads)).encode()
    e.feed_fastq(0, fq, 33)
    u = e.export_unique()
    seqs = [s.decode() for s in u['seq']]
    assert sorted(seqs) == reads
    counts = e.genome_pass_ids(np.arange(len(seqs), dtype=np.uint32), 0, 1, 28, 2)
    kept = {}
    for s, c in zip(seqs, counts):
        c = [int(x) for x in c]
        assert c == o.genome_hits(GENOME_SLOT, s, 0, 1, 28, 2)
        if sum(c) and next(x for x in c if x) == 1:      # lowest occupied stratum holds exactly one alignment
            kept[s] = True
    assert kept == o.a2i_genome_filter(GENOME_SLOT, seqs) and len(kept) > 20
    # -m 3: more than three alignments suppress the read
    sub = [i for i, s in enumerate(seqs) if len(s) <= 25]
    c2 = e.genome_pass_ids(np.array(sub, np.uint32), 0, 0, 25, 0)
    for i, c in zip(sub, c2):
        want = o.genome_hits(GENOME_SLOT, seqs[i], 0, 0, 25, 0)
        assert [int(x) for x in c] == want
    assert any(int(c.sum()) > 3 for c in c2) and any(int(c.sum()) == 1 for c in c2)


def test_long_reads_and_tiny_genome(oracle_mod):
    from mirge_b200.engine import Engine
    rng = np.random.default_rng(9)
    g = ''.join('ACGT'[x] for x in rng.integers(0, 4, 500))
    names, chroms = ['c1', 'c2'], [g, _rc(g[100:300]) + 'ACGTTGCA' * 10]
    o = oracle_mod.Oracle(1, False, 'none')
    o.load_library(GENOME_SLOT, names, chroms)
    e = Engine(1, False, 1 << 12, 128)
    e.set_adapter('none')
    e.load_genome(names, chroms)
    reads = [g[10:10 + n] for n in (16, 33, 64, 65, 100, 128)] + [_rc(g[200:290]), g[100:180][:40] + 'A' + g[141:180], 'ACGTTGCA' * 3, 'ACGT', 'AC']
    for args in ((0, 1, 28, 2, False), (0, 0, 25, 0, False), (1, 2, 28, 0, False)):
        got = e.genome_pass(reads, *args)
        for r, c in zip(reads, got):
            assert [int(x) for x in c] == o.genome_hits(GENOME_SLOT, r, *args), (r, args)
