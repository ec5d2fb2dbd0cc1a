"""Oracle for the genome passes (the CUDA kernel for them is a next-round row):
`bowtie <genome> -n N -a -3 2` over both strands, and the A-to-I "unique best hit" filter
(writeDataToCSV.py:1260-1281).  Known answers derived by hand."""
import numpy as np

GENOME_SLOT = 8


def _rc(s):
    return s[::-1].translate(str.maketrans('ACGT', 'TGCA'))


def _oracle(O, chroms):
    o = O.Oracle(1, False, 'none')
    o.load_library(GENOME_SLOT, ['chr%d' % (i + 1) for i in range(len(chroms))], chroms)
    return o


def test_both_strands_trim3_and_strata(oracle_mod):
    rng = np.random.default_rng(11)
    g = ''.join('ACGT'[x] for x in rng.integers(0, 4, 400))
    o = _oracle(oracle_mod, [g])
    read = g[100:122]
    assert o.genome_hits(GENOME_SLOT, read + 'GG', 0, 1, 28, 2) == [1, 0, 0, 0]            # -3 2 removes the two extra bases
    assert o.genome_hits(GENOME_SLOT, read + 'GG', 0, 1, 28, 0)[0] == 0                     # without -3 they count
    assert o.genome_hits(GENOME_SLOT, _rc(read) + 'AA', 0, 1, 28, 2) == [1, 0, 0, 0]        # reverse strand
    assert o.genome_hits(GENOME_SLOT, _rc(read) + 'AA', 0, 1, 28, 2, norc=True) == [0, 0, 0, 0]
    one = read[:7] + ('A' if read[7] != 'A' else 'C') + read[8:]
    assert o.genome_hits(GENOME_SLOT, one + 'TT', 0, 1, 28, 2) == [0, 1, 0, 0]
    assert o.genome_hits(GENOME_SLOT, one + 'TT', 0, 0, 28, 2) == [0, 0, 0, 0]              # -n 0: the seed covers the whole 22-mer
    two = one[:15] + ('A' if one[15] != 'A' else 'C') + one[16:]
    assert o.genome_hits(GENOME_SLOT, two + 'TT', 0, 1, 28, 2) == [0, 0, 0, 0]              # two seed mismatches > -n 1


def test_seed_length_and_reverse_strand_seed_side(oracle_mod):
    rng = np.random.default_rng(12)
    g = ''.join('ACGT'[x] for x in rng.integers(0, 4, 400))
    o = _oracle(oracle_mod, [g])
    read = g[50:90]                                                                          # 40 nt
    flip = lambda s, i: s[:i] + ('A' if s[i] != 'A' else 'C') + s[i + 1:]
    late = flip(read, 35)                                                                    # mismatch beyond a 25-nt seed
    early = flip(read, 3)
    assert o.genome_hits(GENOME_SLOT, late, 0, 0, 25) == [0, 1, 0, 0]
    assert o.genome_hits(GENOME_SLOT, early, 0, 0, 25) == [0, 0, 0, 0]
    # reverse strand: the read's 5' end (the seed) lies at the right end of the genomic window
    assert o.genome_hits(GENOME_SLOT, _rc(flip(read, 3)), 0, 0, 25) == [0, 1, 0, 0]         # rc: the mismatch is 3 from the read's 3' end
    assert o.genome_hits(GENOME_SLOT, _rc(flip(read, 36)), 0, 0, 25) == [0, 0, 0, 0]        # rc: 3 from the read's 5' end -> inside the seed


def test_a2i_unique_best_hit_filter(oracle_mod):
    rng = np.random.default_rng(13)
    unit = ''.join('ACGT'[x] for x in rng.integers(0, 4, 30))
    spacer = lambda n: ''.join('ACGT'[x] for x in rng.integers(0, 4, n))
    variant = unit[:10] + ('A' if unit[10] != 'A' else 'C') + unit[11:]
    g1 = spacer(50) + unit + spacer(50) + variant + spacer(50)
    g2 = spacer(40) + unit + spacer(40)
    once = spacer(24)
    g1 += once + spacer(20)
    o = _oracle(oracle_mod, [g1, g2])
    r_multi = unit[:22] + 'AA'            # exact twice (two chromosomes) + one 1-mismatch copy
    r_once = once[:22] + 'CC'
    r_var = variant[:22] + 'GG'           # exact once, 1-mismatch twice: best stratum unique -> kept
    assert o.genome_hits(GENOME_SLOT, r_multi, 0, 1, 28, 2) == [2, 1, 0, 0]
    assert o.genome_hits(GENOME_SLOT, r_var, 0, 1, 28, 2) == [1, 2, 0, 0]
    kept = o.a2i_genome_filter(GENOME_SLOT, [r_multi, r_once, r_var, 'T' * 24])
    assert kept == {r_once: True, r_var: True}
    # -m 3 of the unannotated-read pass (__main__.py:538): more than three alignments suppress the read
    assert sum(o.genome_hits(GENOME_SLOT, r_multi[:22], 0, 0, 25)) == 2
