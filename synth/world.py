"""Seeded synthetic world for the annotate-mode pipeline.

Builds a miRge-style library directory
    <root>/<species>/index.Libs/<sp>_<type>.fa   (+ empty <sp>_<type>.1.ebwt marker)
    <root>/<species>/fasta.Libs/<sp>_mirna_SNP_pseudo_<DB>.fa
    <root>/<species>/annotation.Libs/<sp>_merges_<DB>.csv
and synthetic small-RNA FASTQ text carrying the Illumina small-RNA 3' adapter.
bowtie-build is not available offline, so each index gets a sibling 2-line
FASTA that the loader reads (reference layout: src/mirge/__main__.py:110-112,
262-281).  Workload generator only: neither product nor oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
BASES = np.frombuffer(b'ACGT', dtype=np.uint8)

# library type -> (entries, mean length) at scale 1.0 ("human-like")
FULL_SIZES = {
    'hairpin': (1900, 85), 'mature_trna': (430, 75), 'pre_trna': (600, 100), 'snorna': (1000, 130),
    'rrna': (20, 1000), 'ncrna_others': (10000, 300), 'mrna': (20000, 2000), 'spike-in': (50, 22),
}


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libmirge_synth.so')
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.synth_record_bytes.restype = C.c_uint64
        _LIB.synth_record_bytes.argtypes = [C.c_int]
    return _LIB


def _rand_seqs(rng, n, mean_len, jitter=0.2):
    lens = np.maximum(16, (mean_len * (1 + jitter * (rng.random(n) * 2 - 1))).astype(np.int64))
    cat = BASES[rng.integers(0, 4, int(lens.sum()))]
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum(lens)
    return [cat[off[i]:off[i + 1]].tobytes().decode() for i in range(n)]


class World:
    """All libraries as {type: (names, seqs)} plus the expressed-miRNA table."""

    def __init__(self, seed=20180626, scale=1.0, species='synth', db='miRBase', n_expressed=400, snpc_names=False):
        # snpc_names: a miRNA that has a SNP entry is itself named <name>.SNPC, as the reference's -gff code expects
        # (runAnnotationPipeline.py:417-419); off by default so that fixtures and benchmarks keep their names
        rng = np.random.default_rng(np.random.PCG64(seed))
        self.species, self.db, self.seed = species, db, seed
        self.libs = {}
        sz = {k: (max(4, int(round(v[0] * scale))), v[1]) for k, v in FULL_SIZES.items()}
        sz['rrna'] = (max(4, FULL_SIZES['rrna'][0] if scale >= 0.2 else 6), FULL_SIZES['rrna'][1])
        for t in ('mature_trna', 'pre_trna', 'snorna', 'rrna', 'ncrna_others', 'mrna'):
            n, ml = sz[t]
            self.libs[t] = ([f'{species}-{t}-{i + 1}' for i in range(n)], _rand_seqs(rng, n, ml))
        # mature tRNAs carry the CCA tail
        self.libs['mature_trna'] = (self.libs['mature_trna'][0], [s[:-3] + 'CCA' for s in self.libs['mature_trna'][1]])
        n, ml = sz['spike-in']
        self.libs['spike-in'] = ([f'spike-{i + 1}' for i in range(n)], _rand_seqs(rng, n, ml, 0.05))
        # hairpins and the matures cut from them
        nh, _ = sz['hairpin']
        hp_seqs = _rand_seqs(rng, nh, 85, 0.15)
        hp_names = [f'{species}-mir-{i + 1}' for i in range(nh)]
        mir_names, mir_seqs, mature_seqs, merges = [], [], [], []
        self.mir_table = []   # (name, hairpin idx, mature start, mature len)
        self.derives = []     # (canonical miRNA name, stem-loop name) for <sp>_<DB>.gff3
        for i in range(nh):
            h = hp_seqs[i]
            arms = ['5p']
            if rng.random() < 0.42:
                arms.append('3p')
            for arm in arms:
                ln = int(rng.integers(20, 25))
                if arm == '5p':
                    st = int(rng.integers(8, 15))
                else:
                    st = len(h) - int(rng.integers(8, 15)) - ln
                name = f'{species}-miR-{i + 1}-{arm}'
                entry = h[st - 2:st + ln + 6]
                mir_names.append(name)
                mir_seqs.append(entry)
                mature_seqs.append(h[st:st + ln])
                self.mir_table.append((name, i, st, ln))
                self.derives.append((name, hp_names[i]))
                r = rng.random()
                if r < 0.10:      # SNP variant of this entry (one substitution inside the mature)
                    p = 2 + int(rng.integers(1, ln - 1))
                    alt = 'ACGT'.replace(entry[p], '')[int(rng.integers(0, 3))]
                    sname = name + '.SNP1'
                    cname = name
                    if snpc_names:
                        cname = name + '.SNPC'
                        mir_names[-1] = cname
                    mir_names.append(sname)
                    mir_seqs.append(entry[:p] + alt + entry[p + 1:])
                    mature_seqs.append(mir_seqs[-1][2:2 + ln])
                    merges.append(f'{name}/SNP1,{cname},{sname}')
                elif r < 0.13:    # paralog: same mature, different flanks, on its own hairpin
                    left = ''.join('ACGT'[x] for x in rng.integers(0, 4, st))
                    right = ''.join('ACGT'[x] for x in rng.integers(0, 4, len(h) - st - ln))
                    hp_names.append(f'{species}-mir-{i + 1}-2')
                    hp_seqs.append(left + h[st:st + ln] + right)
                    pname = f'{species}-miR-{i + 1}b-{arm}'
                    mir_names.append(pname)
                    mir_seqs.append(hp_seqs[-1][st - 2:st + ln + 6])
                    mature_seqs.append(h[st:st + ln])
                    self.derives.append((pname, hp_names[-1]))
                    merges.append(f'{species}-miR-{i + 1}/{i + 1}b-{arm},{name},{pname}')
        self.libs['hairpin'] = (hp_names, hp_seqs)
        self.libs['mirna'] = (mir_names, mir_seqs)
        self.mature_seqs = mature_seqs
        self.merges = merges
        # expressed miRNAs with Zipf(1.2) abundance
        ne = min(n_expressed, len(self.mir_table))
        pick = rng.choice(len(self.mir_table), ne, replace=False)
        wts = 1.0 / np.arange(1, ne + 1) ** 1.2
        self.expr = [self.mir_table[int(k)] for k in pick]
        self.expr_cdf = np.cumsum(wts / wts.sum())
        self.expr_cdf[-1] = 1.0
        self._cworld = None

    # -- on-disk layout ----------------------------------------------------------
    def write(self, root):
        sp, db = self.species, self.db
        d = os.path.join(root, sp)
        for sub in ('index.Libs', 'fasta.Libs', 'annotation.Libs'):
            os.makedirs(os.path.join(d, sub), exist_ok=True)
        fn = {'mirna': f'{sp}_mirna_{db}', 'hairpin': f'{sp}_hairpin_{db}', 'mrna': f'{sp}_mrna',
              'mature_trna': f'{sp}_mature_trna', 'pre_trna': f'{sp}_pre_trna', 'snorna': f'{sp}_snorna',
              'rrna': f'{sp}_rrna', 'ncrna_others': f'{sp}_ncrna_others', 'spike-in': f'{sp}_spike-in'}
        for t, base in fn.items():
            names, seqs = self.libs[t]
            with open(os.path.join(d, 'index.Libs', base + '.fa'), 'w') as f:
                f.write(''.join(f'>{n}\n{s}\n' for n, s in zip(names, seqs)))
            open(os.path.join(d, 'index.Libs', base + '.1.ebwt'), 'w').close()
        with open(os.path.join(d, 'fasta.Libs', f'{sp}_mirna_SNP_pseudo_{db}.fa'), 'w') as f:
            f.write(''.join(f'>{n}\n{s}\n' for n, s in zip(self.libs['mirna'][0], self.mature_seqs)))
            mature_of = dict(zip(self.libs['mirna'][0], self.mature_seqs))
            for m in self.merges:                      # merged names carry the sequence of their first member
                c = m.split(',')
                f.write(f'>{c[0]}\n{mature_of[c[1]]}\n')
        with open(os.path.join(d, 'annotation.Libs', f'{sp}_merges_{db}.csv'), 'w') as f:
            f.write(''.join(m + '\n' for m in self.merges))
        with open(os.path.join(d, 'annotation.Libs', f'{sp}_{db}.gff3'), 'w') as f:
            f.write(self.gff3_text())
        return root

    def genome(self, total=400_000, n_chrom=4, seed=77):
        """Synthetic genome for the genome passes (library slot 8): n_chrom random chromosomes of `total` bases in all
        with every hairpin planted once (its miRNA reads then map to one locus) and a few hairpins planted twice on
        different chromosomes (multi-mappers the A-to-I filter must drop)."""
        rng = np.random.Generator(np.random.PCG64(seed))
        hp = self.libs['hairpin'][1]
        per = max(total // n_chrom, 4 * sum(len(h) for h in hp) // n_chrom + 1000)
        lut = np.frombuffer(b'ACGT', dtype=np.uint8)
        chroms = [lut[rng.integers(0, 4, per, dtype=np.uint8)].copy() for _ in range(n_chrom)]
        slots = [per // (len(hp) // n_chrom + 2)] * n_chrom
        cursor = [200] * n_chrom
        twice = set(int(x) for x in rng.choice(len(hp), max(1, len(hp) // 50), replace=False))
        for k, h in enumerate(hp):
            for c in ([k % n_chrom, (k + 1) % n_chrom] if k in twice else [k % n_chrom]):
                a = cursor[c] + int(rng.integers(0, max(1, slots[c] - len(h) - 100)))
                if a + len(h) + 100 < per:
                    chroms[c][a:a + len(h)] = np.frombuffer(h.encode(), dtype=np.uint8)
                cursor[c] = min(per - 200, cursor[c] + slots[c])
        return ['chr%d' % (c + 1) for c in range(n_chrom)], [c.tobytes().decode() for c in chroms]

    def write_genome(self, root, total=400_000, n_chrom=4, seed=77):
        """index.Libs/<sp>_genome.fa (+ the .1.ebwt presence marker) and an empty repetitive-element list."""
        sp, db = self.species, self.db
        d = os.path.join(root, sp)
        names, seqs = self.genome(total, n_chrom, seed)
        with open(os.path.join(d, 'index.Libs', f'{sp}_genome.fa'), 'w') as f:
            for n, s_ in zip(names, seqs):
                f.write(f'>{n}\n{s_}\n')
        open(os.path.join(d, 'index.Libs', f'{sp}_genome.1.ebwt'), 'w').close()
        with open(os.path.join(d, 'annotation.Libs', f'{sp}_miRNAs_in_repetitive_element_{db}.csv'), 'w') as f:
            f.write('')
        return names, seqs

    def gff3_text(self):
        """<sp>_<DB>.gff3 in miRBase layout: one miRNA_primary_transcript per hairpin, one miRNA per mature."""
        hp = self.libs['hairpin'][0]
        alias = {n: 'MI%07d' % (k + 1) for k, n in enumerate(hp)}
        out = ['##gff-version 3\n']
        for k, n in enumerate(hp):
            out.append('chr1\t.\tmiRNA_primary_transcript\t%d\t%d\t.\t+\t.\tID=%s;Alias=%s;Name=%s\n'
                       % (1000 * k + 1, 1000 * k + len(self.libs['hairpin'][1][k]), alias[n], alias[n], n))
        for k, (name, loop) in enumerate(self.derives):
            out.append('chr1\t.\tmiRNA\t%d\t%d\t.\t+\t.\tID=MIMAT%07d;Alias=MIMAT%07d;Name=%s;Derives_from=%s\n'
                       % (k + 1, k + 22, k + 1, k + 1, name, alias[loop]))
        return ''.join(out)

    # -- FASTQ ------------------------------------------------------------------
    def _c_world(self):
        if self._cworld is not None:
            return self._cworld

        class SynLib(C.Structure):
            _fields_ = [('seq', C.c_char_p), ('off', C.c_void_p), ('n', C.c_int64)]

        class SynWorld(C.Structure):
            _fields_ = [(k, SynLib) for k in ('hairpin', 'trna', 'pretrna', 'snorna', 'rrna', 'ncrna', 'mrna', 'spike')] + \
                       [('n_expr', C.c_int64), ('expr_hp', C.c_void_p), ('expr_start', C.c_void_p),
                        ('expr_len', C.c_void_p), ('expr_cdf', C.c_void_p)]

        keep = []
        w = SynWorld()
        for field, t in (('hairpin', 'hairpin'), ('trna', 'mature_trna'), ('pretrna', 'pre_trna'), ('snorna', 'snorna'),
                         ('rrna', 'rrna'), ('ncrna', 'ncrna_others'), ('mrna', 'mrna'), ('spike', 'spike-in')):
            seqs = self.libs[t][1]
            cat = ''.join(seqs).encode()
            off = np.zeros(len(seqs) + 1, np.uint64)
            off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
            keep += [cat, off]
            setattr(w, field, SynLib(cat, off.ctypes.data, len(seqs)))
        hp = np.array([e[1] for e in self.expr], np.int32)
        st = np.array([e[2] for e in self.expr], np.int32)
        ln = np.array([e[3] for e in self.expr], np.int32)
        cdf = np.ascontiguousarray(self.expr_cdf, np.float64)
        keep += [hp, st, ln, cdf]
        w.n_expr = len(self.expr)
        w.expr_hp, w.expr_start, w.expr_len, w.expr_cdf = hp.ctypes.data, st.ctypes.data, ln.ctypes.data, cdf.ctypes.data
        self._cworld = (w, keep)
        return self._cworld

    def fastq(self, n_reads, seed=1000, L=50, first_read=0, out=None, threads=0, index6=b'ACAGTG'):
        """Return FASTQ text for n_reads reads as a uint8 numpy array (or fill `out`)."""
        w, _ = self._c_world()
        rec = int(_lib().synth_record_bytes(L))
        if out is None:
            out = np.empty(n_reads * rec, np.uint8)
        assert out.nbytes >= n_reads * rec
        _lib().synth_fastq(C.byref(w), C.c_uint64(seed), C.c_uint64(first_read), C.c_uint64(n_reads), C.c_int(L),
                           C.c_char_p(index6), C.c_void_p(out.ctypes.data), C.c_int(threads))
        return out[:n_reads * rec]

    @staticmethod
    def record_bytes(L=50):
        return int(_lib().synth_record_bytes(L))
