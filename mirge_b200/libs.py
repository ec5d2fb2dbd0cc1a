"""Library directory of `miRge2.0 annotate` (src/mirge/__main__.py:110-112,262-281).

    <lib>/<species>/index.Libs/<sp>_{mirna_<DB>,hairpin_<DB>,mrna,mature_trna,pre_trna,
                                    snorna,rrna,ncrna_others,spike-in}.*.ebwt
    <lib>/<species>/fasta.Libs/<sp>_mirna_SNP_pseudo_<DB>.fa
    <lib>/<species>/annotation.Libs/<sp>_merges_<DB>.csv
    <lib>/<species>/annotation.Libs/<sp>_<DB>.gff3                  (-gff only)

The reference hands the `.ebwt` indexes to bowtie.  The CUDA mapper needs the
reference names and sequences themselves (the reference gets them from
bowtie-inspect at utils/summarize.py:6 and utils/runAnnotationPipeline.py:610-611):
they are read straight from the index files (`.1.ebwt` names, `.3.ebwt`/`.4.ebwt`
2-bit reference; mirge_b200/ebwt.py), unless `<index>.fa` -- what
`bowtie-inspect <index> > <index>.fa` prints -- lies next to the index, which
takes precedence.  The presence check on `.1.ebwt` (:262-267) and its message
are kept.
"""
import os

from . import ebwt

LIB_TYPES = {            # library slot -> index basename suffix
    'mirna': 'mirna_%(db)s', 'hairpin': 'hairpin_%(db)s', 'mrna': 'mrna', 'mature_trna': 'mature_trna',
    'pre_trna': 'pre_trna', 'snorna': 'snorna', 'rrna': 'rrna', 'ncrna_others': 'ncrna_others', 'spike-in': 'spike-in',
}
REQUIRED = ['mirna', 'hairpin', 'mrna', 'mature_trna', 'pre_trna', 'snorna', 'rrna', 'ncrna_others']   # :262


class LibraryError(Exception):
    pass


def read_fasta(path):
    """FASTA -> (names, seqs); a name is the header up to the first whitespace
    (what bowtie reports as RNAME and bowtie-inspect -n lists)."""
    names, seqs, cur = [], [], []
    with open(path) as f:
        for line in f:
            if line.startswith('>'):
                if names:
                    seqs.append(''.join(cur))
                hdr = line[1:].strip()
                names.append(hdr.split()[0] if hdr else '')
                cur = []
            else:
                cur.append(line.strip())
    if names:
        seqs.append(''.join(cur))
    return names, seqs


class LibraryDir:
    def __init__(self, library_path, species, db, spike_in=False):
        self.species, self.db = species, db
        self.index_path = os.path.join(library_path, species, 'index.Libs')
        self.mirna_fa = os.path.join(library_path, species, 'fasta.Libs', '%s_mirna_SNP_pseudo_%s.fa' % (species, db))
        self.merges_csv = os.path.join(library_path, species, 'annotation.Libs', '%s_merges_%s.csv' % (species, db))
        self.gff3 = os.path.join(library_path, species, 'annotation.Libs', '%s_%s.gff3' % (species, db))     # __main__.py:150
        self.repetitive_csv = os.path.join(library_path, species, 'annotation.Libs',
                                           '%s_miRNAs_in_repetitive_element_%s.csv' % (species, db))          # __main__.py:137
        self.genome_base = os.path.join(self.index_path, '%s_genome' % species)                               # __main__.py:270
        self.types = list(REQUIRED) + (['spike-in'] if spike_in else [])

    def index_base(self, t):
        return os.path.join(self.index_path, '%s_%s' % (self.species, LIB_TYPES[t] % {'db': self.db}))

    def check(self):
        """The reference's index presence check (__main__.py:262-267), same message."""
        for t in REQUIRED:
            if not os.path.isfile(self.index_base(t) + '.1.ebwt'):
                raise LibraryError('The bowtie index file of %s_%s.*.ebwt is not located at %s, please check it.'
                                   % (self.species, LIB_TYPES[t] % {'db': self.db}, self.index_path))

    @staticmethod
    def _load_base(base, what):
        if os.path.isfile(base + '.fa'):
            return read_fasta(base + '.fa')
        if ebwt.has_sequences(base):
            try:
                return ebwt.read_index(base)
            except ebwt.EbwtError as e:
                raise LibraryError('%s; `bowtie-inspect %s > %s.fa` next to the index is read instead when present.' % (e, base, base))
        raise LibraryError('%s: neither %s.fa nor readable %s.{1,3,4}.ebwt files were found; create the FASTA with '
                           '`bowtie-inspect %s > %s.fa`.' % (what, base, base, base, base))

    def load(self, t):
        return self._load_base(self.index_base(t), 'library %s' % t)

    def has_genome(self):
        return os.path.isfile(self.genome_base + '.fa') or ebwt.has_sequences(self.genome_base)

    def load_genome(self):
        """Chromosomes of index.Libs/<sp>_genome (the -ai genome passes, writeDataToCSV.py:1260,1488)."""
        return self._load_base(self.genome_base, 'the genome (-ai)')

    def mature_sequences(self):
        """mirNameSeqDic: name -> sequence from fasta.Libs/<sp>_mirna_SNP_pseudo_<DB>.fa (miRNAmerge.py:4-11)."""
        out = {}
        with open(self.mirna_fa) as f:
            line = f.readline()
            while line != '':
                name = line.strip()[1:]
                out[name] = f.readline().strip()
                line = f.readline()
        return out

    def merged_names(self):
        """mirMergedNameDic: member name -> merged name (__main__.py:114-120)."""
        out = {}
        for line in self.merge_lines() or []:
            c = line.strip().split(',')
            for item in c[1:]:
                out[item] = c[0]
        return out

    def repetitive_mirnas(self):
        """removedMiRNA_ai_List (__main__.py:138-144); a missing file means none."""
        out = []
        if os.path.isfile(self.repetitive_csv):
            with open(self.repetitive_csv) as f:
                for line in f:
                    k = line.strip().split(',')[0]
                    if k not in out:
                        out.append(k)
        return out

    def merge_lines(self):
        """Rows of <sp>_merges_<DB>.csv; a missing file skips the merge (miRNAmerge.py:13,42-43)."""
        if not os.path.isfile(self.merges_csv):
            return None
        with open(self.merges_csv) as f:
            return [line for line in f if line.strip()]
