"""bowtie 1 index reader (mirge_b200/ebwt.py) against indexes written in the documented layout: names from .1.ebwt,
sequences from .3.ebwt/.4.ebwt, ambiguous stretches, the library directory falling back to the index files when no
bowtie-inspect FASTA is present.  (UNVERIFIED against a real bowtie-build index: none exists offline.)"""
import os
import struct

import numpy as np
import pytest

from mirge_b200 import ebwt
from mirge_b200.libs import LibraryDir, LibraryError


def _rand(rng, n):
    return ''.join('ACGT'[x] for x in rng.integers(0, 4, n))


def test_round_trip_names_sequences_and_ambiguous_stretches(tmp_path):
    rng = np.random.default_rng(1)
    names = ['hsa-miR-1-5p', 'hsa-let-7a-5p MIMAT0000062 Homo sapiens', 'chrN', 'x']
    seqs = [_rand(rng, 30), _rand(rng, 22), 'NN' + _rand(rng, 13) + 'NNNNN' + _rand(rng, 7) + 'RY' + _rand(rng, 5), 'a' + _rand(rng, 1001)]
    base = str(tmp_path / 'sp_mirna_miRBase')
    ebwt.write_index(base, names, seqs)
    assert ebwt.read_names(base) == ['hsa-miR-1-5p', 'hsa-let-7a-5p', 'chrN', 'x']       # RNAME = header up to the first blank
    got = ebwt.read_sequences(base)
    want = [s.upper() for s in seqs]
    want[2] = want[2].replace('R', 'N').replace('Y', 'N')
    assert got == want
    assert ebwt.read_index(base) == (ebwt.read_names(base), got)
    # 2-bit stream: four bases per byte, first base in the low bits, stretches packed back to back
    raw = open(base + '.4.ebwt', 'rb').read()
    flat = ''.join(s for s in (''.join(c for c in w if c in 'ACGT') for w in want))
    assert len(raw) == (len(flat) + 3) // 4
    assert 'ACGT'[raw[0] & 3] == flat[0] and 'ACGT'[(raw[0] >> 6) & 3] == flat[3]
    n_rec = struct.unpack_from('<I', open(base + '.3.ebwt', 'rb').read(), 4)[0]
    assert n_rec == 1 + 1 + 3 + 1


def test_header_checks(tmp_path):
    base = str(tmp_path / 'bad')
    for ext in ('.1.ebwt', '.3.ebwt', '.4.ebwt'):
        open(base + ext, 'wb').write(b'\x00' * 64)
    with pytest.raises(ebwt.EbwtError):
        ebwt.read_names(base)
    ebwt.write_index(base, ['a'], ['ACGTACGTAC'])
    buf = bytearray(open(base + '.1.ebwt', 'rb').read())
    open(base + '.1.ebwt', 'wb').write(bytes(buf) + b'junk')
    with pytest.raises(ebwt.EbwtError):
        ebwt.read_names(base)


def test_library_dir_reads_index_files_without_fasta(tmp_path, small_world):
    root = str(tmp_path)
    small_world.write(root)
    d = os.path.join(root, 'synth', 'index.Libs')
    ld = LibraryDir(root, 'synth', 'miRBase')
    want = {t: ld.load(t) for t in ld.types}
    for t in ld.types:                                   # replace every FASTA by index files in bowtie's layout
        base = ld.index_base(t)
        ebwt.write_index(base, *want[t])
        os.remove(base + '.fa')
    ld.check()
    for t in ld.types:
        names, seqs = ld.load(t)
        assert names == want[t][0] and seqs == [s.upper() for s in want[t][1]]
    os.remove(ld.index_base('mrna') + '.3.ebwt')
    with pytest.raises(LibraryError):
        ld.load('mrna')
    assert len(os.listdir(d)) > 8
