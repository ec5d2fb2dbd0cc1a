"""Reader for bowtie 1 index files (`<base>.1.ebwt`, `.3.ebwt`, `.4.ebwt`) -- reference names and sequences without
`bowtie-inspect`.

The reference resolves its libraries as bowtie indexes (src/mirge/__main__.py:262-281) and shells out to
`bowtie-inspect` for the names (utils/summarize.py:6) and the sequences (utils/runAnnotationPipeline.py:610-611).
The CUDA mapper needs exactly those two things; this module takes them from the index files directly:

  * names      tail of `.1.ebwt`: after the header and the BWT / ftab / eftab sections, the reference names as
               '\\n'-separated text closed by a NUL byte (Ebwt::readIntoMemory / szsToDisk in bowtie's ebwt.h);
  * sequences  `.3.ebwt` holds one RefRecord (u32 off = ambiguous characters skipped before the stretch, u32 len =
               unambiguous characters in it, u8 first = starts a new reference) per unambiguous stretch, `.4.ebwt`
               the stretches 2 bits per base, four bases per byte, first base in the low bits
               (BitPairReference in bowtie's reference.cpp); no BWT inversion is needed.

UNVERIFIED-OFFLINE: written from the documented layout of bowtie 1.1.x small (32-bit) indexes; no real index is
available in this environment.  `write_index` produces files in the same layout (BWT sections zero-filled: the CUDA
mapper never reads them) so that the reader is at least exercised end to end, and `read_names` cross-checks the
computed section sizes against the position of the text block at the end of the file.  A `<base>.fa` next to the
index (what `bowtie-inspect <base>` prints) always takes precedence (mirge_b200/libs.py).
"""
import os
import struct

import numpy as np


class EbwtError(Exception):
    pass


def _u32s(buf, off, n, be):
    return struct.unpack_from(('>' if be else '<') + '%dI' % n, buf, off), off + 4 * n


def _sizes(length, line_rate, lines_per_side, off_rate, ftab_chars):
    """Section sizes that follow from the header (EbwtParams::init)."""
    bwt_sz = length // 4 + 1
    side_sz = (1 << line_rate) * lines_per_side
    side_bwt_sz = side_sz - 8
    num_side_pairs = (bwt_sz + 2 * side_bwt_sz - 1) // (2 * side_bwt_sz)
    return dict(ebwt_tot=num_side_pairs * 2 * side_sz, ftab_len=(1 << (2 * ftab_chars)) + 1, eftab_len=2 * ftab_chars,
                offs_len=((length + 1) + (1 << off_rate) - 1) >> off_rate)


def read_header(buf):
    if len(buf) < 28:
        raise EbwtError('not a bowtie index: file shorter than its header')
    one = struct.unpack_from('<i', buf, 0)[0]
    be = one != 1
    if be and struct.unpack_from('>i', buf, 0)[0] != 1:
        raise EbwtError('not a bowtie 1 index (endianness marker missing)')
    f = '>' if be else '<'
    length, line_rate, lines_per_side, off_rate, ftab_chars, flags = struct.unpack_from(f + 'Iiiiii', buf, 4)
    if not (0 < line_rate < 16 and 0 < lines_per_side < 64 and 0 <= off_rate < 32 and 0 < ftab_chars < 16):
        raise EbwtError('implausible bowtie index header (large / bowtie2 index?)')
    return dict(be=be, len=length, line_rate=line_rate, lines_per_side=lines_per_side, off_rate=off_rate,
                ftab_chars=ftab_chars, flags=flags)


def read_names(base):
    """Reference names of `<base>.1.ebwt` up to the first whitespace (what bowtie reports as RNAME)."""
    with open(base + '.1.ebwt', 'rb') as f:
        buf = f.read()
    h = read_header(buf)
    be = h['be']
    off = 28
    (n_pat,), off = _u32s(buf, off, 1, be)
    off += 4 * n_pat                                  # plen[]
    (n_frag,), off = _u32s(buf, off, 1, be)
    off += 12 * n_frag                                # rstarts[]: (joined offset, sequence, offset in sequence)
    sz = _sizes(h['len'], h['line_rate'], h['lines_per_side'], h['off_rate'], h['ftab_chars'])
    off += sz['ebwt_tot'] + 4 + 20 + 4 * sz['ftab_len'] + 4 * sz['eftab_len']     # ebwt[], zOff, fchr[5], ftab[], eftab[]
    end = buf.find(b'\0', off)
    if off > len(buf) or end < 0:
        raise EbwtError('%s.1.ebwt: the name block is not where the header says (not a bowtie 1.1.x small index?)' % base)
    if end != len(buf) - 1:
        raise EbwtError('%s.1.ebwt: bytes follow the name block' % base)
    text = buf[off:end].decode('ascii', 'replace')
    names = text.split('\n')
    if names and names[-1] == '':
        names.pop()
    return [n.split()[0] if n.split() else '' for n in names]


_UNPACK = np.array([[ord('ACGT'[(b >> (2 * k)) & 3]) for k in range(4)] for b in range(256)], dtype=np.uint8)


def read_sequences(base):
    """Reference sequences of `<base>.3.ebwt` / `.4.ebwt`, ambiguous stretches as N (trailing ones are not stored)."""
    with open(base + '.3.ebwt', 'rb') as f:
        rec = f.read()
    if len(rec) < 8:
        raise EbwtError('%s.3.ebwt is too short' % base)
    be = struct.unpack_from('<i', rec, 0)[0] != 1
    f4 = '>' if be else '<'
    (n_rec,) = struct.unpack_from(f4 + 'I', rec, 4)
    if 8 + 9 * n_rec > len(rec):
        raise EbwtError('%s.3.ebwt: %d records do not fit the file' % (base, n_rec))
    packed = np.fromfile(base + '.4.ebwt', dtype=np.uint8)
    bases = _UNPACK[packed].reshape(-1)
    seqs, cur, pos = [], None, 0
    for k in range(n_rec):
        off, ln, first = struct.unpack_from(f4 + 'IIB', rec, 8 + 9 * k)
        if first:
            if cur is not None:
                seqs.append(''.join(cur))
            cur = []
        if cur is None:
            raise EbwtError('%s.3.ebwt: the first record does not start a reference' % base)
        if pos + ln > len(bases):
            raise EbwtError('%s.4.ebwt holds fewer bases than .3.ebwt describes' % base)
        cur.append('N' * off)
        cur.append(bases[pos:pos + ln].tobytes().decode('ascii'))
        pos += ln
    if cur is not None:
        seqs.append(''.join(cur))
    return seqs


def read_index(base):
    """(names, sequences) of a bowtie 1 index; references without any unambiguous base carry no records and are skipped
    by bowtie itself, so the two lists must agree in length."""
    names, seqs = read_names(base), read_sequences(base)
    if len(names) != len(seqs):
        raise EbwtError('%s: %d names but %d sequences' % (base, len(names), len(seqs)))
    return names, seqs


def write_index(base, names, seqs, line_rate=6, lines_per_side=1, off_rate=5, ftab_chars=10):
    """Files in the bowtie 1.1.x small-index layout for `names` / `seqs` (test generator; the BWT, ftab and suffix-array
    sections are zero-filled -- nothing on the CUDA path reads them, bowtie itself could not search this index)."""
    stretches, frags, joined = [], [], 0
    packed = bytearray()
    acc, nacc = 0, 0
    for sid, s in enumerate(seqs):
        s = s.upper()
        first, i = True, 0
        while i < len(s):
            j = i
            while j < len(s) and s[j] not in 'ACGT':
                j += 1
            k = j
            while k < len(s) and s[k] in 'ACGT':
                k += 1
            if k > j:
                stretches.append((j - i, k - j, 1 if first else 0))
                frags.append((joined, sid, j))
                joined += k - j
                first = False
                for ch in s[j:k]:
                    acc |= 'ACGT'.index(ch) << (2 * nacc)
                    nacc += 1
                    if nacc == 4:
                        packed.append(acc)
                        acc, nacc = 0, 0
            i = k
    if nacc:
        packed.append(acc)
    length = joined
    sz = _sizes(length, line_rate, lines_per_side, off_rate, ftab_chars)
    with open(base + '.1.ebwt', 'wb') as f:
        f.write(struct.pack('<iIiiiii', 1, length, line_rate, lines_per_side, off_rate, ftab_chars, 0))
        f.write(struct.pack('<I', len(seqs)))
        f.write(struct.pack('<%dI' % len(seqs), *[sum(1 for c in s.upper() if c in 'ACGT') for s in seqs]))
        f.write(struct.pack('<I', len(frags)))
        for fr in frags:
            f.write(struct.pack('<III', *fr))
        f.write(b'\0' * sz['ebwt_tot'])
        f.write(struct.pack('<I', 0))
        f.write(struct.pack('<5I', 0, 0, 0, 0, 0))
        f.write(b'\0' * (4 * sz['ftab_len'] + 4 * sz['eftab_len']))
        f.write(('\n'.join(names) + '\n').encode('ascii') + b'\0')
    with open(base + '.2.ebwt', 'wb') as f:
        f.write(struct.pack('<i', 1) + b'\0' * (4 * sz['offs_len']))
    with open(base + '.3.ebwt', 'wb') as f:
        f.write(struct.pack('<iI', 1, len(stretches)))
        for off, ln, first in stretches:
            f.write(struct.pack('<IIB', off, ln, first))
    with open(base + '.4.ebwt', 'wb') as f:
        f.write(bytes(packed))
    for rev in ('.rev.1.ebwt', '.rev.2.ebwt'):
        open(base + rev, 'wb').close()
    return base


def has_sequences(base):
    return all(os.path.isfile(base + e) and os.path.getsize(base + e) > 8 for e in ('.1.ebwt', '.3.ebwt')) and os.path.isfile(base + '.4.ebwt')
