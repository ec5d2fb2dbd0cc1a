import gzip

from mirge_b200.pipeline import iter_fastq_chunks, peek_max_read_len


def _fq(n, tail_newline=True):
    s = ''.join('@r%d\n%s\n+\n%s\n' % (i, 'ACGT' * (3 + i % 5), 'IIII' * (3 + i % 5)) for i in range(n))
    return s if tail_newline else s[:-1]


def test_chunks_are_record_aligned(tmp_path):
    for tail in (True, False):
        text = _fq(500, tail)
        p = tmp_path / ('a%d.fastq' % tail)
        p.write_text(text)
        for cb in (64, 257, 1000, 1 << 20):
            chunks = list(iter_fastq_chunks(str(p), cb))
            assert b''.join(chunks).decode() == text
            for c in chunks[:-1]:
                assert c.endswith(b'\n') and c.count(b'\n') % 4 == 0 and c.startswith(b'@')
    pz = tmp_path / 'b.fastq.gz'
    with gzip.open(pz, 'wb') as f:
        f.write(_fq(300).encode())
    assert b''.join(iter_fastq_chunks(str(pz), 500)).decode() == _fq(300)
    assert peek_max_read_len([str(pz)]) == 28
