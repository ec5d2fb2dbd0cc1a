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


def test_rank_shards_cover_the_file_once(tmp_path):
    """Multi-GPU CLI: every rank takes a record-aligned byte range (plain files) or every world-th chunk (.gz);
    together the ranks see every record exactly once -- also when quality lines start with '@'."""
    import numpy as np
    from mirge_b200.pipeline import iter_fastq_chunks_shard
    rng = np.random.default_rng(3)
    recs = []
    for i in range(3000):
        n = int(rng.integers(16, 60))
        q = ''.join(rng.choice(list('@+IF#'), n))
        recs.append('@r%d some header\n%s\n+\n%s\n' % (i, ''.join(rng.choice(list('ACGTN'), n)), q))
    text = ''.join(recs)
    p = tmp_path / 's.fastq'
    p.write_text(text)
    pz = tmp_path / 's.fastq.gz'
    with gzip.open(pz, 'wb') as f:
        f.write(text.encode())
    for path in (str(p), str(pz)):
        for world in (1, 2, 3, 8):
            seen = []
            for rank in range(world):
                for c in iter_fastq_chunks_shard(path, rank, world, chunk_bytes=20000):
                    assert c.startswith(b'@r') and c.count(b'\n') % 4 == 0
                    seen += c.decode().split('\n')[0::4][:-1] if c.endswith(b'\n') else c.decode().split('\n')[0::4]
            assert sorted(seen) == sorted(r.split('\n')[0] for r in recs), (path, world)


def test_pinned_reader_streams_whole_records(tmp_path):
    from mirge_b200.pipeline import PinnedChunkReader, record_aligned_ranges
    text = ''.join('@r%d\n%s\n+\n%s\n' % (i, 'ACGT' * (4 + i % 7), '@III' * (4 + i % 7)) for i in range(20000))
    p = tmp_path / 'r.fastq'
    p.write_text(text)
    for cb in (5000, 100000, 1 << 30):
        ranges = record_aligned_ranges(str(p), 0, len(text), cb)
        assert ranges[0][0] == 0 and ranges[-1][1] == len(text) and all(text[a] == '@' for a, _ in ranges)
        r = PinnedChunkReader(str(p), ranges, threads=3)
        got = b''.join(bytes(c) for c in r)
        r.close()
        assert got.decode() == text
