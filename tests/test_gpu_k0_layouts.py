"""K0 (single-pass FASTQ parse + pack, csrc/k0_parse.cuh) against the oracle on text layouts that stress the tiling:
records straddling 16 KiB tiles and feed regions, header lines longer than the 1 KiB halo (global-memory path), reads
of a few bases (more newlines per tile than one pass of the line list holds), CRLF, a missing final newline, chunk
lengths around multiples of the tile size, several chunks per sample.  Every layout runs through the speculative kernel
(k0_parse_spec + k0_verify, falling back to the exact kernel when a tile cannot prove its line phase) and through the
exact chained-scan kernel alone."""
import numpy as np
import pytest

from conftest import ILLUMINA

pytestmark = pytest.mark.gpu


def _collapse(O, fq_chunks, adapter=ILLUMINA, max_read_len=64, device_resident=False, exact_too=True):
    """Runs the default path (speculate + verify, exact fallback when needed) and, with exact_too, the exact chained
    path alone (mirge_debug bit 0x2000)."""
    got = _collapse_mode(O, fq_chunks, adapter, max_read_len, device_resident, 0)
    if exact_too:
        assert _collapse_mode(O, fq_chunks, adapter, max_read_len, device_resident, 0x2000) == got
    return got


def _collapse_mode(O, fq_chunks, adapter, max_read_len, device_resident, mode):
    from mirge_b200.engine import Engine
    o = O.Oracle(1, False, adapter)
    e = Engine(1, False, 1 << 18, max_read_len)
    e.set_adapter(adapter)
    e.set_trim_mode(mode)
    keep = []
    for fq in fq_chunks:
        if len(fq):
            o.feed_fastq(0, fq, 33)
        if device_resident and len(fq):
            import torch
            t = torch.frombuffer(bytearray(fq), dtype=torch.uint8).cuda()
            keep.append(t)
            e.feed_fastq_device(0, t.data_ptr(), len(fq), 33)
        else:
            e.feed_fastq(0, fq, 33)
    e.sync()
    oe, ee = o.export(), e.export_unique()
    want = {}
    for s, c in zip(oe['seqs'], oe['counts']):       # the packed form turns every non-ACGT byte into N
        key = ''.join(ch if ch in 'ACGT' else 'N' for ch in s.upper())
        want[key] = want.get(key, 0) + int(c[0])
    got = {s.decode(): int(c[0]) for s, c in zip(ee['seq'], ee['counts'])}
    assert got == want
    op, ok = o.stats()
    ep, ek, _ = e.stats()
    assert (int(ep[0]), int(ek[0])) == (int(op[0]), int(ok[0]))
    assert np.array_equal(e.export_tables_prelim()['len_hist'][:, 0], o.len_hist(max_read_len)[:, 0])
    return got


def _reads(rng, n, lo=0, hi=50):
    return [''.join('ACGTN'[x] for x in rng.choice(5, int(rng.integers(lo, hi + 1)), p=[.245, .245, .245, .245, .02])) for _ in range(n)]


def _fastq(reads, header=lambda k: 'r%d' % k, eol='\n', final_eol=True, qual=lambda r, k: 'I' * len(r)):
    s = ''.join('@%s%s%s%s+%s%s%s' % (header(k), eol, r, eol, eol, qual(r, k), eol) for k, r in enumerate(reads))
    if not final_eol:
        s = s[:-len(eol)]
    return s.encode()


def test_tile_and_region_boundaries(oracle_mod):
    rng = np.random.default_rng(11)
    reads = _reads(rng, 6000, 16, 50)
    fq = _fastq(reads)
    # every chunk length near a multiple of the 16 KiB tile: cut at record boundaries closest to k * 16384 +- a few records
    starts = [0]
    pos = 0
    for k, r in enumerate(reads):
        pos += len('@r%d\n' % k) + 2 * len(r) + 4
        starts.append(pos)
    starts = np.array(starts)
    for target in (16384, 32768, 5 * 16384):
        i = int(np.searchsorted(starts, target))
        for j in (i - 1, i, i + 1):
            _collapse(oracle_mod, [fq[:starts[j]]])
    _collapse(oracle_mod, [fq])
    _collapse(oracle_mod, [fq], device_resident=True)
    # several chunks of one sample, the last one without its final newline
    cut = [int(starts[1500]), int(starts[1501]), int(starts[4000])]
    parts = [fq[:cut[0]], fq[cut[0]:cut[1]], fq[cut[1]:cut[2]], fq[cut[2]:-1]]
    _collapse(oracle_mod, parts)


def test_exact_multiple_of_tile_without_final_newline(oracle_mod):
    """The virtual newline of an unterminated last line falls into a tile of its own."""
    rng = np.random.default_rng(12)
    for n_reads, extra in ((400, 0), (150, 16384)):
        reads = _reads(rng, n_reads, 20, 50)
        base = _fastq(reads[:-1])
        last = reads[-1]
        h = (-(len(base) + 2 * len(last) + 5)) % 16384 + extra          # header length that makes the text a multiple of the tile
        if h == 0:
            h = 16384
        fq = base + ('@%s\n%s\n+\n%s' % ('h' * h, last, 'I' * len(last))).encode()
        assert len(fq) % 16384 == 0
        _collapse(oracle_mod, [fq])
        _collapse(oracle_mod, [fq], device_resident=True)


def test_long_headers_take_the_global_path(oracle_mod):
    rng = np.random.default_rng(13)
    reads = _reads(rng, 300, 16, 50)
    hdr = lambda k: ('x' * (int(rng.integers(900, 5000)) if k % 3 == 0 else 20)) + str(k)    # noqa: E731
    _collapse(oracle_mod, [_fastq(reads, header=hdr)])
    # a '+' line that repeats a kilobyte header, and CRLF line ends
    s = ''.join('@%s\r\n%s\r\n+%s\r\n%s\r\n' % ('y' * 1500 + str(k), r, 'y' * 1500 + str(k), 'I' * len(r)) for k, r in enumerate(reads))
    _collapse(oracle_mod, [s.encode()])


def test_tiny_reads_need_several_list_passes(oracle_mod):
    """About 4 bytes per line: more than 2044 newlines per 16 KiB tile."""
    rng = np.random.default_rng(14)
    reads = _reads(rng, 30000, 0, 4) + _reads(rng, 50, 20, 40)
    order = rng.permutation(len(reads))
    reads = [reads[i] for i in order]
    fq = _fastq(reads, header=lambda k: '')
    _collapse(oracle_mod, [fq], adapter='none')
    _collapse(oracle_mod, [fq[:-1]], adapter='none')


def test_quality_windows_through_the_single_pass(oracle_mod):
    rng = np.random.default_rng(15)
    reads = _reads(rng, 5000, 10, 50)

    def qual(r, k):
        q = rng.choice(list('#,:FI'), len(r), p=[.2, .2, .2, .2, .2])
        if k % 2:
            q[len(r) - min(len(r), int(rng.integers(0, 20))):] = '#'
        return ''.join(q)
    _collapse(oracle_mod, [_fastq(reads, qual=qual)])
    _collapse(oracle_mod, [_fastq(reads, qual=qual, eol='\r\n', final_eol=False)])


def test_partial_record_is_reported(oracle_mod):
    from mirge_b200.engine import Engine, MirgeError
    rng = np.random.default_rng(16)
    reads = _reads(rng, 2000, 16, 50)
    fq = _fastq(reads)
    cut = fq.rfind(b'\n+\n')            # drop the last record's '+' and quality lines
    e = Engine(1, False, 1 << 16, 64)
    e.set_adapter(ILLUMINA)
    with pytest.raises(MirgeError) as ei:
        e.feed_fastq(0, fq[:cut + 1], 33)
        e.sync()
    assert ei.value.code == -3


def test_speculation_is_rejected_on_ambiguous_text(oracle_mod):
    """Quality lines that look like headers ('@' first) and sequence lines that look like '+' lines make the local
    phase test ambiguous or wrong; k0_verify must notice and the exact path must give the reference's parse."""
    rng = np.random.default_rng(17)
    reads = _reads(rng, 3000, 16, 50)
    # quality strings starting with '@' (a legal quality character, Q31): every quality line looks like a header
    fq = _fastq(reads, qual=lambda r, k: '@' + 'I' * (len(r) - 1))
    _collapse(oracle_mod, [fq])
    # header-like quality AND a '+'-like sequence line in the same record (non-ACGT bytes collapse to N)
    fq2 = _fastq(['+' + r[1:] for r in reads], qual=lambda r, k: '@' + 'I' * (len(r) - 1))
    _collapse(oracle_mod, [fq2], adapter='none')
