"""The per-thread adapter searches of the trim pipeline's slow queue -- k1_match_full (all-column bit-vector search)
and match_back / locate_back (cutadapt's banded DP, cell for cell) -- are __host__ __device__: the very code the
GPU runs is executed here on the CPU (tests/helpers/prefilter_host.cu) and compared with the oracle's
Adapter.match_to restatement read by read.  The GPU tests compare whole tables; this pins the device logic itself."""
import ctypes as C

import numpy as np
import pytest

from conftest import ILLUMINA


def _mutate(rng, s, nsub, nins, ndel):
    s = list(s)
    for _ in range(nsub):
        if s:
            s[int(rng.integers(0, len(s)))] = 'ACGT'[rng.integers(0, 4)]
    for _ in range(nins):
        s.insert(int(rng.integers(0, len(s) + 1)), 'ACGT'[rng.integers(0, 4)])
    for _ in range(ndel):
        if s:
            del s[int(rng.integers(0, len(s)))]
    return ''.join(s)


def _reads(rng, adapter, n, maxlen):
    m = len(adapter)
    out = []
    for _ in range(n):
        kind = int(rng.integers(0, 6))
        ilen = int(rng.integers(0, maxlen))
        if kind == 5:
            unit = adapter[int(rng.integers(0, m)):][:int(rng.integers(1, 6))] or 'A'
            ins = (unit * maxlen)[:ilen]
        else:
            ins = ''.join('ACGT'[x] for x in rng.integers(0, 4, ilen))
        tail = '' if kind == 0 else _mutate(rng, adapter, int(rng.integers(0, 4)), int(rng.integers(0, 3)), int(rng.integers(0, 3)))
        if kind == 2:
            tail = tail[:int(rng.integers(0, m))] + adapter
        r = (ins + tail)[:maxlen]
        if kind == 3 and len(r) > 1:
            r = r[:int(rng.integers(max(1, len(r) - 12), len(r) + 1))]
        if rng.random() < 0.05 and r:
            q = int(rng.integers(0, len(r))); r = r[:q] + 'N' + r[q + 1:]
        out.append(r or 'A')
    return out


@pytest.mark.parametrize('adapter', [ILLUMINA, 'TGGAATTCTCGGGTGCCAAGG', 'AAAAAAAAAAAAAAAA', 'ACACACACACACACACACAC', 'TGGAATTC',
                                     'AGATCGGAAGAGCACACGTCTGAACTCCAGTC', 'AGATCGGAAGAGCACACGTCTGAACTCCAGTCACATCACGATCTCGTAT'])
@pytest.mark.parametrize('rate,mo', [(0.12, 3), (0.2, 1), (0.0, 3)])
def test_slow_queue_searches_match_oracle(host_kernel, oracle_mod, adapter, rate, mo):
    O = oracle_mod
    rng = np.random.default_rng(len(adapter) + int(rate * 100))
    for maxlen, W in ((64, 2), (120, 4)):
        reads = _reads(rng, adapter, 6000, maxlen)
        off = np.zeros(len(reads) + 1, np.uint64)
        off[1:] = np.cumsum([len(r) for r in reads])
        cat = ''.join(reads).encode() + b'\0'
        want = []
        for r in reads:
            mb = O.match_back(adapter, r, rate, mo)
            want.append((1, mb[0], mb[1]) if mb else None)
        for literal in (0, 1):
            got = np.zeros(3 * len(reads), np.int32)
            host_kernel.pfh_match_batch(adapter.encode(), C.c_double(rate), mo, cat, off.ctypes.data_as(C.c_void_p), C.c_uint64(len(reads)), W, literal,
                                        got.ctypes.data_as(C.c_void_p))
            got = got.reshape(-1, 3)
            for k, w in enumerate(want):
                if w is None:
                    assert got[k, 0] == 0, (adapter, reads[k], literal, got[k].tolist())
                else:
                    assert tuple(int(x) for x in got[k]) == w, (adapter, reads[k], literal, got[k].tolist(), w)


def _oracle_trim_len(O, adapters, read, rate, mo):
    """AdapterCutter(times=1): most matches wins, first adapter wins ties (cutadapt modifiers.py; oracle orc_trim_read)."""
    best = None
    for a in adapters:
        mb = O.match_back(a, read, rate, mo)
        if mb and (best is None or mb[1] > best[1]):
            best = mb
    return best[0] if best else len(read)


@pytest.mark.parametrize('spec', [ILLUMINA, 'TGGAATTCTCGGGTGCCAAGG', 'AAAAAAAAAAAAAAAA', 'ACACACACACACACACACAC', 'TGGAATTC',
                                  'AGATCGGAAGAGC,TGGAATTCTCGGGTGCCAAGG', 'AGATCGGAAGAGCACACGTCTGAACTCCAGTCACATCACGATCTCGTAT'])
@pytest.mark.parametrize('rate,mo', [(0.12, 3), (0.2, 1), (0.34, 5)])
def test_general_search_then_slow_search_match_oracle(host_kernel, oracle_mod, spec, rate, mo):
    """k1_trim_fast (windowed bit-vector search with candidate walks), falling back to k1_trim_slow exactly when it
    answers K1_SLOW -- the device code of k1_general_search / k1_slow_search, one lane at a time."""
    O = oracle_mod
    adapters = spec.split(',')
    rng = np.random.default_rng(len(spec) * 3 + int(rate * 100))
    n_fast = n_slow = 0
    for maxlen, W in ((64, 2), (120, 4)):
        reads = _reads(rng, adapters[-1], 6000, maxlen)
        off = np.zeros(len(reads) + 1, np.uint64)
        off[1:] = np.cumsum([len(r) for r in reads])
        cat = ''.join(reads).encode() + b'\0'
        got = np.zeros(2 * len(reads), np.int32)
        assert host_kernel.pfh_trim_batch(spec.encode(), C.c_double(rate), mo, cat, off.ctypes.data_as(C.c_void_p), C.c_uint64(len(reads)), W,
                                          got.ctypes.data_as(C.c_void_p)) == 0
        got = got.reshape(-1, 2)
        for k, r in enumerate(reads):
            assert int(got[k, 0]) == _oracle_trim_len(O, adapters, r, rate, mo), (spec, r, got[k].tolist())
        n_slow += int(got[:, 1].sum())
        n_fast += len(reads) - int(got[:, 1].sum())
    if all(len(a) <= 32 for a in adapters):
        assert n_fast > 0 and (n_slow > 0 or max(len(a) for a in adapters) <= 8)   # both outcomes of the windowed search occur
    else:
        assert n_fast == 0                         # adapters longer than one 32-bit word always take the slow search


class K1Pre(C.Structure):       # mirge_b200/csrc/k1_prefilter.cuh
    _fields_ = [('enabled', C.c_int), ('m', C.c_int), ('icap', C.c_int), ('anchor_rep', C.c_uint32), ('safe_p', C.c_uint32),
                ('safe_len', C.c_uint32 * 49), ('fullreach', C.c_uint8 * 32)]


def _device_tables(adapter, rate, mo):
    from mirge_b200 import _capi
    pre = K1Pre()
    m, ic, sp = C.c_int(), C.c_int(), C.c_uint32()
    pre.enabled = _capi.lib.mirge_prefilter_tables(adapter.encode(), rate, mo, C.byref(m), C.byref(ic), C.byref(sp),
                                                   C.cast(pre.safe_len, C.c_void_p), C.cast(pre.fullreach, C.c_void_p))
    aw = sum('ACGT'.index(c) << (2 * i) for i, c in enumerate(adapter))
    pre.m, pre.icap, pre.safe_p, pre.anchor_rep = m.value, ic.value, sp.value, ((aw & 0xFF) * 0x01010101) & 0xFFFFFFFF
    return pre


def _decide(host_kernel, pre, adapter, rate, mo, reads, W, use_1b):
    off = np.zeros(len(reads) + 1, np.uint64)
    off[1:] = np.cumsum([len(r) for r in reads])
    cat = ''.join(reads).encode() + b'\0'
    got = np.zeros(2 * len(reads), np.int32)
    assert host_kernel.pfh_decide_batch(C.byref(pre), adapter.encode(), C.c_double(rate), mo, cat, off.ctypes.data_as(C.c_void_p),
                                        C.c_uint64(len(reads)), W, int(use_1b), got.ctypes.data_as(C.c_void_p)) == 0
    return got.reshape(-1, 2)


@pytest.mark.parametrize('adapter', [ILLUMINA, 'TGGAATTCTCGGGTGCCAAGG', 'ACACACACACACACACACAC', 'AGATCGGAAGAGCACACGTCTGAACTCCAGTC',
                                     'AACTGTAGGCACCATCAAT', 'TTTTGGAATTTTGGAATTGGAATT'])
@pytest.mark.parametrize('rate,mo', [(0.12, 3), (0.2, 1)])
@pytest.mark.parametrize('use_1b', [0, 1, 2])       # bit 0: the class-1b prefix scan, bit 1: class 1 by the pigeonhole filter (the kernel's way)
def test_whole_read_decision_matches_oracle(host_kernel, oracle_mod, adapter, rate, mo, use_1b):
    """prefilter -> class 0 / class-1 scan / (planned) class-1b scan -> general -> slow search: the trimmed length every
    path of the K1 pipeline arrives at, device code on the host, against Adapter.match_to."""
    O = oracle_mod
    pre = _device_tables(adapter, rate, mo)
    assert pre.enabled == 1
    rng = np.random.default_rng(len(adapter) * 11 + int(rate * 100) + use_1b)
    paths = set()
    for maxlen, W in ((50, 2), (64, 2), (120, 4)):
        reads = _reads(rng, adapter, 3000, maxlen)
        # plus reads that END in an exact adapter prefix behind inserts of every length (the prefilter's home ground)
        m = len(adapter)
        for _ in range(3000):
            ins = ''.join('ACGT'[x] for x in rng.integers(0, 4, int(rng.integers(0, maxlen))))
            if rng.random() < 0.3:
                ins = ins[:len(ins) // 2] + _mutate(rng, adapter, int(rng.integers(0, 4)), int(rng.integers(0, 2)), int(rng.integers(0, 2)))
            reads.append((ins + adapter[:int(rng.integers(0, m + 1))])[:maxlen] or 'A')
        got = _decide(host_kernel, pre, adapter, rate, mo, reads, W, use_1b)
        for k, r in enumerate(reads):
            mb = O.match_back(adapter, r, rate, mo)
            assert int(got[k, 0]) == (mb[0] if mb else len(r)), (adapter, r, got[k].tolist(), mb)
        paths.update(got[:, 1].tolist())
    assert {0, 4} <= paths and 9 not in paths and (not (use_1b & 1) or 3 in paths)


def test_whole_read_decision_on_the_bench_workload(host_kernel, oracle_mod, small_world):
    O = oracle_mod
    pre = _device_tables(ILLUMINA, 0.12, 3)
    fq = bytes(small_world.fastq(40000, seed=1000)).split(b'\n')
    reads = []
    for r in range(40000):
        seq, q = fq[4 * r + 1].decode(), fq[4 * r + 3].decode()
        st, sp = O.quality_trim_index(q, 0, 10, 33)
        reads.append(seq[st:sp] or 'A')
    share = {}
    for use_1b in (False, True, 2):
        got = _decide(host_kernel, pre, ILLUMINA, 0.12, 3, reads, 2, use_1b)
        for k, r in enumerate(reads):
            mb = O.match_back(ILLUMINA, r, 0.12, 3)
            assert int(got[k, 0]) == (mb[0] if mb else len(r)), (r, got[k].tolist(), mb)
        share[use_1b] = np.bincount(got[:, 1], minlength=5) / float(len(reads))
    assert share[False][0] > 0.55 and share[False][4] < 0.15          # what the kernels do today
    assert share[True][4] < 0.07 and share[True][3] > 0.04            # the planned class 1b halves the general search's load
    assert share[2][2] < 0.01 and share[2][1] > 0.2                   # the pigeonhole filter settles nearly every class-1 read
