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
