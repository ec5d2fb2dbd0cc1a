"""The trim kernel settles most reads with exact comparisons and host-built dominance tables
(mirge_b200/csrc/k1_prefilter.cuh).  Here, on the CPU: the tables the product library builds are
(a) re-derived by an independent numpy-free Python DP, and (b) used by a C model of the classifier
whose every settled read must agree with the oracle's Adapter.match_to restatement."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ILLUMINA

HERE = os.path.dirname(os.path.abspath(__file__))
ADAPTERS = [ILLUMINA, 'TGGAATTCTCGGGTGCCAAGG', 'AAAAAAAAAAAAAAAA', 'ACACACACACACACACACAC', 'TGGAATTC',
            'AGATCGGAAGAGCACACGTCTGAACTCCAGTC', 'AGATCGGAAGAGC', 'TTTTGGAATTTTGGAATTGGAATT', 'GATCGATCGGATCGATCGATCG',
            'AACTGTAGGCACCATCAAT', 'AGATCGGAAGAGCACACGTC', 'CTGTCTCTTATACACATCT', 'TCGTATGCCGTCTTCTGCTTG']   # Qiagen, NEBNext, Nextera, Illumina v1.0 small RNA


class PfTables(C.Structure):
    _fields_ = [('m', C.c_int), ('icap', C.c_int), ('thr_m', C.c_int), ('safe_p', C.c_uint32),
                ('safe_len', C.c_uint32 * 49), ('fullreach', C.c_uint8 * 32)]


def thresholds(m, rate, min_overlap):
    thr = []
    for i in range(m + 1):
        c = -1
        if i >= min_overlap:
            c = 0
            while float(c + 1) <= float(i) * rate:
                c += 1
        thr.append(c)
    return thr


def product_tables(adapter, rate=0.12, min_overlap=3):
    from mirge_b200 import _capi
    t = PfTables()
    m, ic, sp = C.c_int(), C.c_int(), C.c_uint32()
    on = _capi.lib.mirge_prefilter_tables(adapter.encode(), rate, min_overlap, C.byref(m), C.byref(ic), C.byref(sp),
                                          C.cast(t.safe_len, C.c_void_p), C.cast(t.fullreach, C.c_void_p))
    assert on >= 0
    t.m, t.icap, t.safe_p = m.value, ic.value, sp.value
    t.thr_m = thresholds(len(adapter), rate, min_overlap)[len(adapter)]
    return on, t


@pytest.fixture(scope='module')
def model(oracle_mod):
    so = os.path.join(HERE, 'helpers', 'libprefilter_model.so')
    src = os.path.join(HERE, 'helpers', 'prefilter_model.c')
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-fopenmp', '-shared', '-fPIC', '-o', so, src, '-lm'])
    return C.CDLL(so)


def py_tables(A, rate, min_overlap):
    """Independent restatement of capi.cu:k1pre_build (dict-based DP over the wildcard read)."""
    m = len(A)
    thr = thresholds(m, rate, min_overlap)
    kmax = thr[m]
    icap = m + kmax + 1
    NEG = -1000
    safe_len = [0] * (icap + 1)
    safe_p, fullreach = 0, [0] * m
    for I in range(icap + 1):
        for l in range(4, m):
            if thr[l] < 0:
                continue
            n = I + l
            G = [[[0] * (kmax + 1) for _ in range(n + 1)] for _ in range(m + 1)]
            for i in range(1, m + 1):
                for c in range(kmax + 1):
                    G[i][0][c] = 0 if c >= i else NEG
            for j in range(1, n + 1):
                for i in range(1, m + 1):
                    eq = j <= I or A[i - 1] == A[j - 1 - I]
                    for c in range(kmax + 1):
                        best = NEG
                        if eq and G[i - 1][j - 1][c] > NEG:
                            best = G[i - 1][j - 1][c] + 1
                        if c > 0:
                            best = max(best, G[i - 1][j - 1][c - 1], G[i - 1][j][c - 1], G[i][j - 1][c - 1])
                        G[i][j][c] = best
            partial = any(thr[i] >= 0 and G[i][n][thr[i]] >= l + 1 for i in range(l + 1, m + 1))
            reach = max([j for j in range(1, n + 1) if G[m][j][kmax] >= l + 1], default=-1)
            if not partial and reach < 0:
                safe_len[I] |= 1 << l
            if I == icap:
                if not partial:
                    safe_p |= 1 << l
                fullreach[l] = max(0, reach - I)
    return icap, safe_p, safe_len, fullreach


@pytest.mark.parametrize('adapter,rate,mo', [('TGGAATTCTCGG', 0.12, 3), ('ACACACACAC', 0.2, 1), ('TGGAATTCTCGGGTGC', 0.12, 3)])
def test_tables_match_independent_dp(adapter, rate, mo):
    on, t = product_tables(adapter, rate, mo)
    assert on == 1
    icap, safe_p, safe_len, fullreach = py_tables(adapter, rate, mo)
    assert (t.icap, t.safe_p) == (icap, safe_p)
    assert list(t.safe_len)[:icap + 1] == safe_len
    assert list(t.fullreach)[:len(adapter)] == fullreach


def test_prefilter_scope():
    assert product_tables('TGGAATT')[0] == 0                                     # shorter than 8: general search only
    assert product_tables('AGATCGGAAGAGCACACGTCTGAACTCCAGTCA')[0] == 0           # longer than one 32-bit word
    on, t = product_tables(ILLUMINA)
    assert on == 1 and t.m == 29
    assert (t.safe_len[22] >> 28) & 1 and (t.safe_len[24] >> 26) & 1             # the bench's common shapes are settled by length
    assert not (t.safe_len[26] >> 24) & 1 and (t.safe_p >> 24) & 1               # a 26-nt insert could hide an adapter: class 1
    assert t.safe_p & 0x7FF == 0                                                 # short exact prefixes prove nothing


@pytest.mark.parametrize('adapter', ADAPTERS)
@pytest.mark.parametrize('rate,mo', [(0.12, 3), (0.2, 1), (0.34, 5), (0.0, 3)])
def test_model_agrees_with_oracle_on_adversarial_reads(model, adapter, rate, mo):
    on, t = product_tables(adapter, rate, mo)
    if not on:
        pytest.skip('prefilter not applicable')
    for maxlen in (40, 64, 128):
        st = np.zeros(8, np.uint64)
        bad = C.create_string_buffer(700)
        model.pf_fuzz(C.byref(t), adapter.encode(), C.c_double(rate), mo, C.c_uint64(300000), C.c_uint64(maxlen + int(rate * 100)),
                      maxlen, st.ctypes.data_as(C.c_void_p), bad)
        assert int(st[4]) == 0, (adapter, rate, mo, maxlen, bad.value)
        assert int(st[5]) == 300000


def test_model_agrees_with_oracle_on_the_bench_workload(model, small_world):
    on, t = product_tables(ILLUMINA)
    fq = bytes(small_world.fastq(300000, seed=1000, L=50))
    st = np.zeros(8, np.uint64)
    model.pf_check_fastq(C.byref(t), ILLUMINA.encode(), fq, C.c_uint64(len(fq)), 33, 10, C.c_double(0.12), 3, st.ctypes.data_as(C.c_void_p))
    assert int(st[5]) == 300000 and int(st[4]) == 0
    assert st[0] > 0.55 * st[5] and st[2] < 0.15 * st[5]          # most reads are settled without DP


# ---------------------------------------------------------------------------------------------------------
# the DEVICE classification function, compiled for the host (tests/helpers/prefilter_host.cu), against the model
# ---------------------------------------------------------------------------------------------------------
class K1Pre(C.Structure):       # mirge_b200/csrc/k1_prefilter.cuh
    _fields_ = [('enabled', C.c_int), ('m', C.c_int), ('icap', C.c_int), ('anchor_rep', C.c_uint32), ('safe_p', C.c_uint32),
                ('safe_len', C.c_uint32 * 49), ('fullreach', C.c_uint8 * 32)]


def _adversarial_batch(rng, adapter, n_reads, maxlen):
    m = len(adapter)
    reads = []
    for _ in range(n_reads):
        ilen = int(rng.integers(0, maxlen + 1))
        kind = int(rng.integers(0, 4))
        ins = []
        while len(ins) < ilen:
            mode = 0 if kind == 0 else int(rng.integers(0, 3))
            if mode == 0:
                ins.append('ACGT'[rng.integers(0, 4)])
            elif mode == 1:
                a = int(rng.integers(0, m)); ins.extend(adapter[a:a + int(rng.integers(1, m + 1))])
            else:
                ins.extend(adapter[:int(rng.integers(1, 9))])          # anchors (and near-anchors) inside the insert
        r = (''.join(ins[:ilen]) + adapter[:int(rng.integers(0, m + 1))])[:maxlen]
        if rng.random() < 0.2:
            r = (r + adapter + 'ACGT'[rng.integers(0, 4)] * int(rng.integers(0, 6)))[:maxlen]
        if rng.random() < 0.1 and r:
            q = int(rng.integers(0, len(r))); r = r[:q] + 'N' + r[q + 1:]
        reads.append(r)
    return reads


@pytest.mark.parametrize('adapter', [ILLUMINA, 'AAAAAAAAAAAAAAAA', 'ACACACACACACACACACAC', 'TGGAATTC', 'AGATCGGAAGAGCACACGTCTGAACTCCAGTC',
                                     'AACTGTAGGCACCATCAAT'])
@pytest.mark.parametrize('maxlen,W', [(50, 2), (64, 2), (128, 4), (70, 4)])
def test_device_classifier_on_host_matches_model(model, host_kernel, adapter, maxlen, W):
    on, t = product_tables(adapter)
    assert on == 1
    aw = sum('ACGT'.index(c) << (2 * i) for i, c in enumerate(adapter))
    pre = K1Pre(enabled=1, m=t.m, icap=t.icap, anchor_rep=((aw & 0xFF) * 0x01010101) & 0xFFFFFFFF, safe_p=t.safe_p)
    for i in range(49):
        pre.safe_len[i] = t.safe_len[i]
    for i in range(32):
        pre.fullreach[i] = t.fullreach[i]
    rng = np.random.default_rng(len(adapter) * 7 + maxlen + W)
    reads = _adversarial_batch(rng, adapter, 20000, maxlen)
    off = np.zeros(len(reads) + 1, np.uint64)
    off[1:] = np.cumsum([len(r) for r in reads])
    cat = ''.join(reads).encode() + b'\0'
    got = np.zeros(3 * len(reads), np.int32)
    want = np.zeros(3 * len(reads), np.int32)
    host_kernel.pfh_classify_batch(C.byref(pre), C.c_uint64(aw), cat, off.ctypes.data_as(C.c_void_p), C.c_uint64(len(reads)), W, got.ctypes.data_as(C.c_void_p))
    model.pf_classify_batch(C.byref(t), adapter.encode(), cat, off.ctypes.data_as(C.c_void_p), C.c_uint64(len(reads)), want.ctypes.data_as(C.c_void_p))
    got, want = got.reshape(-1, 3), want.reshape(-1, 3)
    bad = np.flatnonzero((got != want).any(axis=1))
    assert bad.size == 0, (reads[bad[0]], got[bad[0]].tolist(), want[bad[0]].tolist())
    seen = set(got[:, 0].tolist())
    assert {0, 2} <= seen and (adapter != ILLUMINA or 1 in seen)
