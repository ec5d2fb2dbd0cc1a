// Runs the trim kernel's DEVICE classification function (mirge_b200/csrc/k1_prefilter.cuh:k1_prefilter,
// __host__ __device__) on the CPU, so its bit tricks -- the byte-wise anchor search on the 2-bit words, the
// window extension, the table lookups -- are checked without a GPU against the plain C model
// (tests/helpers/prefilter_model.c).  Test helper only; built by tests/test_prefilter_tables.py with nvcc.
#include "../../mirge_b200/csrc/k1_prefilter.cuh"

template <int W>
static int classify(const K1Pre &pre, u64 aw, const char *read, int n, int *rstart, int *limit)
{
    PSeq<W> p;
    for (int w = 0; w < W; w++) p.s[w] = 0;
    for (int w = 0; w < W / 2; w++) p.nm[w] = 0;
    for (int i = 0; i < n; i++) {                       // the record K0 writes: 2 bits per base, N -> mask bit + code 0
        int c = base_code((unsigned char)read[i]);
        if (c > 3) { p.nm[i >> 6] |= 1ull << (i & 63); c = 0; }
        p.s[i >> 5] |= (u64)c << ((i & 31) * 2);
    }
    int rs = -1, lim = 0;
    const int cls = k1_prefilter<W>(p, n, pre, aw, rs, lim);
    *rstart = rs; *limit = lim;
    return cls;
}

extern "C" int pfh_classify(const K1Pre *pre, unsigned long long aw, const char *read, int n, int W, int *rstart, int *limit)
{
    if (n > 32 * W) return -1;
    return W == 2 ? classify<2>(*pre, aw, read, n, rstart, limit) : classify<4>(*pre, aw, read, n, rstart, limit);
}

// Batch form: reads concatenated, off[k]..off[k+1]; out[3k..3k+2] = class, rstart, limit.
extern "C" void pfh_classify_batch(const K1Pre *pre, unsigned long long aw, const char *reads, const unsigned long long *off,
                                   unsigned long long n_reads, int W, int *out)
{
    for (unsigned long long k = 0; k < n_reads; k++)
        out[3 * k] = pfh_classify(pre, aw, reads + off[k], (int)(off[k + 1] - off[k]), W, out + 3 * k + 1, out + 3 * k + 2);
}
