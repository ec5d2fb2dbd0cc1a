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

// ---------------------------------------------------------------------------------------------------------
// The per-thread adapter searches of the slow queue, on the host: k1_match_full (all-column bit-vector search,
// adapters <= 32 nt) and match_back (exact find + the cell-for-cell banded DP locate_back).
// ---------------------------------------------------------------------------------------------------------
#include "../../mirge_b200/csrc/k1_kernel.cuh"

template <int W>
static int match_one(const char *adapter, double rate, int min_overlap, const char *read, int n, int literal, int *rstart, int *matches)
{
    TrimParams tp;
    memset(&tp, 0, sizeof tp);
    tp.mode = 1; tp.n_adapters = 1; tp.error_rate = rate; tp.min_overlap = min_overlap;
    const int m = (int)strlen(adapter);
    tp.alen[0] = m;
    for (int i = 0; i < m; i++) {
        const int c = base_code((unsigned char)adapter[i]);
        tp.aw[0][i >> 5] |= (u64)c << ((i & 31) * 2);
        tp.peq[0][c] |= 1ull << i;
    }
    signed char thr[MIRGE_MAX_ADAPTER_LEN + 1];
    for (int i = 0; i <= MIRGE_MAX_ADAPTER_LEN; i++) {
        int c = -1;
        if (i >= min_overlap) { c = 0; while ((double)(c + 1) <= (double)i * rate) c++; }
        thr[i] = (signed char)c;
    }
    u32 peq8[8];
    for (int k = 0; k < 8; k++) peq8[k] = (k & 4) ? 0u : (u32)tp.peq[0][k & 3];
    PSeq<W> p;
    for (int w = 0; w < W; w++) p.s[w] = 0;
    for (int w = 0; w < W / 2; w++) p.nm[w] = 0;
    for (int i = 0; i < n; i++) {
        int c = base_code((unsigned char)read[i]);
        if (c > 3) { p.nm[i >> 6] |= 1ull << (i & 63); c = 0; }
        p.s[i >> 5] |= (u64)c << ((i & 31) * 2);
    }
    int rs = -1, mt = 0;
    bool found;
    if (m <= 32 && !literal) { rs = k1_match_full<W>(p, n, tp, thr, 0, peq8, mt); found = rs >= 0; }
    else found = match_back<W>(p, n, tp, 0, rs, mt);
    *rstart = rs; *matches = mt;
    return found ? 1 : 0;
}

// out[3k..3k+2] = found, rstart, matches
extern "C" void pfh_match_batch(const char *adapter, double rate, int min_overlap, const char *reads, const unsigned long long *off,
                                unsigned long long n_reads, int W, int literal, int *out)
{
    for (unsigned long long k = 0; k < n_reads; k++) {
        const int n = (int)(off[k + 1] - off[k]);
        out[3 * k] = W == 2 ? match_one<2>(adapter, rate, min_overlap, reads + off[k], n, literal, out + 3 * k + 1, out + 3 * k + 2)
                            : match_one<4>(adapter, rate, min_overlap, reads + off[k], n, literal, out + 3 * k + 1, out + 3 * k + 2);
    }
}

// ---------------------------------------------------------------------------------------------------------
// The general search (k1_trim_fast: windowed bit-vector search + candidate walks; a "warp" of one lane on the
// host) followed, when it answers K1_SLOW, by k1_trim_slow -- the sequence k1_general_search / k1_slow_search run.
// Adapters may be a comma-separated list.  out[2k] = trimmed length, out[2k+1] = 1 if the slow search was needed.
// ---------------------------------------------------------------------------------------------------------
template <int W>
static void trim_one(const TrimParams &tp, const signed char *thr, const u32 *peq8, const char *read, int n, int *out)
{
    static u64 ring[K1_RING * K1_THREADS];
    static u32 ringh[K1_RING * K1_THREADS];
    PSeq<W> p;
    for (int w = 0; w < W; w++) p.s[w] = 0;
    for (int w = 0; w < W / 2; w++) p.nm[w] = 0;
    for (int i = 0; i < n; i++) {
        int c = base_code((unsigned char)read[i]);
        if (c > 3) { p.nm[i >> 6] |= 1ull << (i & 63); c = 0; }
        p.s[i >> 5] |= (u64)c << ((i & 31) * 2);
    }
    int nf = k1_trim_fast<W>(p, n, tp, thr, ring, ringh, peq8);
    out[1] = nf == K1_SLOW;
    if (nf == K1_SLOW) nf = k1_trim_slow<W>(p, n, tp, thr, peq8, false);
    out[0] = nf;
}

extern "C" int pfh_trim_batch(const char *spec, double rate, int min_overlap, const char *reads, const unsigned long long *off,
                              unsigned long long n_reads, int W, int *out)
{
    TrimParams tp;
    memset(&tp, 0, sizeof tp);
    tp.mode = 1; tp.error_rate = rate; tp.min_overlap = min_overlap;
    const char *q = spec;
    while (*q) {
        const char *e = strchr(q, ',');
        const int m = e ? (int)(e - q) : (int)strlen(q);
        if (tp.n_adapters == MIRGE_MAX_ADAPTERS || m > MIRGE_MAX_ADAPTER_LEN) return -1;
        const int a = tp.n_adapters++;
        tp.alen[a] = m;
        for (int i = 0; i < m; i++) {
            const int c = base_code((unsigned char)q[i]);
            tp.aw[a][i >> 5] |= (u64)c << ((i & 31) * 2);
            tp.peq[a][c] |= 1ull << i;
        }
        q += m;
        if (*q == ',') q++;
    }
    signed char thr[MIRGE_MAX_ADAPTER_LEN + 1];
    for (int i = 0; i <= MIRGE_MAX_ADAPTER_LEN; i++) {
        int c = -1;
        if (i >= min_overlap) { c = 0; while ((double)(c + 1) <= (double)i * rate) c++; }
        thr[i] = (signed char)c;
    }
    u32 peq8[MIRGE_MAX_ADAPTERS * 8];
    for (int k = 0; k < MIRGE_MAX_ADAPTERS * 8; k++) peq8[k] = (k & 4) ? 0u : (u32)tp.peq[k >> 3][k & 3];
    for (unsigned long long k = 0; k < n_reads; k++) {
        const int n = (int)(off[k + 1] - off[k]);
        if (W == 2) trim_one<2>(tp, thr, peq8, reads + off[k], n, out + 2 * k);
        else trim_one<4>(tp, thr, peq8, reads + off[k], n, out + 2 * k);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// The whole per-read decision of the K1 pipeline for ONE adapter, on the host: prefilter -> class 0 / class 1 scan /
// (optionally) the class-1b prefix scan -> general search -> slow search.  out[2k] = trimmed length,
// out[2k+1] = path: 0 class 0, 1 class 1 settled by the scan, 2 class 1 sent on, 3 settled by the 1b scan, 4 general search.
// ---------------------------------------------------------------------------------------------------------
template <int W>
static void decide_one(const K1Pre &pre, const TrimParams &tp, const signed char *thr, const u32 *peq8, const char *read, int n,
                       int use_1b, int *out)
{
    static u64 ring[K1_RING * K1_THREADS];
    static u32 ringh[K1_RING * K1_THREADS];
    PSeq<W> p;
    for (int w = 0; w < W; w++) p.s[w] = 0;
    for (int w = 0; w < W / 2; w++) p.nm[w] = 0;
    for (int i = 0; i < n; i++) {
        int c = base_code((unsigned char)read[i]);
        if (c > 3) { p.nm[i >> 6] |= 1ull << (i & 63); c = 0; }
        p.s[i >> 5] |= (u64)c << ((i & 31) * 2);
    }
    const int m = tp.alen[0];
    int rs = -1, lim = 0, path = 4;
    const int cls = pre.enabled ? k1_prefilter<W>(p, n, pre, tp.aw[0][0], rs, lim) : 2;
    bool general = true;
    if (cls == 0) { general = false; path = 0; }
    else if (cls == 1) {
        const bool scan = k1_full_candidate_scan<W>(p, lim, m, thr[m], peq8);
        u64 bm[4 * (W / 2)];
        const bool piece = k1_piece_hit<W>(p, lim, m, thr[m], tp.aw[0][0], bm, 1);
        if (scan && !piece) { out[0] = -1000; out[1] = 9; return; }     // the filter must never miss a candidate
        general = (use_1b & 2) ? piece : scan;                           // bit 1: what the kernel does (pigeonhole filter)
        path = general ? 2 : 1;
    } else if ((use_1b & 1) && rs >= 0 && thr[n - rs] >= 0) {
        general = k1_prefix_candidate_scan<W>(p, n, n - rs, m, thr, peq8);
        path = general ? 4 : 3;
    }
    int nf = rs;
    if (general) {
        nf = k1_trim_fast<W>(p, n, tp, thr, ring, ringh, peq8);
        if (nf == K1_SLOW) nf = k1_trim_slow<W>(p, n, tp, thr, peq8, false);
    }
    out[0] = nf; out[1] = path;
}

extern "C" int pfh_decide_batch(const K1Pre *pre, const char *adapter, double rate, int min_overlap, const char *reads,
                                const unsigned long long *off, unsigned long long n_reads, int W, int use_1b, int *out)
{
    TrimParams tp;
    memset(&tp, 0, sizeof tp);
    tp.mode = 1; tp.n_adapters = 1; tp.error_rate = rate; tp.min_overlap = min_overlap;
    const int m = (int)strlen(adapter);
    if (m > MIRGE_MAX_ADAPTER_LEN) return -1;
    tp.alen[0] = m;
    for (int i = 0; i < m; i++) {
        const int c = base_code((unsigned char)adapter[i]);
        tp.aw[0][i >> 5] |= (u64)c << ((i & 31) * 2);
        tp.peq[0][c] |= 1ull << i;
    }
    signed char thr[MIRGE_MAX_ADAPTER_LEN + 1];
    for (int i = 0; i <= MIRGE_MAX_ADAPTER_LEN; i++) {
        int c = -1;
        if (i >= min_overlap) { c = 0; while ((double)(c + 1) <= (double)i * rate) c++; }
        thr[i] = (signed char)c;
    }
    u32 peq8[MIRGE_MAX_ADAPTERS * 8];
    for (int k = 0; k < MIRGE_MAX_ADAPTERS * 8; k++) peq8[k] = (k & 4) ? 0u : (u32)tp.peq[k >> 3][k & 3];
    for (unsigned long long k = 0; k < n_reads; k++) {
        const int n = (int)(off[k + 1] - off[k]);
        if (W == 2) decide_one<2>(*pre, tp, thr, peq8, reads + off[k], n, use_1b, out + 2 * k);
        else decide_one<4>(*pre, tp, thr, peq8, reads + off[k], n, use_1b, out + 2 * k);
    }
    return 0;
}
