/*
 * synth_fastq.c -- seeded synthetic small-RNA FASTQ generator (workload for
 * tests and bench.py; neither product nor oracle).  Read mixture follows
 * miRNA / isomiR / mismatch reads planted from the
 * synthetic libraries, other RNA classes, noise and adapter dimers, each
 * followed by the Illumina small-RNA 3' adapter, fixed read length L.
 * Every read draws from its own splitmix64 stream keyed by (seed, read index),
 * so output is identical for any thread count.
 */
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    const char *seq; const uint64_t *off; int64_t n;
} SynLib;

typedef struct {
    SynLib hairpin, trna, pretrna, snorna, rrna, ncrna, mrna, spike;
    int64_t n_expr;             /* expressed miRNAs */
    const int32_t *expr_hp;     /* hairpin index */
    const int32_t *expr_start;  /* mature start within hairpin */
    const int32_t *expr_len;    /* mature length */
    const double *expr_cdf;     /* cumulative Zipf weights, last = 1 */
} SynWorld;

static inline uint64_t sm64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline double urand(uint64_t *s) { return (double)(sm64(s) >> 11) * (1.0 / 9007199254740992.0); }
static inline int irand(uint64_t *s, int lo, int hi) { return lo + (int)(sm64(s) % (uint64_t)(hi - lo + 1)); }
static const char BASES[4] = {'A', 'C', 'G', 'T'};
static inline char other_base(uint64_t *s, char c)
{
    char d; do { d = BASES[sm64(s) & 3]; } while (d == c); return d;
}

static int frag(const SynLib *l, uint64_t *s, int minl, int maxl, int where, char *out)
{
    /* where: 0 anywhere, 1 5' end, 2 3' end */
    if (l->n <= 0) return 0;
    int64_t r = (int64_t)(sm64(s) % (uint64_t)l->n);
    int64_t rl = (int64_t)(l->off[r + 1] - l->off[r]);
    int len = irand(s, minl, maxl);
    if (len > rl) len = (int)rl;
    int64_t st = 0;
    if (where == 0) st = (int64_t)(sm64(s) % (uint64_t)(rl - len + 1));
    else if (where == 2) st = rl - len;
    memcpy(out, l->seq + l->off[r] + st, (size_t)len);
    return len;
}

static const char ADAPTER3[] = "TGGAATTCTCGGGTGCCAAGGAACTCCAGTCAC";
static const char ADAPTER_TAIL[] = "ATCTCGTATGCCGTCTTCTGCTTG";

/* record size for read length L with a fixed-width header "@r%011llu" */
uint64_t synth_record_bytes(int L) { return (uint64_t)(1 + 1 + 11 + 1 + L + 1 + 2 + L + 1); }

int synth_fastq(const SynWorld *w, uint64_t seed, uint64_t first_read, uint64_t n_reads, int L,
                const char *index6, char *out, int nthreads)
{
    uint64_t rec = synth_record_bytes(L);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)n_reads; k++) {
        uint64_t rid = first_read + (uint64_t)k;
        uint64_t st = seed * 0xD1342543DE82EF95ULL + rid * 0xA0761D6478BD642FULL + 0x1234567ULL;
        uint64_t *s = &st;
        char ins[256]; int n = 0;
        double u = urand(s);
        if (u < 0.55 && w->n_expr > 0) {
            /* miRNA-derived */
            double z = urand(s);
            int64_t lo = 0, hi = w->n_expr - 1;
            while (lo < hi) { int64_t mid = (lo + hi) / 2; if (w->expr_cdf[mid] < z) lo = mid + 1; else hi = mid; }
            int hp = w->expr_hp[lo], ms = w->expr_start[lo], ml = w->expr_len[lo];
            const char *h = w->hairpin.seq + w->hairpin.off[hp];
            int hl = (int)(w->hairpin.off[hp + 1] - w->hairpin.off[hp]);
            double v = urand(s);
            int a = ms, b = ms + ml;
            if (v >= 0.65 && v < 0.85) {          /* templated 5'/3' offsets */
                a = ms + irand(s, -2, 2); b = ms + ml + irand(s, -3, 4);
                if (a < 0) a = 0;
                if (b > hl) b = hl;
            }
            n = b - a; memcpy(ins, h + a, (size_t)n);
            if (v >= 0.85 && v < 0.92) {          /* non-templated 3' A/U additions */
                int add = irand(s, 1, 2);
                for (int t = 0; t < add; t++) ins[n++] = (sm64(s) & 1) ? 'A' : 'T';
            } else if (v >= 0.92 && v < 0.97) {   /* one substitution, half A->G at positions 2-16 */
                int done = 0;
                if (sm64(s) & 1) {
                    int cand[32], nc = 0;
                    for (int t = 1; t < 16 && t < n; t++) if (ins[t] == 'A') cand[nc++] = t;
                    if (nc) { ins[cand[sm64(s) % (uint64_t)nc]] = 'G'; done = 1; }
                }
                if (!done) { int p = irand(s, 0, n - 1); ins[p] = other_base(s, ins[p]); }
            } else if (v >= 0.97) {               /* two substitutions */
                int p = irand(s, 0, n - 1), q;
                do { q = irand(s, 0, n - 1); } while (q == p);
                ins[p] = other_base(s, ins[p]); ins[q] = other_base(s, ins[q]);
            }
        } else if (u < 0.58) {
            n = frag(&w->hairpin, s, 26, 40, 0, ins);
        } else if (u < 0.68) {
            int wh = (int)(sm64(s) % 3);
            n = frag(&w->trna, s, 16, 40, wh, ins);
            if (n > 0 && urand(s) < 0.25) { int p = irand(s, 0, n - 1); ins[p] = other_base(s, ins[p]); }
        } else if (u < 0.70) {
            n = frag(&w->pretrna, s, 14, 30, 0, ins);
            while (n > 0 && ins[n - 1] == 'T') n--;   /* keep the planted remainder unambiguous */
            int t = irand(s, 3, 6);
            for (int x = 0; x < t; x++) ins[n++] = 'T';
        } else if (u < 0.73) {
            n = frag(&w->snorna, s, 18, 45, 0, ins);
            if (n > 0 && urand(s) < 0.2) { int p = irand(s, 0, n - 1); ins[p] = other_base(s, ins[p]); }
        } else if (u < 0.79) {
            n = frag(&w->rrna, s, 18, 45, 0, ins);
            if (n > 0 && urand(s) < 0.2) { int p = irand(s, 0, n - 1); ins[p] = other_base(s, ins[p]); }
        } else if (u < 0.81) {
            n = frag(&w->ncrna, s, 18, 45, 0, ins);
            if (n > 0 && urand(s) < 0.2) { int p = irand(s, 0, n - 1); ins[p] = other_base(s, ins[p]); }
        } else if (u < 0.87) {
            n = frag(&w->mrna, s, 16, 50, 0, ins);
            if (n >= 30) {                        /* 0-2 mismatches beyond position 28 */
                int nm = irand(s, 0, 2);
                for (int x = 0; x < nm; x++) { int p = irand(s, 28, n - 1); ins[p] = other_base(s, ins[p]); }
            }
        } else if (u < 0.88) {
            n = frag(&w->spike, s, 200, 200, 0, ins);   /* whole spike-in oligo */
        } else if (u < 0.96) {
            n = irand(s, 16, 45);
            for (int x = 0; x < n; x++) ins[x] = BASES[sm64(s) & 3];
        } else {
            n = irand(s, 0, 15);
            for (int x = 0; x < n; x++) ins[x] = BASES[sm64(s) & 3];
        }
        /* insert + adapter + index + tail, truncated / random-padded to L */
        char full[512]; int fl = 0;
        memcpy(full, ins, (size_t)n); fl = n;
        memcpy(full + fl, ADAPTER3, sizeof(ADAPTER3) - 1); fl += (int)sizeof(ADAPTER3) - 1;
        memcpy(full + fl, index6, 6); fl += 6;
        memcpy(full + fl, ADAPTER_TAIL, sizeof(ADAPTER_TAIL) - 1); fl += (int)sizeof(ADAPTER_TAIL) - 1;
        while (fl < L) full[fl++] = BASES[sm64(s) & 3];
        char *o = out + (uint64_t)k * rec;
        int hl = sprintf(o, "@r%011llu\n", (unsigned long long)rid);
        char *sq = o + hl, *ql = o + hl + L + 3;
        for (int x = 0; x < L; x++) {
            char c = full[x];
            double e = urand(s);
            if (e < 0.001) c = other_base(s, c);
            else if (e < 0.0015) c = 'N';
            sq[x] = c;
        }
        sq[L] = '\n'; sq[L + 1] = '+'; sq[L + 2] = '\n';
        for (int x = 0; x < L; x++) {
            double e = urand(s);
            ql[x] = (e < 0.80) ? 'F' : (e < 0.95 ? ':' : ',');
        }
        if (urand(s) < 0.12) {
            int t = irand(s, 1, 15);
            for (int x = L - t; x < L; x++) if (x >= 0) ql[x] = '#';
        }
        ql[L] = '\n';
    }
    return 0;
}
