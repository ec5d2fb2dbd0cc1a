/* CPU model of the trim kernel's exact-match prefilter (mirge_b200/csrc/k1_prefilter.cuh)
 * driven by the tables the PRODUCT library builds (mirge_prefilter_tables), checked
 * read by read against the oracle's Adapter.match_to restatement.  Test helper only:
 * compiled by tests/test_prefilter_tables.py.
 *   class 0: FULL (whole adapter occurs exactly -> leftmost occurrence), or E (the read
 *            ends with an exact adapter prefix of l bases) with safe_len[I] bit l
 *   class 1: E with safe_p bit l -> accepted when no full-length candidate
 *            (cost(m, j) <= thr_m) ends at a column j <= I + fullreach[l]
 *   class 2: everything else (the kernel runs the general search; nothing to check here) */
#include "../../oracle/mirge_oracle.c"

typedef struct {
    int m, icap, thr_m;
    uint32_t safe_p;
    uint32_t safe_len[49];
    uint8_t fullreach[32];
} PfTables;

static int pf_classify(const PfTables *T, const char *A, const char *R, int n, int *rstart, int *limit)
{
    const int m = T->m;
    for (int p = 0; p + 4 <= n; p++) {
        if (memcmp(R + p, A, 4) != 0) continue;
        int ml = 4;
        while (ml < m && p + ml < n && R[p + ml] == A[ml]) ml++;
        if (ml == m) { *rstart = p; return 0; }
        if (p + ml == n) {
            const int l = ml, I = p;
            *rstart = p;
            if (I <= T->icap && ((T->safe_len[I] >> l) & 1)) return 0;
            if ((T->safe_p >> l) & 1) { *limit = I + T->fullreach[l]; if (*limit > n) *limit = n; return 1; }
            return 2;
        }
    }
    return 2;
}

/* semiglobal minimum cost of the whole adapter ending at column j, for j = 1..limit */
static int pf_full_candidate(const char *A, int m, const char *R, int limit, int thr_m)
{
    int col[65];
    for (int i = 0; i <= m; i++) col[i] = i;
    for (int j = 1; j <= limit; j++) {
        int diag = col[0];
        col[0] = 0;
        for (int i = 1; i <= m; i++) {
            const int up = col[i];
            int v = diag + (A[i - 1] != R[j - 1]);
            if (col[i - 1] + 1 < v) v = col[i - 1] + 1;
            if (up + 1 < v) v = up + 1;
            diag = up; col[i] = v;
        }
        if (col[m] <= thr_m) return 1;
    }
    return 0;
}

/* stats: [0..2] reads per class, [3] class-1 reads sent on to the general search, [4] DISAGREEMENTS, [5] reads */
static int pf_check_read(const PfTables *T, const char *A, const char *R, int n, double rate, int min_overlap, uint64_t *stats)
{
    int rs = -1, limit = 0;
    const int cls = pf_classify(T, A, R, n, &rs, &limit);
    int settled = cls == 0;
    if (cls == 1) {
        if (pf_full_candidate(A, T->m, R, limit, T->thr_m)) __atomic_fetch_add(&stats[3], 1, __ATOMIC_RELAXED);
        else settled = 1;
    }
    __atomic_fetch_add(&stats[5], 1, __ATOMIC_RELAXED);
    __atomic_fetch_add(&stats[cls], 1, __ATOMIC_RELAXED);
    if (!settled) return 0;
    int ors = -1, omt = 0;
    const int found = orc_match_back(A, T->m, R, n, rate, min_overlap, &ors, &omt);
    if (!found || ors != rs) { __atomic_fetch_add(&stats[4], 1, __ATOMIC_RELAXED); return 1; }
    return 0;
}

static uint64_t rng_next(uint64_t *s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }

/* Adversarial reads: inserts made of random bases, adapter pieces and edited adapter copies, then an adapter
 * prefix (exact, sometimes edited, sometimes the whole adapter plus trailing bases). */
void pf_fuzz(const PfTables *T, const char *A, double rate, int min_overlap, uint64_t trials, uint64_t seed, int maxlen,
             uint64_t *stats, char *first_bad)
{
    const int m = T->m;
    #pragma omp parallel
    {
        uint64_t s = seed * 0x9E3779B97F4A7C15ULL + 77 * (uint64_t)omp_get_thread_num() + 1;
        char R[600];
        #pragma omp for schedule(static)
        for (uint64_t t = 0; t < trials; t++) {
            int n = 0;
            const int kind = rng_next(&s) % 8;
            const int ilen = rng_next(&s) % (maxlen + 1);
            while (n < ilen) {
                const int mode = (kind < 2) ? 0 : (int)(rng_next(&s) % 4);
                if (mode == 0) R[n++] = "ACGT"[rng_next(&s) & 3];
                else if (mode == 1) {
                    const int a = rng_next(&s) % m, len = 1 + rng_next(&s) % m;
                    for (int i = 0; i < len && a + i < m && n < ilen; i++) R[n++] = A[a + i];
                } else if (mode == 2) {
                    const int len = 1 + rng_next(&s) % m;
                    for (int i = 0; i < len && n < ilen; i++) {
                        const int r = rng_next(&s) % 16;
                        if (r == 0) R[n++] = "ACGT"[rng_next(&s) & 3];
                        else if (r == 1) { R[n++] = "ACGT"[rng_next(&s) & 3]; i--; }
                        else if (r == 2) { }
                        else R[n++] = A[i];
                    }
                } else R[n++] = A[rng_next(&s) % m];
            }
            int tl = rng_next(&s) % (m + 1);
            if (kind == 7) tl = m;
            for (int i = 0; i < tl && n < maxlen; i++) {
                char c = A[i];
                if (kind == 6 && rng_next(&s) % 12 == 0) c = "ACGT"[rng_next(&s) & 3];
                R[n++] = c;
            }
            if (kind == 7) { const int extra = rng_next(&s) % 8; for (int i = 0; i < extra && n < maxlen; i++) R[n++] = "ACGT"[rng_next(&s) & 3]; }
            if (n > maxlen) n = maxlen;
            if (rng_next(&s) % 40 == 0 && n > 0) R[rng_next(&s) % n] = 'N';
            if (pf_check_read(T, A, R, n, rate, min_overlap, stats)) {
                #pragma omp critical
                { if (!first_bad[0]) { memcpy(first_bad, R, n); first_bad[n] = 0; } }
            }
        }
    }
}

/* FASTQ driver: quality-trim every read like the pipeline (QualityTrimmer(0, qcut, base)), then check it. */
void pf_check_fastq(const PfTables *T, const char *A, const char *buf, uint64_t nbytes, int base, int qcut, double rate,
                    int min_overlap, uint64_t *stats)
{
    uint64_t p = 0;
    while (p < nbytes) {
        uint64_t l[5]; l[0] = p;
        for (int k = 1; k <= 4; k++) {
            const char *q = l[k - 1] < nbytes ? memchr(buf + l[k - 1], '\n', nbytes - l[k - 1]) : NULL;
            l[k] = q ? (uint64_t)(q - buf) + 1 : nbytes + 1;
        }
        const int len = (int)(l[2] - l[1] - 1);
        int st, sp;
        orc_quality_trim_index(buf + l[3], len, 0, qcut, base, &st, &sp);
        char up[512];
        int n = sp - st; if (n < 0) n = 0; if (n > 500) n = 500;
        for (int i = 0; i < n; i++) {
            const char c = (char)toupper((unsigned char)buf[l[1] + st + i]);
            up[i] = (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : 'N';
        }
        pf_check_read(T, A, up, n, rate, min_overlap, stats);
        p = l[4];
    }
}

/* Batch form of the model's classifier (for the comparison with the device function run on the host). */
void pf_classify_batch(const PfTables *T, const char *A, const char *reads, const uint64_t *off, uint64_t n_reads, int *out)
{
    for (uint64_t k = 0; k < n_reads; k++) {
        int rs = -1, lim = 0;
        out[3 * k] = pf_classify(T, A, reads + off[k], (int)(off[k + 1] - off[k]), &rs, &lim);
        out[3 * k + 1] = rs; out[3 * k + 2] = lim;
    }
}
