/*
 * mirge_oracle.c -- CPU ORACLE for the miRge2.0 annotate-mode read pipeline.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mirge_b200/ may import, link or
 * execute this file; it is the checker used by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference (/root/reference, Python 2) ships no tests,
 * fixtures or golden vectors, and its arithmetic lives in two un-vendored
 * third-party programs that are absent from this image:
 *   - cutadapt (setup.py:27 ">=1.11", README.md:50 "v1.11-1.16"; restated: 1.16)
 *   - bowtie 1.1.2 (README.md:49)
 * What follows restates their PUBLISHED algorithms as used at the reference's
 * call sites; each function cites the reference file:line it stands in for.
 *
 *   stage 1a  trim_file        src/mirge/utils/trim_file.py:24-66,89-134
 *             QualityTrimmer   cutadapt/_qualtrim.pyx quality_trim_index
 *             AdapterCutter    cutadapt/adapters.py Adapter.match_to (BACK)
 *                              cutadapt/_align.pyx Aligner.locate
 *   stage 1b  quantReads       src/mirge/utils/quantReads.py:3-24
 *   stage 2   bowtie rounds    src/mirge/utils/runAnnotationPipeline.py:543-554,
 *                              566-599,636-705  (acceptance rules: bowtie manual,
 *                              "-n" / "-v" alignment modes, -l 28 -e 70, FASTA
 *                              input => every base quality 'I' => Maq-rounded 30)
 *   stage 3   summarize        src/mirge/utils/summarize.py:34-66 (integer tallies)
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ctype.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_ADAPTERS 8
#define ORC_MAX_LIBS 10
#define ORC_MAX_READLEN 4096

/* ------------------------------------------------------------------ */
/* stage 1a: cutadapt restatement                                      */
/* ------------------------------------------------------------------ */

/* cutadapt/_qualtrim.pyx quality_trim_index(qualities, cutoff_front,
 * cutoff_back, base), as called by QualityTrimmer(0, 10, base)
 * (trim_file.py:30). */
void orc_quality_trim_index(const char *q, int len, int cutoff_front,
                            int cutoff_back, int base, int *start_out,
                            int *stop_out)
{
    int s = 0, max_qual = 0, start = 0, stop = len, i;
    for (i = 0; i < len; i++) {
        s += cutoff_front - ((int)(unsigned char)q[i] - base);
        if (s < 0) break;
        if (s > max_qual) { max_qual = s; start = i + 1; }
    }
    max_qual = 0; s = 0;
    for (i = len - 1; i >= 0; i--) {
        s += cutoff_back - ((int)(unsigned char)q[i] - base);
        if (s < 0) break;
        if (s > max_qual) { max_qual = s; stop = i; }
    }
    if (start >= stop) { start = 0; stop = 0; }
    *start_out = start; *stop_out = stop;
}

typedef struct { int cost, matches, origin; } OrcEntry;

/* cutadapt/_align.pyx Aligner.locate(query) for flags = BACK
 * (START_WITHIN_SEQ2 | STOP_WITHIN_SEQ2 | STOP_WITHIN_SEQ1): alignment starts
 * at adapter position 0, anywhere in the read; may stop inside the adapter
 * only at the read's end.  Unit costs, Ukkonen band exactly as upstream.
 * out = (refstart, refstop, querystart, querystop, matches, errors);
 * returns 1 if an alignment was found, else 0. */
int orc_locate(const char *s1, int m, const char *s2, int n,
               double max_error_rate, int min_overlap, int out[6])
{
    OrcEntry column[m + 1];
    int i, j;
    int k = (int)(max_error_rate * m);
    /* start_in_query => max_n = n; stop_in_query => min_n = 0 */
    for (i = 0; i <= m; i++) {
        column[i].matches = 0;
        column[i].cost = i;
        column[i].origin = 0;  /* max(0, min_n - i) */
    }
    int best_ref_stop = m, best_query_stop = n, best_cost = m + n,
        best_origin = 0, best_matches = 0;
    int last = (m < k + 1) ? m : k + 1;
    int broke = 0;
    for (j = 1; j <= n; j++) {
        OrcEntry tmp = column[0];
        column[0].origin = j;
        for (i = 1; i <= last; i++) {
            int cost, origin, matches;
            if (s1[i - 1] == s2[j - 1]) {
                cost = tmp.cost; origin = tmp.origin; matches = tmp.matches + 1;
            } else {
                int cost_diag = tmp.cost + 1;
                int cost_deletion = column[i].cost + 1;
                int cost_insertion = column[i - 1].cost + 1;
                if (cost_diag <= cost_deletion && cost_diag <= cost_insertion) {
                    cost = cost_diag; origin = tmp.origin; matches = tmp.matches;
                } else if (cost_insertion <= cost_deletion) {
                    cost = cost_insertion; origin = column[i - 1].origin;
                    matches = column[i - 1].matches;
                } else {
                    cost = cost_deletion; origin = column[i].origin;
                    matches = column[i].matches;
                }
            }
            tmp = column[i];
            column[i].cost = cost; column[i].origin = origin;
            column[i].matches = matches;
        }
        while (last >= 0 && column[last].cost > k) last--;
        if (last < m) {
            last++;
        } else {
            int length = m + (column[m].origin < 0 ? column[m].origin : 0);
            int cost = column[m].cost, matches = column[m].matches;
            if (length >= min_overlap && cost <= length * max_error_rate &&
                (matches > best_matches ||
                 (matches == best_matches && cost < best_cost))) {
                best_matches = matches; best_cost = cost;
                best_origin = column[m].origin; best_ref_stop = m;
                best_query_stop = j;
                if (cost == 0 && matches == m) { broke = 1; break; }
            }
        }
    }
    (void)broke;
    /* max_n == n always here; stop_in_ref => first_i = 0 */
    for (i = m; i >= 0; i--) {
        int length = i + (column[i].origin < 0 ? column[i].origin : 0);
        int cost = column[i].cost, matches = column[i].matches;
        if (length >= min_overlap && cost <= length * max_error_rate &&
            (matches > best_matches ||
             (matches == best_matches && cost < best_cost))) {
            best_matches = matches; best_cost = cost;
            best_origin = column[i].origin; best_ref_stop = i;
            best_query_stop = n;
        }
    }
    if (best_cost == m + n) return 0;
    int start1, start2;
    if (best_origin >= 0) { start1 = 0; start2 = best_origin; }
    else { start1 = -best_origin; start2 = 0; }
    out[0] = start1; out[1] = best_ref_stop; out[2] = start2;
    out[3] = best_query_stop; out[4] = best_matches; out[5] = best_cost;
    return 1;
}

/* Same recurrence WITHOUT the Ukkonen band (all m rows every column); used by
 * tests to show the band is only an optimisation. */
int orc_locate_unbanded(const char *s1, int m, const char *s2, int n,
                        double max_error_rate, int min_overlap, int out[6])
{
    OrcEntry column[m + 1];
    int i, j;
    for (i = 0; i <= m; i++) { column[i].matches = 0; column[i].cost = i; column[i].origin = 0; }
    int best_ref_stop = m, best_query_stop = n, best_cost = m + n, best_origin = 0, best_matches = 0;
    for (j = 1; j <= n; j++) {
        OrcEntry tmp = column[0];
        column[0].origin = j;
        for (i = 1; i <= m; i++) {
            int cost, origin, matches;
            if (s1[i - 1] == s2[j - 1]) { cost = tmp.cost; origin = tmp.origin; matches = tmp.matches + 1; }
            else {
                int cd = tmp.cost + 1, cdel = column[i].cost + 1, cins = column[i - 1].cost + 1;
                if (cd <= cdel && cd <= cins) { cost = cd; origin = tmp.origin; matches = tmp.matches; }
                else if (cins <= cdel) { cost = cins; origin = column[i - 1].origin; matches = column[i - 1].matches; }
                else { cost = cdel; origin = column[i].origin; matches = column[i].matches; }
            }
            tmp = column[i];
            column[i].cost = cost; column[i].origin = origin; column[i].matches = matches;
        }
        {
            int length = m, cost = column[m].cost, matches = column[m].matches;
            if (length >= min_overlap && cost <= length * max_error_rate &&
                (matches > best_matches || (matches == best_matches && cost < best_cost))) {
                best_matches = matches; best_cost = cost; best_origin = column[m].origin;
                best_ref_stop = m; best_query_stop = j;
                if (cost == 0 && matches == m) break;
            }
        }
    }
    for (i = m; i >= 0; i--) {
        int length = i, cost = column[i].cost, matches = column[i].matches;
        if (length >= min_overlap && cost <= length * max_error_rate &&
            (matches > best_matches || (matches == best_matches && cost < best_cost))) {
            best_matches = matches; best_cost = cost; best_origin = column[i].origin;
            best_ref_stop = i; best_query_stop = n;
        }
    }
    if (best_cost == m + n) return 0;
    out[0] = 0; out[1] = best_ref_stop; out[2] = best_origin; out[3] = best_query_stop;
    out[4] = best_matches; out[5] = best_cost;
    return 1;
}

/* cutadapt/adapters.py Adapter.match_to(read) for where = BACK, no wildcards:
 * exact str.find first, else Aligner.locate.  Returns 1 + (rstart, matches). */
int orc_match_back(const char *adapter, int m, const char *read, int n,
                   double max_error_rate, int min_overlap, int *rstart,
                   int *matches)
{
    if (m <= n) {
        const char *p = (n > 0) ? memmem(read, (size_t)n, adapter, (size_t)m) : NULL;
        if (p) { *rstart = (int)(p - read); *matches = m; return 1; }
    }
    int out[6];
    if (!orc_locate(adapter, m, read, n, max_error_rate, min_overlap, out)) return 0;
    *rstart = out[2]; *matches = out[4];
    return 1;
}

typedef struct {
    int mode;               /* 0 = 'none', 1 = BACK adapters, 2 = '+N' unconditional cut */
    int n_adapters;
    char *adapters[ORC_MAX_ADAPTERS];
    int adapter_len[ORC_MAX_ADAPTERS];
    int uncond;             /* N of '+N' */
    double error_rate;      /* trim_file.py:32 */
    int min_overlap;        /* cutadapt default 3 */
    int min_length;         /* trim_file.py:33 */
    int qcut;               /* trim_file.py:30 */
} OrcTrimCfg;

/* Worker.run modifier chain for one read (trim_file.py:52-55):
 * QualityTrimmer(0,10,base) then AdapterCutter(times=1) / UnconditionalCutter,
 * then the min_length test.  seq is matched upper-cased (match_to does
 * read.sequence.upper()).  Returns 1 if kept, and [start,stop) of the kept
 * segment in the ORIGINAL read. */
int orc_trim_read(const OrcTrimCfg *cfg, const char *seq, const char *qual,
                  int len, int base, int *start_out, int *stop_out)
{
    int start, stop;
    orc_quality_trim_index(qual, len, 0, cfg->qcut, base, &start, &stop);
    if (cfg->mode == 2) {
        int L = cfg->uncond, cur = stop - start;
        if (L > 0) { start += (L < cur ? L : cur); }
        else if (L < 0) { int cut = -L; stop -= (cut < cur ? cut : cur); }
    } else if (cfg->mode == 1) {
        int n = stop - start;
        char up[n > 0 ? n : 1];
        for (int i = 0; i < n; i++) up[i] = (char)toupper((unsigned char)seq[start + i]);
        int best_rstart = -1, best_matches = -1;
        for (int a = 0; a < cfg->n_adapters; a++) {
            int rs, mt;
            if (orc_match_back(cfg->adapters[a], cfg->adapter_len[a], up, n,
                               cfg->error_rate, cfg->min_overlap, &rs, &mt)) {
                if (best_rstart < 0 || mt > best_matches) { best_rstart = rs; best_matches = mt; }
            }
        }
        if (best_rstart >= 0) stop = start + best_rstart;
    }
    *start_out = start; *stop_out = stop;
    return (stop - start) >= cfg->min_length;
}

/* ------------------------------------------------------------------ */
/* stage 1b: collapse (quantReads.py:3-24)                              */
/* ------------------------------------------------------------------ */

typedef struct {
    uint64_t *hash;     /* 0 = empty */
    uint32_t *uid;
    uint64_t cap, n;
} OrcMap;

typedef struct {
    char *arena; uint64_t arena_len, arena_cap;
    uint64_t *seq_off; uint32_t *seq_len;
    uint32_t *counts;   /* [n_unique][n_samples] */
    int32_t *round_of;  /* 0 = un-annotated, else round index + 1 */
    int32_t *ref_of, *pos_of, *mm_of;
    uint64_t n_unique, cap_unique;
    OrcMap map;
} OrcUniq;

typedef struct {
    char *seq;              /* concatenated, upper-cased, non-ACGT -> 'N' */
    uint64_t *off;          /* n_refs + 1 */
    int n_refs;
    uint64_t total;
    uint32_t *sorted_pos;   /* positions sorted by 14-mer key */
    uint32_t *sorted_key;
} OrcLib;

typedef struct {
    int n_samples, spike_in;
    int flip_ties;          /* test knob: invert every round's (ref, pos) tie preference (orc_set_tie_policy) */
    OrcTrimCfg trim;
    OrcUniq uq;
    OrcLib *libs[ORC_MAX_LIBS];
    uint64_t *len_hist;     /* [ORC_MAX_READLEN+1][n_samples] */
    uint64_t *processed, *kept;
    char err[256];
} OrcCtx;

static uint64_t fnv1a(const char *s, int n)
{
    uint64_t h = 1469598103934665603ULL;
    for (int i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 1099511628211ULL; }
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ULL; h ^= h >> 32;
    return h ? h : 1;
}

static void map_grow(OrcUniq *u)
{
    uint64_t ncap = u->map.cap ? u->map.cap * 2 : (1u << 16);
    uint64_t *nh = calloc(ncap, sizeof(uint64_t));
    uint32_t *nu = malloc(ncap * sizeof(uint32_t));
    for (uint64_t i = 0; i < u->map.cap; i++) if (u->map.hash[i]) {
        uint64_t p = u->map.hash[i] & (ncap - 1);
        while (nh[p]) p = (p + 1) & (ncap - 1);
        nh[p] = u->map.hash[i]; nu[p] = u->map.uid[i];
    }
    free(u->map.hash); free(u->map.uid);
    u->map.hash = nh; u->map.uid = nu; u->map.cap = ncap;
}

static uint32_t uniq_upsert(OrcCtx *c, const char *s, int n)
{
    OrcUniq *u = &c->uq;
    if ((u->map.n + 1) * 2 > u->map.cap) map_grow(u);
    uint64_t h = fnv1a(s, n), p = h & (u->map.cap - 1);
    while (u->map.hash[p]) {
        if (u->map.hash[p] == h) {
            uint32_t id = u->map.uid[p];
            if (u->seq_len[id] == (uint32_t)n && memcmp(u->arena + u->seq_off[id], s, (size_t)n) == 0) return id;
        }
        p = (p + 1) & (u->map.cap - 1);
    }
    if (u->n_unique == u->cap_unique) {
        uint64_t nc = u->cap_unique ? u->cap_unique * 2 : (1u << 15);
        u->seq_off = realloc(u->seq_off, nc * sizeof(uint64_t));
        u->seq_len = realloc(u->seq_len, nc * sizeof(uint32_t));
        u->counts = realloc(u->counts, nc * (size_t)c->n_samples * sizeof(uint32_t));
        memset(u->counts + u->cap_unique * (size_t)c->n_samples, 0, (nc - u->cap_unique) * (size_t)c->n_samples * sizeof(uint32_t));
        u->cap_unique = nc;
    }
    if (u->arena_len + (uint64_t)n + 1 > u->arena_cap) {
        uint64_t nc = u->arena_cap ? u->arena_cap * 2 : (1u << 20);
        while (nc < u->arena_len + (uint64_t)n + 1) nc *= 2;
        u->arena = realloc(u->arena, nc); u->arena_cap = nc;
    }
    uint32_t id = (uint32_t)u->n_unique++;
    memcpy(u->arena + u->arena_len, s, (size_t)n);
    u->seq_off[id] = u->arena_len; u->seq_len[id] = (uint32_t)n;
    u->arena_len += (uint64_t)n;
    u->map.hash[p] = h; u->map.uid[p] = id; u->map.n++;
    return id;
}

OrcCtx *orc_new(int n_samples, int spike_in)
{
    OrcCtx *c = calloc(1, sizeof(OrcCtx));
    c->n_samples = n_samples; c->spike_in = spike_in;
    c->trim.mode = 0; c->trim.error_rate = 0.12; c->trim.min_overlap = 3;
    c->trim.min_length = 16; c->trim.qcut = 10;
    c->len_hist = calloc((size_t)(ORC_MAX_READLEN + 1) * (size_t)n_samples, sizeof(uint64_t));
    c->processed = calloc((size_t)n_samples, sizeof(uint64_t));
    c->kept = calloc((size_t)n_samples, sizeof(uint64_t));
    return c;
}

static void lib_free(OrcLib *l)
{
    if (!l) return;
    free(l->seq); free(l->off); free(l->sorted_pos); free(l->sorted_key); free(l);
}

void orc_free(OrcCtx *c)
{
    if (!c) return;
    for (int i = 0; i < ORC_MAX_ADAPTERS; i++) free(c->trim.adapters[i]);
    for (int i = 0; i < ORC_MAX_LIBS; i++) lib_free(c->libs[i]);
    free(c->uq.arena); free(c->uq.seq_off); free(c->uq.seq_len); free(c->uq.counts);
    free(c->uq.round_of); free(c->uq.ref_of); free(c->uq.pos_of); free(c->uq.mm_of);
    free(c->uq.map.hash); free(c->uq.map.uid);
    free(c->len_hist); free(c->processed); free(c->kept);
    free(c);
}

const char *orc_last_error(OrcCtx *c) { return c->err; }

/* Adapter spec as handed to trim_file (trim_file.py:34-41; __main__.py:124-127
 * has already expanded 'illumina'):  'none' | '+N' | seq[,seq...] (3' adapters).
 * Adapter.__init__ upper-cases and maps U->T. */
int orc_set_adapter(OrcCtx *c, const char *spec)
{
    OrcTrimCfg *t = &c->trim;
    for (int i = 0; i < ORC_MAX_ADAPTERS; i++) { free(t->adapters[i]); t->adapters[i] = NULL; }
    t->n_adapters = 0;
    if (spec[0] == '+') { t->mode = 2; t->uncond = atoi(spec); return 0; }
    if (strcmp(spec, "none") == 0) { t->mode = 0; return 0; }
    t->mode = 1;
    const char *p = spec;
    while (*p) {
        const char *q = strchr(p, ',');
        size_t n = q ? (size_t)(q - p) : strlen(p);
        if (n == 0 || t->n_adapters == ORC_MAX_ADAPTERS) { snprintf(c->err, sizeof c->err, "bad adapter spec"); return -1; }
        if (isdigit((unsigned char)p[0])) { snprintf(c->err, sizeof c->err, "Adapter sequence starts with a digit"); return -1; }
        char *a = malloc(n + 1);
        for (size_t i = 0; i < n; i++) { char ch = (char)toupper((unsigned char)p[i]); a[i] = (ch == 'U') ? 'T' : ch; }
        a[n] = 0;
        t->adapters[t->n_adapters] = a; t->adapter_len[t->n_adapters] = (int)n; t->n_adapters++;
        p += n; if (*p == ',') p++;
    }
    return 0;
}

/* Test knob: the constants trim_file.py hard-codes (:30,:32,:33) and cutadapt's min_overlap default. */
void orc_set_trim_params(OrcCtx *c, double error_rate, int min_overlap, int min_length, int qcut)
{
    c->trim.error_rate = error_rate; c->trim.min_overlap = min_overlap;
    c->trim.min_length = min_length; c->trim.qcut = qcut;
}

/* trim_file.py:104-109 -- workers are constructed at index 0, so the quality
 * base they use is decided by read #0 alone: 64 iff it has a char > 'J'. */
int orc_detect_phred(const char *buf, uint64_t n)
{
    uint64_t p = 0; int line = 0;
    while (p < n && line < 3) { if (buf[p] == '\n') line++; p++; }
    for (; p < n && buf[p] != '\n'; p++) if ((unsigned char)buf[p] > 74) return 64;
    return 33;
}

/* trim_file (trim_file.py:89-134) + quantReads (quantReads.py:3-24) over one
 * in-memory FASTQ buffer (4-line records).  Returns 0 or -1 on format error. */
int orc_feed_fastq(OrcCtx *c, int sample, const char *buf, uint64_t n, int phred_base, int nthreads)
{
    /* line index */
    uint64_t nl = 0;
    for (uint64_t i = 0; i < n; i++) nl += (buf[i] == '\n');
    if (n > 0 && buf[n - 1] != '\n') nl++;
    if (nl % 4 != 0) { snprintf(c->err, sizeof c->err, "FASTQ line count %llu not a multiple of 4", (unsigned long long)nl); return -1; }
    uint64_t nrec = nl / 4;
    uint64_t *ls = malloc((nl + 1) * sizeof(uint64_t));
    {
        uint64_t li = 0; ls[li++] = 0;
        for (uint64_t i = 0; i < n; i++) if (buf[i] == '\n' && li <= nl) ls[li++] = i + 1;
        if (n > 0 && buf[n - 1] != '\n') ls[nl] = n + 1;
    }
    int32_t *st = malloc(nrec * sizeof(int32_t)), *sp = malloc(nrec * sizeof(int32_t));
    int bad = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t r = 0; r < (int64_t)nrec; r++) {
        uint64_t a0 = ls[4 * r], a1 = ls[4 * r + 1], a2 = ls[4 * r + 2], a3 = ls[4 * r + 3], a4 = ls[4 * r + 4];
        if (buf[a0] != '@' || buf[a2] != '+') { bad = 1; st[r] = 0; sp[r] = -1; continue; }
        int64_t sl = (int64_t)(a2 - 1 - a1), ql = (int64_t)(a4 - 1 - a3);
        while (sl > 0 && buf[a1 + sl - 1] == '\r') sl--;
        while (ql > 0 && buf[a3 + ql - 1] == '\r') ql--;
        if (sl != ql || sl > ORC_MAX_READLEN) { bad = 1; st[r] = 0; sp[r] = -1; continue; }
        int s0, s1;
        int keep = orc_trim_read(&c->trim, buf + a1, buf + a3, (int)sl, phred_base, &s0, &s1);
        st[r] = s0; sp[r] = keep ? s1 : -1;
    }
    if (bad) { free(ls); free(st); free(sp); snprintf(c->err, sizeof c->err, "FASTQ format error"); return -1; }
    uint64_t kept = 0;
    for (uint64_t r = 0; r < nrec; r++) {
        if (sp[r] < 0) continue;
        kept++;
        int len = sp[r] - st[r];
        uint32_t id = uniq_upsert(c, buf + ls[4 * r + 1] + st[r], len);
        c->uq.counts[(size_t)id * c->n_samples + sample]++;
        c->len_hist[(size_t)len * c->n_samples + sample]++;
    }
    c->processed[sample] += nrec; c->kept[sample] += kept;
    free(ls); free(st); free(sp);
    return 0;
}

/* Trim-only entry used by tests: returns per-record (start, stop) or stop=-1. */
int orc_trim_fastq(OrcCtx *c, const char *buf, uint64_t n, int phred_base, int32_t *start, int32_t *stop, uint64_t max_rec)
{
    uint64_t p = 0, r = 0;
    while (p < n && r < max_rec) {
        uint64_t e[4]; uint64_t q = p; int ok = 1;
        uint64_t b[4];
        for (int k = 0; k < 4; k++) {
            b[k] = q;
            while (q < n && buf[q] != '\n') q++;
            e[k] = q; if (q < n) q++; else if (k < 3) ok = 0;
        }
        if (!ok) break;
        int sl = (int)(e[1] - b[1]);
        while (sl > 0 && buf[b[1] + sl - 1] == '\r') sl--;
        int s0, s1;
        int keep = orc_trim_read(&c->trim, buf + b[1], buf + b[3], sl, phred_base, &s0, &s1);
        start[r] = s0; stop[r] = keep ? s1 : -1;
        r++; p = q;
    }
    return (int)r;
}

/* Forget reads and unique sequences; keep adapter and libraries (bench loops). */
void orc_reset(OrcCtx *c)
{
    OrcUniq *u = &c->uq;
    free(u->arena); free(u->seq_off); free(u->seq_len); free(u->counts);
    free(u->round_of); free(u->ref_of); free(u->pos_of); free(u->mm_of);
    free(u->map.hash); free(u->map.uid);
    memset(u, 0, sizeof *u);
    memset(c->len_hist, 0, (size_t)(ORC_MAX_READLEN + 1) * (size_t)c->n_samples * sizeof(uint64_t));
    memset(c->processed, 0, (size_t)c->n_samples * sizeof(uint64_t));
    memset(c->kept, 0, (size_t)c->n_samples * sizeof(uint64_t));
}

uint64_t orc_n_unique(OrcCtx *c) { return c->uq.n_unique; }
uint64_t orc_seq_bytes(OrcCtx *c) { return c->uq.arena_len; }

void orc_stats(OrcCtx *c, uint64_t *processed, uint64_t *kept)
{
    for (int s = 0; s < c->n_samples; s++) { processed[s] = c->processed[s]; kept[s] = c->kept[s]; }
}

/* ------------------------------------------------------------------ */
/* stage 2: bowtie acceptance rules                                    */
/* ------------------------------------------------------------------ */

#define ORC_K 14

static inline int base2(char ch)
{
    switch (ch) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}

/* Library = FASTA entries (names stay with the caller; the oracle speaks ref
 * indices).  bowtie-build upper-cases; non-ACGT stretches are unindexed. */
int orc_load_lib(OrcCtx *c, int lib_id, const char *seq_concat, const uint64_t *offsets, int n_refs)
{
    if (lib_id < 0 || lib_id >= ORC_MAX_LIBS) return -1;
    lib_free(c->libs[lib_id]);
    OrcLib *l = calloc(1, sizeof(OrcLib));
    l->n_refs = n_refs; l->total = offsets[n_refs];
    l->off = malloc(((size_t)n_refs + 1) * sizeof(uint64_t));
    memcpy(l->off, offsets, ((size_t)n_refs + 1) * sizeof(uint64_t));
    l->seq = malloc(l->total + 1);
    for (uint64_t i = 0; i < l->total; i++) {
        char ch = (char)toupper((unsigned char)seq_concat[i]);
        if (ch == 'U') ch = 'T';
        l->seq[i] = (base2(ch) >= 0) ? ch : 'N';
    }
    /* sorted position index keyed by the (A-padded) 14-mer starting at each position */
    uint64_t T = l->total;
    uint32_t *key = malloc((T + 1) * sizeof(uint32_t)), *pos = malloc((T + 1) * sizeof(uint32_t));
    uint32_t *key2 = malloc((T + 1) * sizeof(uint32_t)), *pos2 = malloc((T + 1) * sizeof(uint32_t));
    int r = 0;
    for (uint64_t g = 0; g < T; g++) {
        while (g >= l->off[r + 1]) r++;
        uint32_t kk = 0; int dead = 0;
        for (int t = 0; t < ORC_K; t++) {
            int b = 0;
            if (!dead) {
                if (g + t >= l->off[r + 1]) dead = 1;
                else { b = base2(l->seq[g + t]); if (b < 0) { dead = 1; b = 0; } }
            }
            kk = (kk << 2) | (uint32_t)b;
        }
        key[g] = kk; pos[g] = (uint32_t)g;
    }
    /* LSD radix sort, 2 passes of 14 bits (stable => positions ascending in a key) */
    for (int pass = 0; pass < 2; pass++) {
        int sh = pass * 14; uint32_t mask = (1u << 14) - 1;
        uint64_t *cnt = calloc((1u << 14) + 1, sizeof(uint64_t));
        for (uint64_t g = 0; g < T; g++) cnt[((key[g] >> sh) & mask) + 1]++;
        for (uint32_t b = 0; b < (1u << 14); b++) cnt[b + 1] += cnt[b];
        for (uint64_t g = 0; g < T; g++) { uint64_t d = cnt[(key[g] >> sh) & mask]++; key2[d] = key[g]; pos2[d] = pos[g]; }
        uint32_t *t1 = key; key = key2; key2 = t1; t1 = pos; pos = pos2; pos2 = t1;
        free(cnt);
    }
    free(key2); free(pos2);
    l->sorted_key = key; l->sorted_pos = pos;
    c->libs[lib_id] = l;
    return 0;
}

static uint64_t lower_bound_u32(const uint32_t *a, uint64_t n, uint64_t v)
{
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = (lo + hi) / 2; if ((uint64_t)a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

static int ref_of_gpos(const OrcLib *l, uint64_t g)
{
    int lo = 0, hi = l->n_refs;  /* find r with off[r] <= g < off[r+1] */
    while (hi - lo > 1) { int mid = (lo + hi) / 2; if (l->off[mid] <= g) lo = mid; else hi = mid; }
    return lo;
}

/* One candidate placement of read[0:n] at global library offset g.
 * mode 0 = bowtie "-n N" (seed -l 28, -e 70 with FASTA quals => 30 per
 * mismatch => at most 2 mismatches overall); mode 1 = "-v V".
 * N in the read mismatches everything; a window touching an ambiguous
 * reference base or crossing the entry end is invalid. */
static int verify_at(const OrcLib *l, uint64_t g, const char *read, int n, int mode, int param, int *mm_out)
{
    if (g + (uint64_t)n > l->total) return 0;
    int r = ref_of_gpos(l, g);
    if (g + (uint64_t)n > l->off[r + 1]) return 0;
    int mm_seed = 0, mm_total = 0;
    for (int i = 0; i < n; i++) {
        char rc = l->seq[g + i];
        if (rc == 'N') return 0;
        if (read[i] != rc) { mm_total++; if (i < 28) mm_seed++; if (mm_total > 3) return 0; }
    }
    if (mode == 0) { if (mm_seed > param || 30 * mm_total > 70) return 0; }
    else { if (mm_total > param) return 0; }
    *mm_out = mm_total;
    return 1;
}

/* Exhaustive search for read[0:n] in library l under the round's rule.
 * Tie policy: fewest mismatches, then lowest
 * (ref_idx, pos) -- or highest when pick_highest (rounds 2 and 3, where the
 * reference keeps the LAST line bowtie -a --best --strata prints,
 * runAnnotationPipeline.py:27).  Returns 1 and (gpos, mm) on a hit. */
static int map_read(const OrcLib *l, const char *read, int n, int mode, int param, int pick_highest, uint64_t *gpos_out, int *mm_out)
{
    if (!l || n <= 0) return 0;
    int region = (mode == 0) ? (n < 28 ? n : 28) : n;
    int nseeds = param + 1;
    int slen = region / nseeds; if (slen > ORC_K) slen = ORC_K;
    if (slen < 1) return 0;
    int found = 0, best_mm = 99; uint64_t best_g = 0;
    for (int s = 0; s < nseeds; s++) {
        int so = s * slen;
        uint64_t code = 0; int okseed = 1;
        for (int t = 0; t < slen; t++) { int b = base2(read[so + t]); if (b < 0) { okseed = 0; break; } code = (code << 2) | (uint64_t)b; }
        if (!okseed) continue;   /* a seed holding N cannot match exactly */
        int sh = 2 * (ORC_K - slen);
        uint64_t lo = lower_bound_u32(l->sorted_key, l->total, code << sh);
        uint64_t hi = lower_bound_u32(l->sorted_key, l->total, (code + 1) << sh);
        for (uint64_t x = lo; x < hi; x++) {
            uint64_t p = l->sorted_pos[x];
            if (p < (uint64_t)so) continue;
            uint64_t g = p - (uint64_t)so; int mm;
            if (!verify_at(l, g, read, n, mode, param, &mm)) continue;
            if (!found || mm < best_mm || (mm == best_mm && (pick_highest ? g > best_g : g < best_g))) { found = 1; best_mm = mm; best_g = g; }
        }
    }
    if (found) { *gpos_out = best_g; *mm_out = best_mm; }
    return found;
}

/* Brute force over every library position: pins map_read in tests. */
int orc_map_bruteforce(OrcCtx *c, int lib_id, const char *read, int n, int mode, int param, int pick_highest, int32_t *ref_out, int32_t *pos_out, int32_t *mm_out)
{
    OrcLib *l = c->libs[lib_id];
    int found = 0, best_mm = 99; uint64_t best_g = 0;
    if (!l) return 0;
    for (uint64_t g = 0; g + (uint64_t)n <= l->total; g++) {
        int mm;
        if (!verify_at(l, g, read, n, mode, param, &mm)) continue;
        if (!found || mm < best_mm || (mm == best_mm && (pick_highest ? g > best_g : g < best_g))) { found = 1; best_mm = mm; best_g = g; }
    }
    if (found) { int r = ref_of_gpos(l, best_g); *ref_out = r; *pos_out = (int32_t)(best_g - l->off[r]); *mm_out = best_mm; }
    return found;
}

int orc_map_one(OrcCtx *c, int lib_id, const char *read, int n, int mode, int param, int pick_highest, int32_t *ref_out, int32_t *pos_out, int32_t *mm_out)
{
    OrcLib *l = c->libs[lib_id];
    uint64_t g; int mm;
    if (!map_read(l, read, n, mode, param, pick_highest, &g, &mm)) return 0;
    int r = ref_of_gpos(l, g); *ref_out = r; *pos_out = (int32_t)(g - l->off[r]); *mm_out = mm;
    return 1;
}

/* ------------------------------------------------------------------ */
/* genome passes -- both strands, every valid    */
/* alignment enumerated (bowtie -a)                                        */
/* ------------------------------------------------------------------ */

/* Like verify_at with a configurable seed length (-l): mode 0 = "-n N" (at most N
 * mismatches in the first min(seedlen, n) bases, 30 per mismatch against -e 70 overall),
 * mode 1 = "-v V". */
static int verify_at_l(const OrcLib *l, uint64_t g, const char *read, int n, int mode, int param, int seedlen, int *mm_out)
{
    if (g + (uint64_t)n > l->total) return 0;
    int r = ref_of_gpos(l, g);
    if (g + (uint64_t)n > l->off[r + 1]) return 0;
    int mm_seed = 0, mm_total = 0;
    for (int i = 0; i < n; i++) {
        char rc = l->seq[g + i];
        if (rc == 'N') return 0;
        if (read[i] != rc) { mm_total++; if (i < seedlen) mm_seed++; if (mm_total > 3) return 0; }
    }
    if (mode == 0) { if (mm_seed > param || 30 * mm_total > 70) return 0; }
    else { if (mm_total > param) return 0; }
    *mm_out = mm_total;
    return 1;
}

/* All alignments of read[0 : n - trim3] (bowtie -3) and, unless norc, of its reverse
 * complement, against library lib_id: counts[k] = number of alignments with k mismatches
 * (k = 0..3).  Uses:
 *   writeDataToCSV.py:1260  `-n 1 -f -a -3 2`  A-to-I candidate reads against the genome; the caller keeps
 *                           a read when its lowest occupied stratum holds exactly one alignment (:1273-1281)
 *   writeDataToCSV.py:1488  `-n 0 -f -a -3 2`
 *   __main__.py:538         `-n 0 -m 3 -l 25 -a --best` (unannotated reads, both strands; -m: reads with more
 *                           than 3 alignments are suppressed -- sum(counts) > 3)
 * A reverse-strand alignment is a forward alignment of the reverse complement: bowtie's seed (the read's 5'
 * end) then sits at the RIGHT end of the window, so the seed test applies to the revcomp's LAST seedlen bases.
 * Exhaustive scan (test oracle: small genomes only). */
int orc_map_genome_all(OrcCtx *c, int lib_id, const char *read, int n, int mode, int param, int seedlen, int trim3,
                       int norc, uint32_t counts[4])
{
    OrcLib *l = (lib_id >= 0 && lib_id < ORC_MAX_LIBS) ? c->libs[lib_id] : NULL;
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    n -= trim3;
    if (!l || n <= 0) return 0;
    char fw[n + 1], rc[n + 1];
    for (int i = 0; i < n; i++) {
        char ch = (char)toupper((unsigned char)read[i]);
        fw[i] = (base2(ch) >= 0) ? ch : 'N';
    }
    for (int i = 0; i < n; i++) {
        char ch = fw[n - 1 - i];
        rc[i] = ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : ch == 'T' ? 'A' : 'N';
    }
    int sl = seedlen < n ? seedlen : n;
    uint32_t total = 0;
    for (uint64_t g = 0; g + (uint64_t)n <= l->total; g++) {
        int mm;
        if (verify_at_l(l, g, fw, n, mode, param, sl, &mm)) { counts[mm]++; total++; }
        if (!norc) {
            /* seed = last sl bases of the window: count seed mismatches there */
            int r = ref_of_gpos(l, g);
            if (g + (uint64_t)n > l->off[r + 1]) continue;
            int mm_seed = 0, mm_total = 0, bad = 0;
            for (int i = 0; i < n; i++) {
                char ref = l->seq[g + i];
                if (ref == 'N') { bad = 1; break; }
                if (rc[i] != ref) { mm_total++; if (i >= n - sl) mm_seed++; if (mm_total > 3) { bad = 1; break; } }
            }
            if (bad) continue;
            if (mode == 0) { if (mm_seed > param || 30 * mm_total > 70) continue; }
            else { if (mm_total > param) continue; }
            counts[mm_total]++; total++;
        }
    }
    return (int)total;
}

/* Round table, runAnnotationPipeline.py:574-599.  lib ids:
 * 0 mirna 1 hairpin 2 mature_trna 3 pre_trna 4 snorna 5 rrna 6 ncrna_others
 * 7 mrna 8 (mirna again) 9 spike-in. */
typedef struct { int lib, len_filter, mode, param, trim5, trim3, highest, polyT; } OrcRound;
static const OrcRound ORC_ROUNDS[10] = {
    {0, -26, 0, 0, 0, 0, 0, 0},   /* mirna   -n 0            len < 26 */
    {1,  25, 0, 1, 0, 0, 0, 0},   /* hairpin -n 1            len > 25 */
    {2,   0, 1, 1, 0, 0, 1, 0},   /* mature tRNA -v 1 -a --best --strata */
    {3,   0, 1, 0, 0, 0, 1, 1},   /* pre-tRNA: poly-T strip, -v 0 -a --best --strata */
    {4,   0, 0, 1, 0, 0, 0, 0},   /* snoRNA  -n 1 */
    {5,   0, 0, 1, 0, 0, 0, 0},   /* rRNA    -n 1 */
    {6,   0, 0, 1, 0, 0, 0, 0},   /* ncRNA others -n 1 */
    {7,   0, 0, 0, 0, 0, 0, 0},   /* mRNA    -n 0 */
    {0,   0, 1, 2, 1, 2, 0, 0},   /* mirna -5 1 -3 2 -v 2 --best (isomiR) */
    {9,   0, 0, 0, 0, 0, 0, 0},   /* spike-in -n 0 */
};

/* Tie-policy knob: bowtie picks among equally good hits pseudo-randomly, so the
 * count tables must not depend on WHICH of them is reported.  flip = 1 inverts the (ref, pos) preference
 * of every round (lowest <-> highest within the fewest-mismatch stratum); tests/test_tie_policy.py runs the
 * pipeline under both policies and checks what is and is not invariant. */
void orc_set_tie_policy(OrcCtx *c, int flip) { c->flip_ties = flip ? 1 : 0; }

/* runAnnotationPipeline.py:636-705: rounds in order over the still
 * un-annotated unique sequences (writeSeqToAnnot :543-554), updateAnnotDic
 * :341-345, poly-T special round :664-704. */
int orc_annotate(OrcCtx *c, int nthreads, uint64_t *round_processed, uint64_t *round_aligned)
{
    OrcUniq *u = &c->uq;
    uint64_t U = u->n_unique;
    free(u->round_of); free(u->ref_of); free(u->pos_of); free(u->mm_of);
    u->round_of = calloc(U + 1, sizeof(int32_t)); u->ref_of = calloc(U + 1, sizeof(int32_t));
    u->pos_of = calloc(U + 1, sizeof(int32_t)); u->mm_of = calloc(U + 1, sizeof(int32_t));
    int nrounds = c->spike_in ? 10 : 9;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    for (int ri = 0; ri < nrounds; ri++) {
        const OrcRound R = ORC_ROUNDS[ri];
        const OrcLib *l = c->libs[R.lib];
        uint64_t nproc = 0, naln = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+:nproc, naln)
        for (int64_t id = 0; id < (int64_t)U; id++) {
            if (u->round_of[id] != 0) continue;
            int len = (int)u->seq_len[id];
            if (R.len_filter < 0 && !(len < -R.len_filter)) continue;
            if (R.len_filter > 0 && !(len > R.len_filter)) continue;
            const char *s = u->arena + u->seq_off[id];
            int n = len; const char *q = s;
            if (R.polyT) {
                int t = 0;
                while (t < len && s[len - 1 - t] == 'T') t++;
                if (t < 3) continue;              /* re.search('T{3,}$') :672 */
                n = len - t;
                if (n < 11) continue;             /* :680 */
            } else {
                q = s + R.trim5; n = len - R.trim5 - R.trim3;
            }
            nproc++;
            uint64_t g; int mm;
            if (l && map_read(l, q, n, R.mode, R.param, R.highest ^ c->flip_ties, &g, &mm)) {
                int r = ref_of_gpos(l, g);
                u->round_of[id] = ri + 1; u->ref_of[id] = r;
                u->pos_of[id] = (int32_t)(g - l->off[r]); u->mm_of[id] = mm;
                naln++;
            }
        }
        if (round_processed) round_processed[ri] = nproc;
        if (round_aligned) round_aligned[ri] = naln;
    }
    return 0;
}

/* Export the collapsed table (seqDic, __main__.py:318-321). */
void orc_export(OrcCtx *c, char *seqs, uint64_t *seq_off, uint32_t *seq_len, uint32_t *counts,
                int32_t *round_of, int32_t *ref_of, int32_t *pos_of, int32_t *mm_of)
{
    OrcUniq *u = &c->uq;
    memcpy(seqs, u->arena, u->arena_len);
    memcpy(seq_off, u->seq_off, u->n_unique * sizeof(uint64_t));
    memcpy(seq_len, u->seq_len, u->n_unique * sizeof(uint32_t));
    memcpy(counts, u->counts, u->n_unique * (size_t)c->n_samples * sizeof(uint32_t));
    if (u->round_of) {
        if (round_of) memcpy(round_of, u->round_of, u->n_unique * sizeof(int32_t));
        if (ref_of) memcpy(ref_of, u->ref_of, u->n_unique * sizeof(int32_t));
        if (pos_of) memcpy(pos_of, u->pos_of, u->n_unique * sizeof(int32_t));
        if (mm_of) memcpy(mm_of, u->mm_of, u->n_unique * sizeof(int32_t));
    }
}

void orc_len_hist(OrcCtx *c, uint64_t *out, int max_len)
{
    for (int L = 0; L <= max_len && L <= ORC_MAX_READLEN; L++)
        for (int s = 0; s < c->n_samples; s++)
            out[(size_t)L * c->n_samples + s] = c->len_hist[(size_t)L * c->n_samples + s];
}

/* summarize.py:34-66 integer tallies.  lib_totals[s][12]:
 * 0 trimmedUniq 1 mirnaReads 2 hairpin 3 mature tRNA 4 pre-tRNA 5 snoRNA
 * 6 rRNA 7 ncRNA others 8 mRNA 9 spikeIn 10 remReads 11 (unused);
 * mir_quant / mir_iscan are [n_mirna_refs][n_samples]. */
void orc_tally(OrcCtx *c, int64_t *mir_quant, int64_t *mir_iscan, int64_t *lib_totals)
{
    OrcUniq *u = &c->uq; int S = c->n_samples;
    int nmir = c->libs[0] ? c->libs[0]->n_refs : 0;
    memset(mir_quant, 0, (size_t)nmir * S * sizeof(int64_t));
    memset(mir_iscan, 0, (size_t)nmir * S * sizeof(int64_t));
    memset(lib_totals, 0, (size_t)S * 12 * sizeof(int64_t));
    for (uint64_t id = 0; id < u->n_unique; id++) {
        int ro = u->round_of ? u->round_of[id] : 0;
        for (int s = 0; s < S; s++) {
            int64_t q = u->counts[id * (size_t)S + s];
            if (q == 0) continue;
            int64_t *t = lib_totals + (size_t)s * 12;
            t[0]++;
            if (ro == 1 || ro == 9) {
                t[1] += q;
                mir_quant[(size_t)u->ref_of[id] * S + s] += q;
                if (ro == 1) mir_iscan[(size_t)u->ref_of[id] * S + s] += q;
            } else if (ro >= 2 && ro <= 8) t[ro] += q;
            else if (ro == 10) t[9] += q;
            else t[10] += q;
        }
    }
}
