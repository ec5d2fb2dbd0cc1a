// capi.cu -- context management and the extern "C" entry points declared in
// include/mirge_b200.h.  Host-side plumbing only; the kernels live in
// k0_parse.cuh (FASTQ -> packed records), k1_trim.cuh (trim + collapse),
// k2_map.cuh (mapping rounds) and k3_tally.cuh (segmented tallies).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <ctype.h>
#include <stddef.h>
#include <string>
#include <vector>

#include "../../include/mirge_b200.h"
#include "common.cuh"
#include "scan.cuh"
#include "k0_parse.cuh"
#include "k1_trim.cuh"
#include "k1_kernel.cuh"
#include "k2_map.cuh"
#include "k2_genome.cuh"
#include "k3_tally.cuh"

enum KernelId {
    KID_K0_COUNT = 0, KID_K0_SCAN, KID_K0_LINES,          // 0..2: retired (round 1's three-pass K0)
    KID_K0_PACK, KID_K1_TRIM, KID_K3_KEYS, KID_K3_SORT,
    KID_K3_BOUNDS, KID_K3_REDUCE, KID_K3_FINAL, KID_LIB_BUILD, KID_EXCHANGE, KID_H2D, KID_RESET, KID_GENOME, KID_IMPORT,
    KID_K2_ROUND0 = 16,   // .. KID_K2_ROUND0 + 9
    KID_K1_CLASSIFY = 26, KID_K1_SCAN, KID_K1_GENERAL, KID_K1_SLOW,
    KID_N = 30
};
static const char *KERNEL_NAMES[MIRGE_N_KERNELS] = {
    "", "", "", "k0_parse_pack", "k1_trim_collapse", "k3_make_keys",
    "k3_radix_sort", "k3_seg_bounds", "k3_seg_reduce", "k3_finalize", "lib_build", "exchange", "h2d_copy", "reset_memset", "k2g_genome_all_hits", "exchange_import",
    "k2_round0_mirna_exact", "k2_round1_hairpin", "k2_round2_mature_trna", "k2_round3_pre_trna", "k2_round4_snorna",
    "k2_round5_rrna", "k2_round6_ncrna_others", "k2_round7_mrna", "k2_round8_isomir", "k2_round9_spikein",
    "k1a_classify_collapse", "", "k1c_general_search", "k1d_slow_search", "", ""};

struct DevLib {
    u32 max_ref_len = 0;
    int has_amb = 0;
    u64 *seq = nullptr, *amb = nullptr;
    u32 *ref_off = nullptr, *blk_ref = nullptr;
    int n_refs = 0;
    u32 total = 0;
    int ntabs = 0;
    int tab_k[K2_MAX_TABS];
    u32 *tab_off[K2_MAX_TABS];
    u32 *tab_pos[K2_MAX_TABS];
    u32 *gx_off = nullptr, *gx_pos = nullptr;     // genome slot: one prefix-ordered table for every seed length
    int gx_kmax = 0;
    int ngap = 0;                                 // pair-of-pieces seed sets (k2_map.cuh:GapSet)
    GapSet gap[4];
};

struct ProfSpan { int kid; cudaEvent_t a, b; };

struct mirge_ctx {
    int device = 0, S = 1, spike_in = 0, W = 2, max_len = 64;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;             // H2D copies of mirge_feed_fastq
    std::vector<cudaEvent_t> copy_events;
    int max_warps = 148 * 64;                       // resident warps of the device (pads the warp-chunked queues)
    std::string err;
    // trim configuration
    TrimParams tp;
    K1Thr thr;
    K1Pre pre;
    int qual_cutoff = 10;
    int force_slow = 0;
    // K0 scratch
    unsigned char *d_text = nullptr; size_t text_cap = 0;
    u64 *d_k0ctl = nullptr; size_t k0ctl_cap = 0;   // exact path: tile ticket + tile statuses of the running launch
    u32 *d_k0info = nullptr; size_t k0info_cap = 0; // speculative path: newline count | claim per tile
    K0State *d_k0state = nullptr;
    u64 *d_bases = nullptr, *d_nmask = nullptr; u32 *d_meta = nullptr; size_t rec_cap = 0;
    void *d_q2 = nullptr, *d_q3 = nullptr; u64 *d_qn = nullptr, *d_part = nullptr;   // K1 class queues (records, rec_cap entries each), their fills; exchange partition counters
    // unique table
    void *d_slots = nullptr; u64 table_cap = 0;
    u64 max_unique = 0;
    u64 *d_useq = nullptr, *d_unmask = nullptr; u8 *d_ulen = nullptr; u32 *d_counts = nullptr;
    signed char *d_round_of = nullptr, *d_mm_of = nullptr; int *d_ref_of = nullptr, *d_pos_of = nullptr;
    // counters: [0] n_unique, [1] err flags (low 32 bits), [2..2+S) kept, [2+S..6+S) diagnostics (K1: all-column search, class 1, class 2; K0: regions that took the exact path),
    // [6+S..6+2S) processed
    u64 *d_ctr = nullptr;
    u64 *d_len_hist = nullptr;
    // mapping
    DevLib libs[MIRGE_N_LIBS];
    u32 *d_todo[2] = {nullptr, nullptr};
    u32 *d_todo_long = nullptr;     // third list: ids that skip round 0's length filter (k2_fill_range), later the list past the poly-T round
    u64 *d_todo_n = nullptr;        // [MIRGE_N_ROUNDS + 3]
    u64 *d_round_stats = nullptr;   // [2][MIRGE_N_ROUNDS]
    bool annotated = false;
    bool annotate_while_feeding = false;
    int pair_seeds = 16;                            // MIRGE_OPT_PAIR_SEEDS
    u64 *d_annot = nullptr;         // [0] ids annotated so far, [1] base and [2] size of the range being annotated
    // tallies
    i64 *d_mir_quant = nullptr, *d_mir_iscan = nullptr, *d_lib_totals = nullptr;
    int tally_nmir = -1;
    unsigned char *d_tally = nullptr; size_t tally_cap = 0;   // persistent K3 scratch
    bool tallied = false;
    // profiling
    bool prof = false;
    std::vector<ProfSpan> spans;
    double prof_ms[MIRGE_N_KERNELS];
    uint64_t prof_n[MIRGE_N_KERNELS];
    uint64_t n_launches = 0;
};

static thread_local std::string g_create_err;
static int annot_clear(mirge_ctx *ctx);
template <int W> static int annotate_impl(mirge_ctx *ctx);

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->err = b_;                                                                         \
            return MIRGE_E_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

static int fail(mirge_ctx *ctx, int code, const char *msg) { ctx->err = msg; return code; }

// Brackets `launches` kernel launches of family `kid` (CUDA events when profiling is on).
struct ProfScope {
    mirge_ctx *c; int kid; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(mirge_ctx *c_, int kid_, int launches = 1) : c(c_), kid(kid_)
    {
        c->n_launches += launches;
        if (c->prof) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, c->stream); }
    }
    ~ProfScope()
    {
        if (c->prof) { cudaEventRecord(b, c->stream); c->spans.push_back({kid, a, b}); }
    }
};

template <typename T>
static cudaError_t ensure(T **p, size_t *cap, size_t need)
{
    if (*cap >= need) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    size_t n = need + need / 8 + 64;
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e == cudaSuccess) *cap = n;
    return e;
}

static int grid_for(u64 n, int threads, int max_blocks)
{
    u64 b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > (u64)max_blocks) b = max_blocks;
    return (int)b;
}

// One resident wave: grid = SM count x CTAs that fit per SM (the kernels use grid-stride loops).
template <typename K>
static int wave_grid(K kernel, int threads, u64 n_items)
{
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return grid_for(n_items, threads, sms * per_sm);
}

template <int W>
static TableView<W> table_view(mirge_ctx *c)
{
    TableView<W> T;
    T.slots = (TEntry<W> *)c->d_slots; T.mask = c->table_cap - 1;
    { int lg = 0; while ((1ull << lg) < c->table_cap) lg++; T.shift = 64 - lg; }
    T.useq = c->d_useq; T.unmask = c->d_unmask; T.ulen = c->d_ulen; T.counts = c->d_counts;
    T.max_unique = c->max_unique; T.n_unique = c->d_ctr; T.err = (u32 *)(c->d_ctr + 1); T.S = c->S;
    return T;
}


// ---------------------------------------------------------------------------
// K1 prefilter tables (k1_prefilter.cuh).  For the "wildcard read" W^I + A[0..l):
// G[i][j][c] = most matches any alignment path from row 0 to cell (i, j) can
// collect with cost <= c (wildcard columns match every adapter base).  A candidate
// cell can outrank the exact-prefix candidate E only with >= l + 1 matches at a
// cost within its own threshold thr[i].
// ---------------------------------------------------------------------------
static void k1pre_dp(const unsigned char *A, int m, int I, int l, int K, std::vector<short> &G)
{
    const int n = I + l;
    const short NEG = -1000;
    auto at = [&](int i, int j, int c) -> short & { return G[((size_t)i * (n + 1) + j) * K + c]; };
    for (int j = 0; j <= n; j++) for (int c = 0; c < K; c++) at(0, j, c) = 0;
    for (int i = 1; i <= m; i++) for (int c = 0; c < K; c++) at(i, 0, c) = (c >= i) ? 0 : NEG;
    for (int j = 1; j <= n; j++)
        for (int i = 1; i <= m; i++)
            for (int c = 0; c < K; c++) {
                short best = NEG;
                const bool eq = (j <= I) || (A[i - 1] == A[j - 1 - I]);
                if (eq) { short v = at(i - 1, j - 1, c); if (v > NEG && v + 1 > best) best = v + 1; }
                if (c > 0) {
                    short v = at(i - 1, j - 1, c - 1); if (v > best) best = v;      // mismatch
                    v = at(i - 1, j, c - 1); if (v > best) best = v;                // adapter base unmatched
                    v = at(i, j - 1, c - 1); if (v > best) best = v;                // read base unmatched
                }
                at(i, j, c) = best;
            }
}

static void k1pre_build(const TrimParams &tp, const K1Thr &thr, K1Pre &pre)
{
    memset(&pre, 0, sizeof pre);
    if (tp.mode != 1 || tp.n_adapters != 1) return;
    const int m = tp.alen[0];
    if (m < 8 || m > 32) return;
    const int kmax = thr.v[m];
    if (kmax < 0 || m + kmax + 1 > K1PRE_MAXI) return;
    unsigned char A[32];
    for (int i = 0; i < m; i++) A[i] = (unsigned char)((tp.aw[0][0] >> (2 * i)) & 3ull);
    const int K = kmax + 1, icap = m + kmax + 1;
    std::vector<short> G((size_t)(m + 1) * (icap + m + 1) * K);
    for (int I = 0; I <= icap; I++) {
        for (int l = 4; l < m; l++) {
            if (thr.v[l] < 0) continue;                 // E itself is not an acceptable candidate
            const int n = I + l;
            k1pre_dp(A, m, I, l, K, G);
            auto at = [&](int i, int j, int c) { return G[((size_t)i * (n + 1) + j) * K + c]; };
            bool partial = false;
            int reach = -1;
            for (int i = l + 1; i <= m; i++)
                if (thr.v[i] >= 0 && at(i, n, thr.v[i]) >= l + 1) partial = true;
            for (int j = 1; j <= n; j++)
                if (at(m, j, kmax) >= l + 1) reach = j;
            if (!partial && reach < 0) pre.safe_len[I] |= 1u << l;
            if (I == icap) {                             // inserts this long behave like an unbounded one
                if (!partial) pre.safe_p |= 1u << l;
                pre.fullreach[l] = (u8)(reach > I ? reach - I : 0);
            }
        }
    }
    pre.m = m; pre.icap = icap;
    pre.np = K; pre.np_inv = (65536 + K - 1) / K;                       // pigeonhole pieces of the class-1 filter (k1_piece_hit_warp)
    for (int q = 0; q <= K; q++) pre.pb[q] = (u8)((q * m) / K);
    pre.anchor_rep = (u32)(tp.aw[0][0] & 0xFFull) * 0x01010101u;
    pre.enabled = 1;
}

// Definitions below take C linkage from their declarations in mirge_b200.h.

const char *mirge_last_error(const mirge_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int mirge_create(int device, int n_samples, int spike_in, uint64_t max_unique, int max_read_len, mirge_ctx **out)
{
    if (!out || n_samples < 1 || max_unique < 1 || max_unique > 0xFFFFFFF0ull || max_read_len < 1 || max_read_len > 128) {
        g_create_err = "mirge_create: bad argument (max_read_len must be <= 128)";
        return MIRGE_E_ARG;
    }
    mirge_ctx *ctx = new mirge_ctx();
    *out = ctx;
    ctx->device = device; ctx->S = n_samples; ctx->spike_in = spike_in ? 1 : 0;
    ctx->W = max_read_len <= 64 ? 2 : 4;
    ctx->max_len = ctx->W * 32;
    ctx->max_unique = max_unique;
    {
        int sms = 148, tps = 2048;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        cudaDeviceGetAttribute(&tps, cudaDevAttrMaxThreadsPerMultiProcessor, device);
        if (sms > 0 && tps > 0) ctx->max_warps = sms * (tps / 32);
    }
    memset(&ctx->tp, 0, sizeof ctx->tp);
    memset(&ctx->pre, 0, sizeof ctx->pre);
    ctx->tp.mode = 0; ctx->tp.error_rate = 0.12; ctx->tp.min_overlap = 3; ctx->tp.min_length = 16;
    memset(ctx->prof_ms, 0, sizeof ctx->prof_ms); memset(ctx->prof_n, 0, sizeof ctx->prof_n);
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
    u64 cap = 1024;
    while (cap < 2 * max_unique) cap <<= 1;
    ctx->table_cap = cap;
    const size_t esz = ctx->W == 2 ? sizeof(TEntry<2>) : sizeof(TEntry<4>);
    CK(cudaMalloc(&ctx->d_slots, cap * esz));
    CK(cudaMalloc(&ctx->d_useq, max_unique * ctx->W * sizeof(u64)));
    CK(cudaMalloc(&ctx->d_unmask, max_unique * (ctx->W / 2) * sizeof(u64)));
    CK(cudaMalloc(&ctx->d_ulen, max_unique));
    CK(cudaMalloc(&ctx->d_counts, max_unique * (size_t)n_samples * sizeof(u32)));
    CK(cudaMalloc(&ctx->d_round_of, max_unique));
    CK(cudaMalloc(&ctx->d_mm_of, max_unique));
    CK(cudaMalloc(&ctx->d_ref_of, max_unique * sizeof(int)));
    CK(cudaMalloc(&ctx->d_pos_of, max_unique * sizeof(int)));
    CK(cudaMalloc(&ctx->d_ctr, (6 + 2 * (size_t)n_samples) * sizeof(u64)));
    CK(cudaMalloc(&ctx->d_len_hist, (size_t)(ctx->max_len + 1) * n_samples * sizeof(u64)));
    // to-do lists: every unique once, plus the partly used chunk of every appending warp
    CK(cudaMalloc(&ctx->d_todo[0], max_unique * sizeof(u32)));
    CK(cudaMalloc(&ctx->d_todo[1], max_unique * sizeof(u32)));
    CK(cudaMalloc(&ctx->d_todo_long, max_unique * sizeof(u32)));
    CK(cudaMalloc(&ctx->d_todo_n, (MIRGE_N_ROUNDS + 3) * sizeof(u64)));
    CK(cudaMalloc(&ctx->d_k0state, sizeof(K0State)));
    CK(cudaMalloc(&ctx->d_qn, 4 * sizeof(u64)));
    CK(cudaMalloc(&ctx->d_round_stats, 2 * MIRGE_N_ROUNDS * sizeof(u64)));
    CK(cudaMalloc(&ctx->d_annot, 4 * sizeof(u64)));
    CK(cudaMalloc(&ctx->d_lib_totals, (size_t)n_samples * MIRGE_N_TOTALS * sizeof(i64)));
    CK(cudaMemsetAsync(ctx->d_slots, 0, cap * esz, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_counts, 0, max_unique * (size_t)n_samples * sizeof(u32), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_ctr, 0, (6 + 2 * (size_t)n_samples) * sizeof(u64), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_len_hist, 0, (size_t)(ctx->max_len + 1) * n_samples * sizeof(u64), ctx->stream));
    { int rc = annot_clear(ctx); if (rc) return rc; }
    CK(cudaStreamSynchronize(ctx->stream));
    return MIRGE_OK;
}

static void free_lib(DevLib &l)
{
    cudaFree(l.seq); cudaFree(l.amb); cudaFree(l.ref_off); cudaFree(l.blk_ref);
    for (int t = 0; t < l.ntabs; t++) { cudaFree(l.tab_off[t]); cudaFree(l.tab_pos[t]); }
    cudaFree(l.gx_off); cudaFree(l.gx_pos);
    for (int g = 0; g < l.ngap; g++) { cudaFree((void *)l.gap[g].off); cudaFree((void *)l.gap[g].pos); }
    l = DevLib();
}

void mirge_destroy(mirge_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &s : ctx->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (int i = 0; i < MIRGE_N_LIBS; i++) free_lib(ctx->libs[i]);
    cudaFree(ctx->d_text); cudaFree(ctx->d_k0ctl); cudaFree(ctx->d_k0info); cudaFree(ctx->d_k0state);
    cudaFree(ctx->d_bases); cudaFree(ctx->d_nmask); cudaFree(ctx->d_meta);
    cudaFree(ctx->d_q2); cudaFree(ctx->d_q3); cudaFree(ctx->d_qn); cudaFree(ctx->d_part);
    cudaFree(ctx->d_slots); cudaFree(ctx->d_useq); cudaFree(ctx->d_unmask); cudaFree(ctx->d_ulen); cudaFree(ctx->d_counts);
    cudaFree(ctx->d_round_of); cudaFree(ctx->d_mm_of); cudaFree(ctx->d_ref_of); cudaFree(ctx->d_pos_of);
    cudaFree(ctx->d_ctr); cudaFree(ctx->d_len_hist); cudaFree(ctx->d_todo[0]); cudaFree(ctx->d_todo[1]);
    cudaFree(ctx->d_todo_long); cudaFree(ctx->d_todo_n); cudaFree(ctx->d_round_stats); cudaFree(ctx->d_annot);
    cudaFree(ctx->d_mir_quant); cudaFree(ctx->d_mir_iscan); cudaFree(ctx->d_lib_totals); cudaFree(ctx->d_tally);
    for (auto e : ctx->copy_events) cudaEventDestroy(e);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int mirge_set_stream(mirge_ctx *ctx, void *cuda_stream)
{
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return MIRGE_OK;
}

int mirge_sync(mirge_ctx *ctx)
{
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    u64 flags = 0;
    CK(cudaMemcpy(&flags, ctx->d_ctr + 1, sizeof(u64), cudaMemcpyDeviceToHost));
    const u32 f = (u32)flags;
    if (f & ERRF_PARTIAL) return fail(ctx, MIRGE_E_FORMAT, "FASTQ chunk does not hold a whole number of 4-line records");
    if (f & ERRF_FORMAT) return fail(ctx, MIRGE_E_FORMAT, "FASTQ format error: record does not start with '@'/'+' or sequence and quality lengths differ");
    if (f & ERRF_TOOLONG) return fail(ctx, MIRGE_E_TOOLONG, "a read is longer than max_read_len");
    if (f & ERRF_FULL) return fail(ctx, MIRGE_E_FULL, "unique-sequence table full: raise max_unique");
    return MIRGE_OK;
}

int mirge_set_adapter(mirge_ctx *ctx, const char *spec, double error_rate, int min_overlap, int min_length, int qual_cutoff)
{
    if (!spec || min_length < 0 || min_length > 255) return fail(ctx, MIRGE_E_ARG, "mirge_set_adapter: bad argument");
    TrimParams tp;
    memset(&tp, 0, sizeof tp);
    tp.error_rate = error_rate; tp.min_overlap = min_overlap; tp.min_length = min_length;
    if (spec[0] == '+') {
        tp.mode = 2; tp.uncond = atoi(spec);       // UnconditionalCutter(int(adapter)), trim_file.py:34-35
    } else if (strcmp(spec, "none") == 0) {
        tp.mode = 0;
    } else {
        tp.mode = 1;
        const char *p = spec;
        while (*p) {
            const char *q = strchr(p, ',');
            size_t n = q ? (size_t)(q - p) : strlen(p);
            if (n == 0) return fail(ctx, MIRGE_E_ARG, "The adapter sequence is empty.");
            if (isdigit((unsigned char)p[0])) return fail(ctx, MIRGE_E_ARG, "Adapter sequence starts with a digit.");
            if (n > MIRGE_MAX_ADAPTER_LEN) return fail(ctx, MIRGE_E_ARG, "adapter longer than 64 nt is not supported");
            if (tp.n_adapters == MIRGE_MAX_ADAPTERS) return fail(ctx, MIRGE_E_ARG, "more than 8 adapters are not supported");
            const int a = tp.n_adapters++;
            tp.alen[a] = (int)n;
            for (size_t i = 0; i < n; i++) {
                const int c = base_code((unsigned char)p[i]);
                if (c > 3) return fail(ctx, MIRGE_E_ARG, "adapter must consist of A, C, G, T/U (wildcards and anchors are not supported)");
                tp.aw[a][i >> 5] |= (u64)c << ((i & 31) * 2);
                tp.peq[a][c] |= 1ull << i;
            }
            p += n;
            if (*p == ',') p++;
        }
    }
    ctx->tp = tp;
    for (int i = 0; i <= MIRGE_MAX_ADAPTER_LEN; i++) {
        // cutadapt accepts a candidate of length i iff i >= min_overlap and cost <= i * max_error_rate (double)
        int c = -1;
        if (i >= min_overlap) { c = 0; while ((double)(c + 1) <= (double)i * error_rate) c++; }
        ctx->thr.v[i] = (signed char)c;
    }
    k1pre_build(ctx->tp, ctx->thr, ctx->pre);
    ctx->qual_cutoff = qual_cutoff;
    return MIRGE_OK;
}

int mirge_detect_phred(const char *buf, uint64_t n)
{
    uint64_t p = 0; int line = 0;
    while (p < n && line < 3) { if (buf[p] == '\n') line++; p++; }
    for (; p < n && buf[p] != '\n'; p++) if ((unsigned char)buf[p] > 74) return 64;
    return 33;
}

#define FEED_REGION_HOST (64ull << 20)      // H2D copy granularity: kernels of region r overlap the copy of region r+1
#define FEED_REGION_DEVICE (4096ull << 20)  // text already resident: one launch per chunk (the K1 queue kernels have long tails)

// Core of mirge_feed_fastq{,_device}: the chunk is consumed region by region, without any host
// synchronisation.  With a host source every region's H2D copy is queued up front on the copy
// stream, and K0/K1 of region r run on the compute stream as soon as copy r has landed, so
// transfers overlap the kernels.  A record may straddle a region boundary: K0 reaches back into
// the earlier bytes of the chunk (they are resident) and the line count is carried from launch to
// launch on the device (K0State), as is the number of records K1 has to process.
template <int W>
static int feed_regions(mirge_ctx *ctx, int sample, unsigned char *d_text, uint64_t nbytes, int phred_base,
                        const char *host_src, int last_char)
{
    cudaStream_t st = ctx->stream;
    const u64 region = host_src ? FEED_REGION_HOST : FEED_REGION_DEVICE;
    const int nreg = (int)((nbytes + region - 1) / region);
    if (host_src) {
        if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        while ((int)ctx->copy_events.size() < nreg + 1) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->copy_events.push_back(e);
        }
        // earlier kernels may still read d_text
        CK(cudaEventRecord(ctx->copy_events[nreg], st));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_events[nreg], 0));
        cudaEvent_t pa = nullptr, pb = nullptr;
        if (ctx->prof) { cudaEventCreate(&pa); cudaEventCreate(&pb); cudaEventRecord(pa, ctx->copy_stream); }
        for (int r = 0; r < nreg; r++) {
            const u64 a = (u64)r * region, len = (a + region <= nbytes) ? region : nbytes - a;
            CK(cudaMemcpyAsync(d_text + a, host_src + a, len, cudaMemcpyHostToDevice, ctx->copy_stream));
            CK(cudaEventRecord(ctx->copy_events[r], ctx->copy_stream));
        }
        if (ctx->prof) { cudaEventRecord(pb, ctx->copy_stream); ctx->spans.push_back({KID_H2D, pa, pb}); }
    }
    K0Params kp; kp.phred_base = phred_base; kp.qual_cutoff = ctx->qual_cutoff; kp.max_len = ctx->max_len;
    kp.min_len = ctx->tp.min_length > 0 ? ctx->tp.min_length : 0;
    K1Counters kc; kc.kept = ctx->d_ctr + 2; kc.len_hist = ctx->d_len_hist; kc.slow_reads = ctx->d_ctr + 2 + ctx->S;
    // record buffers: K0 emits only reads of at least min_length bases, and such a 4-line record is at least
    // 6 + 2 * min_length bytes long
    const u64 max_region = nbytes < region ? nbytes : region;
    const u64 rec_min_bytes = 6 + 2 * (u64)kp.min_len;
    const u64 need_rec = max_region / rec_min_bytes + 64;
    if (need_rec > ctx->rec_cap) {
        CK(cudaStreamSynchronize(st));
        cudaFree(ctx->d_bases); cudaFree(ctx->d_nmask); cudaFree(ctx->d_meta);
        cudaFree(ctx->d_q2); cudaFree(ctx->d_q3);
        ctx->d_bases = ctx->d_nmask = nullptr; ctx->d_meta = nullptr; ctx->rec_cap = 0;
        ctx->d_q2 = ctx->d_q3 = nullptr;
        const size_t cap = need_rec + need_rec / 8;
        CK(cudaMalloc(&ctx->d_bases, cap * W * sizeof(u64)));
        CK(cudaMalloc(&ctx->d_nmask, cap * (W / 2) * sizeof(u64)));
        CK(cudaMalloc(&ctx->d_meta, cap * sizeof(u32)));
        // queues: every record once, plus the partly used chunk of every warp that appends
        const size_t qcap = cap + (size_t)K1Q_CHUNK * 2 * ctx->max_warps;
        CK(cudaMalloc(&ctx->d_q2, qcap * sizeof(K1QRec<W>)));
        CK(cudaMalloc(&ctx->d_q3, qcap * sizeof(K1QRec<W>)));
        ctx->rec_cap = cap;
    }
    const u64 max_tiles = (max_region + 1 + K0_TILE - 1) / K0_TILE;
    CK(ensure(&ctx->d_k0ctl, &ctx->k0ctl_cap, max_tiles + 2));
    CK(ensure(&ctx->d_k0info, &ctx->k0info_cap, max_tiles + 2));
    CK(cudaMemsetAsync(ctx->d_k0state, 0, sizeof(K0State), st));      // the chunk starts at line 0
    const int k0_grid = wave_grid(k0_parse_pack<W>, K0_THREADS, 1ull << 40), k0s_grid = wave_grid(k0_parse_spec<W>, K0_THREADS, 1ull << 40);
    const bool exact_only = (ctx->force_slow & 0x2000) != 0;          // test hook: skip the speculation
    u32 *err_flags = (u32 *)(ctx->d_ctr + 1);
    u64 *processed = ctx->d_ctr + 6 + ctx->S + sample;
    for (int r = 0; r < nreg; r++) {
        const u64 a = (u64)r * region, len = (a + region <= nbytes) ? region : nbytes - a;
        if (host_src) CK(cudaStreamWaitEvent(st, ctx->copy_events[r], 0));
        const bool last = r == nreg - 1;
        const int vnl = (last && last_char != '\n') ? 1 : 0;       // unterminated final line
        const u32 ntiles = (u32)((len + vnl + K0_TILE - 1) / K0_TILE);
        { ProfScope ps(ctx, KID_K0_PACK, exact_only ? 1 : 3);
          CK(cudaMemsetAsync(&ctx->d_k0state->nrec, 0, sizeof(K0State) - offsetof(K0State, nrec), st));
          CK(cudaMemsetAsync(ctx->d_k0ctl, 0, ((size_t)ntiles + 1) * sizeof(u64), st));
          if (!exact_only) {
              k0_parse_spec<W><<<grid_for((u64)ntiles * K0_THREADS, K0_THREADS, k0s_grid), K0_THREADS, 0, st>>>(
                  d_text, a, a + len, vnl, ntiles, ctx->d_k0info, ctx->d_k0state, kp, ctx->rec_cap, ctx->d_bases, ctx->d_nmask, ctx->d_meta);
              k0_verify<<<1, 1024, 0, st>>>(ctx->d_k0info, ntiles, ctx->d_k0state, r & 1, last ? 1 : 0, err_flags, processed, ctx->d_ctr + 5 + ctx->S);
          }
          k0_parse_pack<W><<<grid_for((u64)ntiles * K0_THREADS, K0_THREADS, k0_grid), K0_THREADS, 0, st>>>(
              d_text, a, a + len, vnl, last ? 1 : 0, ntiles, ctx->d_k0ctl, ctx->d_k0state, r & 1, kp, ctx->rec_cap,
              ctx->d_bases, ctx->d_nmask, ctx->d_meta, err_flags, processed, exact_only ? 1 : 0); }
        const u64 *nrec = &ctx->d_k0state->nrec;
        const u64 rec_bound = len / rec_min_bytes + 1;              // sizes the grids only
        { ProfScope ps(ctx, KID_K1_TRIM, 3);
          K1Queues Q; Q.n = ctx->d_qn; Q.q2 = ctx->d_q2; Q.q3 = ctx->d_q3;
          const TableView<W> T = table_view<W>(ctx);
          CK(cudaMemsetAsync(ctx->d_qn, 0, 2 * sizeof(u64), st));
          { ProfScope p1(ctx, KID_K1_CLASSIFY, 0);
            k1_trim_collapse<W><<<wave_grid(k1_trim_collapse<W>, K1A_THREADS, rec_bound), K1A_THREADS, 0, st>>>(
                ctx->d_bases, ctx->d_nmask, ctx->d_meta, nrec, sample, ctx->tp, ctx->pre, ctx->thr, T, kc, Q, ctx->force_slow); }
          // the queue fills live on the device: fixed grids, grid-stride loops bounded by the fill
          { ProfScope p1(ctx, KID_K1_GENERAL, 0);
            k1_general_search<W><<<wave_grid(k1_general_search<W>, K1_THREADS, rec_bound), K1_THREADS, 0, st>>>(sample, ctx->tp, ctx->thr, T, kc, Q); }
          { ProfScope p1(ctx, KID_K1_SLOW, 0);
            k1_slow_search<W><<<wave_grid(k1_slow_search<W>, K1_THREADS, rec_bound), K1_THREADS, 0, st>>>(sample, ctx->tp, ctx->thr, T, kc, Q,
                                                            (ctx->force_slow & 2) ? 1 : 0); } }
        if (ctx->annotate_while_feeding && ctx->libs[MIRGE_LIB_MIRNA].total && !last) {
            // the rounds of this region's new sequences run while the next region is being copied
            int rc = annotate_impl<W>(ctx);
            if (rc) return rc;
        }
    }
    CK(cudaGetLastError());
    ctx->annotated = false; ctx->tallied = false;
    return MIRGE_OK;
}

int mirge_feed_fastq_device(mirge_ctx *ctx, int sample, const void *fastq_dev, uint64_t nbytes, int phred_base)
{
    if (sample < 0 || sample >= ctx->S || nbytes >= 0xFFFFFFF0ull || ((uintptr_t)fastq_dev & 15)) return fail(ctx, MIRGE_E_ARG, "mirge_feed_fastq_device: bad argument (chunk must be < 4 GiB and 16-byte aligned)");
    if (nbytes == 0) return MIRGE_OK;
    CK(cudaSetDevice(ctx->device));
    unsigned char last = 0;
    CK(cudaMemcpyAsync(&last, (const unsigned char *)fastq_dev + nbytes - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    unsigned char *t = (unsigned char *)const_cast<void *>(fastq_dev);
    return ctx->W == 2 ? feed_regions<2>(ctx, sample, t, nbytes, phred_base, nullptr, last)
                       : feed_regions<4>(ctx, sample, t, nbytes, phred_base, nullptr, last);
}

int mirge_feed_fastq(mirge_ctx *ctx, int sample, const char *fastq, uint64_t nbytes, int phred_base)
{
    if (sample < 0 || sample >= ctx->S || !fastq || nbytes >= 0xFFFFFFF0ull) return fail(ctx, MIRGE_E_ARG, "mirge_feed_fastq: bad argument (chunk must be < 4 GiB)");
    if (nbytes == 0) return MIRGE_OK;
    CK(cudaSetDevice(ctx->device));
    if (ctx->text_cap < (size_t)nbytes + 64) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(ensure(&ctx->d_text, &ctx->text_cap, (size_t)nbytes + 64));
    }
    const int last = (unsigned char)fastq[nbytes - 1];
    return ctx->W == 2 ? feed_regions<2>(ctx, sample, ctx->d_text, nbytes, phred_base, fastq, last)
                       : feed_regions<4>(ctx, sample, ctx->d_text, nbytes, phred_base, fastq, last);
}

int mirge_get_stats(mirge_ctx *ctx, uint64_t *processed, uint64_t *kept, uint64_t *n_unique)
{
    int rc = mirge_sync(ctx);
    if (rc) return rc;
    std::vector<u64> h(6 + 2 * ctx->S);
    CK(cudaMemcpy(h.data(), ctx->d_ctr, h.size() * sizeof(u64), cudaMemcpyDeviceToHost));
    if (n_unique) *n_unique = h[0] < ctx->max_unique ? h[0] : ctx->max_unique;
    for (int s = 0; s < ctx->S; s++) {
        if (processed) processed[s] = h[6 + ctx->S + s];
        if (kept) kept[s] = h[2 + s];
    }
    return MIRGE_OK;
}

// ---------------------------------------------------------------------------
// libraries
// ---------------------------------------------------------------------------
struct RoundSpec { int lib, len_filter, mode, param, trim5, trim3, highest, polyT; };
// runAnnotationPipeline.py:574-599
static const RoundSpec ROUNDS[MIRGE_N_ROUNDS] = {
    {MIRGE_LIB_MIRNA, -26, 0, 0, 0, 0, 0, 0},        // bowtie mirna -n 0, len < 26
    {MIRGE_LIB_HAIRPIN, 25, 0, 1, 0, 0, 0, 0},       // hairpin -n 1, len > 25
    {MIRGE_LIB_MATURE_TRNA, 0, 1, 1, 0, 0, 1, 0},    // mature tRNA -v 1 -a --best --strata
    {MIRGE_LIB_PRE_TRNA, 0, 1, 0, 0, 0, 1, 1},       // pre-tRNA poly-T round, -v 0 -a --best --strata
    {MIRGE_LIB_SNORNA, 0, 0, 1, 0, 0, 0, 0},         // snoRNA -n 1
    {MIRGE_LIB_RRNA, 0, 0, 1, 0, 0, 0, 0},           // rRNA -n 1
    {MIRGE_LIB_NCRNA_OTHERS, 0, 0, 1, 0, 0, 0, 0},   // ncRNA others -n 1
    {MIRGE_LIB_MRNA, 0, 0, 0, 0, 0, 0, 0},           // mRNA -n 0
    {MIRGE_LIB_MIRNA, 0, 1, 2, 1, 2, 0, 0},          // mirna -5 1 -3 2 -v 2 --best (isomiR)
    {MIRGE_LIB_SPIKEIN, 0, 0, 0, 0, 0, 0, 0},        // spike-in -n 0
};

// One GapSet (k2_map.cuh): m pieces of pl bases, a table per pair of pieces, serving regions rg_lo .. rg_hi.
static int build_gap_set(mirge_ctx *ctx, DevLib &l, const LibView &L, int m, int pl, int rg_lo, int rg_hi)
{
    cudaStream_t st = ctx->stream;
    const int npairs = m * (m - 1) / 2;
    const u64 nb1 = (1ull << (4 * pl)) + 1;
    const u32 npos = l.total - (u32)(m * pl) + 1;
    u32 *off = nullptr, *pos = nullptr, *cursor = nullptr, *scratch = nullptr;
    cudaError_t e = cudaMalloc(&off, (size_t)npairs * nb1 * sizeof(u32));
    if (e == cudaSuccess) e = cudaMalloc(&pos, (size_t)npairs * npos * sizeof(u32));
    if (e == cudaSuccess) e = cudaMalloc(&cursor, nb1 * sizeof(u32));
    if (e == cudaSuccess) e = cudaMalloc(&scratch, scan_scratch_elems(nb1) * sizeof(u32));
    if (e == cudaSuccess) e = cudaMemsetAsync(off, 0, (size_t)npairs * nb1 * sizeof(u32), st);
    for (int sd = 0; sd < npairs && e == cudaSuccess; sd++) {
        const int grid = grid_for(npos, 256, 148 * 8);
        const int oi = gap_piece(m, sd, 0) * pl, oj = gap_piece(m, sd, 1) * pl;
        u32 *o = off + (size_t)sd * nb1;
        k2_gap_pass<<<grid, 256, 0, st>>>(L, pl, oi, oj, npos, 0, o, nullptr);
        device_excl_scan(o, o, nb1, scratch, st);
        e = cudaMemcpyAsync(cursor, o, nb1 * sizeof(u32), cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) k2_gap_pass<<<grid, 256, 0, st>>>(L, pl, oi, oj, npos, 1, cursor, pos + (size_t)sd * npos);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(cursor); cudaFree(scratch);
    if (e != cudaSuccess) {
        cudaFree(off); cudaFree(pos);
        if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return MIRGE_OK; }   // no room: the plain seeds serve every query
        ctx->err = std::string("pair seed table build failed: ") + cudaGetErrorString(e);
        return MIRGE_E_CUDA;
    }
    GapSet &G = l.gap[l.ngap++];
    G.m = m; G.pl = pl; G.rg_lo = rg_lo; G.rg_hi = rg_hi; G.nb1 = (u32)nb1; G.npos = npos; G.off = off; G.pos = pos;
    return MIRGE_OK;
}

// plain seed lengths of a library slot (k-mer CSR tables)
static std::vector<int> seed_lengths_for(int lib_id, u64 total)
{
    // the miRNA library serves the 3-seed isomiR round on reads as short as 13 nt (k = 4..8); the big mRNA library
    // gets a 12-mer table.  The longest seed a read allows: every extra base divides the candidates to verify by four
    if (lib_id == MIRGE_LIB_MIRNA) return {4, 5, 6, 7, 8};
    if (total >= (1u << 20)) return {8, 9, 10, 11, 12};
    return {8, 9, 10, 11};
}

// Which pair sets a library gets, and which region lengths each serves: those where the classic seeds (e + 1 pieces,
// longest table that fits) are expected to offer at least `pair_seeds` candidates per query and clearly more than the
// pair tables do (pair_seeds == 1: every region the set can serve; 0: none).  Expected counts assume a random library.
// Host arithmetic only (mirge_pair_seed_plan exposes it to the CPU tests).  Rows: {m, pl, rg_lo, rg_hi}.
struct GapPlanRow { int m, pl, rg_lo, rg_hi; };
static std::vector<GapPlanRow> plan_gap_rows(int lib_id, u64 total, int pair_seeds)
{
    std::vector<GapPlanRow> rows;
    if (pair_seeds <= 0 || lib_id < 0 || lib_id >= MIRGE_N_LIBS || lib_id == MIRGE_LIB_GENOME) return rows;
    const std::vector<int> ks = seed_lengths_for(lib_id, total);
    for (int e = 1; e <= 2; e++) {
        bool used = false;
        for (int r = 0; r < MIRGE_N_ROUNDS; r++) if (ROUNDS[r].lib == lib_id && ROUNDS[r].param == e) used = true;
        if (!used) continue;
        const int m = e + 2, npairs = m * (m - 1) / 2;
        const int pls[2] = {m == 4 ? 3 : 5, m == 4 ? 4 : 6};
        for (int x = 0; x < 2; x++) {
            const int pl = pls[x], lo = m * pl, hi_max = x == 0 ? m * pls[1] - 1 : 23;
            if (total < (u64)lo) continue;
            const double pair_c = (double)npairs * (double)total / (double)(1ull << (4 * pl));
            auto classic_k = [&](int rg) { int kc = -1; for (int k : ks) if (k <= rg / (e + 1) && (u64)k <= total) kc = k; return kc; };
            int s_lo = -1, s_hi = -1;
            for (int rg = lo; rg <= hi_max; rg++) {
                const int kc = classic_k(rg);
                const double classic_c = kc < 0 ? 0.0 : (double)(e + 1) * (double)total / (double)(1ull << (2 * kc));
                const bool serve = pair_seeds == 1 || (classic_c >= (double)pair_seeds && classic_c > 1.5 * pair_c + 2.0);
                if (serve) { if (s_lo < 0) s_lo = rg; s_hi = rg; }
                else if (s_lo >= 0) break;
            }
            if (s_lo < 0 || rows.size() >= 4) continue;
            // regions no classic table fits (below 16 bases, only with a non-default min_length) would scan every offset:
            // a set that exists anyway takes them too
            while (s_lo > lo && classic_k(s_lo - 1) < 0) s_lo--;
            rows.push_back({m, pl, s_lo, s_hi});
        }
    }
    return rows;
}

int mirge_pair_seed_plan(int lib_id, uint64_t total, int pair_seeds, int *out16)
{
    const std::vector<GapPlanRow> rows = plan_gap_rows(lib_id, total, pair_seeds);
    for (size_t i = 0; i < rows.size() && out16; i++) {
        out16[4 * i] = rows[i].m; out16[4 * i + 1] = rows[i].pl; out16[4 * i + 2] = rows[i].rg_lo; out16[4 * i + 3] = rows[i].rg_hi;
    }
    return (int)rows.size();
}

static int plan_gap_sets(mirge_ctx *ctx, DevLib &l, const LibView &L, int lib_id)
{
    for (const GapPlanRow &g : plan_gap_rows(lib_id, l.total, ctx->pair_seeds)) {
        int rc = build_gap_set(ctx, l, L, g.m, g.pl, g.rg_lo, g.rg_hi);
        if (rc) return rc;
    }
    return MIRGE_OK;
}

static int load_library_impl(mirge_ctx *ctx, int lib_id, int n_refs, const char *seq_ascii, const uint64_t *offsets)
{
    if (lib_id < 0 || lib_id >= MIRGE_N_LIBS || n_refs < 0 || (n_refs > 0 && (!seq_ascii || !offsets)))
        return fail(ctx, MIRGE_E_ARG, "mirge_load_library: bad argument");
    const u64 total = n_refs ? offsets[n_refs] : 0;
    if (total >= 0xFFFFFF00ull) return fail(ctx, MIRGE_E_ARG, "library larger than 4 Gbases is not supported");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(cudaStreamSynchronize(st));
    DevLib &l = ctx->libs[lib_id];
    free_lib(l);
    ProfScope ps(ctx, KID_LIB_BUILD, 0);
    l.n_refs = n_refs; l.total = (u32)total;
    const size_t nw64 = (total + 63) / 64;
    CK(cudaMalloc(&l.seq, (2 * nw64 + 8) * sizeof(u64)));
    CK(cudaMalloc(&l.amb, (nw64 + 8) * sizeof(u64)));
    CK(cudaMemsetAsync(l.seq, 0, (2 * nw64 + 8) * sizeof(u64), st));
    CK(cudaMemsetAsync(l.amb, 0, (nw64 + 8) * sizeof(u64), st));
    std::vector<u32> off32(n_refs + 1);
    for (int i = 0; i <= n_refs; i++) off32[i] = (u32)(n_refs ? offsets[i] : 0);
    l.has_amb = 0;
    for (u64 i = 0; i < total && !l.has_amb; i++) if (base_code((unsigned char)seq_ascii[i]) > 3) l.has_amb = 1;
    l.max_ref_len = 0;
    for (int i = 0; i < n_refs; i++) if (off32[i + 1] - off32[i] > l.max_ref_len) l.max_ref_len = off32[i + 1] - off32[i];
    CK(cudaMalloc(&l.ref_off, (size_t)(n_refs + 1) * sizeof(u32)));
    CK(cudaMemcpyAsync(l.ref_off, off32.data(), (size_t)(n_refs + 1) * sizeof(u32), cudaMemcpyHostToDevice, st));
    {   // entry index at the start of every 64-base block (empty entries are skipped by the forward scan)
        std::vector<u32> blk(total / 64 + 2);
        int r = 0;
        for (size_t b = 0; b < blk.size(); b++) {
            const u64 g = (u64)b * 64;
            while (r + 1 < n_refs && offsets[r + 1] <= g) r++;
            blk[b] = (u32)r;
        }
        CK(cudaMalloc(&l.blk_ref, blk.size() * sizeof(u32)));
        CK(cudaMemcpy(l.blk_ref, blk.data(), blk.size() * sizeof(u32), cudaMemcpyHostToDevice));
    }
    if (total == 0) { CK(cudaStreamSynchronize(st)); return MIRGE_OK; }
    unsigned char *d_ascii = nullptr;
    CK(cudaMalloc(&d_ascii, total));
    CK(cudaMemcpyAsync(d_ascii, seq_ascii, total, cudaMemcpyHostToDevice, st));
    k2_pack_library<<<(unsigned)((nw64 + 127) / 128), 128, 0, st>>>(d_ascii, (u32)total, l.seq, l.amb);
    LibView L; L.seq = l.seq; L.amb = l.amb; L.ref_off = l.ref_off; L.blk_ref = l.blk_ref; L.n_refs = n_refs; L.total = (u32)total; L.max_ref_len = l.max_ref_len; L.has_amb = l.has_amb;
    if (lib_id == MIRGE_LIB_GENOME) {
        // one prefix-ordered table: kmax-mers, first base most significant (k2_genome.cuh)
        int kmax = 6;
        while (kmax < 12 && (1ull << (2 * (kmax + 1))) <= 4 * total) kmax++;
        const u64 nb = 1ull << (2 * kmax);
        u32 *cursor = nullptr, *scratch = nullptr;
        cudaError_t e = cudaMalloc(&l.gx_off, (nb + 1) * sizeof(u32));
        if (e == cudaSuccess) e = cudaMalloc(&cursor, (nb + 1) * sizeof(u32));
        if (e == cudaSuccess) e = cudaMalloc(&l.gx_pos, (total + 1) * sizeof(u32));
        if (e == cudaSuccess) e = cudaMalloc(&scratch, scan_scratch_elems(nb + 1) * sizeof(u32));
        if (e == cudaSuccess) e = cudaMemsetAsync(l.gx_off, 0, (nb + 1) * sizeof(u32), st);
        if (e == cudaSuccess) {
            const int grid = grid_for(total, 256, 148 * 8);
            k2g_index_pass<<<grid, 256, 0, st>>>(L, kmax, 0, l.gx_off, nullptr);
            device_excl_scan(l.gx_off, l.gx_off, nb + 1, scratch, st);
            e = cudaMemcpyAsync(cursor, l.gx_off, (nb + 1) * sizeof(u32), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) k2g_index_pass<<<grid, 256, 0, st>>>(L, kmax, 1, cursor, l.gx_pos);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
        cudaFree(cursor); cudaFree(scratch); cudaFree(d_ascii);
        if (e != cudaSuccess) { ctx->err = std::string("genome index build failed: ") + cudaGetErrorString(e); return MIRGE_E_CUDA; }
        l.gx_kmax = kmax;
        CK(cudaGetLastError());
        return MIRGE_OK;
    }
    const std::vector<int> ks = seed_lengths_for(lib_id, total);
    for (int k : ks) {
        if ((u64)k > total) continue;
        const u64 nb = 1ull << (2 * k);
        u32 *boff = nullptr, *cursor = nullptr, *pos = nullptr, *scratch = nullptr;
        cudaError_t e = cudaMalloc(&boff, (nb + 1) * sizeof(u32));
        if (e == cudaSuccess) e = cudaMalloc(&cursor, (nb + 1) * sizeof(u32));
        if (e == cudaSuccess) e = cudaMalloc(&pos, (total + 1) * sizeof(u32));
        if (e == cudaSuccess) e = cudaMalloc(&scratch, scan_scratch_elems(nb + 1) * sizeof(u32));
        if (e == cudaSuccess) e = cudaMemsetAsync(boff, 0, (nb + 1) * sizeof(u32), st);
        if (e == cudaSuccess) {
            const int grid = grid_for(total, 256, 148 * 8);
            k2_kmer_pass<<<grid, 256, 0, st>>>(L, k, 0, boff, nullptr);
            device_excl_scan(boff, boff, nb + 1, scratch, st);
            e = cudaMemcpyAsync(cursor, boff, (nb + 1) * sizeof(u32), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) k2_kmer_pass<<<grid, 256, 0, st>>>(L, k, 1, cursor, pos);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
        cudaFree(cursor); cudaFree(scratch);
        if (e != cudaSuccess) {
            cudaFree(boff); cudaFree(pos); cudaFree(d_ascii);
            ctx->err = std::string("seed table build failed: ") + cudaGetErrorString(e);
            return MIRGE_E_CUDA;
        }
        const int t = l.ntabs++;
        l.tab_k[t] = k; l.tab_off[t] = boff; l.tab_pos[t] = pos;
    }
    CK(cudaStreamSynchronize(st));
    cudaFree(d_ascii);
    CK(cudaGetLastError());
    { int rc = plan_gap_sets(ctx, l, L, lib_id); if (rc) return rc; }
    { int rc = annot_clear(ctx); if (rc) return rc; }          // a library changed: every sequence is looked at again
    ctx->annotated = false; ctx->tallied = false;
    return MIRGE_OK;
}

int mirge_load_library(mirge_ctx *ctx, int lib_id, int n_refs, const char *seq_ascii, const uint64_t *offsets)
{
    if (lib_id == MIRGE_LIB_GENOME) return fail(ctx, MIRGE_E_ARG, "mirge_load_library: slot 8 is the genome, use mirge_load_genome");
    return load_library_impl(ctx, lib_id, n_refs, seq_ascii, offsets);
}

int mirge_load_genome(mirge_ctx *ctx, int n_refs, const char *seq_ascii, const uint64_t *offsets)
{
    return load_library_impl(ctx, MIRGE_LIB_GENOME, n_refs, seq_ascii, offsets);
}

// Shared tail of the genome passes: queries are packed in d_q* (n of them); counts come back to the host.
template <int W>
static int genome_pass_run(mirge_ctx *ctx, uint64_t n, const u64 *qs, const u64 *qnm, const u8 *qlen, const GenomeParams &P, uint32_t *counts_host)
{
    cudaStream_t st = ctx->stream;
    const DevLib &l = ctx->libs[MIRGE_LIB_GENOME];
    LibView L; L.seq = l.seq; L.amb = l.amb; L.ref_off = l.ref_off; L.blk_ref = l.blk_ref; L.n_refs = l.n_refs; L.total = l.total; L.max_ref_len = l.max_ref_len; L.has_amb = l.has_amb;
    GenomeIndex X; X.off = l.gx_off; X.pos = l.gx_pos; X.kmax = l.gx_kmax;
    u32 *d_counts = nullptr;
    CK(cudaMalloc(&d_counts, (size_t)n * 4 * sizeof(u32)));
    { ProfScope ps(ctx, KID_GENOME);
      k2g_all_hits<W><<<grid_for(n * 32, K2G_THREADS, 148 * 16), K2G_THREADS, 0, st>>>(qs, qnm, qlen, n, L, X, P, d_counts); }
    cudaError_t e = cudaMemcpyAsync(counts_host, d_counts, (size_t)n * 4 * sizeof(u32), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_counts);
    if (e != cudaSuccess) { ctx->err = std::string("genome pass failed: ") + cudaGetErrorString(e); return MIRGE_E_CUDA; }
    CK(cudaGetLastError());
    return MIRGE_OK;
}

static int genome_params(mirge_ctx *ctx, int mode, int param, int seedlen, int trim3, int norc, GenomeParams &P)
{
    if ((mode != 0 && mode != 1) || param < 0 || param > 3 || seedlen < 1 || trim3 < 0) return fail(ctx, MIRGE_E_ARG, "mirge_genome_pass: bad argument");
    if (!ctx->libs[MIRGE_LIB_GENOME].gx_off) return fail(ctx, MIRGE_E_STATE, "mirge_genome_pass: no genome loaded (mirge_load_genome)");
    P.mode = mode; P.param = param; P.seedlen = seedlen; P.trim3 = trim3; P.norc = norc ? 1 : 0;
    return MIRGE_OK;
}

int mirge_genome_pass(mirge_ctx *ctx, uint64_t n, const char *reads, int stride, const uint8_t *lens,
                      int mode, int param, int seedlen, int trim3, int norc, uint32_t *counts)
{
    GenomeParams P;
    int rc = genome_params(ctx, mode, param, seedlen, trim3, norc, P);
    if (rc) return rc;
    if (n == 0) return MIRGE_OK;
    if (!reads || !lens || !counts || stride < 1) return fail(ctx, MIRGE_E_ARG, "mirge_genome_pass: bad argument");
    CK(cudaSetDevice(ctx->device));
    const int W = ctx->W, NW = W / 2;
    std::vector<u64> hs((size_t)n * W, 0), hn((size_t)n * NW, 0);
    for (uint64_t i = 0; i < n; i++) {
        const int len = lens[i];
        if (len > ctx->max_len || len > stride) return fail(ctx, MIRGE_E_TOOLONG, "mirge_genome_pass: a read is longer than max_read_len");
        const char *r = reads + i * (uint64_t)stride;
        for (int b = 0; b < len; b++) {
            int c = base_code((unsigned char)r[b]);
            if (c > 3) { hn[i * NW + (b >> 6)] |= 1ull << (b & 63); c = 0; }
            hs[i * W + (b >> 5)] |= (u64)c << ((b & 31) * 2);
        }
    }
    u64 *qs = nullptr, *qnm = nullptr; u8 *ql = nullptr;
    CK(cudaMalloc(&qs, hs.size() * sizeof(u64)));
    CK(cudaMalloc(&qnm, hn.size() * sizeof(u64)));
    CK(cudaMalloc(&ql, (size_t)n));
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(qs, hs.data(), hs.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(qnm, hn.data(), hn.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ql, lens, (size_t)n, cudaMemcpyHostToDevice, st));
    rc = W == 2 ? genome_pass_run<2>(ctx, n, qs, qnm, ql, P, counts) : genome_pass_run<4>(ctx, n, qs, qnm, ql, P, counts);
    cudaFree(qs); cudaFree(qnm); cudaFree(ql);
    return rc;
}

int mirge_genome_pass_ids(mirge_ctx *ctx, uint64_t n, const uint32_t *ids, int mode, int param, int seedlen, int trim3, int norc,
                          uint32_t *counts)
{
    GenomeParams P;
    int rc = genome_params(ctx, mode, param, seedlen, trim3, norc, P);
    if (rc) return rc;
    if (n == 0) return MIRGE_OK;
    if (!ids || !counts) return fail(ctx, MIRGE_E_ARG, "mirge_genome_pass_ids: bad argument");
    CK(cudaSetDevice(ctx->device));
    u64 U = 0;
    rc = mirge_get_stats(ctx, nullptr, nullptr, (uint64_t *)&U);
    if (rc) return rc;
    const int W = ctx->W, NW = W / 2;
    u64 *qs = nullptr, *qnm = nullptr; u8 *ql = nullptr; u32 *d_ids = nullptr;
    CK(cudaMalloc(&qs, (size_t)n * W * sizeof(u64)));
    CK(cudaMalloc(&qnm, (size_t)n * NW * sizeof(u64)));
    CK(cudaMalloc(&ql, (size_t)n));
    CK(cudaMalloc(&d_ids, (size_t)n * sizeof(u32)));
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(d_ids, ids, (size_t)n * sizeof(u32), cudaMemcpyHostToDevice, st));
    const int g = grid_for(n, 256, 148 * 8);
    ctx->n_launches++;
    if (W == 2) k2g_gather_queries<2><<<g, 256, 0, st>>>(d_ids, n, ctx->d_useq, ctx->d_unmask, ctx->d_ulen, U, qs, qnm, ql);
    else k2g_gather_queries<4><<<g, 256, 0, st>>>(d_ids, n, ctx->d_useq, ctx->d_unmask, ctx->d_ulen, U, qs, qnm, ql);
    rc = W == 2 ? genome_pass_run<2>(ctx, n, qs, qnm, ql, P, counts) : genome_pass_run<4>(ctx, n, qs, qnm, ql, P, counts);
    cudaFree(qs); cudaFree(qnm); cudaFree(ql); cudaFree(d_ids);
    return rc;
}


// Annotation is incremental: dense ids are handed out in arrival order, so "the sequences not looked at yet" is the id
// range [annot[0], n_unique).  k2_take_range claims it on the device (no host round trip); the rounds then run on that
// range only.  With MIRGE_OPT_ANNOTATE_WHILE_FEEDING the rounds of region r run while the copy of region r+1 is in
// flight, so at the end of a feed only the last region's sequences are left for mirge_annotate.
__global__ void k2_take_range(const u64 *n_unique, u64 max_unique, u64 *annot)
{
    u64 hi = *n_unique;
    if (hi > max_unique) hi = max_unique;
    u64 lo = annot[0];
    if (lo > hi) lo = hi;
    annot[1] = lo; annot[2] = hi - lo; annot[0] = hi;
}

// The first to-do lists: the claimed id range written out (the round kernels keep their register budget: one more
// kernel argument cost k2_round eight registers and an eighth of its occupancy).  Rounds 0 and 1 filter by length
// on complementary conditions (len < 26: miRNA, len > 25: hairpin), so the range is dealt into two lists here and each
// of the two rounds only ever sees the sequences it would search; both append their misses to round 2's list.
// split_len == 0: one list (todo_a).
__global__ void k2_fill_range(const u64 *annot, const u8 *__restrict__ ulen, int split_len, u32 *todo_a, u64 *n_a, u32 *todo_b, u64 *n_b)
{
    const u64 base = annot[1], n = annot[2];
    const unsigned lane = threadIdx.x & 31;
    for (u64 i0 = (u64)blockIdx.x * blockDim.x; i0 < n; i0 += (u64)gridDim.x * blockDim.x) {
        const u64 i = i0 + threadIdx.x;
        const bool in = i < n;
        const bool to_b = in && split_len > 0 && (int)ulen[base + i] >= split_len;
        const unsigned ba = __ballot_sync(0xffffffffu, in && !to_b), bb = __ballot_sync(0xffffffffu, to_b);
        u64 sa = 0, sb = 0;
        if (lane == 0) { if (ba) sa = atomicAdd(n_a, (u64)__popc(ba)); if (bb) sb = atomicAdd(n_b, (u64)__popc(bb)); }
        sa = __shfl_sync(0xffffffffu, sa, 0); sb = __shfl_sync(0xffffffffu, sb, 0);
        const unsigned below = (1u << lane) - 1u;
        if (to_b) todo_b[sb + __popc(bb & below)] = (u32)(base + i);
        else if (in) todo_a[sa + __popc(ba & below)] = (u32)(base + i);
    }
}

static int annot_clear(mirge_ctx *ctx)
{
    cudaStream_t st = ctx->stream;
    CK(cudaMemsetAsync(ctx->d_round_of, 0, ctx->max_unique, st));
    CK(cudaMemsetAsync(ctx->d_mm_of, 0, ctx->max_unique, st));              // exports are deterministic for un-annotated sequences too
    CK(cudaMemsetAsync(ctx->d_ref_of, 0, ctx->max_unique * sizeof(int), st));
    CK(cudaMemsetAsync(ctx->d_pos_of, 0, ctx->max_unique * sizeof(int), st));
    CK(cudaMemsetAsync(ctx->d_round_stats, 0, 2 * MIRGE_N_ROUNDS * sizeof(u64), st));
    CK(cudaMemsetAsync(ctx->d_annot, 0, 4 * sizeof(u64), st));
    return MIRGE_OK;
}

template <int W>
static int annotate_impl(mirge_ctx *ctx)
{
    cudaStream_t st = ctx->stream;
    CK(cudaMemsetAsync(ctx->d_todo_n, 0, (MIRGE_N_ROUNDS + 3) * sizeof(u64), st));
    k2_take_range<<<1, 1, 0, st>>>(ctx->d_ctr, ctx->max_unique, ctx->d_annot);
    // rounds 0 and 1 on complementary length classes: each gets its own list, both feed round 2's (d_todo[1], d_todo_n[2])
    const bool split = ROUNDS[0].len_filter < 0 && ROUNDS[1].len_filter == -ROUNDS[0].len_filter - 1;
    u64 *n_long = ctx->d_todo_n + MIRGE_N_ROUNDS + 1, *n_past = ctx->d_todo_n + MIRGE_N_ROUNDS + 2;
    k2_fill_range<<<148 * 4, 256, 0, st>>>(ctx->d_annot, ctx->d_ulen, split ? -ROUNDS[0].len_filter : 0,
                                           split ? ctx->d_todo[0] : ctx->d_todo[1], ctx->d_todo_n, ctx->d_todo_long, n_long);
    ctx->n_launches += 2;
    K2Arrays A;
    A.useq = ctx->d_useq; A.unmask = ctx->d_unmask; A.ulen = ctx->d_ulen;
    A.round_of = ctx->d_round_of; A.ref_of = ctx->d_ref_of; A.pos_of = ctx->d_pos_of; A.mm_of = ctx->d_mm_of;
    const int nrounds = ctx->spike_in ? 10 : 9;
    for (int r = 0; r < nrounds; r++) {
        const RoundSpec &rs = ROUNDS[r];
        const DevLib &l = ctx->libs[rs.lib];
        LibView L; L.seq = l.total ? l.seq : nullptr; L.amb = l.amb; L.ref_off = l.ref_off; L.blk_ref = l.blk_ref; L.n_refs = l.n_refs; L.total = l.total; L.max_ref_len = l.max_ref_len; L.has_amb = l.has_amb;
        RoundParams R;
        memset(&R, 0, sizeof R);
        R.round = r; R.len_filter = rs.len_filter; R.mode = rs.mode; R.param = rs.param; R.trim5 = rs.trim5; R.trim3 = rs.trim3;
        R.highest = rs.highest; R.polyT = rs.polyT; R.ntabs = l.ntabs;
        for (int t = 0; t < l.ntabs; t++) { R.tabs[t].k = l.tab_k[t]; R.tabs[t].bucket_off = l.tab_off[t]; R.tabs[t].pos = l.tab_pos[t]; }
        R.kcons = l.ntabs > 0;
        for (int t = 1; t < l.ntabs; t++) if (l.tab_k[t] != l.tab_k[t - 1] + 1) R.kcons = 0;
        for (int g = 0; g < l.ngap; g++)
            if (rs.param >= 1 && l.gap[g].m == rs.param + 2 && R.ngap < K2_MAX_GAPSETS) { R.gap[R.ngap++] = l.gap[g]; R.gap_pairs = l.gap[g].m * (l.gap[g].m - 1) / 2; }
        const u32 *todo_in = ctx->d_todo[(r + 1) & 1];             // round 0 reads the range list in d_todo[1]
        const u64 *n_in = ctx->d_todo_n + r;
        u32 *todo_out = ctx->d_todo[r & 1];
        u64 *n_out = ctx->d_todo_n + r + 1;
        if (split && r == 0) { todo_in = ctx->d_todo[0]; todo_out = ctx->d_todo[1]; n_out = ctx->d_todo_n + 2; }
        if (split && r == 1) { todo_in = ctx->d_todo_long; n_in = n_long; }
        // the poly-T round only takes sequences ending in TTT: the round before it sends the others straight to the list
        // of the round after it (the third buffer, free again by then), where the poly-T round appends its own misses
        const bool tail_skip = !(ctx->force_slow & 0x1000);
        if (tail_skip && r + 2 < nrounds && ROUNDS[r + 1].polyT && r >= 2) { R.next_tail3T = 1; R.bypass_out = ctx->d_todo_long; R.bypass_n = n_past; }
        if (tail_skip && r >= 3 && rs.polyT && r + 1 < nrounds) { todo_out = ctx->d_todo_long; n_out = n_past; }
        if (tail_skip && r >= 4 && ROUNDS[r - 1].polyT) { todo_in = ctx->d_todo_long; n_in = n_past; }
        ProfScope ps(ctx, KID_K2_ROUND0 + r);
        if (R.ngap > 0)                                     // short queries go through the pair-of-pieces tables
            k2_round_pairs8<W><<<148 * K2_CTAS, 128, 0, st>>>(A, todo_in, n_in, todo_out, n_out, L, R,
                                                      ctx->d_round_stats + r, ctx->d_round_stats + MIRGE_N_ROUNDS + r);
        else if (rs.param > 0 && rs.lib == MIRGE_LIB_MIRNA) // several seeds over consecutive table sizes: balance them
            k2_round_balanced<W><<<148 * K2_CTAS, 128, 0, st>>>(A, todo_in, n_in, todo_out, n_out, L, R,
                                                      ctx->d_round_stats + r, ctx->d_round_stats + MIRGE_N_ROUNDS + r);
        else
            k2_round<W><<<148 * K2_CTAS, 128, 0, st>>>(A, todo_in, n_in, todo_out, n_out, L, R,
                                                       ctx->d_round_stats + r, ctx->d_round_stats + MIRGE_N_ROUNDS + r);
    }
    CK(cudaGetLastError());
    ctx->annotated = true; ctx->tallied = false;
    return MIRGE_OK;
}

int mirge_annotate(mirge_ctx *ctx)
{
    CK(cudaSetDevice(ctx->device));
    {   // a full table would leave the unique counter beyond the arrays the rounds index
        int rc = mirge_sync(ctx);
        if (rc) return rc;
    }
    return ctx->W == 2 ? annotate_impl<2>(ctx) : annotate_impl<4>(ctx);
}

int mirge_set_option(mirge_ctx *ctx, int option, int value)
{
    if (option == MIRGE_OPT_ANNOTATE_WHILE_FEEDING) { ctx->annotate_while_feeding = value != 0; return MIRGE_OK; }
    if (option == MIRGE_OPT_PAIR_SEEDS) { ctx->pair_seeds = value < 0 ? 0 : value; return MIRGE_OK; }   // read when a library is loaded
    return fail(ctx, MIRGE_E_ARG, "mirge_set_option: unknown option");
}

// ---------------------------------------------------------------------------
// tallies
// ---------------------------------------------------------------------------
int mirge_tally(mirge_ctx *ctx)
{
    if (!ctx->annotated) return fail(ctx, MIRGE_E_STATE, "mirge_tally called before mirge_annotate");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    u64 U = 0;
    CK(cudaMemcpyAsync(&U, ctx->d_ctr, sizeof(u64), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (U > ctx->max_unique) U = ctx->max_unique;
    const int S = ctx->S, n_mir = ctx->libs[MIRGE_LIB_MIRNA].n_refs, n_seg = n_mir + MIRGE_N_TOTALS;
    if (ctx->tally_nmir != n_mir) {
        cudaFree(ctx->d_mir_quant); cudaFree(ctx->d_mir_iscan);
        CK(cudaMalloc(&ctx->d_mir_quant, (size_t)(n_mir + 1) * S * sizeof(i64)));
        CK(cudaMalloc(&ctx->d_mir_iscan, (size_t)(n_mir + 1) * S * sizeof(i64)));
        ctx->tally_nmir = n_mir;
    }
    const u32 n_tiles = (u32)((U + RS_TILE - 1) / RS_TILE);
    const size_t hist_n = (size_t)256 * (n_tiles ? n_tiles : 1);
    const size_t max_chunks = (size_t)(U / K3_CHUNK) + n_seg + 1;
    // one persistent scratch arena, carved into 256-byte aligned pieces
    size_t need = 0;
    auto carve = [&need](size_t bytes) { size_t o = need; need += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_keys0 = carve((U + 1) * sizeof(u32)), o_keys1 = carve((U + 1) * sizeof(u32));
    const size_t o_vals0 = carve((U + 1) * sizeof(u32)), o_vals1 = carve((U + 1) * sizeof(u32));
    const size_t o_hist = carve(hist_n * sizeof(u32)), o_scr = carve(scan_scratch_elems(hist_n > (size_t)n_seg + 1 ? hist_n : n_seg + 1) * sizeof(u32));
    const size_t o_ss = carve((size_t)n_seg * sizeof(u32)), o_se = carve((size_t)n_seg * sizeof(u32));
    const size_t o_ch = carve(((size_t)n_seg + 1) * sizeof(u32));
    const size_t o_sq = carve((size_t)n_seg * S * sizeof(i64)), o_si = carve((size_t)n_seg * S * sizeof(i64)), o_su = carve((size_t)n_seg * S * sizeof(i64));
    const size_t o_part = carve(max_chunks * S * 3 * sizeof(i64));
    CK(ensure(&ctx->d_tally, &ctx->tally_cap, need));
    unsigned char *tb = ctx->d_tally;
    u32 *keys[2] = {(u32 *)(tb + o_keys0), (u32 *)(tb + o_keys1)}, *vals[2] = {(u32 *)(tb + o_vals0), (u32 *)(tb + o_vals1)};
    u32 *hist = (u32 *)(tb + o_hist), *scratch = (u32 *)(tb + o_scr), *seg_start = (u32 *)(tb + o_ss), *seg_end = (u32 *)(tb + o_se);
    u32 *chunk_off = (u32 *)(tb + o_ch);
    i64 *seg_q = (i64 *)(tb + o_sq), *seg_i = (i64 *)(tb + o_si), *seg_u = (i64 *)(tb + o_su), *partial = (i64 *)(tb + o_part);
    CK(cudaMemsetAsync(seg_start, 0, (size_t)n_seg * sizeof(u32), st));
    CK(cudaMemsetAsync(seg_end, 0, (size_t)n_seg * sizeof(u32), st));
    const int g = grid_for(U, 256, 148 * 8);
    { ProfScope ps(ctx, KID_K3_KEYS);
      k3_make_keys<<<g, 256, 0, st>>>(ctx->d_round_of, ctx->d_ref_of, ctx->d_ctr, n_mir, keys[0], vals[0]); }
    int bits = 1;
    while ((1 << bits) < n_seg) bits++;
    int cur = 0;
    if (n_tiles) {
        ProfScope ps(ctx, KID_K3_SORT, 0);
        const int gt = (int)(n_tiles < 148u * 8u ? n_tiles : 148u * 8u);
        for (int shift = 0; shift < bits; shift += 8) {
            k3_rs_hist<<<gt, RS_THREADS, 0, st>>>(keys[cur], ctx->d_ctr, shift, n_tiles, hist);
            ctx->n_launches += 2 + device_excl_scan(hist, hist, hist_n, scratch, st);
            k3_rs_scatter<<<gt, RS_THREADS, 0, st>>>(keys[cur], vals[cur], ctx->d_ctr, shift, n_tiles, hist, keys[cur ^ 1], vals[cur ^ 1]);
            cur ^= 1;
        }
    }
    { ProfScope ps(ctx, KID_K3_BOUNDS);
      k3_seg_bounds<<<g, 256, 0, st>>>(keys[cur], ctx->d_ctr, seg_start, seg_end); }
    { ProfScope ps(ctx, KID_K3_REDUCE, 3);
      k3_seg_chunks<<<(n_seg + 1 + 255) / 256, 256, 0, st>>>(seg_start, seg_end, n_seg, chunk_off);
      ctx->n_launches += device_excl_scan(chunk_off, chunk_off, (u64)n_seg + 1, scratch, st);
      k3_chunk_reduce<<<grid_for((u64)max_chunks * 32, 256, 148 * 8), 256, 0, st>>>(vals[cur], seg_start, seg_end, chunk_off, n_seg,
                                                                                   ctx->d_counts, ctx->d_round_of, S, ctx->max_unique, partial);
      k3_seg_reduce<<<grid_for((u64)n_seg * 32, 256, 148 * 8), 256, 0, st>>>(partial, chunk_off, n_seg, S, seg_q, seg_i, seg_u); }
    { ProfScope ps(ctx, KID_K3_FINAL);
      k3_finalize<<<S, 256, 0, st>>>(seg_q, seg_i, seg_u, n_mir, S, ctx->d_mir_quant, ctx->d_mir_iscan, ctx->d_lib_totals); }
    CK(cudaGetLastError());
    ctx->tallied = true;
    return MIRGE_OK;
}

// ---------------------------------------------------------------------------
// export
// ---------------------------------------------------------------------------
__global__ void k_counts_rows(const u32 *__restrict__ counts, u64 cstride, u64 U, int S, u32 *__restrict__ out)
{
    for (u64 it = (u64)blockIdx.x * blockDim.x + threadIdx.x; it < U * (u64)S; it += (u64)gridDim.x * blockDim.x)
        out[it] = counts[(it % S) * cstride + it / S];
}

// out[u * stride + i] = base i of unique u as ASCII (N for ambiguous), zero beyond its length
template <int W>
__global__ void k_unpack_unique(const u64 *__restrict__ useq, const u64 *__restrict__ unmask, const u8 *__restrict__ ulen, u64 U, int stride,
                                unsigned char *__restrict__ out)
{
    const int gpu = (stride + 3) / 4;                       // 4-base groups per unique
    for (u64 it = (u64)blockIdx.x * blockDim.x + threadIdx.x; it < U * (u64)gpu; it += (u64)gridDim.x * blockDim.x) {
        const u64 u = it / gpu;
        const int g = (int)(it % gpu), n = ulen[u];
        const u32 two = g * 4 < 32 * W ? (u32)((useq[u * W + (g >> 3)] >> ((g & 7) * 8)) & 0xFFull) : 0u;
        const u32 nb = g * 4 < 32 * W ? (u32)((unmask[u * (W / 2) + (g >> 4)] >> ((g & 15) * 4)) & 0xFull) : 0u;
        for (int k = 0; k < 4; k++) {
            const int i = 4 * g + k;
            if (i >= stride) break;
            unsigned char c = 0;
            if (i < n) c = ((nb >> k) & 1u) ? 'N' : "ACGT"[(two >> (2 * k)) & 3u];
            out[u * (u64)stride + i] = c;
        }
    }
}

int mirge_export_unique(mirge_ctx *ctx, uint64_t cap, int stride, char *seq, uint8_t *len, uint32_t *counts,
                        int8_t *round_of, int32_t *ref_of, int32_t *pos_of, int8_t *mm_of)
{
    u64 U = 0;
    int rc = mirge_get_stats(ctx, nullptr, nullptr, (uint64_t *)&U);
    if (rc) return rc;
    if (cap < U) return fail(ctx, MIRGE_E_ARG, "mirge_export_unique: capacity smaller than n_unique");
    if (seq && stride < ctx->max_len) return fail(ctx, MIRGE_E_ARG, "mirge_export_unique: stride smaller than max_read_len");
    const int W = ctx->W;
    if (len) CK(cudaMemcpy(len, ctx->d_ulen, U, cudaMemcpyDeviceToHost));
    if (seq && U) {
        // 2-bit -> ASCII on the device (four bases per thread, coalesced), one copy back
        unsigned char *d_out = nullptr;
        CK(cudaMalloc(&d_out, (size_t)U * stride));
        const u64 items = U * (u64)((stride + 3) / 4);
        ctx->n_launches++;
        if (W == 2) k_unpack_unique<2><<<grid_for(items, 256, 148 * 16), 256, 0, ctx->stream>>>(ctx->d_useq, ctx->d_unmask, ctx->d_ulen, U, stride, d_out);
        else k_unpack_unique<4><<<grid_for(items, 256, 148 * 16), 256, 0, ctx->stream>>>(ctx->d_useq, ctx->d_unmask, ctx->d_ulen, U, stride, d_out);
        cudaError_t e = cudaMemcpyAsync(seq, d_out, (size_t)U * stride, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        cudaFree(d_out);
        if (e != cudaSuccess) { ctx->err = std::string("mirge_export_unique failed: ") + cudaGetErrorString(e); return MIRGE_E_CUDA; }
    }
    if (counts && U) {
        // device layout is a column per sample; the caller gets counts[u][s]
        u32 *d_t = nullptr;
        CK(cudaMalloc(&d_t, (size_t)U * ctx->S * sizeof(u32)));
        ctx->n_launches++;
        k_counts_rows<<<grid_for(U * (u64)ctx->S, 256, 148 * 16), 256, 0, ctx->stream>>>(ctx->d_counts, ctx->max_unique, U, ctx->S, d_t);
        cudaError_t e = cudaMemcpyAsync(counts, d_t, (size_t)U * ctx->S * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        cudaFree(d_t);
        if (e != cudaSuccess) { ctx->err = std::string("mirge_export_unique failed: ") + cudaGetErrorString(e); return MIRGE_E_CUDA; }
    }
    if (ctx->annotated) {
        if (round_of) CK(cudaMemcpy(round_of, ctx->d_round_of, U, cudaMemcpyDeviceToHost));
        if (ref_of) CK(cudaMemcpy(ref_of, ctx->d_ref_of, U * sizeof(int), cudaMemcpyDeviceToHost));
        if (pos_of) CK(cudaMemcpy(pos_of, ctx->d_pos_of, U * sizeof(int), cudaMemcpyDeviceToHost));
        if (mm_of) CK(cudaMemcpy(mm_of, ctx->d_mm_of, U, cudaMemcpyDeviceToHost));
    } else {
        if (round_of) memset(round_of, 0, U);
    }
    return MIRGE_OK;
}

int mirge_export_tables(mirge_ctx *ctx, int64_t *mir_quant, int64_t *mir_iscan, int64_t *lib_totals, uint64_t *len_hist,
                        uint64_t *round_processed, uint64_t *round_aligned)
{
    int rc = mirge_sync(ctx);
    if (rc) return rc;
    const int S = ctx->S;
    if ((mir_quant || mir_iscan || lib_totals) && !ctx->tallied) return fail(ctx, MIRGE_E_STATE, "mirge_export_tables called before mirge_tally");
    const size_t nm = (size_t)ctx->libs[MIRGE_LIB_MIRNA].n_refs;
    if (mir_quant && nm) CK(cudaMemcpy(mir_quant, ctx->d_mir_quant, nm * S * sizeof(i64), cudaMemcpyDeviceToHost));
    if (mir_iscan && nm) CK(cudaMemcpy(mir_iscan, ctx->d_mir_iscan, nm * S * sizeof(i64), cudaMemcpyDeviceToHost));
    if (lib_totals) CK(cudaMemcpy(lib_totals, ctx->d_lib_totals, (size_t)S * MIRGE_N_TOTALS * sizeof(i64), cudaMemcpyDeviceToHost));
    if (len_hist) CK(cudaMemcpy(len_hist, ctx->d_len_hist, (size_t)(ctx->max_len + 1) * S * sizeof(u64), cudaMemcpyDeviceToHost));
    if (round_processed) CK(cudaMemcpy(round_processed, ctx->d_round_stats, MIRGE_N_ROUNDS * sizeof(u64), cudaMemcpyDeviceToHost));
    if (round_aligned) CK(cudaMemcpy(round_aligned, ctx->d_round_stats + MIRGE_N_ROUNDS, MIRGE_N_ROUNDS * sizeof(u64), cudaMemcpyDeviceToHost));
    return MIRGE_OK;
}

// ---------------------------------------------------------------------------
// multi-GPU exchange
// ---------------------------------------------------------------------------
int mirge_record_bytes(const mirge_ctx *ctx)
{
    return 8 * (ctx->W + ctx->W / 2) + 8;
}

__global__ void k_excl_scan_small(const u64 *in, u64 *out, int n)
{
    u64 acc = 0;
    for (int i = 0; i < n; i++) { out[i] = acc; acc += in[i]; }
}

int mirge_export_records_device(mirge_ctx *ctx, int n_parts, void *dev_buf, uint64_t cap_records, uint64_t *part_counts_dev)
{
    if (n_parts < 1 || n_parts > 1024 || !part_counts_dev || !dev_buf) return fail(ctx, MIRGE_E_ARG, "mirge_export_records_device: bad argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (!ctx->d_part) CK(cudaMalloc(&ctx->d_part, 2 * 1024 * sizeof(u64)));      // kept: cudaFree would synchronise the device every step
    u64 *d_cnt = ctx->d_part;
    CK(cudaMemsetAsync(d_cnt, 0, 2 * (size_t)n_parts * sizeof(u64), st));
    const int rb = mirge_record_bytes(ctx), g = grid_for(ctx->max_unique, 128, 148 * 8);
    {
        ProfScope ps(ctx, KID_EXCHANGE, 3);
        if (ctx->W == 2) {
            k1_export_records<2><<<g, 128, 0, st>>>(table_view<2>(ctx), n_parts, rb, 1, cap_records, d_cnt, nullptr, nullptr);
            k_excl_scan_small<<<1, 1, 0, st>>>(d_cnt, d_cnt + n_parts, n_parts);
            k1_export_records<2><<<g, 128, 0, st>>>(table_view<2>(ctx), n_parts, rb, 0, cap_records, d_cnt, d_cnt + n_parts, (unsigned char *)dev_buf);
        } else {
            k1_export_records<4><<<g, 128, 0, st>>>(table_view<4>(ctx), n_parts, rb, 1, cap_records, d_cnt, nullptr, nullptr);
            k_excl_scan_small<<<1, 1, 0, st>>>(d_cnt, d_cnt + n_parts, n_parts);
            k1_export_records<4><<<g, 128, 0, st>>>(table_view<4>(ctx), n_parts, rb, 0, cap_records, d_cnt, d_cnt + n_parts, (unsigned char *)dev_buf);
        }
    }
    CK(cudaMemcpyAsync(part_counts_dev, d_cnt, (size_t)n_parts * sizeof(u64), cudaMemcpyDeviceToDevice, st));
    CK(cudaGetLastError());
    return MIRGE_OK;
}

int mirge_reset(mirge_ctx *ctx)
{
    int rc = mirge_reset_unique(ctx);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    CK(cudaMemsetAsync(ctx->d_ctr, 0, (6 + 2 * (size_t)ctx->S) * sizeof(u64), st));
    CK(cudaMemsetAsync(ctx->d_len_hist, 0, (size_t)(ctx->max_len + 1) * ctx->S * sizeof(u64), st));
    CK(cudaMemsetAsync(ctx->d_round_stats, 0, 2 * MIRGE_N_ROUNDS * sizeof(u64), st));
    return MIRGE_OK;
}

uint64_t mirge_launch_count(const mirge_ctx *ctx) { return ctx->n_launches; }

int mirge_debug(mirge_ctx *ctx, int force_slow, uint64_t *slow_reads)
{
    if (force_slow >= 0) ctx->force_slow = force_slow;
    if (slow_reads) {
        int rc = mirge_sync(ctx);
        if (rc) return rc;
        CK(cudaMemcpy(slow_reads, ctx->d_ctr + 2 + ctx->S, sizeof(u64), cudaMemcpyDeviceToHost));
    }
    return MIRGE_OK;
}

int mirge_trim_classes(mirge_ctx *ctx, uint64_t *out4)
{
    int rc = mirge_sync(ctx);
    if (rc) return rc;
    CK(cudaMemcpy(out4, ctx->d_ctr + 2 + ctx->S, 4 * sizeof(u64), cudaMemcpyDeviceToHost));
    return MIRGE_OK;
}

int mirge_reset_unique(mirge_ctx *ctx)
{
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t esz = ctx->W == 2 ? sizeof(TEntry<2>) : sizeof(TEntry<4>);
    ProfScope ps(ctx, KID_RESET);
    CK(cudaMemsetAsync(ctx->d_slots, 0, ctx->table_cap * esz, st));
    // count rows at or beyond the unique counter were never touched since the last clear
    k1_clear_counts<<<grid_for(ctx->max_unique, 256, 148 * 8), 256, 0, st>>>(ctx->d_counts, ctx->d_ctr, ctx->max_unique, ctx->S);
    CK(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(u64), st));
    { int rc = annot_clear(ctx); if (rc) return rc; }          // ids start again at 0: nothing is annotated
    ctx->annotated = false; ctx->tallied = false;
    return MIRGE_OK;
}

int mirge_import_records_device(mirge_ctx *ctx, const void *dev_buf, uint64_t n_records)
{
    if (n_records == 0) return MIRGE_OK;
    CK(cudaSetDevice(ctx->device));
    const int rb = mirge_record_bytes(ctx), g = grid_for(n_records, 128, 148 * 8);
    ProfScope ps(ctx, KID_IMPORT);
    if (ctx->W == 2) k1_import_records<2><<<g, 128, 0, ctx->stream>>>((const unsigned char *)dev_buf, n_records, rb, table_view<2>(ctx));
    else k1_import_records<4><<<g, 128, 0, ctx->stream>>>((const unsigned char *)dev_buf, n_records, rb, table_view<4>(ctx));
    CK(cudaGetLastError());
    ctx->annotated = false; ctx->tallied = false;
    return MIRGE_OK;
}

// Packed integer tables for the multi-GPU reduction (X2): everything a rank contributes to the report.
static void table_pieces(mirge_ctx *ctx, void **ptr, size_t *words)
{
    const size_t nm = (size_t)ctx->libs[MIRGE_LIB_MIRNA].n_refs, S = (size_t)ctx->S;
    ptr[0] = ctx->d_mir_quant; words[0] = nm * S;
    ptr[1] = ctx->d_mir_iscan; words[1] = nm * S;
    ptr[2] = ctx->d_lib_totals; words[2] = S * MIRGE_N_TOTALS;
    ptr[3] = ctx->d_len_hist; words[3] = (size_t)(ctx->max_len + 1) * S;
    ptr[4] = ctx->d_ctr + 2; words[4] = S;                  // kept
    ptr[5] = ctx->d_ctr + 6 + S; words[5] = S;              // processed
}

uint64_t mirge_tables_words(const mirge_ctx *ctx)
{
    void *p[6]; size_t w[6];
    table_pieces(const_cast<mirge_ctx *>(ctx), p, w);
    return w[0] + w[1] + w[2] + w[3] + w[4] + w[5];
}

int mirge_pack_tables_device(mirge_ctx *ctx, void *dst_dev)
{
    if (!ctx->tallied) return fail(ctx, MIRGE_E_STATE, "mirge_pack_tables_device called before mirge_tally");
    CK(cudaSetDevice(ctx->device));
    void *p[6]; size_t w[6];
    table_pieces(ctx, p, w);
    i64 *d = (i64 *)dst_dev;
    for (int k = 0; k < 6; k++) {
        if (w[k]) CK(cudaMemcpyAsync(d, p[k], w[k] * sizeof(i64), cudaMemcpyDeviceToDevice, ctx->stream));
        d += w[k];
    }
    return MIRGE_OK;
}

int mirge_unpack_tables_device(mirge_ctx *ctx, const void *src_dev)
{
    if (!ctx->tallied) return fail(ctx, MIRGE_E_STATE, "mirge_unpack_tables_device called before mirge_tally");
    CK(cudaSetDevice(ctx->device));
    void *p[6]; size_t w[6];
    table_pieces(ctx, p, w);
    const i64 *d = (const i64 *)src_dev;
    for (int k = 0; k < 6; k++) {
        if (w[k]) CK(cudaMemcpyAsync(p[k], d, w[k] * sizeof(i64), cudaMemcpyDeviceToDevice, ctx->stream));
        d += w[k];
    }
    return MIRGE_OK;
}

int mirge_prefilter_tables(const char *adapter, double error_rate, int min_overlap, int *m_out, int *icap_out,
                           uint32_t *safe_p, uint32_t *safe_len, uint8_t *fullreach)
{
    if (!adapter) return MIRGE_E_ARG;
    TrimParams tp;
    memset(&tp, 0, sizeof tp);
    tp.mode = 1; tp.n_adapters = 1; tp.error_rate = error_rate; tp.min_overlap = min_overlap;
    const size_t n = strlen(adapter);
    if (n == 0 || n > MIRGE_MAX_ADAPTER_LEN) return MIRGE_E_ARG;
    tp.alen[0] = (int)n;
    for (size_t i = 0; i < n; i++) {
        const int c = base_code((unsigned char)adapter[i]);
        if (c > 3) return MIRGE_E_ARG;
        tp.aw[0][i >> 5] |= (u64)c << ((i & 31) * 2);
    }
    K1Thr thr;
    for (int i = 0; i <= MIRGE_MAX_ADAPTER_LEN; i++) {
        int c = -1;
        if (i >= min_overlap) { c = 0; while ((double)(c + 1) <= (double)i * error_rate) c++; }
        thr.v[i] = (signed char)c;
    }
    K1Pre pre;
    k1pre_build(tp, thr, pre);
    if (m_out) *m_out = pre.m;
    if (icap_out) *icap_out = pre.icap;
    if (safe_p) *safe_p = pre.safe_p;
    if (safe_len) memcpy(safe_len, pre.safe_len, sizeof pre.safe_len);
    if (fullreach) memcpy(fullreach, pre.fullreach, sizeof pre.fullreach);
    return pre.enabled;
}

uint64_t mirge_hash_sequence(const char *seq, int len)
{
    u64 s[4] = {0, 0, 0, 0}, nm[2] = {0, 0};
    if (len > 128) len = 128;
    for (int i = 0; i < len; i++) {
        int c = base_code((unsigned char)seq[i]);
        if (c > 3) { nm[i >> 6] |= 1ull << (i & 63); c = 0; }
        s[i >> 5] |= (u64)c << ((i & 31) * 2);
    }
    return owner_hash(key_hash_words(s, 4, nm, len));
}

// ---------------------------------------------------------------------------
// measurement
// ---------------------------------------------------------------------------
int mirge_profile_enable(mirge_ctx *ctx, int on)
{
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto &s : ctx->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    ctx->spans.clear();
    memset(ctx->prof_ms, 0, sizeof ctx->prof_ms); memset(ctx->prof_n, 0, sizeof ctx->prof_n);
    ctx->prof = on != 0;
    return MIRGE_OK;
}

int mirge_profile_read(mirge_ctx *ctx, double *ms, uint64_t *launches, const char **names)
{
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto &s : ctx->spans) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) { ctx->prof_ms[s.kid] += t; ctx->prof_n[s.kid]++; }
        cudaEventDestroy(s.a); cudaEventDestroy(s.b);
    }
    ctx->spans.clear();
    for (int k = 0; k < MIRGE_N_KERNELS; k++) {
        if (ms) ms[k] = ctx->prof_ms[k];
        if (launches) launches[k] = ctx->prof_n[k];
        if (names) names[k] = KERNEL_NAMES[k];
    }
    return MIRGE_OK;
}

