// k1_kernel.cuh -- K1: 3' adapter trim + length filter + collapse.
//
// Per read (one thread each, packed record streamed with one 128-bit load):
//   1. apply the quality-trim window computed by K0;
//   2. find the 3' adapter exactly as cutadapt would:
//        prefilter  - exact comparisons plus host-proved dominance tables settle most
//                     reads with no dynamic programming at all (k1_prefilter.cuh);
//        class 1    - reads for which only a full-length candidate in a bounded column range
//                     could outrank the exact prefix: settled in place by the pigeonhole
//                     piece filter (k1_piece_hit), no dynamic programming either;
//        class 2    - general search: Myers' bit-vector edit distance (the <=32-nt
//                     adapter is one machine word) gives the cost of EVERY DP cell of
//                     the last columns; the few candidate end cells cutadapt ranks by
//                     (matches, cost, visiting order) are resolved by walking
//                     cutadapt's own predecessor preference (diagonal, then
//                     insertion, then deletion) backwards until a zero-cost cell,
//                     reading cell costs from the bit-vectors of the last 8 columns
//                     (kept in shared memory);
//        class 3    - reads whose candidates cannot be settled inside that window
//                     (sequencing errors deep inside the adapter, a full inexact
//                     adapter occurrence, adapters > 32 nt): the same search with all
//                     columns kept, or the literal banded DP (locate_back);
//   3. length filter, then warp-collective insert into the unique-sequence table.
// The classes run as three kernels connected by device queues of 32-byte read records (end of this file).
// Every path returns cutadapt's rstart; parity tests compare them read by read.
#pragma once
#include "common.cuh"
#include "k1_trim.cuh"
#include "k1_prefilter.cuh"

#define K1_THREADS 128
#define K1_RING 8

struct K1Counters {
    u64 *kept;        // [S]
    u64 *len_hist;    // [(max_len + 1)][S]
    u64 *slow_reads;  // diagnostics: [0] reads through the all-column search, [1] class-1 reads, [2] class-2 reads
};

// thr[i] = largest cost c with c <= i * error_rate (double), or -1 below min_overlap
struct K1Thr { signed char v[MIRGE_MAX_ADAPTER_LEN + 1]; };

template <int W>
__device__ __forceinline__ void k1_load_record(const u64 *__restrict__ bases, const u64 *__restrict__ nmask,
                                               const u32 *__restrict__ meta, u64 r, PSeq<W> &p, int &n)
{
    if (W == 2) {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(bases) + r);
        p.s[0] = v.x; p.s[1] = v.y;
    } else {
#pragma unroll
        for (int w = 0; w < W; w++) p.s[w] = __ldg(bases + r * W + w);
    }
#pragma unroll
    for (int w = 0; w < W / 2; w++) p.nm[w] = __ldg(nmask + r * (W / 2) + w);
    const u32 mt = __ldg(meta + r);
    const int qstart = (mt >> 8) & 0xFF, qstop = (mt >> 16) & 0xFF;
    n = qstop - qstart;
    pseq_shift<W>(p, qstart);
    pseq_truncate<W>(p, n);
}

// cost of DP cell (i, j) from the ring of column bit-vectors; column 0 is cost i.
__host__ __device__ __forceinline__ int k1_cell_cost(const u64 *ring, int i, int j)
{
    if (j == 0) return i;
    const u64 v = ring[(j & (K1_RING - 1)) * K1_THREADS];
    const u32 low = (i >= 32) ? 0xFFFFFFFFu : ((1u << i) - 1u);
    return bit_popc32((u32)v & low) - bit_popc32((u32)(v >> 32) & low);
}

#define K1_NOMATCH (-1)
#define K1_SLOW (-2)

// Fast adapter search for one adapter of length m <= 32, executed by all 32
// lanes of a warp in lock step (every loop has a warp-uniform trip count, so
// lanes never serialise behind each other's data-dependent work).
//
// Per column j the ring keeps (Pv_j, Mv_j) -- vertical cost deltas, from which any
// cell cost is two popcounts -- and MhS_j, the shifted negative horizontal deltas.
// cutadapt's choice at a mismatching cell (i, j) follows from those bits alone:
//   cost_diag <= cost_ins  <=>  cost(i-1,j-1) <= cost(i-1,j)  <=>  bit i-1 of MhS_j clear
//   cost_diag <= cost_del  <=>  cost(i-1,j-1) <= cost(i,j-1)  <=>  bit i-1 of Mv_{j-1} clear
// and when the diagonal loses, insertion wins iff the MhS_j bit is set
// (cost_ins <= cost_del  <=>  hdelta <= vdelta, with one of them -1).
// Returns rstart >= 0 (matches in matches_out), K1_NOMATCH, or K1_SLOW.
// `ring`/`ringh` point at this thread's slot 0 (stride K1_THREADS between slots).
template <int W>
__host__ __device__ __forceinline__ int k1_match_fast(const PSeq<W> &p, int n, const TrimParams &tp, const signed char *thr, int a,
                                             u64 *ring, u32 *ringh, const u32 *peq8, int &matches_out)
{
    // peq8[c | isN << 2]: match mask of read base c against the adapter (0 for N)
    const int m = tp.alen[a];
    const u64 aw = tp.aw[a][0];
    u32 Pv = (m >= 32) ? 0xFFFFFFFFu : ((1u << m) - 1u), Mv = 0;
    int score = m;
    const int thr_m = thr[m];
    const int win_lo = n - (K1_RING - 1);          // first column kept in the ring
    const int top = m - 1;
    u32 full_mask = 0;                             // bit (j - win_lo): inexact full-length candidate ending at column j
    bool slow = false;
    int exact_j = -1;
    const int nmax = warp_max_all(n);
    for (int j0 = 0; j0 < nmax; j0 += 16) {        // 16 bases (one 32-bit word) per outer step
        u32 bw = 0, nb = 0;
#pragma unroll
        for (int w = 0; w < W; w++) {
            if ((j0 >> 5) == w) {
                bw = (u32)(p.s[w] >> ((j0 & 16) * 2));
                nb = (u32)(p.nm[w >> 1] >> (((w & 1) * 32) + (j0 & 16))) & 0xFFFFu;
            }
        }
        const int jend = (nmax - j0) < 16 ? (nmax - j0) : 16;
        u32 le16 = 0, z16 = 0;                     // bit t: cost(m, j0+t+1) <= thr_m / == 0
#pragma unroll 4
        for (int t = 0; t < jend; t++) {
            const int j = j0 + t + 1;
            const u32 Eq = peq8[(bw & 3u) | (((nb >> t) & 1u) << 2)];
            bw >>= 2;
            if (j <= n) {
                const u32 Xv = Eq | Mv;
                const u32 Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                u32 Ph = Mv | ~(Xh | Pv);
                u32 Mh = Pv & Xh;
                score += (int)((Ph >> top) & 1u) - (int)((Mh >> top) & 1u);
                Ph <<= 1; Mh <<= 1;
                Pv = Mh | ~(Xv | Ph);
                Mv = Ph & Xv;
                le16 |= (u32)(score <= thr_m) << t;
                z16 |= (u32)(score == 0) << t;
                if (j >= win_lo) {
                    const int slot = (j & (K1_RING - 1)) * K1_THREADS;
                    ring[slot] = (u64)Pv | ((u64)Mv << 32);
                    ringh[slot] = Mh;
                }
            }
        }
        // once per 16 columns: leftmost exact occurrence (str.find), full-length candidates
        // inside the window, and inexact full-length occurrences far from the read end
        if (z16 && exact_j < 0) exact_j = j0 + bit_ffs32(z16);
        if (le16) {
            int below = win_lo - 1 - j0;           // columns of this chunk that lie before the window
            below = below < 0 ? 0 : (below > 16 ? 16 : below);
            if (le16 & ((1u << below) - 1u)) slow = true;
            if (below < 16) full_mask |= (le16 >> below) << (j0 + below + 1 - win_lo);
        }
    }
    // cutadapt keeps the candidate with most matches, then lowest cost, then first
    // visited: full-length candidates by ascending end column, then adapter
    // prefixes i = m .. 0 ending at the read's end.
    u64 valid = 0;
    {
        int c = 0;
        for (int i = 1; i <= m; i++) {
            c += (int)((Pv >> (i - 1)) & 1u) - (int)((Mv >> (i - 1)) & 1u);
            if (c <= thr[i]) valid |= 1ull << i;
        }
        if (n >= 1 && ((full_mask >> (n - win_lo)) & 1u)) valid &= ~(1ull << m);   // same cell as the full candidate at column n
    }
    int best_m = 0, best_c = m + n, best_o = K1_NOMATCH, best_v = 0x7fffffff;
    // per-lane state machine: state 0 = fetch next candidate, 1 = walking back, 2 = finished
    int state = (exact_j >= 0 || slow) ? 2 : 0;
    int ci = 0, cj = 0, cc = 0, c0 = 0, mt = 0, visit = 0;
    while (warp_any_all(state != 2)) {
        if (state == 0) {
            int i = 0, j = 0;
            bool have = true;
            if (full_mask) {
                const int t = bit_ffs32(full_mask) - 1;
                full_mask &= full_mask - 1;
                i = m; j = win_lo + t; visit = j;
            } else if (valid) {
                i = 63 - bit_clz64(valid);
                valid &= ~(1ull << i);
                j = n; visit = n + 1 + (m - i);
                if (i < best_m) { have = false; state = 2; }          // matches <= i: nothing left can win
            } else {
                have = false; state = 2;
            }
            if (have) {
                const int c = k1_cell_cost(ring, i, j);
                if (j == n && i == best_m && c >= best_c) {
                    // could at best tie, and is visited later: skip
                } else {
                    ci = i; cj = j; cc = c; c0 = c; mt = 0; state = 1;
                    // common shape: c unmatched adapter bases at the read's end, i.e. c insertions
                    // straight down column j to an exact cell -- settle it without stepping
                    if (c > 0 && c < i) {
                        const u32 chain = ((1u << c) - 1u) << (i - c);           // rows i-c+1 .. i
                        const u32 mhs = ringh[(j & (K1_RING - 1)) * K1_THREADS];
                        const u32 eqj = peq8[pseq_base<W>(p, j - 1) | (pseq_isn<W>(p, j - 1) << 2)];
                        if ((mhs & chain) == chain && (eqj & chain) == 0u) { ci = i - c; cc = 0; }
                    }
                }
            }
        }
        if (state == 1) {
            // one step of cutadapt's predecessor choice, back towards a zero-cost (exact) cell
            int o = -1;
            bool done = false;
            if (ci == 0) { o = cj; done = true; }
            else if (cc == 0) { mt += ci; o = cj - ci; done = true; }
            else if (cj == 0) { o = 0; done = true; }
            else {
                // matches run diagonally: take the whole run of equal bases in one step
                const int d = cj - ci;
                const u32 bit = 1u << (ci - 1);
                int run = 0;
                if (d >= 0) {
                    const u64 x = words_window64<W>(p.s, d) ^ aw;
                    u64 mm = ((x | (x >> 1)) & 0x5555555555555555ull) | spread32(mask_window32<W / 2>(p.nm, d));
                    if (ci < 32) mm &= (1ull << (2 * ci)) - 1ull;
                    run = mm ? ci - 1 - ((63 - bit_clz64(mm)) >> 1) : ci;
                } else if (peq8[pseq_base<W>(p, cj - 1) | (pseq_isn<W>(p, cj - 1) << 2)] & bit) {
                    run = 1;
                }
                if (run > 0) { mt += run; ci -= run; cj -= run; }
                else if (cj < win_lo || (cj - 1 != 0 && cj - 1 < win_lo)) { slow = true; state = 2; }   // needs a column outside the window
                else {
                    const u32 mhs = ringh[(cj & (K1_RING - 1)) * K1_THREADS];
                    const u32 mvp = (cj - 1 == 0) ? 0u : (u32)(ring[((cj - 1) & (K1_RING - 1)) * K1_THREADS] >> 32);
                    cc--;
                    if (mhs & bit) ci--;                  // insertion: from (i-1, j)
                    else if (mvp & bit) cj--;             // deletion: from (i, j-1)
                    else { ci--; cj--; }                  // mismatch: diagonal
                }
            }
            if (done) {
                if (mt > best_m || (mt == best_m && (c0 < best_c || (c0 == best_c && visit < best_v)))) {
                    best_m = mt; best_c = c0; best_o = o; best_v = visit;
                }
                state = 0;
            }
        }
    }
    if (exact_j >= 0) { matches_out = m; return exact_j - m; }
    if (slow) return K1_SLOW;
    matches_out = best_m;
    return best_o;
}

// AdapterCutter(times=1) over all adapters: most matches wins, first wins ties.
// Returns the new length, or K1_SLOW.  Warp-convergent.
template <int W>
__host__ __device__ __forceinline__ int k1_trim_fast(const PSeq<W> &p, int n, const TrimParams &tp, const signed char *thr, u64 *ring, u32 *ringh, const u32 *peq8)
{
    int best_rstart = -1, best_matches = -1;
    bool slow = false;
    for (int a = 0; a < tp.n_adapters; a++) {
        if (tp.alen[a] > 32) { slow = true; continue; }
        int mt = 0;
        const int rs = k1_match_fast<W>(p, n, tp, thr, a, ring, ringh, peq8 + 8 * a, mt);
        if (rs == K1_SLOW) slow = true;
        else if (rs >= 0 && (best_rstart < 0 || mt > best_matches)) { best_rstart = rs; best_matches = mt; }
    }
    if (slow) return K1_SLOW;
    return best_rstart >= 0 ? best_rstart : n;
}

// Per-thread variant of k1_match_fast for the reads the windowed version could not
// settle (walks that leave the 8-column ring, full-length inexact occurrences far
// from the read end): the same bit-vector algorithm, but the vectors of EVERY column
// are kept (thread-local arrays), so nothing is out of reach.  Not warp-collective;
// run by converged threads when the slow queue is drained.
template <int W>
__host__ __device__ int k1_match_full(const PSeq<W> &p, int n, const TrimParams &tp, const signed char *thr, int a,
                             const u32 *peq8, int &matches_out)
{
    const int m = tp.alen[a];
    const u64 aw = tp.aw[a][0];
    u64 cv[32 * W];        // column j -> cv[j-1] = Pv | Mv << 32
    u32 ch[32 * W];        // column j -> ch[j-1] = shifted negative horizontal deltas
    u64 fullc[W / 2];      // bit j-1: inexact full-length candidate ending at column j
#pragma unroll
    for (int w = 0; w < W / 2; w++) fullc[w] = 0;
    u32 Pv = (m >= 32) ? 0xFFFFFFFFu : ((1u << m) - 1u), Mv = 0;
    int score = m;
    const int thr_m = thr[m], top = m - 1;
    for (int j = 1; j <= n; j++) {
        const u32 Eq = peq8[pseq_base<W>(p, j - 1) | (pseq_isn<W>(p, j - 1) << 2)];
        const u32 Xv = Eq | Mv;
        const u32 Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
        u32 Ph = Mv | ~(Xh | Pv);
        u32 Mh = Pv & Xh;
        score += (int)((Ph >> top) & 1u) - (int)((Mh >> top) & 1u);
        Ph <<= 1; Mh <<= 1;
        Pv = Mh | ~(Xv | Ph);
        Mv = Ph & Xv;
        cv[j - 1] = (u64)Pv | ((u64)Mv << 32);
        ch[j - 1] = Mh;
        if (score == 0) { matches_out = m; return j - m; }           // str.find: leftmost exact occurrence
        if (score <= thr_m) {
#pragma unroll
            for (int w = 0; w < W / 2; w++) if (((j - 1) >> 6) == w) fullc[w] |= 1ull << ((j - 1) & 63);
        }
    }
    u64 valid = 0;
    {
        int c = 0;
        for (int i = 1; i <= m; i++) {
            c += (int)((Pv >> (i - 1)) & 1u) - (int)((Mv >> (i - 1)) & 1u);
            if (c <= thr[i]) valid |= 1ull << i;
        }
    }
    int best_m = 0, best_c = m + n, best_o = K1_NOMATCH, best_v = 0x7fffffff;
    int fw = 0;            // full-candidate word being scanned
    for (;;) {
        int i, j, visit;
        while (fw < W / 2) {
            u64 f = 0;
#pragma unroll
            for (int w = 0; w < W / 2; w++) if (fw == w) f = fullc[w];
            if (f) break;
            fw++;
        }
        if (fw < W / 2) {
            u64 f = 0;
#pragma unroll
            for (int w = 0; w < W / 2; w++) if (fw == w) f = fullc[w];
            const int t = bit_ffs64(f) - 1;
#pragma unroll
            for (int w = 0; w < W / 2; w++) if (fw == w) fullc[w] &= fullc[w] - 1;
            i = m; j = fw * 64 + t + 1; visit = j;
        } else if (valid) {
            i = 63 - bit_clz64(valid);
            valid &= ~(1ull << i);
            if (i < best_m) break;
            j = n; visit = n + 1 + (m - i);
        } else {
            break;
        }
        const u64 vj = cv[j - 1];
        const u32 low = (i >= 32) ? 0xFFFFFFFFu : ((1u << i) - 1u);
        const int c = bit_popc32((u32)vj & low) - bit_popc32((u32)(vj >> 32) & low);
        if (j == n && i == best_m && c >= best_c) continue;
        int ci = i, cj = j, cc = c, mt = 0, o = -1;
        for (;;) {
            if (ci == 0) { o = cj; break; }
            if (cc == 0) { mt += ci; o = cj - ci; break; }
            if (cj == 0) { o = 0; break; }
            const int d = cj - ci;
            const u32 bit = 1u << (ci - 1);
            int run = 0;
            if (d >= 0) {
                const u64 x = words_window64<W>(p.s, d) ^ aw;
                u64 mm = ((x | (x >> 1)) & 0x5555555555555555ull) | spread32(mask_window32<W / 2>(p.nm, d));
                if (ci < 32) mm &= (1ull << (2 * ci)) - 1ull;
                run = mm ? ci - 1 - ((63 - bit_clz64(mm)) >> 1) : ci;
            } else if (peq8[pseq_base<W>(p, cj - 1) | (pseq_isn<W>(p, cj - 1) << 2)] & bit) {
                run = 1;
            }
            if (run > 0) { mt += run; ci -= run; cj -= run; continue; }
            const u32 mhs = ch[cj - 1];
            const u32 mvp = (cj - 1 == 0) ? 0u : (u32)(cv[cj - 2] >> 32);
            cc--;
            if (mhs & bit) ci--;
            else if (mvp & bit) cj--;
            else { ci--; cj--; }
        }
        if (mt > best_m || (mt == best_m && (c < best_c || (c == best_c && visit < best_v)))) {
            best_m = mt; best_c = c; best_o = o; best_v = visit;
        }
    }
    matches_out = best_m;
    return best_o;
}

// Slow-queue reads: all-column bit-vector search per adapter (<= 32 nt), else the
// literal banded DP (match_back -> locate_back).
template <int W>
__host__ __device__ __forceinline__ int k1_trim_slow(const PSeq<W> &p, int n, const TrimParams &tp, const signed char *thr, const u32 *peq8,
                                            bool literal)
{
    int best_rstart = -1, best_matches = -1;
    for (int a = 0; a < tp.n_adapters; a++) {
        int rs = -1, mt = 0;
        bool found;
        if (tp.alen[a] <= 32 && !literal) {
            rs = k1_match_full<W>(p, n, tp, thr, a, peq8 + 8 * a, mt);
            found = rs >= 0;
        } else {
            found = match_back<W>(p, n, tp, a, rs, mt);
        }
        if (found && (best_rstart < 0 || mt > best_matches)) { best_rstart = rs; best_matches = mt; }
    }
    return best_rstart >= 0 ? best_rstart : n;
}

// Length filter + collapse for one converged warp: lanes with `on` hold a trimmed read of
// n bases.  wh = this warp's private read-length histogram (shared memory).
template <int W>
__device__ __forceinline__ u32 k1_finish_warp(PSeq<W> &p, int n, bool on, int min_length, const TableView<W> &T, int sample, u32 *wh)
{
    const bool keep = on && n >= min_length;
    if (keep) pseq_truncate<W>(p, n);
    table_add_warp<W>(T, p, n, keep, sample, keep ? pseq_hash<W>(p, n) : 0ull);
    if (keep) atomicAdd(&wh[n], 1u);
    return keep ? 1u : 0u;
}

// ---------------------------------------------------------------------------
// K1 is a short pipeline of three kernels, each one homogeneous (small code, full
// warps, its own occupancy):
//   k1_trim_collapse  every read: prefilter; class 0 -> length filter + upsert; class 1 -> the
//                     pigeonhole filter (k1_piece_hit) settles it on the spot or makes it class 2;
//                     class 2 -> appended to queue 2
//   k1_general_search queue 2: windowed bit-vector search with candidate walks;
//                     unsettled -> queue 3
//   k1_slow_search    queue 3: all-column search (or the literal DP)
// Queue entries are the quality-trimmed reads themselves (one 32-byte record at W = 2), written
// and read coalesced: the queue kernels never go back to the K0 record arrays.  Appends are
// warp-aggregated (chunk reservation, one atomic per 32 entries).
// ---------------------------------------------------------------------------
template <int W> struct K1QRec;
template <> struct alignas(32) K1QRec<2> { u64 s[2]; u64 nm[1]; u32 n; u32 pad; };
template <> struct alignas(64) K1QRec<4> { u64 s[4]; u64 nm[2]; u32 n; u32 pad; u64 pad2; };
#define K1Q_NONE 0xFFFFFFFFu      // n of a padding entry
#define K1Q_CHUNK 32

struct K1Queues {
    u64 *n;           // [2] fills of q2, q3
    void *q2;         // K1QRec<W>[]
    void *q3;
};

// warp-aggregated append of whole records (common.cuh:warpq_slot hands out the slots)
template <int W>
__device__ __forceinline__ void k1q_push(bool push, u64 *fill, K1QRec<W> *q, WarpQ &wq, const PSeq<W> &p, int n)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned b = __ballot_sync(0xffffffffu, push);
    const u32 c = __popc(b);
    if (!c) return;
    const u32 rank = __popc(b & ((1u << lane) - 1u));
    u64 slot;
    if (c <= wq.left) {
        slot = wq.pos + rank;
        wq.pos += c; wq.left -= c;
    } else {
        u64 nb = 0;
        if (lane == 0) nb = atomicAdd(fill, (u64)K1Q_CHUNK);
        nb = __shfl_sync(0xffffffffu, nb, 0);
        const u32 spill = c - wq.left;
        slot = rank < wq.left ? wq.pos + rank : nb + (rank - wq.left);
        wq.pos = nb + spill; wq.left = K1Q_CHUNK - spill;
    }
    if (push) {
        ulonglong2 *d = reinterpret_cast<ulonglong2 *>(q + slot);
        if (W == 2) {
            d[0] = make_ulonglong2(p.s[0], p.s[1]);
            d[1] = make_ulonglong2(p.nm[0], (u64)(u32)n);
        } else {
            d[0] = make_ulonglong2(p.s[0], p.s[1]);
            d[1] = make_ulonglong2(p.s[W - 2], p.s[W - 1]);
            d[2] = make_ulonglong2(p.nm[0], p.nm[W / 2 - 1]);
            d[3] = make_ulonglong2((u64)(u32)n, 0ull);
        }
    }
}
template <int W>
__device__ __forceinline__ void k1q_close(K1QRec<W> *q, const WarpQ &wq)
{
    for (u32 x = threadIdx.x & 31; x < wq.left; x += 32) q[wq.pos + x].n = K1Q_NONE;
}
template <int W>
__device__ __forceinline__ bool k1q_load(const K1QRec<W> *q, u64 i, u64 fill, PSeq<W> &p, int &n)
{
#pragma unroll
    for (int w = 0; w < W; w++) p.s[w] = 0;
#pragma unroll
    for (int w = 0; w < W / 2; w++) p.nm[w] = 0;
    n = 0;
    if (i >= fill) return false;
    const ulonglong2 *d = reinterpret_cast<const ulonglong2 *>(q + i);
    if (W == 2) {
        const ulonglong2 a = d[0], b = d[1];
        if ((u32)b.y == K1Q_NONE) return false;
        p.s[0] = a.x; p.s[1] = a.y; p.nm[0] = b.x; n = (int)(u32)b.y;
    } else {
        const ulonglong2 a = d[0], b = d[1], c = d[2], e = d[3];
        if ((u32)e.x == K1Q_NONE) return false;
        p.s[0] = a.x; p.s[1] = a.y; p.s[W - 2] = b.x; p.s[W - 1] = b.y; p.nm[0] = c.x; p.nm[W / 2 - 1] = c.y; n = (int)(u32)e.x;
    }
    return true;
}

struct K1Shared {
    TrimParams tp;                                  // dynamic indexing by adapter: keep it out of local memory
    signed char thr[MIRGE_MAX_ADAPTER_LEN + 1];
    u32 peq8[MIRGE_MAX_ADAPTERS * 8];
};

// sh_hist: the CTA's read-length histogram, flushed once per CTA (global atomics on the
// few histogram sectors are same-address traffic: keep them rare).
#define K1_HIST(W) (32 * (W) + 1)

template <int W>
__device__ __forceinline__ void k1_setup(K1Shared &sh, u32 *sh_hist, const TrimParams &tp_param, const K1Thr &thr_param)
{
    for (int b = threadIdx.x; b < K1_HIST(W); b += blockDim.x) sh_hist[b] = 0;
    for (int b = threadIdx.x; b <= MIRGE_MAX_ADAPTER_LEN; b += blockDim.x) sh.thr[b] = thr_param.v[b];
    for (int b = threadIdx.x; b < (int)(sizeof(TrimParams) / 4); b += blockDim.x)
        reinterpret_cast<u32 *>(&sh.tp)[b] = reinterpret_cast<const u32 *>(&tp_param)[b];
    if (threadIdx.x < MIRGE_MAX_ADAPTERS * 8)
        sh.peq8[threadIdx.x] = (threadIdx.x & 4) ? 0u : (u32)tp_param.peq[threadIdx.x >> 3][threadIdx.x & 3];
    __syncthreads();
}

template <int W>
__device__ __forceinline__ void k1_teardown(u32 my_kept, u32 my_diag, u64 *diag, const u32 *sh_hist, const K1Counters &ctr, int S, int sample)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_kept += __shfl_down_sync(0xffffffffu, my_kept, d);
        my_diag += __shfl_down_sync(0xffffffffu, my_diag, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (my_kept) atomicAdd(&ctr.kept[sample], (u64)my_kept);
        if (my_diag && diag) atomicAdd(diag, (u64)my_diag);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < K1_HIST(W); b += blockDim.x)
        if (sh_hist[b]) atomicAdd(&ctr.len_hist[(u64)b * S + sample], (u64)sh_hist[b]);
}

// force_slow bits (parity tests): 1 = every read through the all-column search,
// 2 = ... through the literal DP instead, 4 = prefilter off.
#define K1A_THREADS 256
template <int W>
__global__ void __launch_bounds__(K1A_THREADS)
k1_trim_collapse(const u64 *__restrict__ bases, const u64 *__restrict__ nmask, const u32 *__restrict__ meta,
                 const u64 *__restrict__ n_records_ptr, int sample, TrimParams tp_param, K1Pre pre_param, K1Thr thr_param,
                 TableView<W> T, K1Counters ctr, K1Queues Q, int force_slow)
{
    __shared__ u32 sh_hist[K1_HIST(W)];
    __shared__ u32 sh_cid[K1_CC * (K1A_THREADS / 32)], sh_ccnt[K1_CC * (K1A_THREADS / 32)];
    __shared__ K1Pre pre;
    __shared__ u64 sh_bm[4 * (W / 2) * K1A_THREADS];     // base-position masks of the pigeonhole filter
    __shared__ u8 sh_c1lane[K1A_THREADS], sh_c1lim[K1A_THREADS];
    for (int b = threadIdx.x; b < K1_HIST(W); b += K1A_THREADS) sh_hist[b] = 0;
    for (int b = threadIdx.x; b < K1_CC * (K1A_THREADS / 32); b += K1A_THREADS) { sh_cid[b] = 0xFFFFFFFFu; sh_ccnt[b] = 0; }
    u32 *wh = sh_hist;
    u32 *cid = sh_cid + K1_CC * (threadIdx.x / 32), *ccnt = sh_ccnt + K1_CC * (threadIdx.x / 32);
    for (int b = threadIdx.x; b < (int)(sizeof(K1Pre) / 4); b += K1A_THREADS)
        reinterpret_cast<u32 *>(&pre)[b] = reinterpret_cast<const u32 *>(&pre_param)[b];
    __syncthreads();
    const u64 n_records = *n_records_ptr;                  // written by K0 (K0State::nrec)
    const bool use_pre = pre.enabled && !(force_slow & 4);
    const int mode = tp_param.mode, uncond = tp_param.uncond, min_length = tp_param.min_length;
    const u64 aw0 = tp_param.aw[0][0];
    K1QRec<W> *q2 = reinterpret_cast<K1QRec<W> *>(Q.q2), *q3 = reinterpret_cast<K1QRec<W> *>(Q.q3);
    u32 my_kept = 0, my_c1 = 0;
    WarpQ wq2 = {0, 0}, wq3 = {0, 0};
    for (u64 r0 = (u64)blockIdx.x * K1A_THREADS; r0 < n_records; r0 += (u64)gridDim.x * K1A_THREADS) {
        const u64 r = r0 + threadIdx.x;
        const bool live = r < n_records;
        PSeq<W> p;
        int n = 0;
        if (live) k1_load_record<W>(bases, nmask, meta, r, p, n);
        else {
#pragma unroll
            for (int w = 0; w < W; w++) p.s[w] = 0;
#pragma unroll
            for (int w = 0; w < W / 2; w++) p.nm[w] = 0;
        }
        int nf = n, cls = 0, lim = 0;
        if (mode == 2) {
            if (uncond > 0) { const int cut = uncond < n ? uncond : n; pseq_shift<W>(p, cut); nf = n - cut; }
            else if (uncond < 0) { const int cut = (-uncond) < n ? (-uncond) : n; nf = n - cut; }
        } else if (mode == 1) {
            if (force_slow & 3) cls = 3;
            else if (use_pre) cls = k1_prefilter<W>(p, n, pre, aw0, nf, lim);
            else cls = 2;
        }
        if (!live) cls = -1;
        const unsigned c1m = __ballot_sync(0xffffffffu, cls == 1);
        if (c1m) {                                                         // the warp's class-1 reads, piece by piece (warp-collective)
            const bool mine = cls == 1;
            const bool hit = k1_piece_hit_warp<W>(p, mine, c1m, lim, pre, aw0, sh_bm, K1A_THREADS, sh_c1lane, sh_c1lim);
            if (mine) { my_c1++; cls = hit ? 2 : 0; }
        }
        k1q_push<W>(cls == 2, Q.n + 0, q2, wq2, p, n);
        k1q_push<W>(cls == 3, Q.n + 1, q3, wq3, p, n);
        const bool keep = cls == 0 && nf >= min_length;
        if (force_slow & 0x200) { my_kept += keep; continue; }             // diagnostic: classification only
        if (keep) pseq_truncate<W>(p, nf);
        bool lead; u32 id, cnt;
        table_find_warp<W>(T, p, nf, keep, keep ? pseq_hash<W>(p, nf) : 0ull, lead, id, cnt);
        if (force_slow & 0x100) { }                                        // diagnostic: no count updates
        else if (force_slow & 0x800) { if (lead) atomicAdd(&T.counts[(u64)sample * T.max_unique + id], cnt); }   // count cache off
        else count_cached<W>(cid, ccnt, T, sample, lead, id, cnt);
        if (keep) { atomicAdd(&wh[nf], 1u); my_kept++; }
    }
    count_cache_flush<W>(cid, ccnt, T, sample);
    k1q_close<W>(q2, wq2); k1q_close<W>(q3, wq3);
    k1_teardown<W>(my_kept, my_c1, ctr.slow_reads ? ctr.slow_reads + 1 : nullptr, sh_hist, ctr, T.S, sample);
}

template <int W>
__global__ void __launch_bounds__(K1_THREADS)
k1_general_search(int sample, TrimParams tp_param, K1Thr thr_param, TableView<W> T, K1Counters ctr, K1Queues Q)
{
    __shared__ u64 ring[K1_RING * K1_THREADS];
    __shared__ u32 ringh[K1_RING * K1_THREADS];
    __shared__ u32 sh_hist[K1_HIST(W)];
    __shared__ K1Shared sh;
    k1_setup<W>(sh, sh_hist, tp_param, thr_param);
    u32 *wh = sh_hist;
    const u64 n2 = Q.n[0];
    const K1QRec<W> *q2 = reinterpret_cast<const K1QRec<W> *>(Q.q2);
    K1QRec<W> *q3 = reinterpret_cast<K1QRec<W> *>(Q.q3);
    u32 my_kept = 0, my_n = 0;
    WarpQ wq3 = {0, 0};
    for (u64 i0 = (u64)blockIdx.x * K1_THREADS; i0 < n2; i0 += (u64)gridDim.x * K1_THREADS) {
        PSeq<W> ps;
        int ns = 0;
        const bool on = k1q_load<W>(q2, i0 + threadIdx.x, n2, ps, ns);
        my_n += on;
        const int nfs = k1_trim_fast<W>(ps, ns, sh.tp, sh.thr, ring + threadIdx.x, ringh + threadIdx.x, sh.peq8);   // warp-collective
        k1q_push<W>(on && nfs == K1_SLOW, Q.n + 1, q3, wq3, ps, ns);
        my_kept += k1_finish_warp<W>(ps, nfs, on && nfs != K1_SLOW, sh.tp.min_length, T, sample, wh);
    }
    k1q_close<W>(q3, wq3);
    k1_teardown<W>(my_kept, my_n, ctr.slow_reads ? ctr.slow_reads + 2 : nullptr, sh_hist, ctr, T.S, sample);
}

template <int W>
__global__ void __launch_bounds__(K1_THREADS)
k1_slow_search(int sample, TrimParams tp_param, K1Thr thr_param, TableView<W> T, K1Counters ctr, K1Queues Q, int literal)
{
    __shared__ u32 sh_hist[K1_HIST(W)];
    __shared__ K1Shared sh;
    k1_setup<W>(sh, sh_hist, tp_param, thr_param);
    u32 *wh = sh_hist;
    const u64 n3 = Q.n[1];
    const K1QRec<W> *q3 = reinterpret_cast<const K1QRec<W> *>(Q.q3);
    u32 my_kept = 0, my_n = 0;
    for (u64 i0 = (u64)blockIdx.x * K1_THREADS; i0 < n3; i0 += (u64)gridDim.x * K1_THREADS) {
        PSeq<W> ps;
        int ns = 0, nfs = 0;
        const bool on = k1q_load<W>(q3, i0 + threadIdx.x, n3, ps, ns);
        if (on) {
            nfs = k1_trim_slow<W>(ps, ns, sh.tp, sh.thr, sh.peq8, literal != 0);
            my_n++;
        }
        __syncwarp();
        my_kept += k1_finish_warp<W>(ps, nfs, on, sh.tp.min_length, T, sample, wh);
    }
    k1_teardown<W>(my_kept, my_n, ctr.slow_reads, sh_hist, ctr, T.S, sample);
}
