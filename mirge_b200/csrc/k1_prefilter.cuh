// k1_prefilter.cuh -- settles most reads without any dynamic programming.
//
// cutadapt's Adapter.match_to (adapters.py, where = BACK) returns the leftmost exact
// occurrence of the whole adapter when there is one (str.find), and otherwise ranks the
// candidate end cells of Aligner.locate (_align.pyx) by (matches, cost, first visited).
// For the reads sequencers actually produce the answer is decided by EXACT comparisons:
//
//   FULL  the whole adapter occurs exactly            -> rstart = leftmost occurrence
//   E     no FULL, and the read ends with an exact adapter prefix A[0..l), l < m
//         (take the longest).  E is itself a candidate: cell (l, n), cost 0,
//         matches l, origin n - l.  Any candidate X that outranks it needs
//         matches(X) >= l + 1 (same matches means cost > 0, or X = E).
//
// Whether such an X can exist is bounded on the host once per adapter
// (capi.cu:k1pre_build) with a dynamic program over the "wildcard read" W^I + A[0..l)
// -- I = n - l insert bases that match anything, then the exact prefix -- which
// maximises the matches of every alignment path whose cost stays within the
// candidate's threshold.  Any path in a real read maps to a path of the wildcard
// read with cost <= and matches >=, so when the wildcard read admits no X, no real
// read with that (I, l) does:
//
//   safe_len[I] bit l   no X at all                      -> class 0: rstart = n - l
//   safe_p bit l        no X among the adapter prefixes at the read end, whatever
//                       the insert; full-length X must end at a column
//                       <= I + fullreach[l]             -> class 1: no full-length
//                       candidate in those columns -> rstart = n - l.  The kernel decides that
//                       with the pigeonhole piece filter (k1_piece_hit below); the score-only
//                       Myers scan (k1_full_candidate_scan) is the reference the host tests hold
//                       the filter against
//   otherwise                                            -> class 2: general search
//
// The search for FULL / E uses the adapter's first four bases as an anchor: all anchor
// occurrences are found with byte-wise SIMD-in-register compares of the 2-bit words at the
// four base phases, then each occurrence is extended with one XOR of a 64-bit window.
#pragma once
#include "common.cuh"

#define K1PRE_MAXI 48

struct K1Pre {
    int enabled;                      // single adapter, 8 <= m <= 32, tables built
    int m;
    int icap;                         // safe_len is defined for I <= icap (larger inserts: never safe by length)
    u32 anchor_rep;                   // A[0..4) packed in one byte, replicated
    u32 safe_p;
    u32 safe_len[K1PRE_MAXI + 1];
    u8 fullreach[32];
    int np, np_inv;                   // class-1 piece filter: thr[m] + 1 pieces; ceil(65536 / np) (task index -> read, piece)
    u8 pb[36];                        // piece q covers adapter bases pb[q] .. pb[q + 1] - 1
};

// Returns the class (0, 1, 2).  Class 0/1: rstart = start of FULL / E.  Class 1: limit =
// last column whose full-length cost must be checked.
template <int W>
__host__ __device__ __forceinline__ int k1_prefilter(const PSeq<W> &p, int n, const K1Pre &pre, u64 aw, int &rstart, int &limit)
{
    constexpr int NW = 2 * W;
    u32 w[NW + 1];
#pragma unroll
    for (int k = 0; k < W; k++) { w[2 * k] = (u32)p.s[k]; w[2 * k + 1] = (u32)(p.s[k] >> 32); }
    w[NW] = 0;
    u32 c[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) c[k] = 0;
    const u32 rep = pre.anchor_rep;
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int k = 0; k < NW; k++) {
            const u32 y = r ? bit_funnel_r(w[k], w[k + 1], 2 * r) : w[k];
            const u32 x = y ^ rep;
            const u32 z = (x - 0x01010101u) & ~x & 0x80808080u;      // zero bytes (may over-report; every hit is verified)
            c[k] |= z >> (7 - 2 * r);                                 // occurrence at base 16k + 4b + r -> bit 8b + 2r
        }
    }
    // occurrences must leave room for the four anchor bases: positions <= n - 4
    const int V = n - 3;
    u64 cc[W];
#pragma unroll
    for (int k = 0; k < W; k++) {
        u64 v = (u64)c[2 * k] | ((u64)c[2 * k + 1] << 32);
        const int lo = 32 * k;
        if (V <= lo) v = 0;
        else if (V < lo + 32) v &= (1ull << (2 * (V - lo))) - 1ull;
        cc[k] = v;
    }
    const int m = pre.m;
    for (;;) {
        int kk = -1;
        u64 cur = 0;
#pragma unroll
        for (int k = W - 1; k >= 0; k--) if (cc[k]) { kk = k; cur = cc[k]; }
        if (kk < 0) break;
        const int b = bit_ffs64(cur) - 1;
#pragma unroll
        for (int k = 0; k < W; k++) if (k == kk) cc[k] &= cc[k] - 1;
        const int pos = kk * 32 + (b >> 1);
        const u64 x = words_window64<W>(p.s, pos) ^ aw;
        u64 mm = (x | (x >> 1)) & 0x5555555555555555ull;
        const int vl = (n - pos) < m ? (n - pos) : m;
        if (vl < 32) mm |= 1ull << (2 * vl);
        int ml = mm ? ((bit_ffs64(mm) - 1) >> 1) : 32;
        u32 nb = mask_window32<W / 2>(p.nm, pos);
        if (ml < 32) nb &= (1u << ml) - 1u;
        if (nb) ml = bit_ffs32(nb) - 1;
        if (ml >= m) { rstart = pos; return 0; }                      // FULL: leftmost, occurrences are visited in order
        if (pos + ml >= n) {                                          // E with l = ml, I = pos
            rstart = pos;
            if (pos <= pre.icap && ((pre.safe_len[pos] >> ml) & 1u)) return 0;
            if ((pre.safe_p >> ml) & 1u) {
                const int lim = pos + pre.fullreach[ml];
                limit = lim < n ? lim : n;
                return 1;
            }
            return 2;
        }
    }
    return 2;
}

// Class 1: does any column j <= limit hold a full-length candidate (cost(m, j) <= thr_m)?
// Myers' bit-vector column recurrence, score only; warp lock-step (uniform trip count).
template <int W>
__host__ __device__ __forceinline__ bool k1_full_candidate_scan(const PSeq<W> &p, int limit, int m, int thr_m, const u32 *peq8)
{
    u32 Pv = (m >= 32) ? 0xFFFFFFFFu : ((1u << m) - 1u), Mv = 0;
    int score = m;
    const int top = m - 1;
    bool hit = false;
    const int lmax = warp_max_all(limit);
    for (int j0 = 0; j0 < lmax; j0 += 16) {
        u32 bw = 0, nb = 0;
#pragma unroll
        for (int w = 0; w < W; w++) {
            if ((j0 >> 5) == w) {
                bw = (u32)(p.s[w] >> ((j0 & 16) * 2));
                nb = (u32)(p.nm[w >> 1] >> (((w & 1) * 32) + (j0 & 16))) & 0xFFFFu;
            }
        }
        const int jend = (lmax - j0) < 16 ? (lmax - j0) : 16;
        u32 le16 = 0;
#pragma unroll 4
        for (int t = 0; t < jend; t++) {
            const u32 Eq = peq8[(bw & 3u) | (((nb >> t) & 1u) << 2)];
            bw >>= 2;
            const u32 Xv = Eq | Mv;
            const u32 Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            u32 Ph = Mv | ~(Xh | Pv);
            u32 Mh = Pv & Xh;
            score += (int)((Ph >> top) & 1u) - (int)((Mh >> top) & 1u);
            Ph <<= 1; Mh <<= 1;
            Pv = Mh | ~(Xv | Ph);
            Mv = Ph & Xv;
            le16 |= (u32)(score <= thr_m) << t;
        }
        // columns past `limit` ran on padding; drop them
        const int keep = limit - j0;
        if (keep < 16) le16 &= keep <= 0 ? 0u : ((1u << keep) - 1u);
        hit = hit || le16 != 0u;
    }
    return hit;
}

// 32 bits -> even bits in the low half, odd bits in the high half (the classic unshuffle)
__host__ __device__ __forceinline__ u32 deinterleave32(u32 x)
{
    u32 t;
    t = (x ^ (x >> 1)) & 0x22222222u; x ^= t ^ (t << 1);
    t = (x ^ (x >> 2)) & 0x0C0C0C0Cu; x ^= t ^ (t << 2);
    t = (x ^ (x >> 4)) & 0x00F000F0u; x ^= t ^ (t << 4);
    t = (x ^ (x >> 8)) & 0x0000FF00u; x ^= t ^ (t << 8);
    return x;
}

// Class 1 without dynamic programming: the pigeonhole filter.  A full-length candidate ending at a column j <= limit
// is an occurrence of the whole adapter with at most kmax = thr[m] edits (substitutions, insertions, deletions) inside
// read[0..limit).  Cut the adapter into kmax + 1 disjoint pieces: kmax edits touch at most kmax of them, so at least
// one piece occurs EXACTLY in read[0..limit).  No piece occurrence there -> no candidate -> E wins.  (The converse does
// not hold: a hit only sends the read to the general search.)  Piece occurrences are found for all positions at once:
// the 2-bit words are split into a low-bit and a high-bit plane (one bit per position), giving per base x a mask of the
// positions holding x; then for a piece P
//     R_len = all positions,   R_u = mask[P[u]] & (R_{u+1} >> 1)        (R_0 = start positions of P)
// N is packed as A and may only add hits.  bm: this thread's scratch, 4 masks of W / 2 words, element
// (x * W / 2 + w) * stride (shared memory on the device: the mask is selected by the adapter base, a run-time index).
template <int W>
__host__ __device__ __forceinline__ bool k1_piece_hit(const PSeq<W> &p, int limit, int m, int kmax, u64 aw, u64 *bm, int stride)
{
    constexpr int NP = W / 2;                                  // 64-position plane words
#pragma unroll
    for (int w = 0; w < NP; w++) {
        const u32 d0 = deinterleave32((u32)p.s[2 * w]), d1 = deinterleave32((u32)(p.s[2 * w] >> 32));
        const u32 d2 = deinterleave32((u32)p.s[2 * w + 1]), d3 = deinterleave32((u32)(p.s[2 * w + 1] >> 32));
        const u64 L = (u64)((d0 & 0xFFFFu) | (d1 << 16)) | ((u64)((d2 & 0xFFFFu) | (d3 << 16)) << 32);
        const u64 H = (u64)((d0 >> 16) | (d1 & 0xFFFF0000u)) | ((u64)((d2 >> 16) | (d3 & 0xFFFF0000u)) << 32);
        bm[(0 * NP + w) * stride] = ~(L | H);
        bm[(1 * NP + w) * stride] = L & ~H;
        bm[(2 * NP + w) * stride] = H & ~L;
        bm[(3 * NP + w) * stride] = L & H;
    }
    const int np = kmax + 1;
    bool hit = false;
    for (int q = 0; q < np; q++) {
        const int b0 = (q * m) / np, b1 = ((q + 1) * m) / np, len = b1 - b0;
        const int cnt = limit - len + 1;                       // start positions 0 .. limit - len
        if (cnt <= 0) continue;
        u64 acc[NP];
        {
            const int c = (int)((aw >> (2 * (b1 - 1))) & 3ull);
#pragma unroll
            for (int w = 0; w < NP; w++) acc[w] = bm[(c * NP + w) * stride];
        }
        for (int t = b1 - 2; t >= b0; t--) {
            const int c = (int)((aw >> (2 * t)) & 3ull);
#pragma unroll
            for (int w = 0; w < NP; w++) {
                const u64 up = (w + 1 < NP) ? acc[w + 1] << 63 : 0ull;
                acc[w] = bm[(c * NP + w) * stride] & ((acc[w] >> 1) | up);
            }
        }
        u64 any = 0;
#pragma unroll
        for (int w = 0; w < NP; w++) {
            const int lo = 64 * w;
            const u64 pm = (cnt <= lo) ? 0ull : (cnt >= lo + 64) ? ~0ull : ((1ull << (cnt - lo)) - 1ull);
            any |= acc[w] & pm;
        }
        hit = hit || any != 0ull;
    }
    return hit;
}

#ifdef __CUDACC__
// The same filter for all class-1 reads of a warp at once.  Only about a quarter of the reads are class 1, so run lane
// by lane the filter kept three lanes in four idle for a third of the kernel's instructions; here every (read, piece)
// pair is one work item, dealt out over the 32 lanes.  Owners publish their position masks (bm, layout
// [(x * NP + w) * stride + thread]) and their column limit in shared memory; a work item reads its owner's.
// c1m = ballot of the class-1 lanes; lane_list / lim_list: per-thread shared scratch.  Warp-collective.
template <int W>
__device__ __forceinline__ bool k1_piece_hit_warp(const PSeq<W> &p, bool mine, unsigned c1m, int limit, const K1Pre &pre, u64 aw,
                                                  u64 *bm, int stride, u8 *lane_list, u8 *lim_list)
{
    constexpr int NP = W / 2;
    const unsigned lane = threadIdx.x & 31, wbase = threadIdx.x & ~31u;
    const int rank = __popc(c1m & ((1u << lane) - 1u));
    if (mine) {
        u64 *o = bm + threadIdx.x;
#pragma unroll
        for (int w = 0; w < NP; w++) {
            const u32 d0 = deinterleave32((u32)p.s[2 * w]), d1 = deinterleave32((u32)(p.s[2 * w] >> 32));
            const u32 d2 = deinterleave32((u32)p.s[2 * w + 1]), d3 = deinterleave32((u32)(p.s[2 * w + 1] >> 32));
            const u64 L = (u64)((d0 & 0xFFFFu) | (d1 << 16)) | ((u64)((d2 & 0xFFFFu) | (d3 << 16)) << 32);
            const u64 H = (u64)((d0 >> 16) | (d1 & 0xFFFF0000u)) | ((u64)((d2 >> 16) | (d3 & 0xFFFF0000u)) << 32);
            o[(0 * NP + w) * stride] = ~(L | H);
            o[(1 * NP + w) * stride] = L & ~H;
            o[(2 * NP + w) * stride] = H & ~L;
            o[(3 * NP + w) * stride] = L & H;
        }
        lane_list[wbase + rank] = (u8)lane;
        lim_list[threadIdx.x] = (u8)limit;                     // limit <= read length <= 128
    }
    __syncwarp();
    const int np = pre.np, ntask = __popc(c1m) * np;
    bool myhit = false;
    for (int t0 = 0; t0 < ntask; t0 += 32) {
        const int t = t0 + (int)lane;
        bool h = false;
        if (t < ntask) {
            const int k = (t * pre.np_inv) >> 16, q = t - k * np;      // exact for t < 32 * 33
            const int ot = (int)wbase + lane_list[wbase + k];
            const int b0 = pre.pb[q], b1 = pre.pb[q + 1], len = b1 - b0;
            const int cnt = (int)lim_list[ot] - len + 1;             // start positions 0 .. limit - len
            if (cnt > 0 && len > 0) {
                const u64 *o = bm + ot;
                u64 acc[NP];
                {
                    const int c = (int)((aw >> (2 * (b1 - 1))) & 3ull);
#pragma unroll
                    for (int w = 0; w < NP; w++) acc[w] = o[(c * NP + w) * stride];
                }
                for (int tt = b1 - 2; tt >= b0; tt--) {
                    const int c = (int)((aw >> (2 * tt)) & 3ull);
#pragma unroll
                    for (int w = 0; w < NP; w++) {
                        const u64 up = (w + 1 < NP) ? acc[w + 1] << 63 : 0ull;
                        acc[w] = o[(c * NP + w) * stride] & ((acc[w] >> 1) | up);
                    }
                }
                u64 any = 0;
#pragma unroll
                for (int w = 0; w < NP; w++) {
                    const int lo = 64 * w;
                    const u64 pm = (cnt <= lo) ? 0ull : (cnt >= lo + 64) ? ~0ull : ((1ull << (cnt - lo)) - 1ull);
                    any |= acc[w] & pm;
                }
                h = any != 0ull;
            }
        }
        const unsigned hb = __ballot_sync(0xffffffffu, h);
        if (mine) {                                            // my items are bits rank * np - t0 .. + np of this pass
            const int lo = rank * np - t0;
            const int a = lo < 0 ? 0 : lo, b = lo + np > 32 ? 32 : lo + np;
            if (a < b) {
                const unsigned rm = ((b - a) >= 32 ? 0xffffffffu : ((1u << (b - a)) - 1u)) << a;
                if (hb & rm) myhit = true;
            }
        }
    }
    __syncwarp();                                              // the scratch is rewritten by the warp's next tile
    return myhit;
}
#endif

// Class "1b" (NOT wired into the kernels yet; exercised on the host by tests/test_k1_host_compiled.py): reads that
// end with an exact adapter prefix of l bases (l >= min_overlap) which the tables cannot settle -- typically a SHORT
// prefix behind a long insert.  E still wins unless a candidate with more than l matches exists:
//   * any full-length candidate (cost(m, j) <= thr[m] at some column j <= n) -> undecided (returns true);
//   * an acceptable prefix candidate (i, n), i > l, cost c: when rows i-c+1 .. i of the last column were all reached
//     by cutadapt's "insertion" step (bit set in the shifted negative horizontal deltas) from mismatching cells,
//     its path runs straight up the column to the exact cell (i-c, n): matches = i - c <= l (l is the LONGEST exact
//     prefix), so it loses to E.  Any other shape -> undecided.
// Returns true when the general search must decide, false when rstart = n - l stands.
template <int W>
__host__ __device__ __forceinline__ bool k1_prefix_candidate_scan(const PSeq<W> &p, int n, int l, int m, const signed char *thr, const u32 *peq8)
{
    u32 Pv = (m >= 32) ? 0xFFFFFFFFu : ((1u << m) - 1u), Mv = 0;
    int score = m;
    const int top = m - 1, thr_m = thr[m];
    bool hit = false;
    u32 PvL = Pv, MvL = 0, MhL = 0, EqL = 0;          // vectors of this lane's last column
    const int lmax = warp_max_all(n);
    for (int j0 = 0; j0 < lmax; j0 += 16) {
        u32 bw = 0, nb = 0;
#pragma unroll
        for (int w = 0; w < W; w++) {
            if ((j0 >> 5) == w) {
                bw = (u32)(p.s[w] >> ((j0 & 16) * 2));
                nb = (u32)(p.nm[w >> 1] >> (((w & 1) * 32) + (j0 & 16))) & 0xFFFFu;
            }
        }
        const int jend = (lmax - j0) < 16 ? (lmax - j0) : 16;
        u32 le16 = 0;
#pragma unroll 4
        for (int t = 0; t < jend; t++) {
            const u32 Eq = peq8[(bw & 3u) | (((nb >> t) & 1u) << 2)];
            bw >>= 2;
            const u32 Xv = Eq | Mv;
            const u32 Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            u32 Ph = Mv | ~(Xh | Pv);
            u32 Mh = Pv & Xh;
            score += (int)((Ph >> top) & 1u) - (int)((Mh >> top) & 1u);
            Ph <<= 1; Mh <<= 1;
            Pv = Mh | ~(Xv | Ph);
            Mv = Ph & Xv;
            le16 |= (u32)(score <= thr_m) << t;
            if (j0 + t + 1 == n) { PvL = Pv; MvL = Mv; MhL = Mh; EqL = Eq; }
        }
        const int keep = n - j0;
        if (keep < 16) le16 &= keep <= 0 ? 0u : ((1u << keep) - 1u);
        hit = hit || le16 != 0u;
    }
    if (hit || n < 1) return true;
    int c = 0;
    for (int i = 1; i < m; i++) {                      // (m, n) is a full-length cell: covered above
        c += (int)((PvL >> (i - 1)) & 1u) - (int)((MvL >> (i - 1)) & 1u);
        if (i <= l || c > thr[i]) continue;
        if (c <= 0 || c >= i) return true;
        const u32 chain = ((1u << c) - 1u) << (i - c);
        if ((MhL & chain) != chain || (EqL & chain) != 0u) return true;
    }
    return false;
}
