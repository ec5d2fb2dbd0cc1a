// k2_map.cuh -- K2: the mapping rounds (seed-and-verify over 2-bit words).
//
// Replaces the `bowtie` invocations of utils/runAnnotationPipeline.py:590-599
// (:576-586 with -spikeIn), executed at :643 and :688-689, together with
// writeSeqToAnnot (:543-554), parseAlignment (:20-28) and updateAnnotDic(2)
// (:341-352).  Acceptance rules (bowtie 1.1.2 manual):
//   -n N : <= N mismatches in the first min(28, len) bases and, FASTA input
//          giving every base Maq quality 30 against -e 70, <= 2 overall;
//   -v V : <= V mismatches end to end.
// Forward strand only (--norc), no indels.  N in the read always mismatches;
// a placement touching an ambiguous reference base or crossing an entry end
// is invalid.  Tie policy: fewest mismatches, then lowest (ref, pos) -- highest
// for the -a --best --strata rounds whose LAST printed line the reference keeps.
#pragma once
#include "common.cuh"

#define K2_MAX_TABS 6
#ifndef K2_CTAS
#define K2_CTAS 8       // resident CTAs per SM of the round kernels (64 registers); the grid is one wave of them
#endif
#define K2_CTAS_W4 5    // four-word sequences (reads of 65..128 nt): 96 registers without spills instead of 64 with 240 bytes spilled

struct LibView {
    const u64 *seq;        // 2-bit packed concatenation (zero padded by >= 5 words)
    const u64 *amb;        // 1 bit per base, set for non-ACGT (zero padded)
    const u32 *ref_off;    // [n_refs + 1] base offsets
    const u32 *blk_ref;    // [total / 64 + 1] entry holding the first base of each 64-base block
    int n_refs;
    u32 total;
    u32 max_ref_len;       // longest entry: a longer query has no placement
    int has_amb;           // some base of the library is not A, C, G, T (else the ambiguity windows need not be read)
};

struct KmerTab {
    int k;
    const u32 *bucket_off;  // [4^k + 1]
    const u32 *pos;         // global base offsets grouped by k-mer
};

// Pair-of-pieces seed tables for short queries.  A round that tolerates e mismatches in its seed region cuts
// the region's first m*pl bases into m = e + 2 pieces of pl bases: e mismatches leave at least TWO pieces
// intact, so every valid placement is found under the key (piece i, piece j) of one of the m(m-1)/2 pairs.
// A pair key holds 2*pl bases where the classic pigeonhole seed of the same region holds region/(e+1):
// for a 16-base region and e = 2 that is 8 bases against 5, 64 times fewer candidates to verify.
// One table per pair, indexed by placement start (the pieces sit at fixed offsets i*pl and j*pl).
#define K2_MAX_GAPSETS 2
struct GapSet {
    int m, pl;           // pieces, piece length
    int rg_lo, rg_hi;    // region lengths this set serves (rg_lo >= m*pl)
    u32 nb1;             // 4^(2*pl) + 1: stride of the per-pair bucket arrays
    u32 npos;            // placements per pair table: total - m*pl + 1
    const u32 *off;      // [m(m-1)/2][nb1] offsets into the pair's own block of pos
    const u32 *pos;      // [m(m-1)/2][npos] placement starts grouped by key
};
// piece indices of pair `sd` (which = 0: first piece, 1: second), m = 3 or 4, pairs in lexicographic order
__host__ __device__ __forceinline__ int gap_piece(int m, int sd, int which)
{
    const u32 t = (m == 3) ? (which ? 0x221u : 0x100u) : (which ? 0x332321u : 0x211000u);
    return (int)((t >> (4 * sd)) & 15u);
}

struct RoundParams {
    int round;          // 0-based round index; annot slot = round + 1
    int len_filter;     // <0: len < -f ; >0: len > f ; 0: all   (runAnnotationPipeline.py:574/588)
    int mode;           // 0 = -n, 1 = -v
    int param;          // N or V
    int trim5, trim3;   // -5 / -3
    int highest;        // keep the highest (ref, pos) of the best stratum
    int polyT;          // pre-tRNA round: strip trailing T{3,}, remainder >= 11 (:664-686)
    int ntabs;
    int kcons;          // tabs[] hold consecutive seed lengths (the table for a length is found by subtraction)
    int next_tail3T;    // the NEXT round is the poly-T round: misses that do not end in TTT go to bypass_out instead
    u32 *bypass_out;    // ... the list the round after the next reads (the next round appends its own misses there)
    u64 *bypass_n;
    KmerTab tabs[K2_MAX_TABS];   // ascending k
    int ngap, gap_pairs;         // pair-of-pieces sets for this round's mismatch budget, pairs per set
    GapSet gap[K2_MAX_GAPSETS];
};

// ---------------------------------------------------------------------------
// library build kernels
// ---------------------------------------------------------------------------
__global__ void k2_pack_library(const unsigned char *__restrict__ ascii, u32 total, u64 *__restrict__ seq,
                                u64 *__restrict__ amb)
{
    // one thread per 64 bases: two base words + one ambiguity word
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 b0 = t * 64;
    if (b0 >= total) return;
    u64 w0 = 0, w1 = 0, am = 0;
    for (int i = 0; i < 64 && b0 + i < total; i++) {
        int c = base_code(ascii[b0 + i]);
        if (c > 3) { am |= 1ull << i; c = 0; }
        if (i < 32) w0 |= (u64)c << (2 * i); else w1 |= (u64)c << (2 * (i - 32));
    }
    seq[2 * t] = w0; seq[2 * t + 1] = w1; amb[t] = am;
}

__device__ __forceinline__ u64 lib_window64(const u64 *__restrict__ seq, u32 g)
{
    const u32 wi = g >> 5, sh = (g & 31) * 2;
    const u64 lo = __ldg(seq + wi);
    if (sh == 0) return lo;
    return (lo >> sh) | (__ldg(seq + wi + 1) << (64 - sh));
}

// pass 0: count k-mers per bucket; pass 1: scatter positions (cursor = running offsets)
__global__ void k2_kmer_pass(LibView L, int k, int pass, u32 *__restrict__ bucket, u32 *__restrict__ pos)
{
    const u32 mask = (k == 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g + (u64)k <= (u64)L.total; g += (u64)gridDim.x * blockDim.x) {
        const u32 code = (u32)lib_window64(L.seq, (u32)g) & mask;
        if (pass == 0) atomicAdd(&bucket[code], 1u);
        else pos[atomicAdd(&bucket[code], 1u)] = (u32)g;
    }
}

// the same two passes for one pair table of a GapSet: key = piece at g + oi | piece at g + oj
__global__ void k2_gap_pass(LibView L, int pl, int oi, int oj, u32 npos, int pass, u32 *__restrict__ bucket, u32 *__restrict__ pos)
{
    const u32 pm = (1u << (2 * pl)) - 1u;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < (u64)npos; g += (u64)gridDim.x * blockDim.x) {
        const u32 code = ((u32)lib_window64(L.seq, (u32)g + (u32)oi) & pm) | (((u32)lib_window64(L.seq, (u32)g + (u32)oj) & pm) << (2 * pl));
        if (pass == 0) atomicAdd(&bucket[code], 1u);
        else pos[atomicAdd(&bucket[code], 1u)] = (u32)g;
    }
}

// ---------------------------------------------------------------------------
// per-read search
// ---------------------------------------------------------------------------
// entry that holds global base offset g: block table + short forward scan
__device__ __forceinline__ int ref_of_gpos(const LibView &L, u32 g)
{
    int r = (int)__ldg(L.blk_ref + (g >> 6));
    while (__ldg(L.ref_off + r + 1) <= g) r++;
    return r;
}

template <int W>
struct Query {
    u64 s[W];        // read bases
    u64 n2[W];       // read N positions spread onto the even bits
    u64 len2[W];     // even-bit mask of bases [0, n)
    int n;
};

// Mismatch test of the query placed at global offset g.  Returns true when the
// round's rule accepts the placement (entry bounds are checked by the caller).
template <int W>
__device__ __forceinline__ bool verify_at(const LibView &L, const Query<W> &q, u32 g, int mode, int param, int &mm_out)
{
    if ((u64)g + (u64)q.n > (u64)L.total) return false;
    // ambiguity window
    if (L.has_amb) {
        const u32 wi = g >> 6, sh = g & 63;
        int rem = q.n;
#pragma unroll
        for (int w = 0; w < W / 2; w++) {
            if (rem > 0) {
                const u64 lo = __ldg(L.amb + wi + w);
                u64 a = sh ? ((lo >> sh) | (__ldg(L.amb + wi + w + 1) << (64 - sh))) : lo;
                if (rem < 64) a &= (1ull << rem) - 1ull;
                if (a) return false;
            }
            rem -= 64;
        }
    }
    const u32 wi = g >> 5, sh = (g & 31) * 2;
    int mm_total = 0, mm_seed = 0;
    u64 lo = __ldg(L.seq + wi);
#pragma unroll
    for (int w = 0; w < W; w++) {
        if (w * 32 < q.n) {
            const u64 hi = __ldg(L.seq + wi + w + 1);
            const u64 r = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
            lo = hi;
            const u64 x = r ^ q.s[w];
            const u64 d = (((x | (x >> 1)) & 0x5555555555555555ull) | q.n2[w]) & q.len2[w];
            mm_total += __popcll(d);
            if (w == 0) mm_seed = __popcll(d & 0x00FFFFFFFFFFFFFFull);   // first 28 bases
        }
    }
    if (mode == 0) { if (mm_seed > param || mm_total > 2) return false; }
    else { if (mm_total > param) return false; }
    mm_out = mm_total;
    return true;
}

template <int W>
__device__ __forceinline__ void consider(const LibView &L, const Query<W> &q, u32 g, const RoundParams &R, u64 &best)
{
    int mm;
    if (!verify_at<W>(L, q, g, R.mode, R.param, mm)) return;
    const u64 key = ((u64)mm << 32) | (u64)(R.highest ? ~g : g);
    if (key >= best) return;
    const int r = ref_of_gpos(L, g);
    if ((u64)g + (u64)q.n > (u64)__ldg(L.ref_off + r + 1)) return;   // crosses the entry end
    best = key;
}

struct K2Arrays {
    const u64 *useq;      // [U][W]
    const u64 *unmask;    // [U][W/2]
    const u8 *ulen;       // [U]
    signed char *round_of;  // [U] 0 = un-annotated, else round + 1
    int *ref_of;
    int *pos_of;
    signed char *mm_of;
};

// per-warp output buffer of the round kernels (ids) and its flush: one reservation on the list counter, coalesced copy
#define K2_OB 128
__device__ __forceinline__ void k2_flush(const u32 *buf, u32 n, u32 *__restrict__ out, u64 *__restrict__ counter)
{
    const unsigned lane = threadIdx.x & 31;
    __syncwarp();
    u64 b = 0;
    if (lane == 0) b = atomicAdd(counter, (u64)n);
    b = __shfl_sync(0xffffffffu, b, 0);
    for (u32 x = lane; x < n; x += 32) out[b + x] = buf[x];
    __syncwarp();
}

// One thread per still un-annotated unique sequence; a warp works through its 32
// sequences seed by seed.  When every lane has at most one candidate placement it
// verifies its own; otherwise the candidate lists of the 32 lanes are laid end to
// end and verified 32 at a time (k2_round.inc), so one sequence with a long list
// (few-base seeds of the isomiR round, big libraries) does not serialise its warp.  todo_in == nullptr means
// the identity list 0 .. *n_in-1.  Sequences that stay un-annotated (filtered
// out by length, or without a valid placement) are appended to todo_out.
// BAL (the multi-seed rounds on the miRNA library, whose seed tables have consecutive sizes): the seeds share the
// region's spare bases, so the first ones are one base longer.  Those rounds work out of L1 and are issue
// bound, so they are held to 8 CTAs per SM; the plain rounds wait on random library reads and keep the
// compiler's natural register budget (a forced cap serialises the loads of a verification: +40 % time).
// GAP: rounds whose library carries pair-of-pieces tables for their mismatch budget (GapSet above); queries whose
// region a set serves take the m(m-1)/2 pair look-ups instead of the e + 1 classic seeds.
#define K2_ROUND_NAME k2_round
#define K2_LAUNCH_BOUNDS __launch_bounds__(128, (W == 2 ? K2_CTAS : K2_CTAS_W4))
#define K2_BAL 0
#define K2_GAP 0
#include "k2_round.inc"
#undef K2_ROUND_NAME
#undef K2_LAUNCH_BOUNDS
#undef K2_BAL
#undef K2_GAP
#define K2_ROUND_NAME k2_round_balanced
#define K2_LAUNCH_BOUNDS __launch_bounds__(128, (W == 2 ? K2_CTAS : K2_CTAS_W4))
#define K2_BAL 1
#define K2_GAP 0
#include "k2_round.inc"
#undef K2_ROUND_NAME
#undef K2_LAUNCH_BOUNDS
#undef K2_BAL
#undef K2_GAP
#define K2_ROUND_NAME k2_round_pairs8
#define K2_LAUNCH_BOUNDS __launch_bounds__(128, (W == 2 ? K2_CTAS : K2_CTAS_W4))
#define K2_BAL 1
#define K2_GAP 1
#include "k2_round.inc"
#undef K2_ROUND_NAME
#undef K2_LAUNCH_BOUNDS
#undef K2_BAL
#undef K2_GAP
