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

struct LibView {
    const u64 *seq;        // 2-bit packed concatenation (zero padded by >= 5 words)
    const u64 *amb;        // 1 bit per base, set for non-ACGT (zero padded)
    const u32 *ref_off;    // [n_refs + 1] base offsets
    const u32 *blk_ref;    // [total / 64 + 1] entry holding the first base of each 64-base block
    int n_refs;
    u32 total;
};

struct KmerTab {
    int k;
    const u32 *bucket_off;  // [4^k + 1]
    const u32 *pos;         // global base offsets grouped by k-mer
};

struct RoundParams {
    int round;          // 0-based round index; annot slot = round + 1
    int len_filter;     // <0: len < -f ; >0: len > f ; 0: all   (runAnnotationPipeline.py:574/588)
    int mode;           // 0 = -n, 1 = -v
    int param;          // N or V
    int trim5, trim3;   // -5 / -3
    int highest;        // keep the highest (ref, pos) of the best stratum
    int polyT;          // pre-tRNA round: strip trailing T{3,}, remainder >= 11 (:664-686)
    int ntabs;
    int chunked;        // to-do appends: 1 = chunk reservation per warp (lists hold padding), 0 = one atomic per warp and tile
    KmerTab tabs[K2_MAX_TABS];   // ascending k
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
    {
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

#define K2_HEAVY 16     // seed buckets at least this long are verified by the whole warp

template <int W>
__device__ __forceinline__ Query<W> query_bcast(const Query<W> &q, int src)
{
    Query<W> r;
#pragma unroll
    for (int w = 0; w < W; w++) {
        r.s[w] = __shfl_sync(0xffffffffu, q.s[w], src);
        r.n2[w] = __shfl_sync(0xffffffffu, q.n2[w], src);
        r.len2[w] = __shfl_sync(0xffffffffu, q.len2[w], src);
    }
    r.n = __shfl_sync(0xffffffffu, q.n, src);
    return r;
}

// One thread per still un-annotated unique sequence; a warp works through its 32
// sequences seed by seed.  Short seed buckets are verified by the owning lane;
// long ones (few-base seeds of the isomiR round, big libraries) are broadcast
// and verified by all 32 lanes, then min-reduced back to the owner, so one
// expensive sequence does not serialise its warp.  todo_in == nullptr means
// the identity list 0 .. *n_in-1.  Sequences that stay un-annotated (filtered
// out by length, or without a valid placement) are appended to todo_out.
template <int W>
__global__ void __launch_bounds__(128)
k2_round(K2Arrays A, const u32 *__restrict__ todo_in, const u64 *__restrict__ n_in_ptr,
         u32 *__restrict__ todo_out, u64 *__restrict__ n_out_ptr, LibView L, RoundParams R,
         u64 *__restrict__ stat_processed, u64 *__restrict__ stat_aligned)
{
    const u64 n_in = *n_in_ptr;
    const unsigned lane = threadIdx.x & 31;
    u32 my_proc = 0, my_aln = 0;
    const int nseeds = R.param + 1;
    WarpQ wq = {0, 0};
    for (u64 base = (u64)blockIdx.x * blockDim.x; base < n_in; base += (u64)gridDim.x * blockDim.x) {
        const u64 i = base + threadIdx.x;
        u32 id = WARPQ_NONE;
        if (i < n_in) id = todo_in ? todo_in[i] : (u32)i;
        const bool active = id != WARPQ_NONE;                // lists hold padding (common.cuh:warpq_slot)
        bool pass = false;
        Query<W> q;
        q.n = 0;
#pragma unroll
        for (int w = 0; w < W; w++) { q.s[w] = 0; q.n2[w] = 0; q.len2[w] = 0; }
        if (active) {
            const int len = A.ulen[id];
            pass = L.seq != nullptr;
            if (R.len_filter < 0 && !(len < -R.len_filter)) pass = false;
            if (R.len_filter > 0 && !(len > R.len_filter)) pass = false;
            if (pass) {
                PSeq<W> p;
                int n = 0;
#pragma unroll
                for (int w = 0; w < W; w++) p.s[w] = A.useq[(u64)id * W + w];
#pragma unroll
                for (int w = 0; w < W / 2; w++) p.nm[w] = A.unmask[(u64)id * (W / 2) + w];
                if (R.polyT) {
                    int t = 0;
                    while (t < len && pseq_base<W>(p, len - 1 - t) == 3 && !pseq_isn<W>(p, len - 1 - t)) t++;
                    n = len - t;
                    if (t < 3 || n < 11) pass = false;
                    else pseq_truncate<W>(p, n);
                } else {
                    n = len - R.trim5 - R.trim3;
                    if (n <= 0) pass = false;
                    else { pseq_shift<W>(p, R.trim5); pseq_truncate<W>(p, n); }
                }
                if (pass) {
                    q.n = n;
#pragma unroll
                    for (int w = 0; w < W; w++) {
                        q.s[w] = p.s[w];
                        q.n2[w] = spread32((u32)(p.nm[w >> 1] >> ((w & 1) * 32)));
                        const int lo = w * 32;
                        q.len2[w] = (n <= lo) ? 0ull : (n >= lo + 32) ? 0x5555555555555555ull
                                                                       : (0x5555555555555555ull & ((1ull << ((n - lo) * 2)) - 1ull));
                    }
                }
            }
        }
        u64 best = ~0ull;
        int ti = -1;
        if (pass) {
            my_proc++;
            const int region = (R.mode == 0) ? (q.n < 28 ? q.n : 28) : q.n;
            const int maxk = region / nseeds;
            for (int t = 0; t < R.ntabs; t++) if (R.tabs[t].k <= maxk) ti = t;
            if (ti < 0) {
                // no seed table short enough (only with a non-default min_length): scan every offset
                for (u64 g = 0; g + (u64)q.n <= (u64)L.total; g++) consider<W>(L, q, (u32)g, R, best);
            }
        }
        for (int sd = 0; sd < nseeds; sd++) {           // warp-uniform trip count
            u32 lo = 0, hi = 0;
            int so = 0;
            if (pass && ti >= 0) {
                const KmerTab T = R.tabs[ti];
                const u64 kmask = (T.k == 32) ? ~0ull : ((1ull << (2 * T.k)) - 1ull);
                so = sd * T.k;
                if (!(words_window64<W>(q.n2, so) & kmask)) {   // a seed holding N cannot match exactly
                    const u32 code = (u32)(words_window64<W>(q.s, so) & kmask);
                    lo = __ldg(T.bucket_off + code); hi = __ldg(T.bucket_off + code + 1);
                }
            }
            const u32 cnt = hi - lo;
            if (cnt && cnt < K2_HEAVY) {
                const KmerTab T = R.tabs[ti];
                for (u32 x = lo; x < hi; x++) {
                    const u32 ppos = __ldg(T.pos + x);
                    if (ppos >= (u32)so) consider<W>(L, q, ppos - (u32)so, R, best);
                }
            }
            unsigned heavy = __ballot_sync(0xffffffffu, cnt >= K2_HEAVY);
            while (heavy) {
                const int src = __ffs(heavy) - 1;
                heavy &= heavy - 1;
                const Query<W> qb = query_bcast<W>(q, src);
                const u32 blo = __shfl_sync(0xffffffffu, lo, src), bhi = __shfl_sync(0xffffffffu, hi, src);
                const int bso = __shfl_sync(0xffffffffu, so, src), bti = __shfl_sync(0xffffffffu, ti, src);
                const u32 *pos = R.tabs[bti].pos;
                u64 lb = ~0ull;
                for (u32 x = blo + lane; x < bhi; x += 32) {
                    const u32 ppos = __ldg(pos + x);
                    if (ppos >= (u32)bso) consider<W>(L, qb, ppos - (u32)bso, R, lb);
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const u64 o = __shfl_xor_sync(0xffffffffu, lb, d);
                    lb = o < lb ? o : lb;
                }
                if ((int)lane == src && lb < best) best = lb;
            }
        }
        bool hit = false;
        if (pass && best != ~0ull) {
            hit = true;
            const u32 g = R.highest ? ~(u32)best : (u32)best;
            const int r = ref_of_gpos(L, g);
            A.round_of[id] = (signed char)(R.round + 1);
            A.ref_of[id] = r;
            A.pos_of[id] = (int)(g - __ldg(L.ref_off + r));
            A.mm_of[id] = (signed char)(best >> 32);
            my_aln++;
        }
        const bool push = active && !hit;
        if (R.chunked) {
            const u64 at = warpq_slot(push, n_out_ptr, todo_out, wq);
            if (push) todo_out[at] = id;
        } else {
            const unsigned bal = __ballot_sync(0xffffffffu, push);
            if (bal) {
                const int leader = __ffs(bal) - 1;
                u64 b0 = 0;
                if ((int)lane == leader) b0 = atomicAdd(n_out_ptr, (u64)__popc(bal));
                b0 = __shfl_sync(0xffffffffu, b0, leader);
                if (push) todo_out[b0 + __popc(bal & ((1u << lane) - 1u))] = id;
            }
        }
    }
    warpq_close(todo_out, wq);
    // per-warp statistics
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_proc += __shfl_down_sync(0xffffffffu, my_proc, d);
        my_aln += __shfl_down_sync(0xffffffffu, my_aln, d);
    }
    if (lane == 0) {
        if (my_proc) atomicAdd(stat_processed, (u64)my_proc);
        if (my_aln) atomicAdd(stat_aligned, (u64)my_aln);
    }
}
