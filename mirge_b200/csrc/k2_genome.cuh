// k2_genome.cuh -- the genome passes: ALL alignments of a read, both strands, counted per mismatch stratum.
//
// Replaces the three `bowtie <genome> ... -a` invocations of the reference:
//   utils/writeDataToCSV.py:1260   `-n 1 -f -a -3 2`   A-to-I candidates; a read is retained when the lowest occupied
//                                  stratum holds exactly one alignment (:1273-1281)
//   utils/writeDataToCSV.py:1488   `-n 0 -f -a -3 2`
//   __main__.py:538                `-n 0 -m 3 -l 25 -a --best`  un-annotated reads of 16-25 nt with >= 2 copies
//                                  (utils/convert2Fasta.py:124,179); -m 3 suppresses reads with more than 3 alignments
// Acceptance as in k2_map.cuh: -n N -> at most N mismatches inside the seed (the read's first
// min(seedlen, n) bases) and at most 2 overall (FASTA input, -e 70); -v V -> at most V end to end.  A reverse-strand
// alignment is a forward alignment of the read's reverse complement, whose seed then sits at the RIGHT end of the
// window.  N in the read always mismatches, a window touching an ambiguous genome base or crossing a chromosome end
// is no alignment.  The kernel returns counts[k] = number of alignments with k mismatches (k = 0..3); which reads
// survive (-m, unique best stratum) is the caller's policy.
//
// Index: ONE table for every seed length k <= kmax.  Positions are grouped by the kmax-mer that starts at them, the
// code written with the FIRST base in the most significant digits, so the positions whose first k bases equal a
// seed are the contiguous range  off[c << 2(kmax-k)] .. off[(c + 1) << 2(kmax-k)]  (c = the seed's code): a
// 10^8-nt genome needs 4^12 + 1 offsets (67 MB) and one position array (400 MB) instead of a table per k.
// Search: one warp per read.  Pigeonhole seeds (N + 1 or V + 1 disjoint pieces of the seed region, one of which must
// match exactly), lanes stride over the seed's position range, XOR / popcount verification over 2-bit words; a
// placement reached through piece j is counted only when no earlier piece matches it exactly as well.
#pragma once
#include "common.cuh"
#include "k2_map.cuh"

struct GenomeIndex {
    const u32 *off;     // [4^kmax + 1]
    const u32 *pos;     // [total]
    int kmax;
};

// code of the k-mer in the low 2k bits of x (base t at bits 2t), first base most significant
__host__ __device__ __forceinline__ u32 kmer_msb_first(u32 x, int k)
{
#ifdef __CUDA_ARCH__
    u32 y = __brev(x);
#else
    u32 y = x;
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
    y = ((y >> 2) & 0x33333333u) | ((y & 0x33333333u) << 2);
    y = ((y >> 4) & 0x0F0F0F0Fu) | ((y & 0x0F0F0F0Fu) << 4);
    y = ((y >> 8) & 0x00FF00FFu) | ((y & 0x00FF00FFu) << 8);
    y = (y >> 16) | (y << 16);
#endif
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);     // bit reversal swapped the two bits of every base
    return k >= 16 ? y : (y >> (32 - 2 * k));
}

// pass 0: bucket[code] += 1; pass 1: pos[bucket[code]++] = g.  Every position is indexed (the tail is padded with A).
__global__ void k2g_index_pass(LibView L, int kmax, int pass, u32 *__restrict__ bucket, u32 *__restrict__ pos)
{
    const u32 mask = (1u << (2 * kmax)) - 1u;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < (u64)L.total; g += (u64)gridDim.x * blockDim.x) {
        const u32 code = kmer_msb_first((u32)lib_window64(L.seq, (u32)g) & mask, kmax);
        if (pass == 0) atomicAdd(&bucket[code], 1u);
        else pos[atomicAdd(&bucket[code], 1u)] = (u32)g;
    }
}

// queries from the unique table: ids[i] -> packed bases / mask / length
template <int W>
__global__ void k2g_gather_queries(const u32 *__restrict__ ids, u64 n, const u64 *__restrict__ useq, const u64 *__restrict__ unmask,
                                   const u8 *__restrict__ ulen, u64 n_unique_cap, u64 *__restrict__ qs, u64 *__restrict__ qnm, u8 *__restrict__ qlen)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 id = ids[i];
        const bool ok = id < n_unique_cap;
#pragma unroll
        for (int w = 0; w < W; w++) qs[i * W + w] = ok ? useq[id * W + w] : 0ull;
#pragma unroll
        for (int w = 0; w < W / 2; w++) qnm[i * (W / 2) + w] = ok ? unmask[id * (W / 2) + w] : 0ull;
        qlen[i] = ok ? ulen[id] : (u8)0;
    }
}

// reverse complement of the first n bases of p
template <int W>
__device__ __forceinline__ void pseq_revcomp(const PSeq<W> &p, int n, PSeq<W> &r)
{
    // reverse the base order of the whole 32W-base register image, complement, then drop the 32W - n leading bases
    u64 t[W], tn[W / 2];
#pragma unroll
    for (int w = 0; w < W; w++) {
        u64 x = p.s[W - 1 - w];
        x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
        x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
        x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
        x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
        x = (x >> 32) | (x << 32);
        t[w] = ~x;
    }
#pragma unroll
    for (int w = 0; w < W / 2; w++) tn[w] = __brevll(p.nm[W / 2 - 1 - w]);
#pragma unroll
    for (int w = 0; w < W; w++) r.s[w] = t[w];
#pragma unroll
    for (int w = 0; w < W / 2; w++) r.nm[w] = tn[w];
    pseq_shift<W>(r, 32 * W - n);
    pseq_truncate<W>(r, n);
}

template <int W>
__device__ __forceinline__ void query_from(const PSeq<W> &p, int n, Query<W> &q)
{
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

// even-bit mask of bases [a, b) within word w
__device__ __forceinline__ u64 range2(int a, int b, int w)
{
    const int lo = w * 32;
    int x = a - lo, y = b - lo;
    if (x < 0) x = 0;
    if (y > 32) y = 32;
    if (y <= x) return 0ull;
    const u64 hi = (y >= 32) ? 0x5555555555555555ull : (0x5555555555555555ull & ((1ull << (2 * y)) - 1ull));
    const u64 lw = (x <= 0) ? 0ull : (0x5555555555555555ull & ((1ull << (2 * x)) - 1ull));
    return hi & ~lw;
}

// Mismatch picture of q placed at g: d[w] = even-bit mask of mismatching bases.  false when the window leaves the
// library or touches an ambiguous base.
template <int W>
__device__ __forceinline__ bool diff_at(const LibView &L, const Query<W> &q, u32 g, u64 *d)
{
    if ((u64)g + (u64)q.n > (u64)L.total) return false;
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
    u64 lo = __ldg(L.seq + wi);
#pragma unroll
    for (int w = 0; w < W; w++) {
        d[w] = 0;
        if (w * 32 < q.n) {
            const u64 hi = __ldg(L.seq + wi + w + 1);
            const u64 r = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
            lo = hi;
            const u64 x = r ^ q.s[w];
            d[w] = (((x | (x >> 1)) & 0x5555555555555555ull) | q.n2[w]) & q.len2[w];
        }
    }
    return true;
}

struct GenomeParams {
    int mode;       // 0 = -n, 1 = -v
    int param;      // N or V
    int seedlen;    // -l
    int trim3;      // -3
    int norc;       // --norc
};

#define K2G_THREADS 128

// counts[q][k]: alignments of query q with k mismatches.
template <int W>
__global__ void __launch_bounds__(K2G_THREADS)
k2g_all_hits(const u64 *__restrict__ qs, const u64 *__restrict__ qnm, const u8 *__restrict__ qlen, u64 n_queries,
             LibView L, GenomeIndex X, GenomeParams P, u32 *__restrict__ counts)
{
    const unsigned lane = threadIdx.x & 31;
    const u64 warp0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 qi = warp0; qi < n_queries; qi += nwarps) {
        PSeq<W> fw;
#pragma unroll
        for (int w = 0; w < W; w++) fw.s[w] = qs[qi * W + w];
#pragma unroll
        for (int w = 0; w < W / 2; w++) fw.nm[w] = qnm[qi * (W / 2) + w];
        const int n = (int)qlen[qi] - P.trim3;
        u32 c[4] = {0, 0, 0, 0};
        if (n > 0 && L.total > 0) {
            pseq_truncate<W>(fw, n);
            const int sl = P.mode == 0 ? (P.seedlen < n ? P.seedlen : n) : n;
            const int nseeds = P.param + 1;
            const int plen = sl / nseeds;                          // piece length (the last piece may be longer; unused tail)
            const int k = plen < X.kmax ? plen : X.kmax;
            for (int strand = 0; strand < (P.norc ? 1 : 2); strand++) {
                PSeq<W> p = fw;
                if (strand) pseq_revcomp<W>(fw, n, p);
                Query<W> q;
                query_from<W>(p, n, q);
                const int r0 = (strand && P.mode == 0) ? n - sl : 0;      // seed region [r0, r0 + sl)
                u64 seedm[W];
#pragma unroll
                for (int w = 0; w < W; w++) seedm[w] = range2(r0, r0 + sl, w);
                if (k < 4) {
                    // no usable seed (reads of a few bases): every offset, lanes strided
                    for (u64 g = lane; g + (u64)n <= (u64)L.total; g += 32) {
                        u64 d[W];
                        if (!diff_at<W>(L, q, (u32)g, d)) continue;
                        int mt = 0, ms = 0;
#pragma unroll
                        for (int w = 0; w < W; w++) { mt += __popcll(d[w]); ms += __popcll(d[w] & seedm[w]); }
                        if (P.mode == 0 ? (ms > P.param || mt > 2) : (mt > P.param)) continue;
                        const int r = ref_of_gpos(L, (u32)g);
                        if (g + (u64)n > (u64)__ldg(L.ref_off + r + 1)) continue;
#pragma unroll
                        for (int t = 0; t < 4; t++) c[t] += (mt == t) ? 1u : 0u;
                    }
                    continue;
                }
                for (int j = 0; j < nseeds; j++) {
                    const int so = r0 + j * plen;                  // the piece's first k bases are looked up
                    const u64 kmask = (1ull << (2 * k)) - 1ull;
                    if (words_window64<W>(q.n2, so) & kmask) continue;        // a piece holding N cannot match exactly
                    const u32 code = kmer_msb_first((u32)(words_window64<W>(q.s, so) & kmask), k);
                    const int up = 2 * (X.kmax - k);
                    const u32 lo = __ldg(X.off + ((u64)code << up)), hi = __ldg(X.off + (((u64)code + 1ull) << up));
                    for (u32 x = lo + lane; x < hi; x += 32) {
                        const u32 ppos = __ldg(X.pos + x);
                        if (ppos < (u32)so) continue;
                        const u32 g = ppos - (u32)so;
                        u64 d[W];
                        if (!diff_at<W>(L, q, g, d)) continue;
                        int mt = 0, ms = 0;
#pragma unroll
                        for (int w = 0; w < W; w++) { mt += __popcll(d[w]); ms += __popcll(d[w] & seedm[w]); }
                        if (P.mode == 0 ? (ms > P.param || mt > 2) : (mt > P.param)) continue;
                        bool dup = false;                          // reached through an earlier piece as well?
                        for (int i = 0; i < j; i++) {
                            const int si = r0 + i * plen;
                            u64 hit = 0;
#pragma unroll
                            for (int w = 0; w < W; w++) hit |= d[w] & range2(si, si + k, w);
                            dup = dup || hit == 0ull;
                        }
                        if (dup) continue;
                        const int r = ref_of_gpos(L, g);
                        if ((u64)g + (u64)n > (u64)__ldg(L.ref_off + r + 1)) continue;
#pragma unroll
                        for (int t = 0; t < 4; t++) c[t] += (mt == t) ? 1u : 0u;
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 4; t++) {
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) c[t] += __shfl_xor_sync(0xffffffffu, c[t], dd);
        }
        if (lane == 0) {
#pragma unroll
            for (int t = 0; t < 4; t++) counts[qi * 4 + t] = c[t];
        }
    }
}
