// k0_parse.cuh -- FASTQ text in HBM -> packed 2-bit read records (K0), ONE pass over the text.
//
// Replaces the reader half of trim_file (utils/trim_file.py:102, cutadapt
// seqio.FastqReader) and cutadapt's QualityTrimmer index computation
// (_qualtrim.pyx quality_trim_index, used at trim_file.py:30): the quality
// string is consumed here, once, and only the (start, stop) trim window
// travels on to K1.
//
// Which of a tile's lines are headers, sequences, '+' or quality lines depends on the number of
// newlines BEFORE the tile -- a prefix sum over the whole text.  A chained scan settles it exactly
// but moves its frontier only ~32 tiles per L2 round trip, far too slowly for 16 KiB tiles at HBM
// speed; a counting pre-pass would read the text twice.  So K0 speculates and verifies:
//
//   k0_parse_spec   no dependency between tiles.  Per 16 KiB tile: (1) coalesced 128-bit loads bring
//                   the tile plus a 1 KiB halo (the bytes before it) into shared memory, '\n' masks of
//                   the 16-byte units computed on the way (SIMD byte compare); (2) one block scan ranks
//                   the tile's newlines and every newline drops the start of the following line into a
//                   shared-memory list (the halo contributes its last four); (3) the tile CLAIMS its
//                   phase: the first of its first four newlines that ends a well-formed record ('@' and
//                   '+' in place, sequence and quality equally long); (4) one thread per record whose
//                   quality line ends in the tile packs it straight out of shared memory -- aligned
//                   32-bit loads realigned by funnel shift, four bases per step (2-bit codes and
//                   non-ACGT flags on all four bytes at once), quality_trim_index on both ends, format
//                   checks -> bases / nmask / meta; output slots come from one atomic per tile (K1
//                   does not care about record order).
//   k0_verify       one small CTA: exclusive scan of the tiles' newline counts -> the true phase of
//                   every tile; all claims right -> the launch stands (line count carried on, error
//                   flags committed); any claim wrong or undecided (pathological or malformed text,
//                   records longer than the halo at a tile start) -> the region is parsed again by
//   k0_parse_pack   the exact path: the same tile code behind a chained scan with warp-wide decoupled
//                   look-back, seeded with the line count carried over from the previous launch.  It is
//                   launched after every k0_verify and returns at once unless asked for.
// Records longer than the halo (kilobyte headers) take a byte-wise path through global memory; tiles
// with more than 256 records (reads of a few bases) are processed in several passes.
#pragma once
#include "common.cuh"

#define K0_THREADS 256
#define K0_TILE 16384
#define K0_HALO 1024
#define K0_UNITS ((K0_TILE + K0_HALO) / 16)
#define K0_RPP 160                  // records per pass: one per thread in the per-record phases
#define K0_CAP (4 * K0_RPP + 4)      // line-start list entries per pass
#define K0_MW (K0_TILE / K0_THREADS / 32)   // 32-bit newline-mask words of the bytes one thread ranks
#define K0_NONE 0xFFFFu             // list entry: the line starts before the halo

#define K0_ST_AGG (1ull << 62)      // tile status: newline count of this tile
#define K0_ST_PREFIX (2ull << 62)   // tile status: newline count of the chunk up to and including this tile
#define K0_ST_VAL ((1ull << 62) - 1ull)

struct K0Params {
    int phred_base;     // 33 | 64
    int qual_cutoff;    // QualityTrimmer(0, cutoff, base), trim_file.py:30
    int max_len;        // 32 * W
    int min_len;        // reads shorter than this are counted but not emitted (the length filter would drop them anyway)
};

// Carried from launch to launch (the regions of one chunk); lines[] alternates so that a launch
// reads the slot its predecessor wrote while it writes the other one.
struct K0State {
    u64 lines[2];       // newlines of the chunk before the region
    u64 nrec;           // records packed by the last launch (K1 reads it)
    u32 err_spec;       // format flags raised by the speculative launch (committed by k0_verify)
    u32 need_exact;     // k0_verify: the speculation failed, k0_parse_pack has to redo the region
};

#define K0_CLAIM_NONE 4u        // none of the tile's (fewer than four) newlines ends a record
#define K0_CLAIM_UNDECIDED 5u   // a candidate could not be tested inside the halo

// 0x80 in every byte of x that is zero (exact for every byte, no borrow between bytes)
__device__ __forceinline__ u32 zero_bytes80(u32 x)
{
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}

__device__ __forceinline__ u32 nl_mask_words(u32 x, u32 y, u32 z, u32 w)
{
    const u32 f0 = zero_bytes80(x ^ 0x0A0A0A0Au), f1 = zero_bytes80(y ^ 0x0A0A0A0Au);
    const u32 f2 = zero_bytes80(z ^ 0x0A0A0A0Au), f3 = zero_bytes80(w ^ 0x0A0A0A0Au);
    if (!(f0 | f1 | f2 | f3)) return 0;                  // most 16-byte units hold no newline
    u32 m = 0;
    const u32 f[4] = {f0, f1, f2, f3};
#pragma unroll
    for (int k = 0; k < 4; k++) m |= ((((f[k] >> 7) * 0x01020408u) >> 24) & 0xFu) << (4 * k);   // gather the 4 flag bits
    return m;
}

// Four ASCII bytes -> 2-bit codes (A=0 C=1 G=2 T/U=3, either case; 8 bits) and non-ACGTU flags (4 bits).
// The code comes from bits 1-2 of the byte; a byte is a base iff it equals the letter its code stands for: the
// expected letters are fetched with one PRMT (the codes as selector nibbles) and compared with bit 0 forced to one
// (T and U differ there only), and bit 0 must be set unless the code is 3 ('@', 'B', 'F' are not A, C, G).
__device__ __forceinline__ void k0_code4(u32 v, u32 &packed8, u32 &nbits4)
{
    const u32 c = ((v >> 1) ^ (v >> 2)) & 0x03030303u;
    const u32 t = c | (c >> 4);
    const u32 sel = (t & 0x00FFu) | ((t >> 8) & 0xFF00u);
    const u32 expect = __byte_perm(0x55474341u, 0u, sel);
    const u32 up = v & 0xDFDFDFDFu;
    const u32 z = zero_bytes80((up | 0x01010101u) ^ expect);
    const u32 ok80 = z & (((up | (c & (c >> 1))) & 0x01010101u) << 7);
    const u32 okm = (ok80 >> 7) | (ok80 >> 6);           // 0x03 per valid byte
    packed8 = ((c & okm) * 0x01041040u) >> 24;
    nbits4 = ((((~ok80) & 0x80808080u) >> 7) * 0x01020408u) >> 24 & 0xFu;
}

__device__ __forceinline__ u64 k0_ld_status(const u64 *p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void k0_st_status(u64 *p, u64 v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// Packs one record.  F(off) returns the aligned 32-bit word at byte offset off (multiple of 4) and
// B(off) one byte; offsets a0..a4 are the starts of the record's four lines and of the line after it.
template <int W, typename F32, typename F8>
__device__ __forceinline__ void k0_pack_record(F32 ld32, F8 ld8, u32 a0, u32 a1, u32 a2, u32 a3, u32 a4, const K0Params &prm,
                                               u64 *s, u64 *nm, u32 &meta, u32 &bad)
{
    const int NW = W / 2;
    bad = 0;
    if (ld8(a0) != '@' || ld8(a2) != '+') bad |= ERRF_FORMAT;       // cutadapt FormatError
    int sl = (int)(a2 - 1 - a1), ql = (int)(a4 - 1 - a3);
    while (sl > 0 && ld8(a1 + sl - 1) == '\r') sl--;
    while (ql > 0 && ld8(a3 + ql - 1) == '\r') ql--;
    if (sl != ql) bad |= ERRF_FORMAT;
    if (sl > prm.max_len) { bad |= ERRF_TOOLONG; sl = 0; ql = 0; }
#pragma unroll
    for (int w = 0; w < W; w++) s[w] = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) nm[w] = 0;
    if (!bad && sl > 0) {
        const u32 t0 = a1 & ~3u;
        const u32 sh8 = (a1 & 3u) * 8u;
        u32 wprev = ld32(t0);
        const int ngroups = (sl + 3) >> 2;
#pragma unroll
        for (int w = 0; w < W; w++) {
            u64 acc = 0;
            u32 nacc = 0;
            for (int gi = w * 8; gi < w * 8 + 8 && gi < ngroups; gi++) {
                const u32 wnext = ld32(t0 + 4u * (u32)(gi + 1));
                const u32 v = __funnelshift_r(wprev, wnext, sh8);       // bytes a1+4gi .. a1+4gi+3
                wprev = wnext;
                u32 packed, nbits;
                k0_code4(v, packed, nbits);
                const int left = sl - gi * 4;
                if (left < 4) { packed &= (1u << (2 * left)) - 1u; nbits &= (1u << left) - 1u; }
                acc |= (u64)packed << ((gi & 7) * 8);
                nacc |= nbits << ((gi & 7) * 4);
            }
            s[w] = acc;
            nm[w >> 1] |= (u64)nacc << ((w & 1) * 32);
        }
    }
    // quality_trim_index(qualities, cutoff_front=0, cutoff_back, base)
    int start = 0, stop = sl;
    if (!bad) {
        int acc = 0, mx = 0;
        for (int i = 0; i < sl; i++) {
            acc += 0 - ((int)ld8(a3 + i) - prm.phred_base);
            if (acc < 0) break;
            if (acc > mx) { mx = acc; start = i + 1; }
        }
        acc = 0; mx = 0;
        for (int i = sl - 1; i >= 0; i--) {
            acc += prm.qual_cutoff - ((int)ld8(a3 + i) - prm.phred_base);
            if (acc < 0) break;
            if (acc > mx) { mx = acc; stop = i; }
        }
        if (start >= stop) { start = 0; stop = 0; }
    }
    meta = (u32)sl | ((u32)start << 8) | ((u32)stop << 16);
}


// ---- tile building blocks shared by the speculative and the exact kernel -------------------------

// 1. halo + tile into shared memory, newline mask per 16-byte unit.  Bytes at or beyond reg_end are not
// touched (the next region may not have landed yet); vnl puts a virtual newline at byte reg_end.
__device__ __forceinline__ void k0_load_tile(const unsigned char *__restrict__ text, i64 halo_start, u64 reg_end, int vnl,
                                             unsigned char *buf, u16 *m16)
{
    constexpr int PER = K0_TILE / 16 / K0_THREADS;       // tile units per thread
    if (halo_start >= 0 && (u64)halo_start + K0_HALO + K0_TILE <= reg_end) {
        // interior tile: every load is issued before the first one is consumed (the copy is latency bound otherwise)
        const uint4 *src = reinterpret_cast<const uint4 *>(text + halo_start);
        uint4 v[PER + 1];
#pragma unroll
        for (int k = 0; k < PER; k++) v[k] = __ldg(src + K0_HALO / 16 + k * K0_THREADS + threadIdx.x);
        v[PER] = make_uint4(0, 0, 0, 0);
        if (threadIdx.x < K0_HALO / 16) v[PER] = __ldg(src + threadIdx.x);
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int u = K0_HALO / 16 + k * K0_THREADS + threadIdx.x;
            *reinterpret_cast<uint4 *>(buf + 16 * u) = v[k];
            m16[u] = (u16)nl_mask_words(v[k].x, v[k].y, v[k].z, v[k].w);
        }
        if (threadIdx.x < K0_HALO / 16) {
            *reinterpret_cast<uint4 *>(buf + 16 * threadIdx.x) = v[PER];
            m16[threadIdx.x] = (u16)nl_mask_words(v[PER].x, v[PER].y, v[PER].z, v[PER].w);
        }
        return;
    }
    for (int u = threadIdx.x; u < K0_UNITS; u += K0_THREADS) {
        const i64 g0 = halo_start + 16 * (i64)u;
        uint4 v = make_uint4(0, 0, 0, 0);
        u32 m = 0;
        if (g0 >= 0 && (u64)g0 < reg_end) {
            if ((u64)g0 + 16 <= reg_end) {
                v = __ldg(reinterpret_cast<const uint4 *>(text + g0));
            } else {                                     // the chunk's last, partial unit: byte loads only
                u32 w4[4] = {0, 0, 0, 0};
                for (int b = 0; (u64)g0 + b < reg_end; b++) w4[b >> 2] |= (u32)text[g0 + b] << ((b & 3) * 8);
                v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
            m = nl_mask_words(v.x, v.y, v.z, v.w);
            const u64 rem = reg_end - (u64)g0;
            if (rem < 16) m &= (1u << rem) - 1u;
        }
        if (vnl && g0 <= (i64)reg_end && (i64)reg_end < g0 + 16) m |= 1u << (u32)((i64)reg_end - g0);
        *reinterpret_cast<uint4 *>(buf + 16 * u) = v;
        m16[u] = (u16)m;
    }
}

// 3. line-start list of one pass: entry o + 4 - pass_lo = shared offset of the line that follows the
// tile's o-th newline; in pass 0 entries 0..3 come from the halo's last four newlines.
__device__ __forceinline__ void k0_fill_list(u32 pass_lo, u32 ex, const u32 *mw, u64 tile_start, const u16 *m16, u16 *lst)
{
    const unsigned tid = threadIdx.x;
    u32 o = ex;
#pragma unroll
    for (int k = 0; k < K0_MW; k++) {
        u32 m = mw[k];
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const u32 li = o + 4 - pass_lo;
            if (li < K0_CAP) lst[li] = (u16)(K0_HALO + (K0_TILE / K0_THREADS) * tid + 32 * k + b + 1);
            o++;
        }
    }
    if (pass_lo == 0 && tid == 32) {
        int found = 0;
        if (tile_start == 0) { lst[3] = (u16)K0_HALO; found = 1; }     // the chunk's first line
        else {
            for (int u = K0_HALO / 16 - 1; u >= 0 && found < 4; u--) {
                u32 m = m16[u];
                while (m && found < 4) {
                    const int b = 31 - __clz(m);
                    m &= ~(1u << b);
                    lst[3 - found] = (u16)(16 * u + b + 1);
                    found++;
                }
            }
        }
        for (; found < 4; found++) lst[3 - found] = K0_NONE;
    }
}

// Per-pass staging area in shared memory: the records of one pass in their final layout.
template <int W>
struct K0Stage {
    u64 bases[K0_RPP * W];
    u64 nmask[K0_RPP * (W / 2)];
    u32 meta[K0_RPP];
    u16 a1[K0_RPP];         // shared offset of the record's sequence line
    u8 sl[K0_RPP];          // bases to pack (0: none); K0_SL_DONE: phase A filled the stage itself (global path)
};
#define K0_SL_DONE 0xFFu

// 4. the records whose terminating newline has ordinal o in [pass_lo, pass_lo + K0_CAP - 4) and o = o0 (mod 4).
// Records are emitted in no particular order (K1 does not care): the pass takes its output slots with one atomic.
// A read shorter than min_len can never survive the length filter (trimming only shortens), so it is checked and
// counted but not emitted -- which also bounds the records of a region by bytes / (6 + 2 * min_len).  Three phases:
//   A  one thread per record: format checks, lengths, quality_trim_index -> meta, (a1, sl) for phase B; a block scan
//      of the emit flags compacts the pass
//   B  one thread per 8 bases: two aligned-word groups -> 16 bits of bases + 8 N flags, stored in place in the stage
//      (short critical path, nearly all lanes busy -- a thread per record leaves half of them idle for 13 steps)
//   C  the stage goes out as three flat, coalesced copies
template <int W>
__device__ __forceinline__ void k0_pack_pass(const unsigned char *__restrict__ text, u64 reg_end, i64 delta, const unsigned char *buf,
                                             const u16 *lst, u32 pass_lo, u32 o0, u32 T, u64 *slot_counter, u64 rec_cap, const K0Params &prm,
                                             K0Stage<W> &st, u32 *sh_scan, u64 *sh_slot, u64 *__restrict__ bases, u64 *__restrict__ nmask,
                                             u32 *__restrict__ meta, u32 *err_word)
{
    constexpr int NW = W / 2;
    const unsigned tid = threadIdx.x;
    const u32 lo = pass_lo > o0 ? pass_lo + ((o0 + 4 - (pass_lo & 3)) & 3) : o0;
    u32 hi = pass_lo + K0_CAP - 4;
    if (hi > T) hi = T;
    const u32 nrec_all = hi > lo ? (hi - lo + 3) >> 2 : 0u;    // <= K0_RPP
    if (nrec_all == 0) return;                                 // block-uniform
    // ---- A
    u32 emit = 0, mt = 0, a1s = 0, sls = 0;
    u64 sq[W], nq[W / 2];
    if (tid < nrec_all) {
        const u32 li = lo + 4 * tid + 4 - pass_lo;
        u32 l[5];
#pragma unroll
        for (int k = 0; k < 5; k++) l[k] = lst[li - 4 + k];
        u32 bad = 0;
        if (l[0] != K0_NONE) {
            const u32 a0 = l[0], a1 = l[1], a2 = l[2], a3 = l[3], a4 = l[4];
            if (buf[a0] != '@' || buf[a2] != '+') bad |= ERRF_FORMAT;       // cutadapt FormatError
            int sl = (int)(a2 - 1 - a1), ql = (int)(a4 - 1 - a3);
            while (sl > 0 && buf[a1 + sl - 1] == '\r') sl--;
            while (ql > 0 && buf[a3 + ql - 1] == '\r') ql--;
            if (sl != ql) bad |= ERRF_FORMAT;
            if (sl > prm.max_len) { bad |= ERRF_TOOLONG; sl = 0; }
            if (bad) sl = 0;
            // quality_trim_index(qualities, cutoff_front=0, cutoff_back, base)
            int start = 0, stop = sl;
            {
                int acc = 0, mx = 0;
                for (int i = 0; i < sl; i++) {
                    acc += 0 - ((int)buf[a3 + i] - prm.phred_base);
                    if (acc < 0) break;
                    if (acc > mx) { mx = acc; start = i + 1; }
                }
                acc = 0; mx = 0;
                for (int i = sl - 1; i >= 0; i--) {
                    acc += prm.qual_cutoff - ((int)buf[a3 + i] - prm.phred_base);
                    if (acc < 0) break;
                    if (acc > mx) { mx = acc; stop = i; }
                }
                if (start >= stop) { start = 0; stop = 0; }
            }
            mt = (u32)sl | ((u32)start << 8) | ((u32)stop << 16);
            a1s = a1; sls = (u32)sl;
            emit = sl >= prm.min_len ? 1u : 0u;
        } else {
            // a line of this record starts before the halo: find the missing starts in global memory, pack byte-wise
            u64 g[5];
            g[4] = (u64)((i64)l[4] + delta);
            for (int k = 3; k >= 0; k--) {
                if (l[k] != K0_NONE) { g[k] = (u64)((i64)l[k] + delta); continue; }
                u64 p = g[k + 1] - 1;                // the newline that ends line k
                while (p > 0 && text[p - 1] != '\n') p--;
                g[k] = p;
            }
            const u64 gb = g[0] & ~3ull;
            const unsigned char *tb = text + gb;
            const u64 lim = reg_end;
            auto ld8 = [tb, gb, lim](u32 off) { return gb + off < lim ? (u32)tb[off] : 0u; };
            auto ld32 = [ld8](u32 off) { return ld8(off) | (ld8(off + 1) << 8) | (ld8(off + 2) << 16) | (ld8(off + 3) << 24); };
#pragma unroll
            for (int w = 0; w < W; w++) sq[w] = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) nq[w] = 0;
            if (g[4] - gb >= 0xFFFFFF00ull) bad = ERRF_FORMAT;
            else k0_pack_record<W>(ld32, ld8, (u32)(g[0] - gb), (u32)(g[1] - gb), (u32)(g[2] - gb), (u32)(g[3] - gb), (u32)(g[4] - gb),
                                   prm, sq, nq, mt, bad);
            sls = K0_SL_DONE;
            emit = (int)(mt & 0xFFu) >= prm.min_len ? 1u : 0u;
        }
        if (bad) atomicOr(err_word, bad);
    }
    u32 nrec;
    const u32 e = block_excl_scan(emit, sh_scan, &nrec);       // compacted index of this thread's record
    if (emit) {
        st.meta[e] = mt;
        st.a1[e] = (u16)a1s;
        st.sl[e] = (u8)sls;
        if (sls == K0_SL_DONE) {
#pragma unroll
            for (int w = 0; w < W; w++) st.bases[e * W + w] = sq[w];
#pragma unroll
            for (int w = 0; w < NW; w++) st.nmask[e * NW + w] = nq[w];
        }
    }
    if (tid == 0) *sh_slot = nrec ? atomicAdd(slot_counter, (u64)nrec) : 0ull;
    __syncthreads();
    const u64 slotp = *sh_slot;
    if (nrec == 0) return;                                     // block-uniform
    // ---- B
    {
        u16 *sb = reinterpret_cast<u16 *>(st.bases);
        u8 *sn = reinterpret_cast<u8 *>(st.nmask);
        for (u32 it = tid; it < nrec * (4 * W); it += K0_THREADS) {
            const u32 rec = it / (4 * W), pr = it % (4 * W);
            const u32 sl = st.sl[rec];
            if (sl == K0_SL_DONE) continue;
            u32 p16 = 0, n8 = 0;
            const int left = (int)sl - 8 * (int)pr;
            if (left > 0) {
                const u32 a = (u32)st.a1[rec] + 8 * pr;
                const u32 t0 = a & ~3u, sh8 = (a & 3u) * 8u;
                const u32 w0 = *reinterpret_cast<const u32 *>(buf + t0), w1 = *reinterpret_cast<const u32 *>(buf + t0 + 4);
                const u32 w2 = *reinterpret_cast<const u32 *>(buf + t0 + 8);
                u32 pa, na, pb, nb;
                k0_code4(__funnelshift_r(w0, w1, sh8), pa, na);
                k0_code4(__funnelshift_r(w1, w2, sh8), pb, nb);
                p16 = pa | (pb << 8);
                n8 = na | (nb << 4);
                if (left < 8) { p16 &= (1u << (2 * left)) - 1u; n8 &= (1u << left) - 1u; }
            }
            sb[it] = (u16)p16;
            sn[it] = (u8)n8;
        }
    }
    __syncthreads();
    // ---- C
    {
        u32 n_out = nrec;
        if (slotp + nrec > rec_cap) {
            if (tid == 0) atomicOr(err_word, ERRF_FORMAT);
            n_out = slotp < rec_cap ? (u32)(rec_cap - slotp) : 0u;
        }
        for (u32 i = tid; i < n_out * W; i += K0_THREADS) bases[slotp * W + i] = st.bases[i];
        for (u32 i = tid; i < n_out * NW; i += K0_THREADS) nmask[slotp * NW + i] = st.nmask[i];
        for (u32 i = tid; i < n_out; i += K0_THREADS) meta[slotp + i] = st.meta[i];
    }
}

// Does the tile's newline with list index li end a well-formed record?  (first-record test of the claim)
__device__ __forceinline__ u32 k0_test_candidate(const unsigned char *buf, const u16 *lst, u32 li, bool chunk_start)
{
    u32 l[5];
#pragma unroll
    for (int k = 0; k < 5; k++) l[k] = lst[li - 4 + k];
    if (l[0] == K0_NONE) return chunk_start ? 0u : 2u;   // fewer than four lines before it in the chunk: no record end; else cannot tell inside the halo
    if (buf[l[0]] != '@' || buf[l[2]] != '+') return 0u;
    int sl = (int)(l[2] - 1 - l[1]), ql = (int)(l[4] - 1 - l[3]);
    while (sl > 0 && buf[l[1] + sl - 1] == '\r') sl--;
    while (ql > 0 && buf[l[3] + ql - 1] == '\r') ql--;
    return sl == ql ? 1u : 0u;
}

// text: start of the chunk; the launch covers bytes [reg_start, reg_end) of it (reg_start a multiple
// of K0_TILE); vnl: the chunk's last line has no '\n' -- a virtual one sits at byte reg_end (last
// region only).  info[t] = newline count of tile t | claim << 20.  state->nrec and state->err_spec
// are zeroed before the launch.
template <int W>
__global__ void __launch_bounds__(K0_THREADS, 8)
k0_parse_spec(const unsigned char *__restrict__ text, u64 reg_start, u64 reg_end, int vnl, u32 ntiles,
              u32 *__restrict__ info, K0State *__restrict__ state, K0Params prm, u64 rec_cap,
              u64 *__restrict__ bases, u64 *__restrict__ nmask, u32 *__restrict__ meta)
{
    __shared__ __align__(16) unsigned char buf[K0_HALO + K0_TILE + 16];
    __shared__ __align__(16) u16 m16[K0_UNITS + 8];
    __shared__ u16 lst[K0_CAP];
    __shared__ u32 sh_scan[33];
    __shared__ u32 sh_claim;
    __shared__ u64 sh_slot0;
    __shared__ K0Stage<W> stage;
    const unsigned tid = threadIdx.x;
    for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const u64 tile_start = reg_start + (u64)tile * K0_TILE;
        const i64 halo_start = (i64)tile_start - K0_HALO;
        k0_load_tile(text, halo_start, reg_end, vnl, buf, m16);
        __syncthreads();
        u32 mw[K0_MW], cnt = 0;                                                             // this thread's K0_TILE / K0_THREADS bytes
#pragma unroll
        for (int k = 0; k < K0_MW; k++) {
            mw[k] = *reinterpret_cast<const u32 *>(m16 + K0_HALO / 16 + 2 * K0_MW * tid + 2 * k);
            cnt += __popc(mw[k]);
        }
        u32 T;
        const u32 ex = block_excl_scan(cnt, sh_scan, &T);
        u32 o0 = K0_CLAIM_NONE;
        for (u32 pass_lo = 0; pass_lo == 0 || pass_lo < T; pass_lo += K0_CAP - 4) {
            k0_fill_list(pass_lo, ex, mw, tile_start, m16, lst);
            __syncthreads();
            if (pass_lo == 0) {
                if (tid < 32) {                              // the claim: first of the first four newlines that ends a record
                    const u32 r = (tid < 4 && tid < T) ? k0_test_candidate(buf, lst, tid + 4, tile_start == 0) : 0u;
                    const unsigned okm = __ballot_sync(0xffffffffu, r == 1u), unk = __ballot_sync(0xffffffffu, r == 2u);
                    u32 claim = okm ? (u32)(__ffs(okm) - 1) : K0_CLAIM_NONE;
                    if (unk && (!okm || (u32)(__ffs(unk) - 1) < claim)) claim = K0_CLAIM_UNDECIDED;
                    if (tid == 0) {
                        sh_claim = claim;
                        info[tile] = T | (claim << 20);
                    }
                }
                __syncthreads();
                o0 = sh_claim;
            }
            if (o0 < 4)
                k0_pack_pass<W>(text, reg_end, halo_start, buf, lst, pass_lo, o0, T, &state->nrec, rec_cap, prm, stage, sh_scan, &sh_slot0,
                                bases, nmask, meta, &state->err_spec);
            __syncthreads();
        }
    }
}

// True phase of every tile from the exclusive scan of the newline counts; commits or rejects the
// speculative launch.  One CTA.
__global__ void __launch_bounds__(1024)
k0_verify(const u32 *__restrict__ info, u32 ntiles, K0State *__restrict__ state, int par, int last_region,
          u32 *__restrict__ err_flags, u64 *__restrict__ processed, u64 *__restrict__ n_exact)
{
    __shared__ u32 sh_scan[33];
    const u64 lines_in = state->lines[par];
    u64 carry = 0;
    int bad = 0;
    constexpr u32 ITEMS = 16;                                        // consecutive tiles per thread and round
    for (u32 base = 0; base < ntiles; base += 1024 * ITEMS) {
        const u32 t0 = base + threadIdx.x * ITEMS;
        u32 v[ITEMS], sum = 0;
#pragma unroll
        for (u32 k = 0; k < ITEMS; k++) { v[k] = (t0 + k < ntiles) ? info[t0 + k] : 0u; sum += v[k] & 0xFFFFFu; }
        u32 total;
        u64 G = lines_in + carry + block_excl_scan(sum, sh_scan, &total);
#pragma unroll
        for (u32 k = 0; k < ITEMS; k++) {
            const u32 T = v[k] & 0xFFFFFu, claim = v[k] >> 20;
            const u32 o0 = (u32)((3ull - (G & 3ull)) & 3ull);
            const u32 expect = o0 < T ? o0 : K0_CLAIM_NONE;
            if (t0 + k < ntiles && claim != expect) bad = 1;
            G += T;
        }
        carry += total;
    }
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        if (bad) { state->need_exact = 1u; state->nrec = 0; atomicAdd(n_exact, 1ull); }     // the exact kernel emits again
        else {
            const u64 lines_out = lines_in + carry;
            state->need_exact = 0u;
            state->lines[par ^ 1] = lines_out;
            if (state->err_spec) atomicOr(err_flags, state->err_spec);
            atomicAdd(processed, (lines_out >> 2) - (lines_in >> 2));      // every record counts, emitted or not
            if (last_region && (lines_out & 3ull)) atomicOr(err_flags, ERRF_PARTIAL);
        }
    }
}

// The exact path (runs only when k0_verify asks for it).  ctl[0] = tile ticket, ctl[1 + t] = status of
// tile t (zeroed before the launch).  par: which K0State::lines slot holds the count carried in.
// last_region: the chunk ends with this launch, so the line count must be a multiple of four.
template <int W>
__global__ void __launch_bounds__(K0_THREADS)
k0_parse_pack(const unsigned char *__restrict__ text, u64 reg_start, u64 reg_end, int vnl, int last_region, u32 ntiles,
              u64 *__restrict__ ctl, K0State *__restrict__ state, int par, K0Params prm, u64 rec_cap,
              u64 *__restrict__ bases, u64 *__restrict__ nmask, u32 *__restrict__ meta, u32 *__restrict__ err_flags,
              u64 *__restrict__ processed, int always)
{
    __shared__ __align__(16) unsigned char buf[K0_HALO + K0_TILE + 16];
    __shared__ __align__(16) u16 m16[K0_UNITS + 8];
    __shared__ u16 lst[K0_CAP];
    __shared__ u32 sh_scan[33];
    __shared__ u32 sh_tile;
    __shared__ u64 sh_excl;
    __shared__ u64 sh_slot0;
    __shared__ K0Stage<W> stage;
    if (!always && !state->need_exact) return;
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const u64 lines_in = state->lines[par];
    const u64 rec0 = lines_in >> 2;
    u64 *status = ctl + 1;
    for (;;) {
        if (tid == 0) sh_tile = (u32)atomicAdd(ctl, 1ull);
        __syncthreads();
        const u32 tile = sh_tile;
        if (tile >= ntiles) break;
        const u64 tile_start = reg_start + (u64)tile * K0_TILE;
        const i64 halo_start = (i64)tile_start - K0_HALO;
        k0_load_tile(text, halo_start, reg_end, vnl, buf, m16);
        __syncthreads();
        // rank the tile's newlines; chained scan over the tiles
        u32 mw[K0_MW], cnt = 0;                                                             // this thread's K0_TILE / K0_THREADS bytes
#pragma unroll
        for (int k = 0; k < K0_MW; k++) {
            mw[k] = *reinterpret_cast<const u32 *>(m16 + K0_HALO / 16 + 2 * K0_MW * tid + 2 * k);
            cnt += __popc(mw[k]);
        }
        u32 T;
        const u32 ex = block_excl_scan(cnt, sh_scan, &T);
        if (wid == 0) {
            u64 excl = lines_in;
            if (tile > 0) {
                if (lane == 0) k0_st_status(status + tile, K0_ST_AGG | (u64)T);
                excl = 0;
                i64 look = (i64)tile - 1;
                for (;;) {
                    const i64 idx = look - lane;
                    u64 s = idx >= 0 ? k0_ld_status(status + idx) : K0_ST_PREFIX;      // below tile 0: nothing to add
                    const unsigned pm = __ballot_sync(0xffffffffu, (s >> 62) == 2ull);
                    const unsigned im = __ballot_sync(0xffffffffu, (s >> 62) == 0ull);
                    const int fp = pm ? __ffs(pm) - 1 : 32;
                    const unsigned need = fp < 31 ? ((2u << fp) - 1u) : 0xffffffffu;
                    if (im & need) { __nanosleep(200); continue; }                     // a tile in the window has not published yet
                    u64 val = ((int)lane <= fp) ? (s & K0_ST_VAL) : 0ull;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
                    excl += val;
                    if (fp < 32) break;
                    look -= 32;
                }
            }
            if (lane == 0) {
                k0_st_status(status + tile, K0_ST_PREFIX | (excl + T));
                sh_excl = excl;
            }
        }
        for (u32 pass_lo = 0; pass_lo == 0 || pass_lo < T; pass_lo += K0_CAP - 4) {
            k0_fill_list(pass_lo, ex, mw, tile_start, m16, lst);
            __syncthreads();
            const u64 G = sh_excl;
            const u32 o0 = (u32)((3ull - (G & 3ull)) & 3ull);        // first record-terminating newline of the tile
            k0_pack_pass<W>(text, reg_end, halo_start, buf, lst, pass_lo, o0, T, &state->nrec, rec_cap, prm, stage, sh_scan, &sh_slot0,
                            bases, nmask, meta, err_flags);
            __syncthreads();
        }
        if (tile == ntiles - 1 && tid == 0) {                // hand the line count on to the next launch
            const u64 lines_out = sh_excl + T;
            state->lines[par ^ 1] = lines_out;
            atomicAdd(processed, (lines_out >> 2) - rec0);        // every record counts, emitted or not
            if (last_region && (lines_out & 3ull)) atomicOr(err_flags, ERRF_PARTIAL);
        }
        __syncthreads();
    }
}
