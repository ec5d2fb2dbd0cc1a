// k0_parse.cuh -- FASTQ text in HBM -> packed 2-bit read records (K0).
//
// Replaces the reader half of trim_file (utils/trim_file.py:102, cutadapt
// seqio.FastqReader) and cutadapt's QualityTrimmer index computation
// (_qualtrim.pyx quality_trim_index, used at trim_file.py:30): the quality
// string is consumed here, once, and only the (start, stop) trim window
// travels on to K1.
//
//   k0_count_newlines : per-tile newline counts (128-bit loads, SIMD byte compare)
//   scan (scan.cuh)   : exclusive scan of tile counts
//   k0_line_starts    : byte offset of every line start
//   k0_pack           : one thread per 4-line record -> bases / nmask / meta
#pragma once
#include "common.cuh"

#define K0_THREADS 256
#define K0_STEPS 16                                  // 16-byte units per thread per tile
#define K0_TILE_BYTES (K0_THREADS * 16 * K0_STEPS)   // 64 KiB

// 16-bit mask of '\n' bytes in the 16-byte unit `u` of the text; bytes at or
// beyond nbytes never match.
__device__ __forceinline__ u32 nl_mask16(const uint4 *text16, u64 u, u64 nbytes)
{
    u64 byte0 = u * 16ull;
    if (byte0 >= nbytes) return 0;
    uint4 v = __ldg(text16 + u);
    u32 m = 0;
    u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        u32 eq = __vcmpeq4(w[k], 0x0A0A0A0Au);              // 0xFF per matching byte
        u32 bits = ((eq & 0x01010101u) * 0x01020408u) >> 24;  // gather 4 flag bits
        m |= (bits & 0xFu) << (4 * k);
    }
    u64 rem = nbytes - byte0;
    if (rem < 16) m &= (1u << rem) - 1u;
    return m;
}

__global__ void __launch_bounds__(K0_THREADS)
k0_count_newlines(const uint4 *__restrict__ text16, u64 nbytes, u32 *__restrict__ tile_counts)
{
    __shared__ u32 sh[33];
    u64 unit0 = (u64)blockIdx.x * (K0_THREADS * K0_STEPS);
    u32 c = 0;
#pragma unroll 4
    for (int s = 0; s < K0_STEPS; s++) c += __popc(nl_mask16(text16, unit0 + (u64)s * K0_THREADS + threadIdx.x, nbytes));
    u32 total;
    block_excl_scan(c, sh, &total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

// tile_offsets = exclusive scan of tile_counts.  The k-th newline of this region
// (0-based) puts the start of the following line into line_start[idx_base + k];
// offsets are relative to the whole chunk (byte_base = offset of this region).
__global__ void __launch_bounds__(K0_THREADS)
k0_line_starts(const uint4 *__restrict__ text16, u64 nbytes, const u32 *__restrict__ tile_offsets,
               u32 *__restrict__ line_start, u32 idx_base, u64 byte_base)
{
    __shared__ u32 sh[33];
    u64 unit0 = (u64)blockIdx.x * (K0_THREADS * K0_STEPS);
    u32 base = tile_offsets[blockIdx.x] + idx_base;
    for (int s = 0; s < K0_STEPS; s++) {
        u64 u = unit0 + (u64)s * K0_THREADS + threadIdx.x;
        u32 m = nl_mask16(text16, u, nbytes);
        u32 total;
        u32 pre = block_excl_scan(__popc(m), sh, &total);
        u32 idx = base + pre;
        while (m) {
            int b = __ffs(m) - 1;
            m &= m - 1;
            line_start[idx++] = (u32)(byte_base + u * 16ull + b + 1);
        }
        base += total;
    }
}

struct K0Params {
    int phred_base;     // 33 | 64
    int qual_cutoff;    // QualityTrimmer(0, cutoff, base), trim_file.py:30
    int max_len;        // 32 * W
};

// One thread per record.  Record r owns lines 4r .. 4r+3.
template <int W>
__global__ void __launch_bounds__(256)
k0_pack(const unsigned char *__restrict__ text, const u32 *__restrict__ line_start, u64 n_records,
        K0Params prm, u64 *__restrict__ bases, u64 *__restrict__ nmask, u32 *__restrict__ meta,
        u32 *__restrict__ err_flags)
{
    const int NW = W / 2;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_records; r += (u64)gridDim.x * blockDim.x) {
        const u32 *ls = line_start + 4 * r;
        u32 a0 = ls[0], a1 = ls[1], a2 = ls[2], a3 = ls[3], a4 = ls[4];
        u32 bad = 0;
        if (text[a0] != '@' || text[a2] != '+') bad |= ERRF_FORMAT;   // cutadapt FormatError
        int sl = (int)(a2 - 1 - a1), ql = (int)(a4 - 1 - a3);
        while (sl > 0 && text[a1 + sl - 1] == '\r') sl--;
        while (ql > 0 && text[a3 + ql - 1] == '\r') ql--;
        if (sl != ql) bad |= ERRF_FORMAT;
        if (sl > prm.max_len) { bad |= ERRF_TOOLONG; sl = 0; ql = 0; }
        u64 s[W], nm[NW];
#pragma unroll
        for (int w = 0; w < W; w++) s[w] = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) nm[w] = 0;
        if (!bad && sl > 0) {
            // four bases per step: aligned 32-bit loads, realigned with a funnel shift;
            // 2-bit codes and the non-ACGT flags are computed on all four bytes at once
            const u32 *t32 = reinterpret_cast<const u32 *>(text) + (a1 >> 2);
            const u32 sh8 = (a1 & 3u) * 8u;
            u32 wprev = __ldg(t32);
            const int ngroups = (sl + 3) >> 2;
#pragma unroll
            for (int w = 0; w < W; w++) {
                u64 acc = 0;
                u32 nacc = 0;
                for (int gi = w * 8; gi < w * 8 + 8 && gi < ngroups; gi++) {
                    const u32 wnext = __ldg(t32 + gi + 1);
                    const u32 v = __funnelshift_r(wprev, wnext, sh8);       // bytes a1+4gi .. a1+4gi+3
                    wprev = wnext;
                    const u32 up = v & 0xDFDFDFDFu;
                    const u32 ok = __vcmpeq4(up, 0x41414141u) | __vcmpeq4(up, 0x43434343u) | __vcmpeq4(up, 0x47474747u) |
                                   __vcmpeq4(up, 0x54545454u) | __vcmpeq4(up, 0x55555555u);        // A C G T U
                    u32 code = (((v >> 1) ^ (v >> 2)) & 0x03030303u) & ok;   // A=0 C=1 G=2 T/U=3, others 0
                    u32 packed = (code * 0x01041040u) >> 24;                  // 4 x 2 bits
                    u32 nbits = (((~ok) & 0x01010101u) * 0x01020408u) >> 24 & 0xFu;
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
                acc += 0 - ((int)text[a3 + i] - prm.phred_base);
                if (acc < 0) break;
                if (acc > mx) { mx = acc; start = i + 1; }
            }
            acc = 0; mx = 0;
            for (int i = sl - 1; i >= 0; i--) {
                acc += prm.qual_cutoff - ((int)text[a3 + i] - prm.phred_base);
                if (acc < 0) break;
                if (acc > mx) { mx = acc; stop = i; }
            }
            if (start >= stop) { start = 0; stop = 0; }
        }
        if (bad) atomicOr(err_flags, bad);
#pragma unroll
        for (int w = 0; w < W; w++) bases[r * W + w] = s[w];
#pragma unroll
        for (int w = 0; w < NW; w++) nmask[r * NW + w] = nm[w];
        meta[r] = (u32)sl | ((u32)start << 8) | ((u32)stop << 16);
    }
}
