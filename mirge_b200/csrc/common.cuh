// common.cuh -- shared device helpers: 2-bit packed sequences, hashing, scans.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;
typedef long long i64;

#define MIRGE_MAX_ADAPTERS 8
#define MIRGE_MAX_ADAPTER_LEN 64

// device error flags (DevCounters::err)
#define ERRF_FORMAT 1u
#define ERRF_TOOLONG 2u
#define ERRF_FULL 4u
#define ERRF_PARTIAL 8u     // the chunk does not end on a record boundary

// Bit helpers usable from host code too (tests/helpers/prefilter_host.cu runs the
// prefilter's device function on the CPU); on the device they are the intrinsics.
__host__ __device__ __forceinline__ int bit_ffs32(u32 x)
{
#ifdef __CUDA_ARCH__
    return __ffs((int)x);
#else
    return __builtin_ffs((int)x);
#endif
}
__host__ __device__ __forceinline__ int bit_ffs64(u64 x)
{
#ifdef __CUDA_ARCH__
    return __ffsll((long long)x);
#else
    return __builtin_ffsll((long long)x);
#endif
}
__host__ __device__ __forceinline__ int bit_popc32(u32 x)
{
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
__host__ __device__ __forceinline__ int bit_clz64(u64 x)     // x != 0 on the host side
{
#ifdef __CUDA_ARCH__
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
// Warp votes; on the host a "warp" is the single calling thread (the per-lane logic is what the host tests run).
__host__ __device__ __forceinline__ int warp_max_all(int v)
{
#ifdef __CUDA_ARCH__
    return __reduce_max_sync(0xffffffffu, v);
#else
    return v;
#endif
}
__host__ __device__ __forceinline__ bool warp_any_all(bool p)
{
#ifdef __CUDA_ARCH__
    return __any_sync(0xffffffffu, p) != 0;
#else
    return p;
#endif
}
__host__ __device__ __forceinline__ u32 bit_funnel_r(u32 lo, u32 hi, int sh)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    return (u32)(((((u64)hi) << 32) | (u64)lo) >> (sh & 31));
#endif
}

// ---------------------------------------------------------------------------
// Packed sequence: base i lives in bits [2i, 2i+1] of word i/32; A=0 C=1 G=2 T=3.
// Ambiguity mask: bit i of word i/64 set when base i is not ACGT (packed as 0).
// W = number of 64-bit base words (2 => <=64 nt, 4 => <=128 nt); NW = W/2.
// ---------------------------------------------------------------------------
template <int W>
struct PSeq {
    u64 s[W];
    u64 nm[W / 2];
};

__host__ __device__ __forceinline__ int base_code(unsigned char c)
{
    c &= 0xDF;  // upper-case
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : (c == 'T' || c == 'U') ? 3 : 4;
}

template <int W>
__host__ __device__ __forceinline__ int pseq_base(const PSeq<W> &p, int j)
{
    u64 w = p.s[0];
#pragma unroll
    for (int a = 1; a < W; a++) if ((j >> 5) == a) w = p.s[a];
    return (int)((w >> ((j & 31) * 2)) & 3ull);
}
template <int W>
__host__ __device__ __forceinline__ int pseq_isn(const PSeq<W> &p, int j)
{
    u64 w = p.nm[0];
#pragma unroll
    for (int a = 1; a < W / 2; a++) if ((j >> 6) == a) w = p.nm[a];
    return (int)((w >> (j & 63)) & 1ull);
}

// keep bases [0, n): clear everything at and above base n
template <int W>
__device__ __forceinline__ void pseq_truncate(PSeq<W> &p, int n)
{
#pragma unroll
    for (int w = 0; w < W; w++) {
        int lo = w * 32;
        if (n <= lo) p.s[w] = 0;
        else if (n < lo + 32) p.s[w] &= (1ull << ((n - lo) * 2)) - 1ull;
    }
#pragma unroll
    for (int w = 0; w < W / 2; w++) {
        int lo = w * 64;
        if (n <= lo) p.nm[w] = 0;
        else if (n < lo + 64) p.nm[w] &= (1ull << (n - lo)) - 1ull;
    }
}

// drop the first b bases (b >= 0).  Word selection is written as unrolled
// compare-selects so the arrays stay in registers (no runtime indexing).
template <int W>
__device__ __forceinline__ void pseq_shift(PSeq<W> &p, int b)
{
    if (b <= 0) return;
    if constexpr (W == 2) {
        // two words: plain selects (the generic compare-select chains below are turned into indexed local-memory
        // accesses by the compiler, which parks the whole sequence on the stack)
        const bool up = b >= 32;
        const int bs = (b & 31) * 2;
        const u64 lo = up ? p.s[1] : p.s[0], hi = up ? 0ull : p.s[1];
        p.s[0] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
        p.s[1] = up ? 0ull : (hi >> bs);
        p.nm[0] = b >= 64 ? 0ull : (p.nm[0] >> b);
        return;
    }
    {
        int ws = b >> 5, bs = (b & 31) * 2;
        u64 t[W];
#pragma unroll
        for (int w = 0; w < W; w++) {
            u64 lo = 0ull, hi = 0ull;
#pragma unroll
            for (int a = 0; a < W; a++) {
                if (a == w + ws) lo = p.s[a];
                if (a == w + ws + 1) hi = p.s[a];
            }
            t[w] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
        }
#pragma unroll
        for (int w = 0; w < W; w++) p.s[w] = t[w];
    }
    {
        const int NW = W / 2;
        int ws = b >> 6, bs = b & 63;
        u64 t[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) {
            u64 lo = 0ull, hi = 0ull;
#pragma unroll
            for (int a = 0; a < NW; a++) {
                if (a == w + ws) lo = p.nm[a];
                if (a == w + ws + 1) hi = p.nm[a];
            }
            t[w] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
        }
#pragma unroll
        for (int w = 0; w < NW; w++) p.nm[w] = t[w];
    }
}

// 64-bit hash of (packed key, ambiguity mask, length); identical on host and device.
__host__ __device__ __forceinline__ u64 mix64(u64 x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
// Hash of a key: only the words its length occupies take part, so W=2 and W=4
// contexts (and the host) agree on the hash -- and the owner -- of a sequence.
__host__ __device__ __forceinline__ u64 key_hash_words(const u64 *s, int W, const u64 *nm, int len)
{
    u64 h = 0x9E3779B97F4A7C15ULL ^ ((u64)len * 0xD6E8FEB86659FD93ULL);
    const int nb = (len + 31) >> 5, nn = (len + 63) >> 6;
#pragma unroll
    for (int w = 0; w < 4; w++)
        if (w < W && w < nb) { h = (h ^ s[w]) * 0xFF51AFD7ED558CCDULL; h ^= h >> 32; }
#pragma unroll
    for (int w = 0; w < 2; w++)
        if (w < W / 2 && w < nn) { h = (h ^ nm[w]) * 0xC4CEB9FE1A85EC53ULL; h ^= h >> 29; }
    return mix64(h);
}
// Owner rank of a key in the multi-GPU exchange = owner_hash(h) % n_ranks; re-mixed so
// it is independent of the slot bits (low) and the fingerprint bits (high) of h.
__host__ __device__ __forceinline__ u64 owner_hash(u64 h) { return mix64(h ^ 0xA5A5F00DDEADBEEFULL); }
template <int W>
__device__ __forceinline__ u64 pseq_hash(const PSeq<W> &p, int len)
{
    return key_hash_words(p.s, W, p.nm, len);
}

// bit i of x -> bit 2i of the result
__host__ __device__ __forceinline__ u64 spread32(u32 x)
{
    u64 v = x;
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

// 64-bit window (32 bases) of a W-word 2-bit array starting at base offset `so`
template <int W>
__host__ __device__ __forceinline__ u64 words_window64(const u64 *a, int so)
{
    const int ws = so >> 5, bs = (so & 31) * 2;
    if constexpr (W == 2) {     // plain selects keep the words in registers (see pseq_shift)
        const u64 a0 = a[0], a1 = a[1];
        const u64 lo2 = ws == 0 ? a0 : (ws == 1 ? a1 : 0ull), hi2 = ws == 0 ? a1 : 0ull;
        return bs ? ((lo2 >> bs) | (hi2 << (64 - bs))) : lo2;
    }
    u64 lo = 0ull, hi = 0ull;
#pragma unroll
    for (int w = 0; w < W; w++) {
        if (w == ws) lo = a[w];
        if (w == ws + 1) hi = a[w];
    }
    return bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
}

// 32-bit window of a 1-bit-per-base mask (NW words) starting at base offset `so`
template <int NW>
__host__ __device__ __forceinline__ u32 mask_window32(const u64 *a, int so)
{
    const int ws = so >> 6, bs = so & 63;
    u64 lo = 0ull, hi = 0ull;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        if (w == ws) lo = a[w];
        if (w == ws + 1) hi = a[w];
    }
    return (u32)(bs ? ((lo >> bs) | (hi << (64 - bs))) : lo);
}

// ---------------------------------------------------------------------------
// Device work queues (K1 class queues, K2 to-do lists).  A warp reserves WARPQ_CHUNK consecutive entries with one atomic and hands
// them out to its lanes; what it does not use is padded with WARPQ_NONE (consumers skip it).
// Atomics on the three fill counters -- one L2 sector -- would otherwise serialise the
// whole kernel (same-address atomics retire at about one per nanosecond).
#define WARPQ_CHUNK 128
#define WARPQ_NONE 0xFFFFFFFFu
struct WarpQ { u64 pos; u32 left; };       // warp-uniform

template <typename E>
__device__ __forceinline__ u64 warpq_slot(bool push, u64 *fill, E *q, WarpQ &wq)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned b = __ballot_sync(0xffffffffu, push);
    const u32 c = __popc(b);
    if (!c) return 0;
    const u32 rank = __popc(b & ((1u << lane) - 1u));
    if (c <= wq.left) {
        const u64 slot = wq.pos + rank;
        wq.pos += c; wq.left -= c;
        return slot;
    }
    // the old chunk takes the first `left` entries, a fresh chunk the rest
    u64 nb = 0;
    if (lane == 0) nb = atomicAdd(fill, (u64)WARPQ_CHUNK);
    nb = __shfl_sync(0xffffffffu, nb, 0);
    const u32 spill = c - wq.left;
    const u64 slot = rank < wq.left ? wq.pos + rank : nb + (rank - wq.left);
    wq.pos = nb + spill; wq.left = WARPQ_CHUNK - spill;
    return slot;
}
template <typename E>
__device__ __forceinline__ void warpq_close(E *q, const WarpQ &wq)
{
    for (u32 x = threadIdx.x & 31; x < wq.left; x += 32) q[wq.pos + x] = (E)~(E)0;
}

// ---------------------------------------------------------------------------
// block-wide exclusive scan of one u32 per thread (blockDim.x multiple of 32, <= 1024)
// ---------------------------------------------------------------------------
__device__ __forceinline__ u32 warp_incl_scan(u32 v)
{
    unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += t;
    }
    return v;
}
// returns exclusive prefix; *total = block sum.  sh must hold 33 u32.
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *sh, u32 *total)
{
    unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    u32 inc = warp_incl_scan(v);
    if (lane == 31) sh[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        u32 x = (lane < nw) ? sh[lane] : 0;
        u32 xi = warp_incl_scan(x);
        sh[lane] = xi - x;
        if (lane == 31) sh[32] = xi;
    }
    __syncthreads();
    u32 r = inc - v + sh[wid];
    *total = sh[32];
    __syncthreads();
    return r;
}
