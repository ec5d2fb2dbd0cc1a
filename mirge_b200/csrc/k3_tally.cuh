// k3_tally.cuh -- K3: per-library and per-miRNA tallies.
//
// Replaces the seqKey x sample double loop of utils/summarize.py:34-66:
// trimmedUniq, the fixed-priority per-library read totals and mirDic
// 'quant' / 'iscan'.  A unique sequence is annotated by at most one round
// (annotated sequences leave the pool, runAnnotationPipeline.py:546), so the
// if/elif chain collapses to a switch on round_of.
#pragma once
#include "common.cuh"

// lib_totals columns (per sample)
#define TOT_UNIQ 0
#define TOT_MIRNA 1
#define TOT_SPIKE 9
#define TOT_REM 10

__device__ __forceinline__ int totals_column(int ro)
{
    // ro: 0 none, 1 exact miRNA, 2 hairpin, 3 mature tRNA, 4 pre-tRNA, 5 snoRNA,
    //     6 rRNA, 7 ncRNA others, 8 mRNA, 9 isomiR, 10 spike-in
    if (ro == 1 || ro == 9) return TOT_MIRNA;
    if (ro >= 2 && ro <= 8) return ro;
    if (ro == 10) return TOT_SPIKE;
    return TOT_REM;
}

// ---------------------------------------------------------------------------
// Atomic-free segmented reduction.
//
// Every unique gets a segment key: its miRNA reference index for rounds 1/9,
// else n_mir + totals column.  Keys are ordered with a stable LSD radix sort
// whose histograms and ranks come from warp match/ballot operations (no
// atomics), segment bounds are written by the unique thread that sees a key
// change, and one warp reduces each segment with shuffles.
// ---------------------------------------------------------------------------
#define RS_THREADS 256
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)

__global__ void __launch_bounds__(256)
k3_make_keys(const signed char *__restrict__ round_of, const int *__restrict__ ref_of, const u64 *__restrict__ n_unique_ptr,
             int n_mir, u32 *__restrict__ keys, u32 *__restrict__ vals)
{
    const u64 U = *n_unique_ptr;
    for (u64 id = (u64)blockIdx.x * blockDim.x + threadIdx.x; id < U; id += (u64)gridDim.x * blockDim.x) {
        const int ro = round_of[id];
        keys[id] = (ro == 1 || ro == 9) ? (u32)ref_of[id] : (u32)(n_mir + totals_column(ro));
        vals[id] = (u32)id;
    }
}

// digit histogram of one tile per block -> hist[digit * n_tiles + tile]
__global__ void __launch_bounds__(RS_THREADS)
k3_rs_hist(const u32 *__restrict__ keys, const u64 *__restrict__ n_ptr, int shift, u32 n_tiles, u32 *__restrict__ hist)
{
    __shared__ u32 wh[RS_THREADS / 32][256];
    const u64 n = *n_ptr;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (u32 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int b = lane; b < 256; b += 32) wh[wid][b] = 0;
        __syncwarp();
        const u64 base = (u64)tile * RS_TILE;
        for (int it = 0; it < RS_ITEMS; it++) {
            const u64 i = base + (u64)it * RS_THREADS + threadIdx.x;
            const bool ok = i < n;
            const u32 d = ok ? ((keys[i] >> shift) & 255u) : 256u + lane;   // inactive lanes never group
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            if (ok && lane == (unsigned)(__ffs(peers) - 1)) wh[wid][d] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        {
            u32 tot = 0;
#pragma unroll
            for (int w = 0; w < RS_THREADS / 32; w++) tot += wh[w][threadIdx.x];
            if ((u64)tile * RS_TILE < n) hist[(u64)threadIdx.x * n_tiles + tile] = tot;
            else hist[(u64)threadIdx.x * n_tiles + tile] = 0;
        }
        __syncthreads();
    }
}

// stable scatter using the scanned histogram
__global__ void __launch_bounds__(RS_THREADS)
k3_rs_scatter(const u32 *__restrict__ keys, const u32 *__restrict__ vals, const u64 *__restrict__ n_ptr, int shift,
              u32 n_tiles, const u32 *__restrict__ hist_scanned, u32 *__restrict__ keys_out, u32 *__restrict__ vals_out)
{
    __shared__ u32 run[256];                       // next free slot per digit for this tile
    __shared__ u32 wcnt[RS_THREADS / 32][256];     // per-warp digit counts of the current item row
    const u64 n = *n_ptr;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (u32 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        run[threadIdx.x] = hist_scanned[(u64)threadIdx.x * n_tiles + tile];
        __syncthreads();
        const u64 base = (u64)tile * RS_TILE;
        for (int it = 0; it < RS_ITEMS; it++) {
            for (int b = lane; b < 256; b += 32) wcnt[wid][b] = 0;
            __syncwarp();
            const u64 i = base + (u64)it * RS_THREADS + threadIdx.x;
            const bool ok = i < n;
            const u32 key = ok ? keys[i] : 0u;
            const u32 d = ok ? ((key >> shift) & 255u) : 256u + lane;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const u32 rank_in_warp = __popc(peers & ((1u << lane) - 1u));
            if (ok && rank_in_warp == 0) wcnt[wid][d] = __popc(peers);
            __syncthreads();
            u32 before = 0;
            if (ok) {
                for (unsigned w = 0; w < wid; w++) before += wcnt[w][d];
                const u32 dst = run[d] + before + rank_in_warp;
                keys_out[dst] = key;
                vals_out[dst] = vals[i];
            }
            __syncthreads();
            {
                u32 tot = 0;
#pragma unroll
                for (int w = 0; w < RS_THREADS / 32; w++) tot += wcnt[w][threadIdx.x];
                run[threadIdx.x] += tot;
            }
            __syncthreads();
        }
    }
}

// seg_start[k] = first sorted index holding key k; seg_end[k] = one past the last
__global__ void __launch_bounds__(256)
k3_seg_bounds(const u32 *__restrict__ keys, const u64 *__restrict__ n_ptr, u32 *__restrict__ seg_start, u32 *__restrict__ seg_end)
{
    const u64 n = *n_ptr;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u32 k = keys[i];
        if (i == 0 || keys[i - 1] != k) seg_start[k] = (u32)i;
        if (i + 1 == n || keys[i + 1] != k) seg_end[k] = (u32)(i + 1);
    }
}

// Segments are cut into chunks of K3_CHUNK sorted entries so that the few huge
// segments (un-annotated, mRNA ...) spread over the whole grid.
#define K3_CHUNK 1024

// chunks per segment (one thread per segment); n_chunks[n_seg] = 0 so that the
// exclusive scan leaves the total there.
__global__ void k3_seg_chunks(const u32 *__restrict__ seg_start, const u32 *__restrict__ seg_end, int n_seg, u32 *__restrict__ n_chunks)
{
    const int seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg < n_seg) n_chunks[seg] = (seg_end[seg] - seg_start[seg] + K3_CHUNK - 1) / K3_CHUNK;
    else if (seg == n_seg) n_chunks[seg] = 0;
}

// One warp per chunk: partial[g][s] = {reads, reads of exact-miRNA uniques
// (round_of == 1), uniques with a non-zero count} over the chunk's entries.
__global__ void __launch_bounds__(256)
k3_chunk_reduce(const u32 *__restrict__ vals, const u32 *__restrict__ seg_start, const u32 *__restrict__ seg_end,
                const u32 *__restrict__ chunk_off, int n_seg, const u32 *__restrict__ counts,
                const signed char *__restrict__ round_of, int S, u64 cstride, i64 *__restrict__ partial)
{
    const unsigned lane = threadIdx.x & 31;
    const u32 total = chunk_off[n_seg];
    const u32 warps = (gridDim.x * blockDim.x) >> 5;
    for (u32 g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < total; g += warps) {
        int lo = 0, hi = n_seg;            // segment with chunk_off[seg] <= g < chunk_off[seg + 1]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (chunk_off[mid] <= g) lo = mid; else hi = mid; }
        const u32 a = seg_start[lo] + (g - chunk_off[lo]) * K3_CHUNK;
        const u32 e = seg_end[lo];
        const u32 b = a + K3_CHUNK < e ? a + K3_CHUNK : e;
        for (int s = 0; s < S; s++) {
            i64 q = 0, qi = 0, qu = 0;
            for (u32 x = a + lane; x < b; x += 32) {
                const u32 id = vals[x];
                const u32 c = counts[(u64)s * cstride + id];             // counts[S][max_unique]
                q += c;
                if (round_of[id] == 1) qi += c;
                qu += (c != 0);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                q += __shfl_down_sync(0xffffffffu, q, d);
                qi += __shfl_down_sync(0xffffffffu, qi, d);
                qu += __shfl_down_sync(0xffffffffu, qu, d);
            }
            if (lane == 0) { i64 *o = partial + ((u64)g * S + s) * 3; o[0] = q; o[1] = qi; o[2] = qu; }
        }
    }
}

// One warp per segment sums its chunks' partials.  seg_q / seg_i / seg_u are [n_seg][S].
__global__ void __launch_bounds__(256)
k3_seg_reduce(const i64 *__restrict__ partial, const u32 *__restrict__ chunk_off, int n_seg, int S,
              i64 *__restrict__ seg_q, i64 *__restrict__ seg_i, i64 *__restrict__ seg_u)
{
    const unsigned lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int seg = blockIdx.x * warps_per_block + (threadIdx.x >> 5); seg < n_seg; seg += gridDim.x * warps_per_block) {
        const u32 a = chunk_off[seg], b = chunk_off[seg + 1];
        for (int s = 0; s < S; s++) {
            i64 q = 0, qi = 0, qu = 0;
            for (u32 g = a + lane; g < b; g += 32) {
                const i64 *o = partial + ((u64)g * S + s) * 3;
                q += o[0]; qi += o[1]; qu += o[2];
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                q += __shfl_down_sync(0xffffffffu, q, d);
                qi += __shfl_down_sync(0xffffffffu, qi, d);
                qu += __shfl_down_sync(0xffffffffu, qu, d);
            }
            if (lane == 0) { seg_q[(u64)seg * S + s] = q; seg_i[(u64)seg * S + s] = qi; seg_u[(u64)seg * S + s] = qu; }
        }
    }
}

// Fold the segment sums into the output tables: one 256-thread block per sample.
__global__ void __launch_bounds__(256)
k3_finalize(const i64 *__restrict__ seg_q, const i64 *__restrict__ seg_i, const i64 *__restrict__ seg_u,
            int n_mir, int S, i64 *__restrict__ mir_quant, i64 *__restrict__ mir_iscan,
            i64 *__restrict__ lib_totals)
{
    __shared__ i64 red[2][8];
    const int s = blockIdx.x;
    i64 uniq = 0, mirna = 0;
    for (int m = threadIdx.x; m < n_mir; m += blockDim.x) {
        const i64 q = seg_q[(u64)m * S + s];
        mir_quant[(u64)m * S + s] = q;
        mir_iscan[(u64)m * S + s] = seg_i[(u64)m * S + s];
        mirna += q;
        uniq += seg_u[(u64)m * S + s];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        uniq += __shfl_down_sync(0xffffffffu, uniq, d);
        mirna += __shfl_down_sync(0xffffffffu, mirna, d);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = uniq; red[1][threadIdx.x >> 5] = mirna; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uniq = 0; mirna = 0;
        for (int w = 0; w < 8; w++) { uniq += red[0][w]; mirna += red[1][w]; }
        i64 *t = lib_totals + (u64)s * 12;
        for (int c = 0; c < 12; c++) { t[c] = seg_q[(u64)(n_mir + c) * S + s]; uniq += seg_u[(u64)(n_mir + c) * S + s]; }
        t[TOT_MIRNA] = mirna;
        t[TOT_UNIQ] = uniq;
    }
}
