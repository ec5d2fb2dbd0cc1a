// scan.cuh -- multi-level exclusive scan of a u32 array (hand-written; used for
// newline ranks in K0 and for k-mer bucket offsets at library load).
#pragma once
#include "common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_BLOCK (SCAN_THREADS * SCAN_ITEMS)

__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_sums(const u32 *__restrict__ in, u64 n, u32 *__restrict__ block_sums)
{
    __shared__ u32 sh[33];
    u64 base = (u64)blockIdx.x * SCAN_BLOCK + (u64)threadIdx.x * SCAN_ITEMS;
    u32 v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) v += in[base + k];
    u32 total;
    block_excl_scan(v, sh, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// out[i] = block_offsets[block] + exclusive prefix within the block (in == out allowed)
__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_apply(const u32 *in, u32 *out, u64 n, const u32 *__restrict__ block_offsets)
{
    __shared__ u32 sh[33];
    u64 base = (u64)blockIdx.x * SCAN_BLOCK + (u64)threadIdx.x * SCAN_ITEMS;
    u32 x[SCAN_ITEMS];
    u32 v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { x[k] = (base + k < n) ? in[base + k] : 0; v += x[k]; }
    u32 total;
    u32 pre = block_excl_scan(v, sh, &total) + (block_offsets ? block_offsets[blockIdx.x] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = pre; pre += x[k]; }
}

static inline u64 scan_scratch_elems(u64 n)
{
    u64 tot = 0;
    while (n > 1) { n = (n + SCAN_BLOCK - 1) / SCAN_BLOCK; tot += n + 1; if (n == 1) break; }
    return tot + 2;
}

// Exclusive scan of in[0..n) into out[0..n).  scratch must hold scan_scratch_elems(n) u32.
// Returns the number of kernels launched.
static int device_excl_scan(const u32 *in, u32 *out, u64 n, u32 *scratch, cudaStream_t st)
{
    if (n == 0) return 0;
    u64 nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (nb == 1) {
        scan_block_apply<<<1, SCAN_THREADS, 0, st>>>(in, out, n, nullptr);
        return 1;
    }
    u32 *sums = scratch;
    scan_block_sums<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, sums);
    int k = device_excl_scan(sums, sums, nb, scratch + nb + 1, st);
    scan_block_apply<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, sums);
    return k + 2;
}
