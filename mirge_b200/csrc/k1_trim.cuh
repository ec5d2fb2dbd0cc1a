// k1_trim.cuh -- K1 building blocks: the unique-sequence table (warp-collective upsert,
// count cache), cutadapt's literal banded DP, the multi-GPU exchange records.
//
// Replaces, per read, the Worker.run modifier chain of utils/trim_file.py:52-55
// (cutadapt AdapterCutter / UnconditionalCutter + min_length) and the dict
// upsert of utils/quantReads.py:9-21.  Input is the packed record stream K0
// wrote (bases via one 128-bit load per read at W=2, ambiguity mask, trim
// window); output is the device hash table of unique sequences with per-sample
// counts, the read-length histogram and the kept counter.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------
// unique-sequence table: open addressing, linear probing, key stored inline.
// tag = fp23 << 9 | len << 1 | ready ; 0 = empty.
// ---------------------------------------------------------------------------
template <int W> struct TEntry;
template <> struct alignas(32) TEntry<2> { u64 s[2]; u64 nm[1]; u32 tag; u32 id; };
template <> struct alignas(64) TEntry<4> { u64 s[4]; u64 nm[2]; u32 tag; u32 id; u64 pad; };

template <int W>
struct TableView {
    TEntry<W> *slots;
    u64 mask;            // capacity - 1 (capacity is a power of two)
    int shift;           // 64 - log2(capacity): the home slot is the TOP bits of the hash, so that a range of hash
                         // values (a bucket of the windowed upsert, k1_kernel.cuh) is a contiguous window of the table
    u64 *useq;           // dense copies by unique id: [max_unique][W]
    u64 *unmask;         // [max_unique][W/2]
    u8 *ulen;            // [max_unique]
    u32 *counts;         // [S][max_unique] (column per sample: a rank's own column stays dense however many samples the job has)
    u64 max_unique;
    u64 *n_unique;       // device counter
    u32 *err;            // device error flags
    int S;
};

// Table slots are written and read concurrently by different SMs.  Publication
// follows the PTX memory model's release/acquire pattern: the inserting thread
// writes the key, then st.release.gpu's the tag; probing threads ld.acquire.gpu
// the tag and read the key with strong (relaxed.gpu) loads, so a READY tag
// implies the key words are visible.
__device__ __forceinline__ void ldcg_u64x2(const void *p, u64 &a, u64 &b)
{
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ u64 ldcg_u64(const void *p)
{
    u64 a;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return a;
}
__device__ __forceinline__ u32 ld_acquire_u32(const u32 *p)
{
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(u32 *p, u32 v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// Key compare of a published entry; the id rides along in the same round trip (it shares a 16-byte unit with the
// tag), so a hit costs the tag load plus ONE more access to the sector.
template <int W>
__device__ __forceinline__ bool entry_key_equal(const TEntry<W> *e, const PSeq<W> &k, u32 &id_out)
{
    bool eq = true;
#pragma unroll
    for (int w = 0; w < W; w += 2) {
        u64 a, b;
        ldcg_u64x2(&e->s[w], a, b);
        eq = eq && (a == k.s[w]) && (b == k.s[w + 1]);
    }
    if (W == 2) {
        u64 c, d;
        ldcg_u64x2(&e->nm[0], c, d);            // nm | tag, id
        eq = eq && (c == k.nm[0]);
        id_out = (u32)(d >> 32);
    } else {
        u64 c, d, t, pad;
        ldcg_u64x2(&e->nm[0], c, d);
        ldcg_u64x2(&e->tag, t, pad);
        eq = eq && (c == k.nm[0]) && (d == k.nm[W / 2 - 1]);
        id_out = (u32)(t >> 32);
    }
    return eq;
}

// Warp-collective find-or-insert + count (the collapse of quantReads.py:9-16): every
// lane of a converged warp calls it; lanes with `active` hold a key to be counted once.
//   1. duplicates inside the warp are folded first (__match_any_sync on the 64-bit hash,
//      keys compared with the group leader's): one lane per distinct key probes, and adds
//      the group's size -- hot sequences cost one global atomic per warp, not per read;
//   2. leaders probe until they hit, claim an empty slot, or meet a slot with their
//      fingerprint that is still being published (they retry it after step 3);
//   3. claiming lanes take their dense ids with one atomic, write their keys, execute ONE
//      fence for the warp and publish their tags (the release side of the protocol above).
// Pull the home slot of hash h into L2 ahead of the probe.
template <int W>
__device__ __forceinline__ void table_prefetch(const TableView<W> &T, u64 h)
{
    asm volatile("prefetch.global.L2 [%0];" :: "l"(T.slots + (h >> T.shift)));
}

// FOLD = false (exchange import: records of one sender are distinct): no in-warp folding.  (Opening the probe with
// the CAS instead of a load was measured slower: 1.85 vs 1.37 ms for 7.9 M inserts.)
template <int W, bool FOLD = true>
__device__ __forceinline__ void table_find_warp(const TableView<W> &T, const PSeq<W> &key, int len, bool active, u64 h,
                                                bool &lead_out, u32 &id_out, u32 &cnt_out)
{
    const unsigned lane = threadIdx.x & 31, below = (1u << lane) - 1u;
    if (!active) h = 0ull;
    lead_out = false; id_out = 0xFFFFFFFFu; cnt_out = 0;
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (!act) return;
    bool lead = active;
    u32 cnt = 1u;
    if (FOLD) {
        const unsigned grp = __match_any_sync(0xffffffffu, h) & act;
        const int leader = active ? (__ffs(grp) - 1) : (int)lane;
        bool same = (__shfl_sync(0xffffffffu, len, leader) == len);
#pragma unroll
        for (int w = 0; w < W; w++) same = same && (__shfl_sync(0xffffffffu, key.s[w], leader) == key.s[w]);
#pragma unroll
        for (int w = 0; w < W / 2; w++) same = same && (__shfl_sync(0xffffffffu, key.nm[w], leader) == key.nm[w]);
        const unsigned folded = __ballot_sync(0xffffffffu, active && same);
        // a 64-bit hash collision between different keys of one warp: the odd lane inserts on its own
        lead = active && ((int)lane == leader || !same);
        cnt = ((int)lane == leader) ? (u32)__popc(grp & folded) : 1u;
    }
    const u32 want = (((((u32)h) & 0x7FFFFFu) | 1u) << 9) | ((u32)len << 1);   // busy form (ready bit clear); fingerprint = low hash bits
    u64 slot = h >> T.shift;
    u64 probes = 0;
    int state = lead ? 1 : 0;           // 1 probing, 2 claimed a slot, 0 done
    u32 id = 0xFFFFFFFFu;
    for (;;) {
        if (state == 1) {
            for (;;) {
                if (probes > T.mask) { atomicOr(T.err, ERRF_FULL); state = 0; break; }
                TEntry<W> *e = T.slots + slot;
                u32 t = ld_acquire_u32(&e->tag);
                if (t == 0) {
                    const u32 old = atomicCAS(&e->tag, 0u, want);
                    if (old == 0) { state = 2; break; }
                    t = old;
                }
                if ((t | 1u) == (want | 1u)) {
                    if (!(t & 1u)) break;                               // unpublished: look again after this warp's own publishes
                    u32 eid;
                    if (entry_key_equal<W>(e, key, eid)) {
                        id = eid;
                        state = 0;
                        break;
                    }
                }
                slot = (slot + 1) & T.mask;
                probes++;
            }
        }
        const unsigned cl = __ballot_sync(0xffffffffu, state == 2);
        if (cl) {
            const int first = __ffs(cl) - 1;
            u64 base = 0;
            if ((int)lane == first) base = atomicAdd(T.n_unique, (u64)__popc(cl));
            base = __shfl_sync(0xffffffffu, base, first);
            if (state == 2) {
                const u64 nid = base + __popc(cl & below);
                TEntry<W> *e = T.slots + slot;
                if (nid < T.max_unique) id = (u32)nid; else atomicOr(T.err, ERRF_FULL);
#pragma unroll
                for (int w = 0; w < W; w++) e->s[w] = key.s[w];
#pragma unroll
                for (int w = 0; w < W / 2; w++) e->nm[w] = key.nm[w];
                e->id = id;
                // release fence: __threadfence() is membar.gl = fence.sc.gpu, a sequentially consistent fence the protocol
                // does not need (it was 15 stall cycles per issued instruction in an upsert-only kernel)
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(&e->tag), "r"(want | 1u) : "memory");   // publish
                if (id != 0xFFFFFFFFu) {                                // dense copies: read by later kernels only
#pragma unroll
                    for (int w = 0; w < W; w++) T.useq[nid * W + w] = key.s[w];
#pragma unroll
                    for (int w = 0; w < W / 2; w++) T.unmask[nid * (W / 2) + w] = key.nm[w];
                    T.ulen[nid] = (u8)len;
                }
                state = 0;
            }
        }
        if (!__any_sync(0xffffffffu, state == 1)) break;
    }
    lead_out = lead && id != 0xFFFFFFFFu; id_out = id; cnt_out = cnt;
}

template <int W>
__device__ __forceinline__ void table_add_warp(const TableView<W> &T, const PSeq<W> &key, int len, bool active, int sample, u64 h)
{
    bool lead; u32 id, cnt;
    table_find_warp<W>(T, key, len, active, h, lead, id, cnt);
    if (lead) atomicAdd(&T.counts[(u64)sample * T.max_unique + id], cnt);
}

// Per-warp write-back cache of pending count increments (shared memory, direct mapped
// by id): the few very abundant sequences of a small-RNA sample would otherwise send
// millions of atomics to the same few addresses.  cid/ccnt: this warp's K1_CC entries.
#define K1_CC 128
template <int W>
__device__ __forceinline__ void count_cached(u32 *cid, u32 *ccnt, const TableView<W> &T, int sample, bool lead, u32 id, u32 cnt)
{
    const unsigned lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(0xffffffffu, lead);
    const u32 slot = id & (K1_CC - 1);
    while (todo) {
        const bool pend = (todo >> lane) & 1u;
        const unsigned grp = __match_any_sync(0xffffffffu, pend ? slot : (u32)(K1_CC + lane));
        const bool act = pend && ((unsigned)(__ffs(grp) - 1) == lane);       // one lane per cache entry and round
        if (act) {
            const u32 c = cid[slot];
            if (c == id) ccnt[slot] += cnt;
            else {
                const u32 oc = ccnt[slot];
                if (oc) atomicAdd(&T.counts[(u64)sample * T.max_unique + c], oc);
                cid[slot] = id; ccnt[slot] = cnt;
            }
        }
        __syncwarp();
        todo &= ~__ballot_sync(0xffffffffu, act);
    }
}
template <int W>
__device__ __forceinline__ void count_cache_flush(u32 *cid, u32 *ccnt, const TableView<W> &T, int sample)
{
    __syncwarp();
    for (int b = threadIdx.x & 31; b < K1_CC; b += 32)
        if (ccnt[b]) atomicAdd(&T.counts[(u64)sample * T.max_unique + cid[b]], ccnt[b]);
}

// ---------------------------------------------------------------------------
// adapter matching
// ---------------------------------------------------------------------------
struct TrimParams {
    int mode;            // 0 none, 1 3' adapters, 2 unconditional cut (+N)
    int n_adapters;
    int uncond;
    int min_overlap;
    int min_length;
    int pad_;
    double error_rate;
    int alen[MIRGE_MAX_ADAPTERS];
    u64 aw[MIRGE_MAX_ADAPTERS][2];     // packed adapter bases
    u64 peq[MIRGE_MAX_ADAPTERS][4];    // bit i set iff adapter[i] == base
};

// cutadapt _align.pyx Aligner.locate for a BACK adapter, cell for cell
// (see oracle/mirge_oracle.c:orc_locate).  Column entries are packed as
// cost | matches << 8 | origin << 16.
template <int W>
__host__ __device__ bool locate_back(const PSeq<W> &p, int n, const u64 *aw, int m, double rate, int min_overlap,
                            int &rstart, int &matches_out)
{
    u32 col[MIRGE_MAX_ADAPTER_LEN + 1];
    const int k = (int)(rate * m);
    for (int i = 0; i <= m; i++) col[i] = (u32)i;
    int best_cost = m + n, best_matches = 0, best_origin = 0;
    int last = m < k + 1 ? m : k + 1;
    for (int j = 1; j <= n; j++) {
        const int cj = pseq_isn<W>(p, j - 1) ? 4 : pseq_base<W>(p, j - 1);
        u32 tmp = col[0];
        u32 prev_new = (u32)j << 16;
        col[0] = prev_new;
        for (int i = 1; i <= last; i++) {
            const int ai = (int)((aw[(i - 1) >> 5] >> (((i - 1) & 31) * 2)) & 3ull);
            const u32 cur = col[i];
            u32 e;
            if (ai == cj) {
                e = tmp + (1u << 8);
            } else {
                const u32 cd = (tmp & 0xFF) + 1, cdel = (cur & 0xFF) + 1, cins = (prev_new & 0xFF) + 1;
                if (cd <= cdel && cd <= cins) e = tmp + 1;
                else if (cins <= cdel) e = prev_new + 1;
                else e = cur + 1;
            }
            tmp = cur;
            col[i] = e;
            prev_new = e;
        }
        while (last >= 0 && (int)(col[last] & 0xFF) > k) last--;
        if (last < m) {
            last++;
        } else {
            const u32 e = col[m];
            const int cost = e & 0xFF, matches = (e >> 8) & 0xFF;
            if (m >= min_overlap && (double)cost <= (double)m * rate &&
                (matches > best_matches || (matches == best_matches && cost < best_cost))) {
                best_matches = matches; best_cost = cost; best_origin = (int)(e >> 16);
                if (cost == 0 && matches == m) break;
            }
        }
    }
    for (int i = m; i >= 0; i--) {
        const u32 e = col[i];
        const int cost = e & 0xFF, matches = (e >> 8) & 0xFF;
        if (i >= min_overlap && (double)cost <= (double)i * rate &&
            (matches > best_matches || (matches == best_matches && cost < best_cost))) {
            best_matches = matches; best_cost = cost; best_origin = (int)(e >> 16);
        }
    }
    if (best_cost == m + n) return false;
    rstart = best_origin; matches_out = best_matches;
    return true;
}

// Adapter.match_to (cutadapt adapters.py), where = BACK, no wildcards:
// leftmost exact occurrence (str.find) first, else Aligner.locate.
template <int W>
__host__ __device__ __forceinline__ bool match_back(const PSeq<W> &p, int n, const TrimParams &tp, int a,
                                           int &rstart, int &matches)
{
    const int m = tp.alen[a];
    if (m <= n) {
        u64 D = 0;
        const u64 top = 1ull << (m - 1);
        for (int j = 0; j < n; j++) {
            const u64 pe = pseq_isn<W>(p, j) ? 0ull : tp.peq[a][pseq_base<W>(p, j)];
            D = ((D << 1) | 1ull) & pe;
            if (D & top) { rstart = j - m + 1; matches = m; return true; }
        }
    }
    return locate_back<W>(p, n, tp.aw[a], m, tp.error_rate, tp.min_overlap, rstart, matches);
}

// (Loading the next record and prefetching its home slot while the current one is probed was measured no faster:
// 1.54 vs 1.37 ms for 7.9 M inserts, 47 vs 32 registers.)
// Exchange records (multi-GPU): W + W/2 u64 key words, then u32 (len | sample << 8), then the u32
// read count of that sample -- 32 bytes at W = 2, one DRAM sector.  A sequence seen in k samples of the
// sending rank travels as k records.
template <int W>
__global__ void __launch_bounds__(128)
k1_import_records(const unsigned char *__restrict__ buf, u64 n_records, int rec_bytes, TableView<W> T)
{
    for (u64 r0 = (u64)blockIdx.x * blockDim.x; r0 < n_records; r0 += (u64)gridDim.x * blockDim.x) {
        const u64 r = r0 + threadIdx.x;
        const bool on = r < n_records;
        PSeq<W> p;
#pragma unroll
        for (int w = 0; w < W; w++) p.s[w] = 0;
#pragma unroll
        for (int w = 0; w < W / 2; w++) p.nm[w] = 0;
        int len = 0, smp = 0;
        u32 c = 0;
        if (on) {
            const u64 *q = reinterpret_cast<const u64 *>(buf + r * (u64)rec_bytes);
            if (W == 2) {
                const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(q), b = *reinterpret_cast<const ulonglong2 *>(q + 2);
                p.s[0] = a.x; p.s[1] = a.y; p.nm[0] = b.x;
                len = (int)(b.y & 0xFFull); smp = (int)((b.y >> 8) & 0xFFFFFFull); c = (u32)(b.y >> 32);
            } else {
#pragma unroll
                for (int w = 0; w < W; w++) p.s[w] = q[w];
#pragma unroll
                for (int w = 0; w < W / 2; w++) p.nm[w] = q[W + w];
                const u64 t = q[W + W / 2];
                len = (int)(t & 0xFFull); smp = (int)((t >> 8) & 0xFFFFFFull); c = (u32)(t >> 32);
            }
        }
        // records of one sender and sample are distinct; the same key arrives from several senders / samples
        bool lead; u32 id, cnt;
        table_find_warp<W, false>(T, p, len, on, on ? pseq_hash<W>(p, len) : 0ull, lead, id, cnt);
        if (lead && c && smp < T.S) atomicAdd(&T.counts[(u64)smp * T.max_unique + id], c);
    }
}

// Write every (unique, sample) pair with a non-zero count as an exchange record, grouped by owner =
// hash % n_parts.  Pass 1 (count_only) fills part_counts; after an exclusive scan into part_cursor,
// pass 2 scatters.  The number of uniques is read on the device (no host round trip).
template <int W>
__global__ void __launch_bounds__(128)
k1_export_records(TableView<W> T, int n_parts, int rec_bytes, int count_only, u64 cap_records,
                  u64 *__restrict__ part_counts, u64 *__restrict__ part_cursor, unsigned char *__restrict__ out)
{
    u64 n_unique = *T.n_unique;
    if (n_unique > T.max_unique) n_unique = T.max_unique;
    const int lane = threadIdx.x & 31;
    for (u64 id0 = (u64)blockIdx.x * blockDim.x; id0 < n_unique; id0 += (u64)gridDim.x * blockDim.x) {
        const u64 id = id0 + threadIdx.x;
        const bool on = id < n_unique;
        u64 s[W], nm[W / 2];
        int len = 0, part = -1;
        if (on) {
#pragma unroll
            for (int w = 0; w < W; w++) s[w] = T.useq[id * W + w];
#pragma unroll
            for (int w = 0; w < W / 2; w++) nm[w] = T.unmask[id * (W / 2) + w];
            len = T.ulen[id];
            part = (int)(owner_hash(key_hash_words(s, W, nm, len)) % (u64)n_parts);
        }
        for (int x = 0; x < T.S; x++) {                      // warp-uniform trip count
            const u32 c = on ? T.counts[(u64)x * T.max_unique + id] : 0u;
            // few destinations, many records: lanes bound for the same rank share one atomic
            const unsigned peers = __match_any_sync(0xffffffffu, c ? part : -1 - lane);
            if (!c) continue;
            const int leader = __ffs(peers) - 1;
            const u32 rank = __popc(peers & ((1u << lane) - 1u));
            if (count_only) { if (lane == leader) atomicAdd(&part_counts[part], (u64)__popc(peers)); continue; }
            u64 dst = 0;
            if (lane == leader) dst = atomicAdd(&part_cursor[part], (u64)__popc(peers));
            dst = __shfl_sync(peers, dst, leader) + rank;
            if (dst >= cap_records) { atomicOr(T.err, ERRF_FULL); continue; }
            u64 *q = reinterpret_cast<u64 *>(out + dst * (u64)rec_bytes);
            const u64 tail = (u64)len | ((u64)x << 8) | ((u64)c << 32);
            if (W == 2) {
                *reinterpret_cast<ulonglong2 *>(q) = make_ulonglong2(s[0], s[1]);
                *reinterpret_cast<ulonglong2 *>(q + 2) = make_ulonglong2(nm[0], tail);
            } else {
#pragma unroll
                for (int w = 0; w < W; w++) q[w] = s[w];
#pragma unroll
                for (int w = 0; w < W / 2; w++) q[W + w] = nm[w];
                q[W + W / 2] = tail;
            }
        }
    }
}

// counts[*][id] = 0 for the ids in use (cheaper than a memset of the whole array when S is large)
__global__ void k1_clear_counts(u32 *__restrict__ counts, const u64 *__restrict__ n_unique_ptr, u64 max_unique, int S)
{
    u64 n = *n_unique_ptr;
    if (n > max_unique) n = max_unique;
    for (int s = 0; s < S; s++)
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) counts[(u64)s * max_unique + i] = 0;
}
