// CUDA kernels of the liftover path (sm_100a).  Launch wrappers at the bottom (kernels.hpp).
//
// Per batch:   pair_count -> scan -> pair_fill -> scan -> lift_pairs -> read_finalize -> scan -> emit_records
// Once:        table_count -> scan -> table_fill   (flat ReadToRefTreeMap per contig segment)
//
// This is HBM/latency-bound integer work: no tensor cores.  Grids are sized from the work (>> 148 SMs x resident CTAs
// for real batches); scratch traffic is thread-private and stays in L1/L2.
#include <cuda_runtime.h>

#include <algorithm>

#include "assemble.cuh"
#include "assemble_bam.cuh"
#include "bgzf_store.cuh"
#include "kernels.hpp"
#include "lift_device.cuh"
#include "lift_warp.cuh"
#include "pair_bodies.cuh"
#include "table_warp.cuh"

namespace ptl {

// =================================================================================================== scans
namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t shfl_up_t(uint32_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ uint64_t shfl_up_t(uint64_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ uint2 shfl_up_t(uint2 v, int d) {
    return make_uint2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ uint32_t shfl_down_t(uint32_t v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ uint64_t shfl_down_t(uint64_t v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ uint2 shfl_down_t(uint2 v, int d) {
    return make_uint2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ uint32_t add_t(uint32_t a, uint32_t b) { return a + b; }
__device__ __forceinline__ uint64_t add_t(uint64_t a, uint64_t b) { return a + b; }
__device__ __forceinline__ uint2 add_t(uint2 a, uint2 b) { return make_uint2(a.x + b.x, a.y + b.y); }
template <class T> __device__ __forceinline__ T zero_t();
template <> __device__ __forceinline__ uint32_t zero_t<uint32_t>() { return 0u; }
template <> __device__ __forceinline__ uint64_t zero_t<uint64_t>() { return 0ull; }
template <> __device__ __forceinline__ uint2 zero_t<uint2>() { return make_uint2(0u, 0u); }

// Exclusive scan of one tile per block, in place; block totals to sums[blockIdx] (if sums != nullptr).
// `n_dev` (optional): the live length is min(n, *n_dev + 1); tiles past it are skipped (the pair arrays are sized by
// capacity, the pair count only exists on the device).
template <class T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_kernel(T* a, uint64_t n, T* sums, const unsigned long long* n_dev) {
    __shared__ T warp_tot[SCAN_THREADS / 32];
    if (n_dev) { const uint64_t live = uint64_t(*n_dev) + 1ull; if (live < n) n = live; }
    if (uint64_t(blockIdx.x) * SCAN_TILE >= n) return;
    const uint64_t base = uint64_t(blockIdx.x) * SCAN_TILE + uint64_t(threadIdx.x) * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T run = zero_t<T>();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? a[base + k] : zero_t<T>();
        run = add_t(run, v[k]);
    }
    // inclusive warp scan of per-thread totals
    T inc = run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = shfl_up_t(inc, d);
        if (lane >= d) inc = add_t(inc, o);
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    T warp_off = zero_t<T>();
    T block_tot = zero_t<T>();
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < warp) warp_off = add_t(warp_off, warp_tot[w]);
        block_tot = add_t(block_tot, warp_tot[w]);
    }
    T o = shfl_up_t(inc, 1);
    T excl = add_t(warp_off, lane ? o : zero_t<T>());
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) a[base + k] = excl;
        excl = add_t(excl, v[k]);
    }
    if (sums && threadIdx.x == 0) sums[blockIdx.x] = block_tot;
}

template <class T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(T* a, uint64_t n, const T* sums) {
    const T off = sums[blockIdx.x];
    const uint64_t base = uint64_t(blockIdx.x) * SCAN_TILE + uint64_t(threadIdx.x) * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) a[base + k] = add_t(a[base + k], off);
}

// Second (and last) pass for a moderate number of tiles: every block reduces the totals of the tiles before it
// (L2-resident, <= a few thousand values) instead of waiting for a separate scan-of-sums kernel.
template <class T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_add_reduce_kernel(T* a, uint64_t n, const T* sums, const unsigned long long* n_dev) {
    __shared__ T warp_tot[SCAN_THREADS / 32];
    if (n_dev) { const uint64_t live = uint64_t(*n_dev) + 1ull; if (live < n) n = live; }
    if (uint64_t(blockIdx.x) * SCAN_TILE >= n) return;
    T acc = zero_t<T>();
    for (uint32_t i = threadIdx.x; i < blockIdx.x; i += SCAN_THREADS) acc = add_t(acc, sums[i]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        T o = shfl_down_t(acc, d);
        acc = add_t(acc, o);
    }
    if (lane == 0) warp_tot[warp] = acc;
    __syncthreads();
    T off = zero_t<T>();
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) off = add_t(off, warp_tot[w]);
    const uint64_t base = uint64_t(blockIdx.x) * SCAN_TILE + uint64_t(threadIdx.x) * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) a[base + k] = add_t(a[base + k], off);
}

}  // namespace

template <class T>
void exclusive_scan_inplace(T* a, uint64_t n, void* tmp, size_t tmp_bytes, cudaStream_t st, uint64_t* launches,
                            const unsigned long long* n_dev) {
    if (n == 0) return;
    const uint64_t blocks = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (blocks == 1) {
        scan_tile_kernel<T><<<1, SCAN_THREADS, 0, st>>>(a, n, nullptr, n_dev);
        ++*launches;
        return;
    }
    T* sums = static_cast<T*>(tmp);
    const size_t used = ((blocks * sizeof(T)) + 255) & ~size_t(255);
    scan_tile_kernel<T><<<unsigned(blocks), SCAN_THREADS, 0, st>>>(a, n, sums, n_dev);
    ++*launches;
    if (blocks <= 8192) {
        scan_add_reduce_kernel<T><<<unsigned(blocks), SCAN_THREADS, 0, st>>>(a, n, sums, n_dev);
        ++*launches;
        return;
    }
    // (very large inputs: classic three-pass scan over the full capacity; n_dev is not needed for correctness)
    exclusive_scan_inplace<T>(sums, blocks, static_cast<char*>(tmp) + used, tmp_bytes - used, st, launches, nullptr);
    scan_add_kernel<T><<<unsigned(blocks), SCAN_THREADS, 0, st>>>(a, n, sums);
    ++*launches;
}
template void exclusive_scan_inplace<uint32_t>(uint32_t*, uint64_t, void*, size_t, cudaStream_t, uint64_t*, const unsigned long long*);
template void exclusive_scan_inplace<uint64_t>(uint64_t*, uint64_t, void*, size_t, cudaStream_t, uint64_t*, const unsigned long long*);
template void exclusive_scan_inplace<uint2>(uint2*, uint64_t, void*, size_t, cudaStream_t, uint64_t*, const unsigned long long*);

size_t scan_tmp_bytes(uint64_t n) {
    size_t total = 0;
    while (n > uint64_t(SCAN_TILE)) {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        total += ((n * 16) + 255) & ~size_t(255);
    }
    return total + 256;
}

// =================================================================================================== segment tables
// a7: get_read_segment_to_ref_pos_tree_map (read_to_ref_map.rs:101-137), ONE WARP per segment, lanes over CIGAR ops
// (contig->reference CIGARs have 10^3 - 10^6 ops; a thread per segment walked the longest one alone).  Same table as
// pair_bodies.cuh: table_build_body, which stays the scalar statement of the algorithm (host emulation, tests/emul).
// Per round of 32 ops: warp scans give every op its read / reference position and the length of the match stretch in
// front of it; a non-match op behind a non-empty stretch CLOSES a run and owns two entries, (run start -> ref pos) and
// (run end -> None); a run that starts at the key where the previous one ended overwrites that None (:111-119), which
// is one entry less from there on (a prefix count of such overwrites).  The None entries of a round are stored first,
// the Some entries after a __syncwarp, so the overwrite wins.  Built once per run.
namespace {
__global__ void __launch_bounds__(128) table_build_kernel(DevStatic S, uint32_t* counts, TabEntry* out) {
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g < S.n_segments) table_build_warp_body(S, g, threadIdx.x & 31u, counts, out);
}
}  // namespace

void launch_table_build(const DevStatic& S, uint32_t* counts, TabEntry* out, cudaStream_t st) {
    if (!S.n_segments) return;
    table_build_kernel<<<(S.n_segments + 3) / 4, 128, 0, st>>>(S, counts, out);
}

// =================================================================================================== per-batch kernels
// The per-index logic lives in pair_bodies.cuh (shared with the host emulation of tests/emul); the kernels here own the
// thread mapping, the sentinel elements of the scans and the warp-aggregated counters.
namespace {

// a3 (count).  One thread per read walks its 1..k split segments.
__global__ void __launch_bounds__(128) pair_count_kernel(DevStatic S, DevBatch B, DevWork W, DevTotals* T) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B.n_reads) {
        if (r == B.n_reads) {
            W.rseg_pair_begin[B.n_rsegs] = 0;
            totals_reset(T);  // first kernel of the batch: nothing has touched the totals yet
        }
        return;
    }
    pair_count_body(S, B, W, T, r);
}

// a3 (count), one WARP per read: lanes over the CIGAR ops and over the contig's segments (same results as
// pair_count_body).  For 100 kb reads a thread per read walks ~700 ops alone and the batch has too few reads to fill the
// GPU (configs[4]: 0.33 ms for 29 k read segments); launch_lift picks this kernel from the batch's mean CIGAR length.
__global__ void __launch_bounds__(128) pair_count_warp_kernel(DevStatic S, DevBatch B, DevWork W, DevTotals* T) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (r >= B.n_reads) {
        if (r == B.n_reads && lane == 0) {
            W.rseg_pair_begin[B.n_rsegs] = 0;
            totals_reset(T);
        }
        return;
    }
    pair_count_warp_body(S, B, W, T, r, lane);
}

// a3 (fill).  One thread per read segment.
__global__ void __launch_bounds__(128) pair_fill_kernel(DevStatic S, DevBatch B, DevWork W, DevTotals* T) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == B.n_rsegs) {
        const uint32_t np = W.rseg_pair_begin[B.n_rsegs];
        T->n_pairs = np;
        if (np > W.pair_cap) atomicOr(&T->overflow, OVF_PAIRS);
        else W.pair_slot_begin[np] = 0;
    }
    if (s < B.n_rsegs) pair_fill_body(S, B, W, s);
}

// Work order of the lift kernel.  A warp of lift_pairs_kernel runs as long as its longest lane (natural order: 38 events
// for a mean of 22, ncu r04a: 15 of 32 lanes active), so the pairs of every group of kOrderGroup CONSECUTIVE pairs
// (adjacent CIGARs and overlapping table ranges: they stay hot in L2 and mostly in one L1) are bucket-sorted by work
// (strand of the contig segment, CIGAR length) and handed to the lift kernel as tiles of 32 pairs of near-equal length.
// Each tile is its own 32-thread block there, so a short tile gives its SM slot back as soon as it is done (sorting the
// warps of a 128-thread block changed nothing: the block lives as long as its longest warp; sorting the whole batch cost
// the L1 locality, profiles/r03a).
#ifndef LIFT_SORT
#define LIFT_SORT 1
#endif
#ifndef LIFT_ORDER_GROUP
#define LIFT_ORDER_GROUP 512
#endif
constexpr uint32_t kOrderGroup = LIFT_ORDER_GROUP;
constexpr uint32_t kOrderBuckets = 64;  // (reverse strand ? 32 : 0) + min(ops / 2, 31); parked / trivial pairs in bucket 0
__global__ void __launch_bounds__(256) pair_order_kernel(DevStatic S, DevBatch B, DevWork W, const DevTotals* T) {
    __shared__ uint32_t hist[kOrderBuckets];
    const uint32_t n_pairs = min(uint32_t(T->n_pairs), W.pair_cap);
    const uint32_t base = blockIdx.x * kOrderGroup;
    if (base >= n_pairs) return;
    const uint32_t cnt = min(kOrderGroup, n_pairs - base);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x < kOrderBuckets) hist[threadIdx.x] = 0u;
    __syncthreads();
    constexpr uint32_t kPer = kOrderGroup / 256;
    uint32_t key[kPer], rank[kPer];
#pragma unroll
    for (uint32_t k = 0; k < kPer; ++k) {
        const uint32_t i = threadIdx.x + k * 256u;
        key[k] = 0u;
        rank[k] = 0u;
        if (i < cnt) {
            const uint32_t n = B.rseg_cigar_len[W.pair_rseg[base + i]];
            const bool rev = S.seg_is_fwd[W.pair_seg[base + i]] == 0;
            key[k] = (n > W.long_ops) ? 0u : (rev ? 32u : 0u) + min(n >> 1, 31u);
            rank[k] = atomicAdd(&hist[key[k]], 1u);
        }
    }
    __syncthreads();
    if (warp == 0) {  // exclusive prefix over the buckets in DESCENDING key order: lane owns keys 63 - lane and 31 - lane
        const uint32_t hi = hist[63u - lane], lo = hist[31u - lane];
        uint32_t inc_hi = hi, inc_lo = lo;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t a = __shfl_up_sync(FULL, inc_hi, d), b2 = __shfl_up_sync(FULL, inc_lo, d);
            if (int(lane) >= d) { inc_hi += a; inc_lo += b2; }
        }
        const uint32_t total_hi = __shfl_sync(FULL, inc_hi, 31);
        hist[63u - lane] = inc_hi - hi;
        hist[31u - lane] = total_hi + inc_lo - lo;
    }
    __syncthreads();
#pragma unroll
    for (uint32_t k = 0; k < kPer; ++k) {
        const uint32_t i = threadIdx.x + k * 256u;
        if (i < cnt) W.pair_order[base + hist[key[k]] + rank[k]] = base + i;
    }
}

// a4 + a5 + a6 + a8: one thread per pair, one warp (= one block) per tile of 32 pairs; the stages are warp-collective.
// The lifted CIGARs of a tile are written to a shared-memory pool (StagePool), then moved by the warp, as one flat index
// space with coalesced stores, into the tile's dense output region: the thread-private, 7x-sparse scratch slots are no
// longer written for them (ncu r04a: DRAM traffic 2.6x the algorithmic bytes), and the record emission reads dense CIGARs.
#ifndef LIFT_COPY_FLAT
#define LIFT_COPY_FLAT 1  // 0: one source lane at a time (4x fewer instructions, but 32 dependent rounds per tile: 2 % slower - the kernel is bound by the
                         // length of a warp's dependent chain, not by issue slots)
#endif
#ifndef LIFT_STAGE_WORDS
#define LIFT_STAGE_WORDS 1024  // shared-memory words per tile for the staged outputs (32 per lane)
#endif
#ifndef LIFT_MIN_BLOCKS
#define LIFT_MIN_BLOCKS 32  // 64 registers: 32 tiles per SM
#endif
template <bool kAllStages>
__global__ void __launch_bounds__(32, LIFT_MIN_BLOCKS) lift_pairs_kernel(DevStatic S, DevBatch B, DevWork W, DevTotals* T, uint32_t stage_mask) {
    __shared__ uint32_t stage_pool[LIFT_STAGE_WORDS ? LIFT_STAGE_WORDS : 1];
    const uint32_t n_pairs = min(uint32_t(T->n_pairs), W.pair_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) T->scratch_needed = W.pair_slot_begin[n_pairs];  // total op slots of the batch (capacity feedback)
    const uint32_t i = blockIdx.x * 32u + threadIdx.x;
    if (blockIdx.x * 32u >= n_pairs) return;  // (block-uniform; the grid covers the pair capacity)
    const uint32_t lane = threadIdx.x;
    const bool valid = i < n_pairs;
    const uint32_t p = valid ? (LIFT_SORT ? W.pair_order[i] : i) : 0u;
    uint32_t a = 0, b = 0;
    const StagePool stage{LIFT_STAGE_WORDS ? stage_pool : nullptr, LIFT_STAGE_WORDS};
    const LiftOut out = lift_pair_body<kAllStages>(S, B, W, T, p, valid, stage_mask, a, b, stage);
    // ---- staged outputs of the tile -> its dense region (coalesced), or the pairs' own slots if it does not fit
    if (__any_sync(FULL, out.staged)) {
        const uint32_t n = out.staged ? out.n : 0u;
        uint32_t incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, incl, d);
            if (int(lane) >= d) incl += o;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        const uint32_t excl = incl - n;
        const uint64_t dense_at = W.dense_off + uint64_t(blockIdx.x) * kLiftTileOut;
        if (total <= kLiftTileOut) {
            if (out.staged) W.pair_out_off[p] = dense_at + excl;
            const uint32_t src_off = out.staged ? uint32_t(out.src - stage.pool) : 0u;
            uint32_t* const dense = W.scratch + dense_at;
            __syncwarp();
#if LIFT_COPY_FLAT
            for (uint32_t j0 = 0; j0 < total; j0 += 32u) {
                const uint32_t j = j0 + lane;
                uint32_t lo = 0, hi = 31;  // owner = last lane whose exclusive prefix <= j
#pragma unroll
                for (int it = 0; it < 5; ++it) {
                    const uint32_t mid = (lo + hi + 1u) >> 1;
                    const uint32_t ex_mid = __shfl_sync(FULL, excl, mid);
                    if (ex_mid <= j) lo = mid; else hi = mid - 1u;
                }
                const uint32_t s2 = __shfl_sync(FULL, src_off, lo);
                const uint32_t e2 = __shfl_sync(FULL, excl, lo);
                if (j < total) dense[j] = stage.pool[s2 + (j - e2)];
            }
#else
            // one source lane at a time, its ops spread over the lanes: coalesced both ways, 3 shuffles + a short loop per pair
            for (uint32_t sl = 0; sl < 32u; ++sl) {
                const uint32_t n_s = __shfl_sync(FULL, n, sl);
                const uint32_t s2 = __shfl_sync(FULL, src_off, sl);
                const uint32_t e2 = __shfl_sync(FULL, excl, sl);
                for (uint32_t j = lane; j < n_s; j += 32u) dense[e2 + j] = stage.pool[s2 + j];
            }
#endif
        } else if (out.staged) {  // (a tile of long CIGARs)
            uint32_t* dst = W.scratch + out.slot0;
            for (uint32_t j = 0; j < n; ++j) dst[j] = out.src[j];
            W.pair_out_off[p] = out.slot0;
        }
    }
    // roofline arithmetic: input ops walked + base bytes compared (warp-aggregated atomics)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_down_sync(FULL, a, d);
        b += __shfl_down_sync(FULL, b, d);
    }
    if (lane == 0) {
        if (a) atomicAdd(&T->n_in_ops, (unsigned long long)a);
        if (b) atomicAdd(&T->n_base_bytes, (unsigned long long)b);
    }
}

// The two warp-per-pair worklists of a batch in ONE launch (their lengths are only known on the device: a fixed grid of
// warps strides over the joint index space):
//   long_list      pairs with more than long_ops CIGAR ops: a5 + a6 + a8 + a9 with lanes over ops (lift_long_pair_body)
//   simplify_list  pairs lifted by lift_pairs_kernel whose CIGAR holds a mixed I/D run: a9 (simplify_warp_pair_body)
#ifndef LIFT_LONG_MIN_BLOCKS
#define LIFT_LONG_MIN_BLOCKS 6  // 80 registers, no spills (sweep on the stress workload: 4 -> 1.50 ms, 6 -> 1.43 ms, 8 spills -> 1.37 ms)
#endif
#ifndef SIMPLIFY_THREAD
#define SIMPLIFY_THREAD 1  // 1: the simplify worklist runs one thread per pair (simplify_pairs_kernel); 0: one warp per pair, in warp_pairs_kernel
#endif
__global__ void __launch_bounds__(128, LIFT_LONG_MIN_BLOCKS) warp_pairs_kernel(DevStatic S, DevBatch B, DevWork W, DevTotals* T, uint32_t stage_mask) {
    const uint32_t n_long = min(T->n_long, W.pair_cap);
    const uint32_t n_simp = (!SIMPLIFY_THREAD && (stage_mask & 6u) == 6u) ? min(T->n_simplify, W.pair_cap) : 0u;  // (long pairs simplify inline)
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t t = warp; t < n_long + n_simp; t += n_warps) {
        if (t < n_long) lift_long_pair_body(S, B, W, T, W.long_list[t], lane, stage_mask);
        else simplify_warp_pair_body(S, B, W, T, W.simplify_list[t - n_long], lane);
    }
}

// a9 over the simplify worklist, one thread per listed pair (pair_bodies.cuh: simplify_thread_pair_body).
__global__ void __launch_bounds__(128) simplify_pairs_kernel(DevStatic S, DevBatch B, DevWork W, DevTotals* T) {
    const uint32_t n_simp = min(T->n_simplify, W.pair_cap);
    if (blockIdx.x * blockDim.x >= n_simp) return;  // (block-uniform; the grid covers the pair capacity)
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t b = 0;
    simplify_thread_pair_body(S, B, W, T, i, i < n_simp, b);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) b += __shfl_down_sync(FULL, b, d);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&T->n_base_bytes, (unsigned long long)b);
}

// a10 (field part).  One thread per read.
__global__ void __launch_bounds__(128) read_finalize_kernel(DevStatic S, DevBatch B, DevWork W, DevTotals* T, int do_finish) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == B.n_reads) W.read_counts[B.n_reads] = make_uint2(0u, 0u);
    uint32_t tot = (r < B.n_reads) ? read_finalize_body(S, B, W, T, r, do_finish) : 0u;
    // one atomic per warp, not per read (a same-address atomic per thread cost 140 us per 200k reads, ncu r01)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tot += __shfl_down_sync(0xffffffffu, tot, d);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(&T->n_lifted, (unsigned long long)tot);
}

// Record assembly (:245-282 field updates).  A warp owns 32 consecutive reads: lane l writes the SoA fields of read l's
// records (adjacent lanes -> adjacent records -> coalesced), then the warp copies the CIGAR ops of each record from
// its scratch slot into the dense pool cooperatively (coalesced both ways) and reduces the reference span for
// end / bin (:278-279).
// The arrays live in ONE compact arena laid out from the batch's own totals (device_types.hpp: result_layout); thread 0
// also publishes the totals as the arena header, so one D2H copy brings everything the host needs.
// `rpw` reads per warp (a power of two <= 32; lanes >= rpw own no read and only help with the copies): 32 for HiFi-sized
// CIGARs; for 100 kb reads with hundreds of ops per record a warp per 32 reads left 0.2 waves of warps on the GPU, each
// copying 22 k ops through 700 dependent iterations (configs[4]: finalize_emit 0.49 ms for 29 k records).
__global__ void __launch_bounds__(256) emit_records_kernel(DevStatic S, DevBatch B, DevWork W, char* arena, uint64_t arena_cap, DevTotals* T,
                                                            uint32_t stage_mask, uint32_t rpw) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (lane < rpw) ? warp * rpw + lane : 0xffffffffu;
    if (uint64_t(warp) * rpw >= B.n_reads) return;
    const uint2 total = W.read_counts[B.n_reads];
    const ResultLayout L = result_layout(B.n_reads, total.x, total.y);
    const bool fits = L.total <= arena_cap;
    if (r == 0) {
        T->n_records = total.x;
        T->n_cigar_out = total.y;
        if (!fits) T->overflow |= OVF_RESULT;  // every other writer of the totals finished in earlier kernels
        *reinterpret_cast<DevTotals*>(arena) = *T;
    }
    if (!fits) return;
    const DevResult R = DevResult::view(arena, L);
    const bool live = r < B.n_reads;
    uint32_t k = 0, p = 0, p1 = 0, primary = 0xffffffffu;
    uint64_t op_at = 0;
    uint16_t flag0 = 0;
    bool has_recs = false;
    if (live) {
        const uint2 base = W.read_counts[r];
        const uint2 next = W.read_counts[r + 1];
        R.read_rec_begin[r] = base.x;
        if (r == B.n_reads - 1) {
            R.read_rec_begin[B.n_reads] = total.x;
            R.rec_cigar_begin[total.x] = total.y;
        }
        k = base.x;
        op_at = base.y;
        flag0 = B.read_flag[r];
        primary = W.read_primary[r];
        const uint32_t s0 = min(B.read_seg_begin[r], B.n_rsegs), s1 = min(max(B.read_seg_begin[r + 1], s0), B.n_rsegs);
        if (next.x > base.x) {
            if (primary == 0xffffffffu) {  // unmapped fallback (:317-335)
                emit_unmapped_record(B, R, r, k, s0, op_at);
            } else {
                has_recs = true;
                p = W.rseg_pair_begin[s0];
                p1 = min(W.rseg_pair_begin[s1], W.pair_cap);
            }
        }
    }
    // round j: every lane contributes its j-th lifted pair (most reads have exactly one); the ops of the <= 32 records
    // of a round are copied as ONE flat index space so that every lane moves an op per iteration (independent loads,
    // coalesced stores) instead of one record at a time
    for (;;) {
        uint32_t n = 0;
        uint64_t src_off = 0, dst_off = 0;
        bool mine = false;
        if (has_recs) {
            while (p < p1 && W.pair_status[p] != ST_LIFTED) ++p;
            if (p < p1) {
                mine = true;
                n = W.pair_n_out[p];
                src_off = W.pair_out_off[p];
                dst_off = op_at;
                emit_lifted_record(S, B, W, R, p, k, flag0, primary, op_at, stage_mask);
                ++k;
                op_at += n;
                ++p;
            } else {
                has_recs = false;
            }
        }
        if (!__any_sync(0xffffffffu, mine)) break;
        // inclusive prefix of op counts over the lanes
        uint32_t incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (int(lane) >= d) incl += o;
        }
        const uint32_t total_ops = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t excl = incl - n;
        // the destination is contiguous across the lanes of a round only if every read has a single record; use each
        // record's own (src, dst) pair to stay general
        for (uint32_t base = 0; base < total_ops; base += 32) {
            const uint32_t i = base + lane;
            // owner = last lane whose exclusive prefix <= i (binary search over the warp's registers)
            uint32_t lo = 0, hi = 31;
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                const uint32_t ex_mid = __shfl_sync(0xffffffffu, excl, mid);
                if (ex_mid <= i) lo = mid; else hi = mid - 1;
            }
            const uint64_t s2 = __shfl_sync(0xffffffffu, src_off, lo);
            const uint64_t d2 = __shfl_sync(0xffffffffu, dst_off, lo);
            const uint32_t e2 = __shfl_sync(0xffffffffu, excl, lo);
            if (i < total_ops) R.cigar[d2 + (i - e2)] = W.scratch[s2 + (i - e2)];
        }
    }
}

__global__ void totals_init_kernel(DevTotals* T, DevTotals* header) { totals_reset(T); *header = *T; }

}  // namespace

void launch_lift(const DevStatic& S, const DevBatch& B, const DevWork& W, char* arena, uint64_t arena_cap, DevTotals* T, uint32_t stage_mask,
                 void* scan_tmp, size_t scan_tmp_bytes_, cudaStream_t st, uint64_t* launches, StageEvents* ev) {
    auto mark = [&](int i) { if (ev) cudaEventRecord(ev->e[i], st); };
    const int do_finish = (stage_mask == 7u);
    mark(0);
    if (B.n_reads == 0) {
        totals_init_kernel<<<1, 1, 0, st>>>(T, reinterpret_cast<DevTotals*>(arena));
        ++*launches;
        for (int i = 1; i < StageEvents::N; ++i) mark(i);
        return;
    }
    const uint64_t ops_per_read = B.n_cigar / std::max<uint64_t>(B.n_reads, 1);  // picks kernel shapes only, never results
    if (ops_per_read > 96) pair_count_warp_kernel<<<unsigned((uint64_t(B.n_reads) + 1 + 3) / 4), 128, 0, st>>>(S, B, W, T);
    else pair_count_kernel<<<(B.n_reads + 1 + 127) / 128, 128, 0, st>>>(S, B, W, T);
    ++*launches;
    exclusive_scan_inplace<uint32_t>(W.rseg_pair_begin, uint64_t(B.n_rsegs) + 1, scan_tmp, scan_tmp_bytes_, st, launches, nullptr);
    pair_fill_kernel<<<(B.n_rsegs + 1 + 127) / 128, 128, 0, st>>>(S, B, W, T);
    ++*launches;
    // the pair count is only known on the device: scan / launch over the capacity, kernels clamp to n_pairs
    exclusive_scan_inplace<uint64_t>(W.pair_slot_begin, uint64_t(W.pair_cap) + 1, scan_tmp, scan_tmp_bytes_, st, launches, &T->n_pairs);
    if (LIFT_SORT) {
        pair_order_kernel<<<(W.pair_cap + kOrderGroup - 1) / kOrderGroup, 256, 0, st>>>(S, B, W, T);
        ++*launches;
    }
    mark(1);
    if (stage_mask == 7u) lift_pairs_kernel<true><<<(W.pair_cap + 31u) / 32u, 32, 0, st>>>(S, B, W, T, stage_mask);
    else lift_pairs_kernel<false><<<(W.pair_cap + 31u) / 32u, 32, 0, st>>>(S, B, W, T, stage_mask);
    ++*launches;
    if ((stage_mask & 2u) || stage_mask == 1u) {
        const unsigned blocks = unsigned(std::min<uint64_t>((uint64_t(W.pair_cap) + 3) / 4, 148ull * 16));
        warp_pairs_kernel<<<blocks, 128, 0, st>>>(S, B, W, T, stage_mask);
        ++*launches;
        if (SIMPLIFY_THREAD && (stage_mask & 6u) == 6u) {
            simplify_pairs_kernel<<<(W.pair_cap + 127u) / 128u, 128, 0, st>>>(S, B, W, T);
            ++*launches;
        }
    }
    mark(2);
    read_finalize_kernel<<<(B.n_reads + 1 + 127) / 128, 128, 0, st>>>(S, B, W, T, do_finish);
    ++*launches;
    exclusive_scan_inplace<uint2>(W.read_counts, uint64_t(B.n_reads) + 1, scan_tmp, scan_tmp_bytes_, st, launches, nullptr);
    // reads per warp of the record emission from the mean CIGAR length of the batch (any value gives the same result)
    const uint32_t rpw = ops_per_read > 256 ? 1u : ops_per_read > 96 ? 4u : 32u;
    const uint64_t emit_warps = (uint64_t(B.n_reads) + rpw - 1) / rpw;
    emit_records_kernel<<<unsigned((emit_warps + 7) / 8), 256, 0, st>>>(S, B, W, arena, arena_cap, T, stage_mask, rpw);
    ++*launches;
    mark(3);
}

// =================================================================================================== record assembly
namespace {
__global__ void __launch_bounds__(256) assemble_sizes_kernel(uint32_t n_records, uint32_t n_reads, const uint32_t* read_rec_begin,
                                                             const uint32_t* read_seq_len, const uint64_t* read_qual_off, uint64_t qual_bytes,
                                                             uint32_t* rec_read, uint64_t* seq_begin, uint64_t* qual_begin, unsigned int* error) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n_records) return;
    if (k == n_records) { seq_begin[k] = 0; qual_begin[k] = 0; return; }
    // the read that owns record k: last r with read_rec_begin[r] <= k (as bam_rec_layout; the fallback record of a read
    // without segments has no rec_read_segment of its own to go through)
    uint32_t lo = 0, hi = n_reads;
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (read_rec_begin[mid] <= k) lo = mid; else hi = mid;
    }
    const uint32_t r = lo;
    const uint64_t len = read_seq_len[r];
    // caller memory: a quality range outside the pool is reported, never read (the writer only runs on a clean error word)
    if (k == read_rec_begin[r] && (read_qual_off[r] > qual_bytes || len > qual_bytes - read_qual_off[r])) atomicOr(error, 16u);
    rec_read[k] = r;
    seq_begin[k] = (((len + 1) >> 1) + 15) & ~15ull;
    qual_begin[k] = (len + 15) & ~15ull;
}
// One block per output record (22.5 KB in, 22.5 KB out for a 15 kb read): HBM-bound streaming.
__global__ void __launch_bounds__(256) assemble_records_kernel(AsmArgs A) { assemble_record_body(A, blockIdx.x, threadIdx.x, blockDim.x); }
}  // namespace

void launch_assemble_sizes(uint32_t n_records, uint32_t n_reads, const uint32_t* read_rec_begin, const uint32_t* read_seq_len,
                           const uint64_t* read_qual_off, uint64_t qual_bytes, uint32_t* rec_read, uint64_t* seq_begin, uint64_t* qual_begin,
                           unsigned int* error, void* scan_tmp, size_t scan_tmp_bytes_, cudaStream_t st, uint64_t* launches) {
    assemble_sizes_kernel<<<(n_records + 1 + 255) / 256, 256, 0, st>>>(n_records, n_reads, read_rec_begin, read_seq_len, read_qual_off, qual_bytes, rec_read,
                                                                        seq_begin, qual_begin, error);
    ++*launches;
    exclusive_scan_inplace<uint64_t>(seq_begin, uint64_t(n_records) + 1, scan_tmp, scan_tmp_bytes_, st, launches, nullptr);
    exclusive_scan_inplace<uint64_t>(qual_begin, uint64_t(n_records) + 1, scan_tmp, scan_tmp_bytes_, st, launches, nullptr);
}
void launch_assemble_records(const AsmArgs& A, cudaStream_t st, uint64_t* launches) {
    if (!A.n_records) return;
    assemble_records_kernel<<<A.n_records, 256, 0, st>>>(A);
    ++*launches;
}


// =================================================================================================== record assembly, whole BAM records
namespace {
__global__ void __launch_bounds__(128) bam_read_prep_kernel(BamAsmArgs A) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < A.n_reads) bam_read_prep_body(A, r);
}
__global__ void __launch_bounds__(128) bam_rec_prep_kernel(BamAsmArgs A) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < A.n_records) bam_rec_prep_body(A, k);
}
__global__ void __launch_bounds__(128) bam_rec_size_kernel(BamAsmArgs A) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > A.n_records) return;
    if (k == A.n_records) { A.rec_begin[k] = 0ull; return; }
    const BamRecLayout L = bam_rec_layout(A, k);
    if (L.ps_n > 0xffu || L.sa_n > 0xffffffu) atomicOr(A.error, 4u);  // (descriptor field widths: a 255-byte contig name, 16 MB of SA text)
    if (L.name_n > 254u) atomicOr(A.error, 8u);                        // (l_read_name is a u8 that counts the NUL)
    L.pack(A.rec_desc + 2 * size_t(k));
    A.rec_begin[k] = L.total;
}
// One warp per output record: everything but bases and qualities.
__global__ void __launch_bounds__(128) bam_write_meta_kernel(BamAsmArgs A) {
    const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k < A.n_records) bam_write_meta_body(A, k, threadIdx.x & 31u, 32u);
}
// One block per output record: HBM-bound streaming of bases and qualities (assemble_bam.cuh).
__global__ void __launch_bounds__(256, 6) bam_write_kernel(BamAsmArgs A) { bam_write_bases_body(A, blockIdx.x, threadIdx.x, blockDim.x); }
}  // namespace

void launch_bam_sizes(const BamAsmArgs& A, void* scan_tmp, size_t scan_tmp_bytes_, cudaStream_t st, uint64_t* launches) {
    if (A.n_reads) { bam_read_prep_kernel<<<(A.n_reads + 127) / 128, 128, 0, st>>>(A); ++*launches; }
    if (A.n_records) { bam_rec_prep_kernel<<<(A.n_records + 127) / 128, 128, 0, st>>>(A); ++*launches; }
    bam_rec_size_kernel<<<(A.n_records + 1 + 127) / 128, 128, 0, st>>>(A);
    ++*launches;
    exclusive_scan_inplace<uint64_t>(A.rec_begin, uint64_t(A.n_records) + 1, scan_tmp, scan_tmp_bytes_, st, launches, nullptr);
}
void launch_bam_write(const BamAsmArgs& A, cudaStream_t st, cudaStream_t st_meta, uint64_t* launches) {
    if (!A.n_records) return;
    // the two kernels write disjoint bytes of the same records: the small latency-bound one runs beside the streaming one
    bam_write_meta_kernel<<<(A.n_records + 3) / 4, 128, 0, st_meta>>>(A);
    bam_write_kernel<<<A.n_records, 256, 0, st>>>(A);
    *launches += 2;
}

// =================================================================================================== BGZF framing, level 0
namespace {
__global__ void __launch_bounds__(256, 4) bgzf_store_kernel(BgzfArgs A) {
    extern __shared__ __align__(16) uint32_t bgzf_smem[];
    bgzf_store_init(A, threadIdx.x, bgzf_smem);
    for (uint64_t b = blockIdx.x; b < A.n_blocks; b += gridDim.x) bgzf_store_block_body(A, b, threadIdx.x, bgzf_smem);
}
}  // namespace

namespace {
__global__ void __launch_bounds__(256, 4) bam_frame_kernel(FrameArgs F) {
    extern __shared__ __align__(16) uint32_t bgzf_smem[];
    bgzf_store_init(F.Z, threadIdx.x, bgzf_smem);
    for (uint64_t b = blockIdx.x; b < F.Z.n_blocks; b += gridDim.x) bgzf_frame_block_body(F, b, threadIdx.x, bgzf_smem);
}
}  // namespace

// ptl_frame_records: the small fields of every record into the sparse scratch (one warp per record), then one persistent
// thread block per BGZF block producing its payload from the sources and its CRC from what it has just written.
void launch_bam_frame(const FrameArgs& F, cudaStream_t st, uint64_t* launches) {
    if (F.A.n_records) {
        bam_write_meta_kernel<<<(F.A.n_records + 3) / 4, 128, 0, st>>>(F.A);
        ++*launches;
    }
    if (!F.Z.n_blocks) return;
    const int smem = int(kFrameSmemWords * 4u);
    cudaFuncSetAttribute(bam_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    bam_frame_kernel<<<unsigned(std::min<uint64_t>(F.Z.n_blocks, 148ull * 4)), 256, smem, st>>>(F);
    ++*launches;
}

void launch_bgzf_store(const BgzfArgs& A, cudaStream_t st, uint64_t* launches) {
    if (!A.n_blocks) return;
    const int smem = int(kBgzfSmemWords * 4u);
    cudaFuncSetAttribute(bgzf_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    bgzf_store_kernel<<<unsigned(std::min<uint64_t>(A.n_blocks, 148ull * 4)), 256, smem, st>>>(A);  // persistent: 4 thread blocks per SM
    ++*launches;
}

}  // namespace ptl
