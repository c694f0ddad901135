// CUDA kernels of the liftover path (sm_100a).  Launch wrappers at the bottom (kernels.hpp).
//
// Per batch:   pair_count -> scan -> pair_fill -> scan -> lift_pairs -> read_finalize -> scan -> emit_records
// Once:        table_count -> scan -> table_fill   (flat ReadToRefTreeMap per contig segment)
//
// This is HBM/latency-bound integer work: no tensor cores.  Grids are sized from the work (>> 148 SMs x resident CTAs
// for real batches); scratch traffic is thread-private and stays in L1/L2.
#include <cuda_runtime.h>

#include <algorithm>

#include "kernels.hpp"
#include "lift_device.cuh"

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
// a7: get_read_segment_to_ref_pos_tree_map (lib/rust-vc-utils/src/bam_utils/read_to_ref_map.rs:101-137), flattened.
// A run of M/=/X ops (any other op ends it) with total length > 0 yields (run_start_read_pos -> run_start_ref_pos) and
// (run_end_read_pos -> None); the None is overwritten when the next run starts at the same read_pos (:111-119).
// One thread per segment walks its CIGAR; `out == nullptr` counts.  Built once per run.
namespace {
__global__ void table_build_kernel(DevStatic S, uint32_t* counts, int2* out) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= S.n_segments) return;
    const uint32_t* c = S.seg_cigar + S.seg_cigar_begin[g];
    const uint32_t n = uint32_t(S.seg_cigar_begin[g + 1] - S.seg_cigar_begin[g]);
    int64_t ref_pos = S.seg_pos[g];
    uint64_t read_pos = 0, match_len = 0;
    uint32_t w = 0;
    int2* o = out ? out + S.seg_tab_begin[g] : nullptr;
    // The overwrite decision must not depend on reading `out` (the count pass has none): track the last None key.
    uint32_t last_key = 0xffffffffu;
    bool have_last = false;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t x = c[i];
        if (op_is_match(x & 0xfu)) {
            match_len += x >> 4;
        } else if (match_len > 0) {
            const uint32_t k0 = uint32_t(read_pos - match_len);
            if (have_last && last_key == k0) --w;
            if (o) { o[w] = make_int2(int(k0), int32_t(ref_pos - int64_t(match_len))); }
            ++w;
            if (o) { o[w] = make_int2(int(uint32_t(read_pos)), -1); }
            ++w;
            last_key = uint32_t(read_pos);
            have_last = true;
            match_len = 0;
        }
        ref_pos += op_ref_adv(x);
        read_pos += op_read_adv(x);
    }
    if (match_len > 0) {
        const uint32_t k0 = uint32_t(read_pos - match_len);
        if (have_last && last_key == k0) --w;
        if (o) { o[w] = make_int2(int(k0), int32_t(ref_pos - int64_t(match_len))); }
        ++w;
        if (o) { o[w] = make_int2(int(uint32_t(read_pos)), -1); }
        ++w;
    }
    if (!out) counts[g] = w;
}
}  // namespace

void launch_table_build(const DevStatic& S, uint32_t* counts, int2* out, cudaStream_t st) {
    if (!S.n_segments) return;
    table_build_kernel<<<(S.n_segments + 127) / 128, 128, 0, st>>>(S, counts, out);
}

// =================================================================================================== per-batch kernels
namespace {

__device__ __forceinline__ void totals_reset(DevTotals* T) {
    T->n_pairs = 0; T->scratch_needed = 0; T->n_records = 0; T->n_cigar_out = 0; T->n_lifted = 0; T->n_errors = 0;
    T->first_error_read = 0x7fffffffffffffffLL; T->first_error_status = 0; T->overflow = 0; T->n_in_ops = 0; T->n_base_bytes = 0;
    T->n_simplify = 0;
}

// a3 (count): get_contig_split_segments_from_read_mapping (src/read_alignment_scanner.rs:80-103) + get_cigar_ref_offset.
// One thread per read walks its 1..k split segments.
__global__ void __launch_bounds__(128) pair_count_kernel(DevStatic S, DevBatch B, DevWork W, DevTotals* T) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B.n_reads) {
        if (r == B.n_reads) {
            W.rseg_pair_begin[B.n_rsegs] = 0;
            totals_reset(T);  // first kernel of the batch: nothing has touched the totals yet
        }
        return;
    }
    for (uint32_t s = B.read_seg_begin[r]; s < B.read_seg_begin[r + 1]; ++s) {
        W.rseg_read[s] = r;
        const uint32_t* c = B.cigar + B.rseg_cigar_begin[s];
        const uint32_t n = B.rseg_cigar_len[s];
        int64_t ref_len = 0;
        uint32_t n_id = 0, read_len = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t x = c[i];
            ref_len += op_ref_adv(x);
            read_len += op_read_adv(x);
            n_id += ((x & 0xfu) == OP_I || (x & 0xfu) == OP_D) ? 1u : 0u;
        }
        W.rseg_ref_len[s] = ref_len;
        W.rseg_n_id[s] = n_id;
        W.rseg_read_len[s] = read_len;  // get_cigar_read_offset(cigar, ignore_hard_clip=false)
        const int64_t start = B.rseg_pos[s], end = start + ref_len;
        const uint32_t ctg = B.rseg_contig[s];
        uint32_t cnt = 0;
        for (uint32_t g = S.contig_seg_begin[ctg]; g < S.contig_seg_begin[ctg + 1]; ++g) {
            // IntRange::intersect_range with the segment as `self`: other.end >= self.start && other.start < self.end
            if (end >= int64_t(S.seg_so_start[g]) && start < int64_t(S.seg_so_end[g])) ++cnt;
        }
        W.rseg_pair_begin[s] = cnt;
    }
}

__device__ __forceinline__ uint32_t lower_bound_key(const int2* tab, uint32_t lo, uint32_t hi, int64_t key) {
    while (lo < hi) {  // first index with tab.key >= key
        const uint32_t mid = (lo + hi) >> 1;
        if (int64_t(uint32_t(tab[mid].x)) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// a3 (fill): pair list in (read segment, contig segment index) order + the scratch-slot bound of each pair.
__global__ void __launch_bounds__(128) pair_fill_kernel(DevStatic S, DevBatch B, DevWork W, DevTotals* T) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B.n_rsegs) {
        if (s == B.n_rsegs) {
            const uint32_t np = W.rseg_pair_begin[B.n_rsegs];
            T->n_pairs = np;
            if (np > W.pair_cap) atomicOr(&T->overflow, OVF_PAIRS);
            else W.pair_slot_begin[np] = 0;
        }
        return;
    }
    uint32_t p = W.rseg_pair_begin[s];
    const uint32_t p_end = W.rseg_pair_begin[s + 1];
    if (p == p_end) return;
    const int64_t start = B.rseg_pos[s], ref_len = W.rseg_ref_len[s], end = start + ref_len;
    const uint32_t ctg = B.rseg_contig[s];
    const uint32_t n_in = B.rseg_cigar_len[s];
    for (uint32_t g = S.contig_seg_begin[ctg]; g < S.contig_seg_begin[ctg + 1]; ++g) {
        if (!(end >= int64_t(S.seg_so_start[g]) && start < int64_t(S.seg_so_end[g]))) continue;
        if (p < W.pair_cap) {
            W.pair_rseg[p] = s;
            W.pair_seg[p] = g;
            // contig interval walked by the liftover, in the orientation of the segment's table
            const bool fwd = S.seg_is_fwd[g] != 0;
            const int64_t a = fwd ? start : int64_t(S.contig_len[ctg]) - end;
            const uint32_t t0 = S.seg_tab_begin[g], t1 = S.seg_tab_begin[g + 1];
            const uint32_t n_keys = lower_bound_key(S.table, t0, t1, a + ref_len) - lower_bound_key(S.table, t0, t1, a);
            // op-slot bounds, in stored (compressed) ops:
            //   shifted    <= n_in + n_id + 1            (each I/D op can split one match block in two)
            //   lifted     <= shifted + 2 n_keys          (one extra piece and one gap-D per table key in range)
            //   simplified <= lifted + 2 (n_id + n_keys)  (a mixed cluster grows by <= 2 ops and owns >= 1 D op)
            const uint32_t n_id = W.rseg_n_id[s];
            const uint32_t n_shift = fwd ? n_in : n_in + n_id + 1u;
            //   buffer B doubles as the cluster list of the left shift (3 words per cluster), buffer A keeps 4 words per
            //   mixed cluster of the simplify stage at its top end
            const uint32_t cap_b = max(n_shift + 2u * n_keys + 4u, fwd ? 0u : 3u * n_id + 4u);
            const uint32_t cap_a = cap_b + 6u * (n_id + n_keys) + 8u;
            W.pair_cap_b[p] = cap_b;
            W.pair_slot_begin[p] = uint64_t(cap_a) + cap_b;  // [0,cap_b) = buffer B, [cap_b, cap_b+cap_a) = buffer A
        }
        ++p;
    }
}

// a4 + a5 + a6 + a8 + a9: get_liftover_alignment_for_read_and_contig_segment (src/read_alignment_scanner.rs:136-288),
// one thread per pair; the stages are warp-collective (all 32 lanes enter, idle lanes carry active = false) so that
// the latency-bound base fetches of a warp are issued together (see LeftShifter).
#ifndef LIFT_MIN_BLOCKS
#define LIFT_MIN_BLOCKS 8  // 64 registers, 32 warps per SM (sweep in profiles/: 8 -> 0.558 ms, 1 -> 0.579 ms, 12 spills)
#endif
__global__ void __launch_bounds__(128, LIFT_MIN_BLOCKS) lift_pairs_kernel(DevStatic S, DevBatch B, DevWork W, DevTotals* T, uint32_t stage_mask) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_pairs = min(uint32_t(T->n_pairs), W.pair_cap);
    const bool valid = p < n_pairs;
    if (p == 0) T->scratch_needed = W.pair_slot_begin[n_pairs];  // total op slots of the batch (capacity feedback)
    PairCounters cnt;
    uint32_t n_in_ops = 0;
    int status = ST_NONE, err = 0;
    uint32_t cpos = 0;   // position on the contig strand the segment's table is written in
    int64_t rpos = 0;    // position on the reference once lifted
    bool need_flip = false, contig_fwd = true, usable = false;
    uint32_t s = 0, g = 0, ctg = 0, seq_len = 0, cap_a = 0, cap_b = 0;
    uint32_t* buf_a = nullptr;
    uint32_t* buf_b = nullptr;
    OpSource cur{nullptr, 0, false};
    ReadBases read{nullptr, 0, false};
    uint64_t slot0 = 0;
    if (valid) {
        s = W.pair_rseg[p];
        g = W.pair_seg[p];
        const uint32_t r = W.rseg_read[s];
        slot0 = W.pair_slot_begin[p];
        const uint64_t slot1 = W.pair_slot_begin[p + 1];
        contig_fwd = S.seg_is_fwd[g] != 0;
        const bool rec_rev = (B.read_flag[r] & 0x10) != 0;
        const bool changes_strand = (rec_rev == (B.rseg_is_fwd[s] != 0));
        need_flip = (!contig_fwd) != changes_strand;
        if (slot1 > W.scratch_cap) {
            atomicOr(&T->overflow, OVF_SCRATCH);
            err = ST_ERR_CAPACITY;
        } else {
            usable = true;
            cap_b = W.pair_cap_b[p];
            cap_a = uint32_t(slot1 - slot0) - cap_b;
            buf_b = W.scratch + slot0;
            buf_a = buf_b + cap_b;
            ctg = B.rseg_contig[s];
            seq_len = B.read_seq_len[r];
            read = ReadBases{B.seq4 + B.read_seq_off[r], seq_len, need_flip};
            cur = OpSource{B.cigar + B.rseg_cigar_begin[s], B.rseg_cigar_len[s], false};
            n_in_ops = cur.n;
            const int64_t pos = B.rseg_pos[s];  // validated on the host: 0 <= pos < 2^31
            status = ST_LIFTED;
            if (contig_fwd) {
                cpos = uint32_t(pos);
            } else {
                // reverse-strand contig segment: flip onto the contig's reverse strand (:162-167)
                const int64_t rev = int64_t(S.contig_len[ctg]) - (pos + W.rseg_ref_len[s]);
                if (rev < 0) { err = ST_ERR_BOUNDS; usable = false; }  // read runs past the contig end (invalid input)
                cpos = uint32_t(rev);
                cur.reversed = true;
            }
        }
    }
    bool cur_is_a = false, cur_is_raw = true;
    bool simplify_is_identity = false;
    uint32_t span = 0;  // reference span of the final CIGAR (end = pos + span, :278)

    // ---- a5: left-shift on the contig's reverse strand (:168-175)
    {
        bool go = usable && !contig_fwd && (stage_mask & 1u);
        uint64_t rev_off = ~0ull;
        if (go) {
            rev_off = S.contig_rev_off[ctg];
            if (rev_off == ~0ull) { err = ST_ERR_BOUNDS; go = false; usable = false; }  // Option::unwrap on None (:174)
        }
        if (__any_sync(FULL, go)) {
            OpSink sink(buf_a, go ? cap_a : 0u);
            const uint32_t shifted = run_left_shift_warp(go, cur, cpos, go ? S.rev_pool + rev_off : nullptr,
                                                         go ? uint32_t(S.contig_len[ctg]) : 0u, read, buf_b, sink, cnt, err);
            if (go) {
                cpos = shifted;
                span = sink.ref_span;
                if (sink.overflow) err = ST_ERR_CAPACITY;
                cur = OpSource{buf_a, sink.n, false};
                cur_is_a = true;
                cur_is_raw = false;
            }
        }
    }
    rpos = cpos;
    // ---- a6: liftover (:179-183) + length check (:204-229).  The lifted CIGAR consumes exactly the read bases of the
    //      segment CIGAR (every read-consuming op is re-emitted as M/I/S; the left shift preserves them too), so the
    //      reference's check `seq_len == read length of the lifted CIGAR` is decided by the input CIGAR's read length.
    if (usable && !err && (stage_mask & 2u)) {
        OpSink sink(buf_b, cap_b);
        int64_t lifted_pos = 0;
        const bool some = run_liftover(cur, cpos, S.table, S.seg_tab_begin[g], S.seg_tab_begin[g + 1], sink, &lifted_pos);
        if (sink.overflow) err = ST_ERR_CAPACITY;
        else if (!some) status = ST_NONE;
        else if (W.rseg_read_len[s] != seq_len) err = ST_ERR_LENGTH;
        // simplify_alignment_indels rewrites only I/D runs that hold both kinds; on a cleaned + compressed CIGAR without
        // such a run it is the identity (single-kind runs are already one op, edges are already clean), so it is skipped
        simplify_is_identity = !sink.mixed_cluster;
        span = sink.ref_span;
        rpos = lifted_pos;
        cur = OpSource{buf_b, sink.n, false};
        cur_is_a = false;
        cur_is_raw = false;
    }
    // ---- a9: simplify (:236-243).  Only the few pairs whose lifted CIGAR holds a mixed I/D run need it: they are
    //      appended to a worklist and finished by simplify_pairs_kernel with all lanes busy (inline, the stage ran at
    //      1.9 active threads per instruction and cost 22 % of this kernel, ncu r01g).
    bool deferred = false;
    if (usable && !err && status == ST_LIFTED && (stage_mask & 4u) && !simplify_is_identity) {
        if (stage_mask & 2u) {
            deferred = true;
            W.simplify_list[atomicAdd(&T->n_simplify, 1u)] = p;
        }
    }
    {   // stage tests without the liftover stage: simplify inline, (raw input or A) -> A
        const bool go = usable && !err && status == ST_LIFTED && (stage_mask & 4u) && !(stage_mask & 2u);
        if (__any_sync(FULL, go)) {
            const uint8_t* ref = nullptr;
            uint64_t ref_len = 0;
            uint32_t* rec = nullptr;
            if (go) {
                const int32_t chrom = S.seg_chrom[g];
                ref = S.ref + S.chrom_off[chrom];
                ref_len = S.chrom_off[chrom + 1] - S.chrom_off[chrom];
                if (cur_is_a) {  // shift without liftover: move the input out of the way
                    const uint32_t n = min(cur.n, cap_b);
                    for (uint32_t i = 0; i < n; ++i) buf_b[i] = buf_a[i];
                    cur = OpSource{buf_b, n, false};
                }
                const uint32_t n_rec = 4u * ((cap_a - cap_b - 8u) / 6u);  // 4 words x (n_id + n_keys) possible mixed clusters
                rec = buf_a + (cap_a - n_rec);
            }
            OpSink sink(buf_a, go ? uint32_t(rec - buf_a) : 0u);
            const int64_t simp = run_simplify_warp(go, cur, rpos, ref, ref_len, read, rec, sink, cnt, err);
            if (go) {
                rpos = simp;
                span = sink.ref_span;
                if (sink.overflow) err = ST_ERR_CAPACITY;
                cur = OpSource{buf_a, sink.n, false};
                cur_is_a = true;
                cur_is_raw = false;
            }
        }
    }
    if (usable && !err && status == ST_LIFTED && cur_is_raw) {
        // stage tests with every stage disabled for this pair: hand the (possibly reversed) input back verbatim
        const uint32_t n = min(cur.n, cap_a);
        for (uint32_t i = 0; i < n; ++i) { const uint32_t c = cur.get(i); buf_a[i] = c; span += op_ref_adv(c); }
        cur = OpSource{buf_a, n, false};
    }
    if (valid) {
        if (err) status = err;
        const bool ok = (status == ST_LIFTED);
        if (ok && deferred) status = ST_PENDING_SIMPLIFY;
        W.pair_status[p] = int8_t(status);
        W.pair_flip[p] = need_flip;
        W.pair_pos[p] = ok ? rpos : 0;
        W.pair_n_out[p] = ok ? cur.n : 0u;
        W.pair_out_off[p] = ok ? uint64_t(cur.p - W.scratch) : slot0;
        W.pair_bin[p] = ok ? reg2bin(rpos, rpos + int64_t(span)) : uint16_t(0);  // bam_reg2bin(pos, end) (:278-279)
    }
    // roofline arithmetic: input ops walked + base bytes compared (warp-aggregated atomics)
    uint32_t a = n_in_ops, b = cnt.base_bytes;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_down_sync(FULL, a, d);
        b += __shfl_down_sync(FULL, b, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (a) atomicAdd(&T->n_in_ops, (unsigned long long)a);
        if (b) atomicAdd(&T->n_base_bytes, (unsigned long long)b);
    }
}

// a9 for the worklist of pairs whose lifted CIGAR holds a mixed I/D run: dense, warp-collective, B -> A.
__global__ void __launch_bounds__(128) simplify_pairs_kernel(DevStatic S, DevBatch B, DevWork W, DevTotals* T) {
    const uint32_t n = min(T->n_simplify, W.pair_cap);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    PairCounters cnt;
    for (uint32_t base = warp * 32u; base < n; base += n_warps * 32u) {
        const uint32_t t = base + lane;
        const bool active = t < n;
        uint32_t p = 0, cap_b = 0, n_in = 0;
        uint32_t* buf_a = nullptr;
        uint32_t* rec = nullptr;
        const uint8_t* ref = nullptr;
        uint64_t ref_len = 0;
        int64_t rpos = 0;
        OpSource cur{nullptr, 0, false};
        ReadBases read{nullptr, 0, false};
        uint32_t sink_cap = 0;
        if (active) {
            p = W.simplify_list[t];
            const uint32_t s = W.pair_rseg[p], g = W.pair_seg[p], r = W.rseg_read[s];
            const uint64_t slot0 = W.pair_slot_begin[p], slot1 = W.pair_slot_begin[p + 1];
            cap_b = W.pair_cap_b[p];
            const uint32_t cap_a = uint32_t(slot1 - slot0) - cap_b;
            uint32_t* buf_b = W.scratch + slot0;
            buf_a = buf_b + cap_b;
            n_in = W.pair_n_out[p];
            cur = OpSource{buf_b, n_in, false};
            rpos = W.pair_pos[p];
            read = ReadBases{B.seq4 + B.read_seq_off[r], B.read_seq_len[r], W.pair_flip[p] != 0};
            const int32_t chrom = S.seg_chrom[g];
            ref = S.ref + S.chrom_off[chrom];
            ref_len = S.chrom_off[chrom + 1] - S.chrom_off[chrom];
            const uint32_t n_rec = 4u * ((cap_a - cap_b - 8u) / 6u);
            rec = buf_a + (cap_a - n_rec);
            sink_cap = uint32_t(rec - buf_a);
        }
        int err = 0;
        OpSink sink(buf_a, sink_cap);
        const int64_t simp = run_simplify_warp(active, cur, rpos, ref, ref_len, read, rec, sink, cnt, err);
        if (active) {
            if (sink.overflow) err = ST_ERR_CAPACITY;
            W.pair_status[p] = int8_t(err ? err : ST_LIFTED);
            W.pair_pos[p] = err ? 0 : simp;
            W.pair_n_out[p] = err ? 0u : sink.n;
            W.pair_out_off[p] = uint64_t(buf_a - W.scratch);
            W.pair_bin[p] = err ? uint16_t(0) : reg2bin(simp, simp + int64_t(sink.ref_span));
        }
    }
    uint32_t b = cnt.base_bytes;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) b += __shfl_down_sync(FULL, b, d);
    if (lane == 0 && b) atomicAdd(&T->n_base_bytes, (unsigned long long)b);
}

// a10 (field part): finish_remapped_alignment_set (src/read_alignment_scanner.rs:310-366): record counts per read,
// primary = first max MAPQ (:338-346), unmapped fallback when nothing lifted (:317-335).  One thread per read.
__global__ void __launch_bounds__(128) read_finalize_kernel(DevStatic S, DevBatch B, DevWork W, DevTotals* T, int do_finish) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lifted = 0, ops = 0, primary = 0xffffffffu;
    int best_mapq = -1, first_err = 0;
    if (r == B.n_reads) W.read_counts[B.n_reads] = make_uint2(0u, 0u);
    const bool live = r < B.n_reads;
    const uint32_t s0 = live ? B.read_seg_begin[r] : 0u, s1 = live ? B.read_seg_begin[r + 1] : 0u;
    const uint32_t p0 = live ? W.rseg_pair_begin[s0] : 0u, p1 = live ? min(W.rseg_pair_begin[s1], W.pair_cap) : 0u;
    for (uint32_t p = p0; p < p1; ++p) {
        const int st = W.pair_status[p];
        if (st < 0) { if (!first_err) first_err = st; continue; }
        if (st != ST_LIFTED) continue;
        ++lifted;
        ops += W.pair_n_out[p];
        const int mq = S.seg_mapq[W.pair_seg[p]];
        if (mq > best_mapq) { best_mapq = mq; primary = p; }
    }
    if (first_err) {
        // the reference panics here; report, and emit the unmapped fallback so the batch stays well-formed
        atomicAdd(&T->n_errors, 1ull);
        const long long packed = (static_cast<long long>(r) << 8) | (first_err & 0xff);
        atomicMin(&T->first_error_read, packed);
        lifted = 0; ops = 0; primary = 0xffffffffu;
    }
    uint32_t n_rec = lifted;
    if (lifted == 0 && (do_finish || first_err)) n_rec = 1;  // a panicking read always yields the fallback record
    if (!do_finish) primary = 0xffffffffu - 1u;  // stage tests: no primary is chosen, no fallback
    if (live) {
        W.read_counts[r] = make_uint2(n_rec, ops);
        W.read_primary[r] = (lifted == 0) ? 0xffffffffu : primary;
    }
    // one atomic per warp, not per read (a same-address atomic per thread cost 140 us per 200k reads, ncu r01)
    uint32_t tot = lifted;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tot += __shfl_down_sync(0xffffffffu, tot, d);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(&T->n_lifted, (unsigned long long)tot);
}

// Record assembly (:245-282 field updates).  A warp owns 32 consecutive reads: lane l writes the SoA fields of read l's
// records (adjacent lanes -> adjacent records -> coalesced), then the warp copies the CIGAR ops of each record from
// its scratch slot into the dense pool cooperatively (coalesced both ways) and reduces the reference span for
// end / bin (:278-279).
__global__ void __launch_bounds__(256) emit_records_kernel(DevStatic S, DevBatch B, DevWork W, DevResult R, DevTotals* T,
                                                            uint32_t stage_mask) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = warp * 32u + lane;
    if (warp * 32u >= B.n_reads) return;
    const uint2 total = W.read_counts[B.n_reads];
    if (r == 0) {
        T->n_records = total.x;
        T->n_cigar_out = total.y;
        if (total.x > R.rec_cap) atomicOr(&T->overflow, OVF_RECORDS);
        if (total.y > R.cigar_cap) atomicOr(&T->overflow, OVF_CIGAR);
    }
    if (total.x > R.rec_cap || total.y > R.cigar_cap) return;
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
        const uint32_t s0 = B.read_seg_begin[r], s1 = B.read_seg_begin[r + 1];
        if (next.x > base.x) {
            if (primary == 0xffffffffu) {  // unmapped fallback (:317-335)
                uint16_t f = uint16_t((flag0 | 0x4) & ~0x800);
                uint8_t flip = 0;
                if (f & 0x10) { f ^= 0x10; flip = 1; }
                R.rec_status[k] = 0;
                R.rec_read_segment[k] = s0;
                R.rec_contig_segment[k] = 0xffffffffu;
                R.rec_tid[k] = -1;
                R.rec_pos[k] = -1;
                R.rec_mapq[k] = 255;
                R.rec_flag[k] = f;
                R.rec_bin[k] = B.read_bin[r];
                R.rec_need_flip[k] = flip;
                R.rec_cigar_begin[k] = op_at;
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
                const uint32_t g = W.pair_seg[p], s = W.pair_rseg[p];
                const uint8_t flip = W.pair_flip[p];
                uint16_t f = uint16_t(flag0 ^ (flip ? 0x10 : 0));
                f |= 0x800;
                if (p == primary) f &= ~0x800;
                R.rec_status[k] = 1;
                R.rec_read_segment[k] = s;
                R.rec_contig_segment[k] = g - S.contig_seg_begin[B.rseg_contig[s]];
                R.rec_tid[k] = (stage_mask & 2u) ? S.seg_chrom[g] : -2;
                R.rec_pos[k] = W.pair_pos[p];
                R.rec_mapq[k] = S.seg_mapq[g];
                R.rec_flag[k] = f;
                R.rec_bin[k] = W.pair_bin[p];
                R.rec_need_flip[k] = flip;
                R.rec_cigar_begin[k] = op_at;
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

__global__ void totals_init_kernel(DevTotals* T) { totals_reset(T); }

}  // namespace

void launch_lift(const DevStatic& S, const DevBatch& B, const DevWork& W, const DevResult& R, DevTotals* T, uint32_t stage_mask,
                 void* scan_tmp, size_t scan_tmp_bytes_, cudaStream_t st, uint64_t* launches, StageEvents* ev) {
    auto mark = [&](int i) { if (ev) cudaEventRecord(ev->e[i], st); };
    const int do_finish = (stage_mask == 7u);
    mark(0);
    if (B.n_reads == 0) {
        totals_init_kernel<<<1, 1, 0, st>>>(T);
        ++*launches;
        for (int i = 1; i < StageEvents::N; ++i) mark(i);
        return;
    }
    pair_count_kernel<<<(B.n_reads + 1 + 127) / 128, 128, 0, st>>>(S, B, W, T);
    ++*launches;
    exclusive_scan_inplace<uint32_t>(W.rseg_pair_begin, uint64_t(B.n_rsegs) + 1, scan_tmp, scan_tmp_bytes_, st, launches, nullptr);
    pair_fill_kernel<<<(B.n_rsegs + 1 + 127) / 128, 128, 0, st>>>(S, B, W, T);
    ++*launches;
    // the pair count is only known on the device: scan / launch over the capacity, kernels clamp to n_pairs
    exclusive_scan_inplace<uint64_t>(W.pair_slot_begin, uint64_t(W.pair_cap) + 1, scan_tmp, scan_tmp_bytes_, st, launches, &T->n_pairs);
    mark(1);
    lift_pairs_kernel<<<(W.pair_cap + 127) / 128, 128, 0, st>>>(S, B, W, T, stage_mask);
    ++*launches;
    if ((stage_mask & 6u) == 6u) {
        // worklist length is only known on the device: a fixed grid strides over it (a few % of the pairs)
        const unsigned blocks = unsigned(std::min<uint64_t>((uint64_t(W.pair_cap) + 127) / 128, 148ull * 16));
        simplify_pairs_kernel<<<blocks, 128, 0, st>>>(S, B, W, T);
        ++*launches;
    }
    mark(2);
    read_finalize_kernel<<<(B.n_reads + 1 + 127) / 128, 128, 0, st>>>(S, B, W, T, do_finish);
    ++*launches;
    exclusive_scan_inplace<uint2>(W.read_counts, uint64_t(B.n_reads) + 1, scan_tmp, scan_tmp_bytes_, st, launches, nullptr);
    emit_records_kernel<<<(B.n_reads + 255) / 256, 256, 0, st>>>(S, B, W, R, T, stage_mask);
    ++*launches;
    mark(3);
}

}  // namespace ptl
