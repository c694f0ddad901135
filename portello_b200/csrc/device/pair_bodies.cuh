// Per-index bodies of the liftover kernels (sm_100a): what ONE thread does for its segment / read / read segment / pair.
// kernels.cu wraps each body in a __global__ kernel; tests/emul compiles the same bodies for the host under a one-lane
// SIMT shim so that the kernel logic is parity-tested against the oracle without a GPU (test infrastructure only - the
// product has no CPU path).
#pragma once
#include <cstdint>

#include "device_types.hpp"
#include "lift_device.cuh"

namespace ptl {

__device__ __forceinline__ void totals_reset(DevTotals* T) {
    T->n_pairs = 0; T->scratch_needed = 0; T->n_records = 0; T->n_cigar_out = 0; T->n_lifted = 0; T->n_errors = 0;
    T->first_error_read = 0x7fffffffffffffffLL; T->first_error_status = 0; T->overflow = 0; T->n_in_ops = 0; T->n_base_bytes = 0;
    T->n_simplify = 0; T->n_long = 0;
}

// =================================================================================================== segment tables
// a7: get_read_segment_to_ref_pos_tree_map (lib/rust-vc-utils/src/bam_utils/read_to_ref_map.rs:101-137), flattened.
// A run of M/=/X ops (any other op ends it) with total length > 0 yields (run_start_read_pos -> run_start_ref_pos) and
// (run_end_read_pos -> None); the None is overwritten when the next run starts at the same read_pos (:111-119).
// `out == nullptr` counts.  `gap` of a Some entry: see TabEntry.
__device__ __forceinline__ void table_build_body(const DevStatic& S, uint32_t g, uint32_t* counts, TabEntry* out) {
    const uint32_t* c = S.seg_cigar + S.seg_cigar_begin[g];
    const uint32_t n = uint32_t(S.seg_cigar_begin[g + 1] - S.seg_cigar_begin[g]);
    int64_t ref_pos = S.seg_pos[g];
    uint64_t read_pos = 0, match_len = 0;
    uint32_t w = 0;
    TabEntry* o = out ? out + S.seg_tab_begin[g] : nullptr;
    // The overwrite decision must not depend on reading `out` (the count pass has none): track the last None key.
    uint32_t last_key = 0xffffffffu;
    bool have_last = false;
    int64_t prev_run_ref_end = 0;
    auto close_run = [&]() {
        const uint32_t k0 = uint32_t(read_pos - match_len);
        const int64_t v0 = ref_pos - int64_t(match_len);
        if (have_last && last_key == k0) --w;
        if (o) {
            const int64_t d = have_last ? v0 - prev_run_ref_end : 0;
            o[w] = TabEntry{k0, int32_t(v0), d > 0 ? uint32_t(d) : 0u, 0u};
        }
        ++w;
        if (o) o[w] = TabEntry{uint32_t(read_pos), -1, 0u, 0u};
        ++w;
        last_key = uint32_t(read_pos);
        have_last = true;
        prev_run_ref_end = ref_pos;
        match_len = 0;
    };
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t x = c[i];
        if (op_is_match(x & 0xfu)) match_len += x >> 4;
        else if (match_len > 0) close_run();
        ref_pos += op_ref_adv(x);
        read_pos += op_read_adv(x);
    }
    if (match_len > 0) close_run();
    if (!out) counts[g] = w;
}

// =================================================================================================== pair enumeration
// a3 (count): get_contig_split_segments_from_read_mapping (src/read_alignment_scanner.rs:80-103) + get_cigar_ref_offset,
// for the 1..k split segments of read r.
// Also validates what the reference's slice indexing would panic on (a malformed batch, not a malformed alignment): a
// segment whose contig, CIGAR range, position or base range lies outside its pool gets no pairs and raises OVF_INVALID.
__device__ __forceinline__ void pair_count_body(const DevStatic& S, const DevBatch& B, const DevWork& W, DevTotals* T, uint32_t r) {
    const uint32_t s0 = B.read_seg_begin[r], s1 = B.read_seg_begin[r + 1];
    bool bad = s0 > s1 || s1 > B.n_rsegs || !range_in_pool(B.read_seq_off[r], (uint64_t(B.read_seq_len[r]) + 1u) / 2u, B.seq4_bytes);
    for (uint32_t s = s0; s < s1 && s < B.n_rsegs; ++s) {
        W.rseg_read[s] = r;
        const uint64_t c0 = B.rseg_cigar_begin[s];
        const uint32_t n = B.rseg_cigar_len[s];
        const int64_t pos = B.rseg_pos[s];
        if (bad || B.rseg_contig[s] >= S.n_contigs || !range_in_pool(c0, n, B.n_cigar) || pos < 0 || pos > 0x7fffffffLL) {
            bad = true;
            W.rseg_ref_len[s] = 0;
            W.rseg_n_id[s] = 0;
            W.rseg_read_len[s] = 0;
            W.rseg_pair_begin[s] = 0;
            continue;
        }
        const uint32_t* c = B.cigar + c0;
        int64_t ref_len = 0;
        uint32_t n_id = 0, read_len = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t x = c[i];
            ref_len += op_ref_adv(x);
            read_len += op_read_adv(x);
            n_id += op_is_match(x & 0xfu) ? 0u : 1u;  // I/D ops, and every other non-match op (each can end a match block)
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
    if (bad) atomicOr(&T->overflow, OVF_INVALID);
}

__device__ __forceinline__ uint32_t lower_bound_key(const TabEntry* tab, uint32_t lo, uint32_t hi, int64_t key) {
    while (lo < hi) {  // first index with tab.key >= key
        const uint32_t mid = (lo + hi) >> 1;
        if (int64_t(tab[mid].key) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// a3 (fill): pair list of read segment s in (read segment, contig segment index) order + the scratch-slot bound of each pair.
__device__ __forceinline__ void pair_fill_body(const DevStatic& S, const DevBatch& B, const DevWork& W, uint32_t s) {
    uint32_t p = W.rseg_pair_begin[s];
    const uint32_t p_end = W.rseg_pair_begin[s + 1];
    if (p == p_end) return;
    const int64_t start = B.rseg_pos[s], ref_len = W.rseg_ref_len[s], end = start + ref_len;
    const uint32_t ctg = B.rseg_contig[s];
    const uint32_t n_in = B.rseg_cigar_len[s];
    // the per-segment facts of the pair descriptors, loaded up front (independent of the table searches below, whose
    // dependent loads they overlap with; loaded after them they cost pair_fill_kernel 30 % more time)
    const uint32_t r = W.rseg_read[s];
    const uint16_t d_flag = B.read_flag[r];
    const uint32_t d_seq_len = B.read_seq_len[r];
    const uint64_t d_seq_off = B.read_seq_off[r];
    const bool d_seg_fwd = B.rseg_is_fwd[s] != 0;
    const uint32_t d_read_len = W.rseg_read_len[s];
    const uint64_t d_cigar_begin = B.rseg_cigar_begin[s];
    const uint64_t d_rev_off = S.contig_rev_off[ctg];
    const uint32_t d_contig_len = uint32_t(S.contig_len[ctg]);
    uint32_t d_w0 = 0, d_nw = 0;
    if (B.rseg_win_begin) { d_w0 = B.rseg_win_begin[s]; d_nw = B.rseg_win_begin[s + 1] - d_w0; }
    for (uint32_t g = S.contig_seg_begin[ctg]; g < S.contig_seg_begin[ctg + 1]; ++g) {
        if (!(end >= int64_t(S.seg_so_start[g]) && start < int64_t(S.seg_so_end[g]))) continue;
        if (p < W.pair_cap) {
            W.pair_rseg[p] = s;
            W.pair_seg[p] = g;
            // contig interval walked by the liftover, in the orientation of the segment's table
            const bool fwd = S.seg_is_fwd[g] != 0;
            const int64_t a = fwd ? start : int64_t(S.contig_len[ctg]) - end;
            const uint32_t t0 = S.seg_tab_begin[g], t1 = S.seg_tab_begin[g + 1];
            const uint32_t tab_lo = lower_bound_key(S.table, t0, t1, a);
            const uint32_t n_keys = lower_bound_key(S.table, t0, t1, a + ref_len) - tab_lo;
            W.pair_tab_lo[p] = tab_lo;  // the liftover starts its table cursor here instead of searching again
            // op-slot bounds, in stored (compressed) ops:
            //   shifted    <= n_in + n_id + 1            (each I/D op can split one match block in two)
            //   lifted     <= shifted + 2 n_keys          (one extra piece and one gap-D per table key in range)
            //   simplified <= lifted + 2 (n_id + n_keys)  (a mixed cluster grows by <= 2 ops and owns >= 1 D op)
            const uint32_t n_id = W.rseg_n_id[s];
            const uint32_t n_shift = fwd ? n_in : n_in + n_id + 1u;
            //   buffer B doubles as the cluster list of the left shift (3 words per cluster), buffer A keeps 4 words per
            //   mixed cluster of the simplify stage at its top end
            //   (a long pair on a reverse-strand segment keeps the 6-word event records of the warp left shift there:
            //   one per I/D cluster and per other non-match op, plus the end)
            uint32_t cap_b = max(n_shift + 2u * n_keys + 4u, fwd ? 0u : 3u * n_id + 4u);
            if (!fwd && n_in > W.long_ops) cap_b = max(cap_b, 6u * (n_id + 1u) + 8u);
            //   (a long pair simplifies on the warp path: A = [lifted ops + 3 per I/D cluster | 3 words per mixed cluster])
            const uint32_t cap_a = cap_b + (n_in > W.long_ops ? 9u * (n_id + n_keys) + 16u : 6u * (n_id + n_keys) + 8u);
            W.pair_cap_b[p] = cap_b;
            W.pair_slot_begin[p] = uint64_t(cap_a) + cap_b;  // [0,cap_b) = buffer B, [cap_b, cap_b+cap_a) = buffer A
            // everything lift_pairs_kernel needs before the walk, in one sector (two for a reverse-strand pair)
            const bool rec_rev = (d_flag & 0x10) != 0;
            const bool changes_strand = (rec_rev == d_seg_fwd);
            const bool need_flip = (!fwd) != changes_strand;
            const uint32_t flags = (fwd ? kPdContigFwd : 0u) | (need_flip ? kPdNeedFlip : 0u) | (a < 0 ? kPdErrBounds : 0u) |
                                   (d_read_len != d_seq_len ? kPdErrLength : 0u) | ((!fwd && d_rev_off == ~0ull) ? kPdNoRevSeq : 0u);
            W.pair_desc[p] = PairDesc{d_cigar_begin, n_in, uint32_t(a), tab_lo, t0, t1, flags | (min(cap_b, 0xffffffu) << 8)};
            if (!fwd) W.pair_desc_rev[p] = PairDescRev{d_seq_off, d_rev_off, d_contig_len, d_seq_len, d_w0, d_nw};
        }
        ++p;
    }
}

// =================================================================================================== the pair
// a4 + a5 + a6 + a8 (+ a9 inline in stage tests): get_liftover_alignment_for_read_and_contig_segment
// (src/read_alignment_scanner.rs:136-288) for pair p.  Warp-collective: all 32 lanes enter (idle lanes carry
// valid = false) so that the latency-bound base fetches of a warp are issued together (see run_left_shift_warp).
// Adds the pair's roofline counters (input ops walked, base bytes compared) to n_in_ops / n_base_bytes.
// kAllStages: the production instantiation (stage_mask == PTL_STAGE_ALL); the stage-test paths (a stage switched off,
// simplify inline, verbatim hand-back) compile away, which is worth registers in the hot kernel.
//
// Staging (the CUDA kernel only; `pool == nullptr` keeps every buffer in the pair's own scratch slot, which is what the
// host emulation runs): the lifted CIGAR of the 32 pairs of a warp is written into a shared-memory pool, lanes taking
// consecutive pieces of it sized by their slot bound (a warp scan); a pair whose piece does not fit keeps its global slot.
// The caller then moves the staged outputs of the warp into a dense region with coalesced stores (LiftOut says where the
// pair's final ops are), so that the record emission reads dense, adjacent CIGARs instead of 7x-sparse scratch slots.
struct StagePool {
    uint32_t* pool = nullptr;  // shared memory, one per warp
    uint32_t words = 0;
};
struct LiftOut {
    bool staged = false;            // the final ops sit in the pool: the caller must store them and set pair_out_off
    uint32_t n = 0;                 // number of final ops (0 unless the pair was lifted here)
    const uint32_t* src = nullptr;  // where they are (pool or slot)
    uint64_t slot0 = 0;             // the pair's scratch slot (spill target), capacity cap_b ops
    uint32_t cap_b = 0;
};
template <bool kAllStages>
__device__ __forceinline__ LiftOut lift_pair_body(const DevStatic& S, const DevBatch& B, const DevWork& W, DevTotals* T, uint32_t p, bool valid,
                                                  uint32_t stage_mask_in, uint32_t& n_in_ops, uint32_t& n_base_bytes,
                                                  const StagePool stage = StagePool{}) {
    const uint32_t stage_mask = kAllStages ? 7u : stage_mask_in;
    PairCounters cnt;
    int status = ST_NONE, err = 0;
    uint32_t cpos = 0;   // position on the contig strand the segment's table is written in
    int64_t rpos = 0;    // position on the reference once lifted
    bool need_flip = false, contig_fwd = true, usable = false;
    uint32_t cap_a = 0, cap_b = 0;
    uint32_t* buf_a = nullptr;
    uint32_t* buf_b = nullptr;
    OpSource cur{nullptr, 0, false};
    ReadBases read{nullptr, 0, false};
    uint64_t slot0 = 0;
    PairDesc d{0, 0, 0, 0, 0, 0, 0};
    PairDescRev dr{0, ~0ull, 0, 0, 0, 0};
    if (valid) {
        // one sector says everything about the pair (pair_fill_body packed it); the slot bounds sit next to each other
        d = W.pair_desc[p];
        slot0 = W.pair_slot_begin[p];
        const uint64_t slot1 = W.pair_slot_begin[p + 1];
        contig_fwd = (d.flags_cap_b & kPdContigFwd) != 0u;
        need_flip = (d.flags_cap_b & kPdNeedFlip) != 0u;
        if (slot1 > W.scratch_cap) {
            atomicOr(&T->overflow, OVF_SCRATCH);
            err = ST_ERR_CAPACITY;
        } else {
            usable = true;
            cap_b = d.flags_cap_b >> 8;
            cap_a = uint32_t(slot1 - slot0) - cap_b;
            buf_b = W.scratch + slot0;
            buf_a = buf_b + cap_b;
            cur = OpSource{B.cigar + d.cigar_begin, d.n_ops, false};
            n_in_ops += cur.n;
            status = ST_LIFTED;
            cpos = d.cpos;  // reverse-strand contig segment: contig_len - end, on the contig's reverse strand (:162-167)
            if (!contig_fwd) {
                if (d.flags_cap_b & kPdErrBounds) { err = ST_ERR_BOUNDS; usable = false; }  // read runs past the contig end (invalid input)
                cur.reversed = true;
                dr = W.pair_desc_rev[p];
                read = ReadBases{B.seq4 + dr.seq_off, dr.seq_len, need_flip};
            } else if (!kAllStages) {  // (stage tests simplify forward pairs inline: they need the read's bases too)
                const uint32_t r = W.rseg_read[W.pair_rseg[p]];
                read = ReadBases{B.seq4 + B.read_seq_off[r], B.read_seq_len[r], need_flip};
            }
        }
    }
    bool cur_is_a = false, cur_is_raw = true;
    // ---- long CIGARs: left shift and liftover run in the warp-cooperative kernel (lift_warp.cuh), lanes over ops
    //      (the lane stays in this function: the stages below are warp-collective)
    const bool parked = ((stage_mask & 2u) || stage_mask == 1u) && usable && !err && cur.n > W.long_ops;
    if (parked) {
        W.long_list[atomicAdd(&T->n_long, 1u)] = p;
        usable = false;
    }
    bool simplify_is_identity = false;
    uint32_t span = 0;  // reference span of the final CIGAR (end = pos + span, :278)

    // ---- a5: left-shift on the contig's reverse strand (:168-175)
    {
        bool go = usable && !contig_fwd && (stage_mask & 1u);
        if (go && (d.flags_cap_b & kPdNoRevSeq)) { err = ST_ERR_BOUNDS; go = false; usable = false; }  // Option::unwrap on None (:174)
        if (__any_sync(FULL, go)) {
            OpSink sink(buf_a, go ? cap_a : 0u);
            const uint64_t* win = (go && B.rseg_win_begin) ? B.indel_win + dr.win_begin : nullptr;
            const uint32_t n_win = (go && B.rseg_win_begin) ? dr.n_win : 0u;
            const uint32_t shifted = run_left_shift_warp(go, cur, cpos, go ? S.rev_pool + dr.rev_off : nullptr,
                                                         go ? dr.contig_len : 0u, read, win, n_win, buf_b, sink, cnt, err);
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
    uint32_t* lift_buf = buf_b;
    bool staged = false;
    if (stage.pool) {  // (warp-uniform branch; every lane takes part in the scan)
        const bool want = usable && !err && (stage_mask & 2u);
        const uint32_t need = want ? cap_b : 0u;
        uint32_t incl = need;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, incl, d);
            if (int(threadIdx.x & 31u) >= d) incl += o;
        }
        if (want && incl <= stage.words) {
            lift_buf = stage.pool + (incl - need);
            staged = true;
        }
    }
    if (usable && !err && (stage_mask & 2u)) {
        OpSink sink(lift_buf, cap_b);
        int64_t lifted_pos = 0;
        const bool some = run_liftover(cur, cpos, S.table, d.t0, d.t1, d.tab_lo, sink, &lifted_pos);
        if (sink.overflow) err = ST_ERR_CAPACITY;
        else if (!some) status = ST_NONE;
        else if (d.flags_cap_b & kPdErrLength) err = ST_ERR_LENGTH;
        // simplify_alignment_indels rewrites only I/D runs that hold both kinds; on a cleaned + compressed CIGAR without
        // such a run it is the identity (single-kind runs are already one op, edges are already clean), so it is skipped
        simplify_is_identity = !sink.mixed_cluster;
        span = sink.ref_span;
        rpos = lifted_pos;
        cur = OpSource{lift_buf, sink.n, false};
        cur_is_a = false;
        cur_is_raw = false;
    }
    // ---- a9: simplify (:236-243).  Only the few pairs whose lifted CIGAR holds a mixed I/D run need it: they are
    //      appended to a worklist and finished warp-per-pair by warp_pairs_kernel (lift_warp.cuh) (inline, the stage ran at
    //      1.9 active threads per instruction and cost 22 % of this kernel, ncu r01g).
    bool deferred = false;
    if (usable && !err && status == ST_LIFTED && (stage_mask & 4u) && !simplify_is_identity) {
        if (stage_mask & 2u) {
            deferred = true;
            W.simplify_list[atomicAdd(&T->n_simplify, 1u)] = p;
        }
    }
    if (!kAllStages) {   // stage tests without the liftover stage: simplify inline, (raw input or A) -> A
        const bool go = usable && !err && status == ST_LIFTED && (stage_mask & 4u) && !(stage_mask & 2u);
        if (__any_sync(FULL, go)) {
            const uint8_t* ref = nullptr;
            uint64_t ref_len = 0;
            uint32_t* rec = nullptr;
            if (go) {
                const int32_t chrom = S.seg_chrom[W.pair_seg[p]];
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
    if (!kAllStages && usable && !err && status == ST_LIFTED && cur_is_raw) {
        // stage tests with every stage disabled for this pair: hand the (possibly reversed) input back verbatim
        const uint32_t n = min(cur.n, cap_a);
        for (uint32_t i = 0; i < n; ++i) { const uint32_t c = cur.get(i); buf_a[i] = c; span += op_ref_adv(c); }
        cur = OpSource{buf_a, n, false};
    }
    LiftOut out;
    if (parked) {
        W.pair_status[p] = int8_t(ST_PENDING_LIFT);
        W.pair_flip[p] = need_flip;
        W.pair_pos[p] = cpos;
        W.pair_n_out[p] = cur.n;
    } else if (valid) {
        if (err) status = err;
        const bool ok = (status == ST_LIFTED);
        if (ok && deferred) status = ST_PENDING_SIMPLIFY;
        W.pair_status[p] = int8_t(status);
        W.pair_flip[p] = need_flip;
        W.pair_pos[p] = ok ? rpos : 0;
        W.pair_n_out[p] = ok ? cur.n : 0u;
        // final ops in the pool (cur.p == lift_buf when the liftover was the last stage to run here): the caller stores them
        out.staged = ok && staged && cur.p == lift_buf;
        if (!out.staged) W.pair_out_off[p] = (ok && !staged) ? uint64_t(cur.p - W.scratch) : slot0;
        W.pair_bin[p] = ok ? reg2bin(rpos, rpos + int64_t(span)) : uint16_t(0);  // bam_reg2bin(pos, end) (:278-279)
        out.n = ok ? cur.n : 0u;
        out.src = cur.p;
        out.slot0 = slot0;
        out.cap_b = cap_b;
    }
    n_base_bytes += cnt.base_bytes;
    return out;
}

// a9 for a pair parked on the simplify worklist by lift_pair_body (its lifted CIGAR holds a mixed I/D run), ONE THREAD per
// listed pair: the worklist is dense, so every lane of a warp has work (the same stage inline in the lift kernel ran at 1.9
// active lanes, and a warp per pair spent 32 lanes on ~22 ops).  Warp-collective (run_simplify_warp: the base probes of
// the k-th mixed cluster of all lanes are issued together).  `i` indexes simplify_list; lanes past its end carry active = false.
__device__ __forceinline__ void simplify_thread_pair_body(const DevStatic& S, const DevBatch& B, const DevWork& W, DevTotals* T, uint32_t i,
                                                          bool active, uint32_t& n_base_bytes) {
    uint32_t p = 0, cap_a = 0, cap_b = 0, n_rec = 0;
    uint64_t slot0 = 0;
    uint32_t* buf_a = nullptr;
    uint32_t* rec = nullptr;
    OpSource cur{nullptr, 0, false};
    ReadBases read{nullptr, 0, false};
    const uint8_t* ref = nullptr;
    uint64_t ref_len = 0;
    int64_t rpos = 0;
    if (active) {
        p = W.simplify_list[i];
        const uint32_t s = W.pair_rseg[p], g = W.pair_seg[p];
        const uint32_t r = W.rseg_read[s];
        slot0 = W.pair_slot_begin[p];
        cap_b = W.pair_cap_b[p];
        cap_a = uint32_t(W.pair_slot_begin[p + 1] - slot0) - cap_b;
        buf_a = W.scratch + slot0 + cap_b;
        // buffer A: [simplified ops | 4 words per mixed cluster]; sized by pair_fill_body from (n_id + n_keys) possible clusters
        n_rec = 4u * ((cap_a - cap_b - 8u) / 6u);
        rec = buf_a + (cap_a - n_rec);
        cur = OpSource{W.scratch + W.pair_out_off[p], W.pair_n_out[p], false};  // where lift_pairs_kernel left the lifted CIGAR
        read = ReadBases{B.seq4 + B.read_seq_off[r], B.read_seq_len[r], W.pair_flip[p] != 0};
        const int32_t chrom = S.seg_chrom[g];
        ref = S.ref + S.chrom_off[chrom];
        ref_len = S.chrom_off[chrom + 1] - S.chrom_off[chrom];
        rpos = W.pair_pos[p];
    }
    PairCounters cnt;
    int err = 0;
    OpSink sink(buf_a, active ? cap_a - n_rec : 0u);
    const int64_t out_pos = run_simplify_warp(active, cur, rpos, ref, ref_len, read, rec, sink, cnt, err);
    if (active) {
        int status = err;
        if (!status && sink.overflow) status = ST_ERR_CAPACITY;
        W.pair_status[p] = int8_t(status ? status : ST_LIFTED);
        W.pair_pos[p] = status ? 0 : out_pos;
        W.pair_n_out[p] = status ? 0u : sink.n;
        W.pair_out_off[p] = slot0 + cap_b;
        W.pair_bin[p] = status ? uint16_t(0) : reg2bin(out_pos, out_pos + int64_t(sink.ref_span));
    }
    n_base_bytes += cnt.base_bytes;
}

// =================================================================================================== per-read finish
// a10 (field part): finish_remapped_alignment_set (src/read_alignment_scanner.rs:310-366) for read r: record counts,
// primary = first max MAPQ (:338-346), unmapped fallback when nothing lifted (:317-335).  Returns the lifted count.
__device__ __forceinline__ uint32_t read_finalize_body(const DevStatic& S, const DevBatch& B, const DevWork& W, DevTotals* T, uint32_t r,
                                                       int do_finish) {
    uint32_t lifted = 0, ops = 0, primary = 0xffffffffu;
    int best_mapq = -1, first_err = 0;
    // (clamped: a malformed batch is rejected by the host afterwards, but must not be dereferenced out of bounds here)
    const uint32_t s0 = min(B.read_seg_begin[r], B.n_rsegs), s1 = min(max(B.read_seg_begin[r + 1], s0), B.n_rsegs);
    const uint32_t p0 = W.rseg_pair_begin[s0], p1 = min(W.rseg_pair_begin[s1], W.pair_cap);
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
    W.read_counts[r] = make_uint2(n_rec, ops);
    W.read_primary[r] = (lifted == 0) ? 0xffffffffu : primary;
    return lifted;
}

// Record fields of the unmapped fallback of read r at record k (:317-335).
__device__ __forceinline__ void emit_unmapped_record(const DevBatch& B, const DevResult& R, uint32_t r, uint32_t k, uint32_t s0, uint64_t op_at) {
    uint16_t f = uint16_t((B.read_flag[r] | 0x4) & ~0x800);
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
}

// Record fields of lifted pair p at record k (:245-282 field updates).
__device__ __forceinline__ void emit_lifted_record(const DevStatic& S, const DevBatch& B, const DevWork& W, const DevResult& R, uint32_t p,
                                                   uint32_t k, uint16_t flag0, uint32_t primary, uint64_t op_at, uint32_t stage_mask) {
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
}

}  // namespace ptl
