// Per-pair device algorithms of the liftover path (sm_100a).  One thread owns one (read segment x contig segment)
// pair and runs the stages as streaming transducers over BAM-encoded CIGAR ops:
//
//     [left_shift]  ->  liftover  ->  [simplify]
//
// Every stage writes through an `OpSink`, which fuses the reference's two post-passes
// (clean_up_cigar_edge_indels + compress_cigar, lib/rust-vc-utils/src/bam_utils/cigar/mod.rs:204-291) into the
// write: leading-edge conversion and run merging happen as ops are pushed, the trailing edge is fixed in place at
// finish().  Integer arithmetic only; bases are compared as the exact bytes the reference compares.
//
// Execution shape (ncu r01a-r01c, profiles/): written naively, the base fetches of the homology walk were issued by ONE
// lane at a time (1.0 threads/inst, 64% of stall samples) because lanes reach their indel clusters at different op
// indices; a fully flattened state machine fixed that but tripled the instruction count.  The shape kept here:
//   - op walks are tight nested loops over L1-resident ops (ALU-bound, short divergent arms),
//   - everything that needs a DRAM/L2 round trip (homology walk, cluster trimming) is hoisted into a warp-converged
//     middle phase over a per-lane list of recorded clusters, so 32 lanes x 4 loads are in flight per warp,
//   - the liftover merge of op boundaries and table keys is one flat event loop (see run_liftover); the table block
//     under the walk lives in registers and the next entry is prefetched.
// Contig coordinates are 32-bit (BAM l_seq is int32); reference positions are widened only where they leave the loop.
#pragma once
#include <cstdint>

#include "cigar_ops.cuh"
#include "device_types.hpp"

namespace ptl {

// ------------------------------------------------------------------------------------------------------------------
// Read bases: BAM 4-bit packed, optionally viewed reverse-complemented (need_flipped_read_alignment,
// src/read_alignment_scanner.rs:153-157,170-173,238-241).  Returns the ASCII byte the reference would compare:
// rust-htslib's "=ACMGRSVTWYHKDBN" decode, then comp_base (lib/rust-vc-utils/src/seq_util.rs:1-15) when flipped.
struct ReadBases {
    const uint8_t* seq4;
    uint32_t len;
    bool flip;
    __device__ __forceinline__ uint8_t at(uint32_t idx) const {
        const uint32_t j = flip ? (len - 1u - idx) : idx;
        const uint32_t byte = seq4[j >> 1];
        const uint32_t nib = (j & 1u) ? (byte & 0xfu) : (byte >> 4);
        return decode(nib, flip);
    }
    // nibble -> ASCII, branch-free: two 16-byte tables held as 64-bit immediates.
    //   forward : "=ACMGRSVTWYHKDBN"   (rust-htslib decode)
    //   flipped : comp_base of the above = "NTGNCNNNANNNNNNN"  (A<->T, C<->G, everything else incl. '=' and IUPAC -> 'N')
    static __device__ __forceinline__ uint8_t decode(uint32_t nib, bool flipped) {
        const unsigned long long f0 = 0x565352474D43413DULL, f1 = 0x4E42444B48595754ULL;  // V S R G M C A = | N B D K H Y W T
        const unsigned long long r0 = 0x4E4E4E434E47544EULL, r1 = 0x4E4E4E4E4E4E4E41ULL;  // N N N C N G T N | N N N N N N N A
        const unsigned long long lo = flipped ? r0 : f0, hi = flipped ? r1 : f1;
        const unsigned long long w = (nib & 8u) ? hi : lo;
        return uint8_t(w >> ((nib & 7u) * 8u));
    }
};

// ------------------------------------------------------------------------------------------------------------------
// Streaming clean_up_cigar_edge_indels + compress_cigar.
// The open (not yet stored) run is kept as (pop, plen) in two registers; an empty sink holds "M of length 0", so the
// first alignment match merges into it and anything else replaces it (a zero-length run is never stored).
struct OpSink {
    uint32_t* buf;
    uint32_t cap;
    uint32_t n = 0;               // ops stored so far, counting the ones that did not fit (overflow <=> n > cap)
    uint32_t pop = 0, plen = 0;   // op code and length of the open run
    int32_t last_match_idx = -1;  // index (in buf) of the last alignment-match op, counting the open run
    bool seen_match = false;
    bool overflow = false;        // valid after finish() / close_unfinished()
    bool mixed_cluster = false;   // two adjacent stored ops are an I and a D: simplify_alignment_indels would rewrite them
    uint32_t lead_del_shift = 0;  // return value of clean_up_cigar_edge_indels
    uint32_t ref_span = 0;        // reference bases consumed by the stored ops = get_alignment_end - pos (after finish())

    __device__ __forceinline__ OpSink(uint32_t* b, uint32_t c) : buf(b), cap(c) {}

    __device__ __forceinline__ void flush() {
        if (plen) {
            if (n < cap) buf[n] = (plen << 4) | pop;
            ++n;
            if ((kRefMask >> pop) & 1u) ref_span += plen;
        }
    }
    // Branch-free: lanes of a warp push different ops at the same time, so every arm of a branchy push ran with a
    // handful of active lanes (ncu r02a: 7-10 threads per instruction on these lines).  An empty push (len == 0) falls
    // out as "merge nothing into the open run".
    __device__ __forceinline__ void push(uint32_t op, uint32_t len) {
        // leading edge (cigar/mod.rs:278-280): before the first alignment match I -> S, D -> SoftClip(0) (+ shift)
        const bool lead = !seen_match;
        const bool drop = lead && op == OP_D;
        lead_del_shift += drop ? len : 0u;
        len = drop ? 0u : len;
        op = (lead && op == OP_I) ? uint32_t(OP_S) : op;
        const bool match = op_is_match(op) && len != 0u;
        seen_match |= match;
        push_steady(op, len);
    }
    // The same without the leading-edge rules: valid once an alignment match has been pushed (seen_match).
    __device__ __forceinline__ void push_steady(uint32_t op, uint32_t len) {
        // compress_cigar (cigar/mod.rs:204-228): merge into the open run or start a new one
        const bool same = (pop == op) || len == 0u;
        const bool store = !same && plen != 0u;
        if (store && n < cap) buf[n] = (plen << 4) | pop;
        n += store ? 1u : 0u;
        ref_span += (store && ((kRefMask >> pop) & 1u)) ? plen : 0u;
        // an I/D run of the compressed CIGAR holds both kinds iff two adjacent stored ops are {I, D}
        mixed_cluster |= !same && (((1u << pop) | (1u << op)) == ((1u << OP_I) | (1u << OP_D)));
        // a Pad after a Pad is dropped, not merged (cigar/mod.rs:208-215)
        plen = same ? plen + ((op != OP_P) ? len : 0u) : len;
        pop = same ? pop : op;
        last_match_idx = (op_is_match(op) && len != 0u) ? int32_t(n) : last_match_idx;
    }
    // a stage that ends without finish() (liftover to None): only the capacity verdict is needed
    __device__ __forceinline__ void close_unfinished() { overflow = n > cap; }
    // trailing edge (cigar/mod.rs:282-288) + re-merge of what the conversion made adjacent
    __device__ __forceinline__ void finish() {
        flush();
        plen = 0;
        pop = 0;
        overflow = n > cap;
        if (overflow || last_match_idx < 0) return;  // no match at all: the leading pass already converted everything
        const uint32_t start = uint32_t(last_match_idx) + 1u;
        uint32_t w = start, prev = NO_OP;
        for (uint32_t i = start; i < n; ++i) {
            uint32_t c = buf[i];
            uint32_t op = c & 0xfu;
            if (op == OP_D) { ref_span -= c >> 4; continue; }
            if (op == OP_I) { op = OP_S; c = (c & ~0xfu) | OP_S; }
            if (prev != NO_OP && (prev & 0xfu) == op) {
                if (op != OP_P) prev += c & ~0xfu;
            } else {
                if (prev != NO_OP) buf[w++] = prev;
                prev = c;
            }
        }
        if (prev != NO_OP) buf[w++] = prev;
        n = w;
    }
};

// A stage input: `n` ops at `p`, visited forwards or backwards (the CIGAR reversal of :167 is just a stride).
struct OpSource {
    const uint32_t* p;
    uint32_t n;
    bool reversed;
    __device__ __forceinline__ uint32_t get(uint32_t i) const { return reversed ? p[n - 1u - i] : p[i]; }
};

struct PairCounters {
    uint32_t base_bytes = 0;
};

// ------------------------------------------------------------------------------------------------------------------
// n_bytes (<= 8) bytes at an arbitrary address, little-endian in the result, with one aligned 64-bit load, or two when
// the range straddles an 8-byte boundary.  An aligned word that holds at least one requested byte never leaves the
// allocation granule (256 B device, 4 KB pinned host), so no byte outside the buffer's allocation is touched.
__device__ __forceinline__ uint64_t load_bytes_le(const uint8_t* a, uint32_t n_bytes) {
    const unsigned long long addr = reinterpret_cast<unsigned long long>(a);
    const uint64_t* base = reinterpret_cast<const uint64_t*>(addr & ~7ull);
    const uint32_t off = uint32_t(addr & 7ull);
    uint64_t w = base[0] >> (off * 8u);
    if (off + n_bytes > 8u) w |= base[1] << (64u - off * 8u);  // off > 0 here
    return w;
}

// One round trip of the left homology walk: up to 8 (ref, read) base pairs are fetched with ONE word load per side (two
// if the window straddles an 8-byte boundary) and compared in order.  The first 16 read bases of a walk come from the
// cluster's indel window when the batch carries one (`have_win`; no read-side memory access at all); otherwise, and
// beyond 16 bases, they are read from the packed bases (in zero-copy mode one PCIe read per probe).
__device__ __forceinline__ void walk_homology8(const uint8_t* __restrict__ ref_seq, const ReadBases& read, uint32_t ref_end, uint32_t read_end,
                                               uint32_t limit, uint64_t win, bool have_win, uint32_t& hom, bool& walking, PairCounters& cnt) {
    const uint32_t nv = min(8u, limit - hom);  // >= 1 while walking
    // reference side: ASCII bytes ref_seq[rhi - q], q = 0..nv-1
    const uint32_t rhi = ref_end - 1u - hom, rlo = rhi - (nv - 1u);
    const uint64_t rw = load_bytes_le(ref_seq + rlo, nv);
    // read side, as the nibble of walk step hom + q in bits [4q, 4q+4) of `nibs`
    uint32_t nibs;
    if (have_win && hom < 16u) {  // hom is 0 or 8 here
        nibs = uint32_t(win >> (4u * hom));
    } else {
        // 4-bit bases at stored positions jhi - q (forward view) or jlo + q (reverse-complement view)
        uint32_t jlo, jhi;
        if (!read.flip) { jhi = read_end - 1u - hom; jlo = jhi - (nv - 1u); }
        else { jlo = read.len - read_end + hom; jhi = jlo + (nv - 1u); }
        const uint32_t b0 = jlo >> 1;
        const uint64_t qw = load_bytes_le(read.seq4 + b0, (jhi >> 1) - b0 + 1u);
        nibs = 0;
#pragma unroll
        for (uint32_t q = 0; q < 8; ++q) {
            const uint32_t j = read.flip ? jlo + q : jhi - q;  // (steps >= nv are never compared)
            nibs |= (uint32_t(qw >> ((8u * ((j >> 1) - b0) + ((j & 1u) ? 0u : 4u)) & 63u)) & 0xfu) << (4u * q);
        }
    }
#pragma unroll
    for (uint32_t q = 0; q < 8; ++q) {
        if (walking && q < nv) {
            const uint32_t rb = uint32_t(rw >> (8u * (nv - 1u - q))) & 0xffu;
            const uint32_t nib = (nibs >> (4u * q)) & 0xfu;
            cnt.base_bytes += 2;
            if (rb != uint32_t(ReadBases::decode(nib, read.flip))) walking = false;
            else ++hom;
        }
    }
    if (hom >= limit) walking = false;
}

// a5: left_shift_indels (lib/rust-vc-utils/.../shift_indels/left_shift_indels.rs:17-39) with CigarShiftBuilder in Left
// mode (cigar_indel_shifter.rs:45-164) and the left walk of get_indel_breakend_homology_info
// (indel_breakend_homology.rs:35-49).
//
// Three phases per lane, the middle one warp-converged (the base compares are the only latency-bound part):
//   A  walk the ops, record each indel cluster (ref_end, read_end, walk limit) in `clus`      -- ALU + L1 hits
//   B  for k = 0,1,2..: every lane with a k-th cluster walks its homology                     -- 32 lanes x 4 loads in flight
//   C  walk the ops again and emit, consuming the recorded homologies                          -- ALU + stores
// Only min(match_block, homology) matters (:124) and match_block_k <= gap_k + homology_{k-1}, so phase A can bound the
// walk of cluster k by limit_k = min(max_left_k, gap_k + limit_{k-1}) without knowing the homologies: same result.
// `clus` needs 3 words per cluster.  Warp-collective.  Returns the shifted position.
// `win` / `n_win`: the indel windows of the read segment (cluster order of this walk), used iff n_win == the cluster count.
__device__ __forceinline__ uint32_t run_left_shift_warp(bool active, const OpSource& in, uint32_t ref_pos, const uint8_t* __restrict__ ref_seq,
                                                        uint32_t ref_len, const ReadBases& read, const uint64_t* __restrict__ win, uint32_t n_win,
                                                        uint32_t* __restrict__ clus, OpSink& sink, PairCounters& cnt, int& err) {
    // ---- phase A
    uint32_t n_clus = 0;
    if (active) {
        uint32_t ref_head = ref_pos, read_head = 0, gap = 0, prev_limit = 0;
        uint32_t blk_ref = 0, blk_read = 0, del = 0, ins = 0;
        bool in_indel = false;
        auto close = [&]() {
            const uint32_t ref_end = blk_ref + del, read_end = blk_read + ins;
            const uint32_t max_left = min(blk_ref, blk_read);
            uint32_t limit = min(max_left, gap + prev_limit);
            // first access is ref_seq[ref_end-1] / read_seq[read_end-1]; later ones only move down and stay >= 0
            if (max_left > 0 && (ref_end > ref_len || read_end > read.len)) { err = ST_ERR_BOUNDS; limit = 0; }
            clus[3 * n_clus] = ref_end;
            clus[3 * n_clus + 1] = read_end;
            clus[3 * n_clus + 2] = limit;
            ++n_clus;
            prev_limit = limit;
            gap = 0;
            in_indel = false;
            del = 0;
            ins = 0;
        };
        // op i sits at q[i * step]; the load of op i+1 is issued before op i is looked at
        const uint32_t* q = in.reversed ? in.p + (in.n ? in.n - 1u : 0u) : in.p;
        const int step = in.reversed ? -1 : 1;
        uint32_t c_next = in.n ? q[0] : 0u;
        for (uint32_t i = 0; i < in.n; ++i) {
            const uint32_t c = c_next;
            c_next = q[int64_t(min(i + 1u, in.n - 1u)) * step];
            const uint32_t op = c & 0xfu, len = c >> 4;
            if (op == OP_D || op == OP_I) {
                if (len > 0) {
                    if (!in_indel) { in_indel = true; blk_ref = ref_head; blk_read = read_head; }
                    if (op == OP_D) del += len; else ins += len;
                }
            } else {
                if (in_indel) close();
                if (op_is_match(op)) gap += len;
                else { gap = 0; prev_limit = 0; }  // add_other flushes the match block: no carry across it
            }
            ref_head += op_ref_adv(c);
            read_head += op_read_adv(c);
        }
        if (in_indel) close();
    }
    // ---- phase B
    const bool have_win = (n_win == n_clus) && n_clus > 0;
    for (uint32_t k = 0; __any_sync(FULL, k < n_clus); ++k) {
        uint32_t hom = 0, limit = 0, ref_end = 0, read_end = 0;
        uint64_t w = 0;
        if (k < n_clus) {
            ref_end = clus[3 * k];
            read_end = clus[3 * k + 1];
            limit = clus[3 * k + 2];
            if (have_win && limit > 0) w = win[k];
        }
        bool walking = limit > 0;
        // 8 bases per round trip (almost every walk ends in the first): long walks only happen in repeats or reducible
        // I/D clusters, and they hold the whole warp
        while (__any_sync(FULL, walking)) {
            if (walking) walk_homology8(ref_seq, read, ref_end, read_end, limit, w, have_win, hom, walking, cnt);
        }
        if (k < n_clus) clus[3 * k] = hom;
    }
    // ---- phase C: every iteration ends in the same push sequence M I D M op (empty pushes cost a compare)
    if (active) {
        uint32_t match_block = 0, del = 0, ins = 0, k = 0;
        bool in_indel = false;
        const uint32_t* q = in.reversed ? in.p + (in.n ? in.n - 1u : 0u) : in.p;
        const int step = in.reversed ? -1 : 1;
        uint32_t c_next = in.n ? q[0] : uint32_t(OP_S);
        for (uint32_t i = 0; i <= in.n; ++i) {
            // the sentinel behaves like get_cigar's add_other(None): closes a trailing cluster, flushes the match block
            const uint32_t c = c_next;
            c_next = (i + 1u < in.n) ? q[int64_t(i + 1u) * step] : uint32_t(OP_S);
            const uint32_t op = c & 0xfu, len = c >> 4;
            uint32_t m1 = 0, e_ins = 0, e_del = 0, m2 = 0, o_len = 0;
            if (op == OP_D || op == OP_I) {
                if (len > 0) {
                    in_indel = true;
                    if (op == OP_D) del += len; else ins += len;
                }
            } else {
                if (in_indel) {  // end_indel (:101-148)
                    const uint32_t actual = min(match_block, clus[3 * k]);
                    ++k;
                    m1 = match_block - actual;
                    match_block = actual;
                    e_ins = ins;  // nImD order (:141-147)
                    e_del = del;
                    in_indel = false;
                    del = 0;
                    ins = 0;
                }
                if (op_is_match(op)) {
                    match_block += len;
                } else {  // add_other (:155-164)
                    m2 = match_block;
                    match_block = 0;
                    o_len = len;
                }
            }
            sink.push(OP_M, m1);
            sink.push(OP_I, e_ins);
            sink.push(OP_D, e_del);
            sink.push(OP_M, m2);
            sink.push(op, o_len);
        }
        sink.finish();
    }
    return ref_pos + sink.lead_del_shift;
}

// ------------------------------------------------------------------------------------------------------------------
// a6: liftover_read_alignment (src/liftover_read_alignment.rs:137-223) + update_ref2_cigar_segment (:35-133).
//
// The reference re-searches its BTreeMap for every reference-consuming op.  Read-op boundaries and table keys both
// advance monotonically, so one binary search per pair plus a forward merge visits exactly the same (piece, block)
// sequence.  The merge is ONE flat loop (ncu r01q: the nested op/piece loops ran at 10 active lanes per instruction;
// a warp paid max-ops x max-pieces-per-op iterations): every iteration handles one event of one lane,
//     [fetch the next op]  ->  [cross a table key]  ->  [one piece, or the verbatim I/S/H op]  ->  ONE sink push
// so a warp runs max(n_ops + keys inside ops) iterations and all lanes share the sink code.  The current block lives in
// registers and the NEXT table entry is prefetched (one 128-bit load, a crossing only moves registers).
//
// `ref2_end_pos` (:91-100) is not tracked: pieces tile the walked contig interval contiguously, so when the walk enters a
// Some block the previous Some piece ended exactly at the end of the previous aligned run, and the pushed deletion is
// the entry's precomputed `gap` (TabEntry) - valid iff some Some piece was seen before (the walk touched that run).
// Returns true if ref2_start_pos was set (Some); *out_pos = start + leading-deletion shift.
__device__ __forceinline__ bool run_liftover(const OpSource& in, uint32_t pos, const TabEntry* __restrict__ tab, uint32_t t0, uint32_t t1,
                                             uint32_t hint, OpSink& sink, int64_t* out_pos) {
    constexpr uint32_t INF = 0xffffffffu;
    // block cursor: the current block is the greatest key <= the walk position (blk_v: >= 0 Some, -1 None, -2 before
    // the first key); (nk, nv, ngap) = the prefetched NEXT entry, ti = its index
    uint32_t ti, nk = INF, ngap = 0, blk_k = 0, pgap = 0;
    int32_t blk_v = -2, nv = -1;
    {
        // first index with key > pos.  `hint` = first index with key >= the start of the pair's contig interval (from the
        // pair enumeration, which searched the table for the slot bounds anyway); pos is that start, or a little
        // beyond it when the left shift dropped a leading deletion, so the cursor is 0-1 steps away - the 13 dependent
        // loads of a binary search per pair were 5 % of the kernel's stall samples (ncu s5b).
        uint32_t lo = min(max(hint, t0), t1);
        bool have_prev = lo > t0, have_next = lo < t1;
        TabEntry prev{0, 0, 0, 0}, next{0, 0, 0, 0};
        if (have_prev) prev = tab[lo - 1];  // (both loads are in flight together)
        if (have_next) next = tab[lo];
        while (have_next && next.key <= pos) {
            prev = next;
            have_prev = true;
            ++lo;
            have_next = lo < t1;
            if (have_next) next = tab[lo];
        }
        while (have_prev && prev.key > pos) {  // (never taken for a valid hint)
            next = prev;
            have_next = true;
            --lo;
            have_prev = lo > t0;
            if (have_prev) prev = tab[lo - 1];
        }
        ti = lo;
        if (have_prev) {
            blk_k = prev.key;
            blk_v = prev.val;
        }
        if (have_next) {
            nk = next.key;
            nv = next.val;
            ngap = next.gap;
        }
    }
    bool start_set = false, some_seen = false, is_match = false;
    int32_t start = 0;  // reference positions fit int32 (BAM)
    uint32_t p = pos, e = pos;  // [p, e) = what is left of the open reference-consuming op; e == p: none open
    uint32_t i = 0, main_op = 0;
    // op i sits at q[i * step] (the CIGAR reversal of a stage test without the left shift is just a stride)
    const uint32_t* q = in.reversed ? in.p + (in.n ? in.n - 1u : 0u) : in.p;
    const int step = in.reversed ? -1 : 1;
    uint32_t c_next = in.n ? q[0] : 0u;  // software prefetch: the op the NEXT fetch will consume
    // One event of the walk up to its push: [fetch the next op] -> [cross a table key] -> [the bounds of one piece].
    // Returns false when the ops are exhausted.
    auto step_event = [&](uint32_t& m_op, uint32_t& m_len, uint32_t& plen, bool& piece) -> bool {
        const bool fetch = e == p;
        if (fetch && i == in.n) return false;
        // The prefetch is issued unconditionally at the top of the event and handed over at its END: written inside
        // the fetch arm, the compiler merged the loaded value into c_next right there and every event waited for
        // its own load (ncu s5b: 11 % of the kernel's stall samples on that one move).
        const uint32_t i_after = i + (fetch ? 1u : 0u);
        const uint32_t c_pf = q[int64_t(min(i_after, in.n - 1u)) * step];
        const uint32_t c = c_next;
        c_next = c_pf;
        if (fetch) {  // fetch the next op
            i = i_after;
            const uint32_t op = c & 0xfu, len = c >> 4;
            if ((kRefMask >> op) & 1u) {
                e = p + len;  // an empty op opens nothing: no piece (get_ref_range of an empty interval)
                is_match = op_is_match(op);
                main_op = is_match ? uint32_t(OP_M) : op;  // D stays D, N stays N, M/=/X become M (:103-107)
            } else if (op != OP_P) {  // I/S/H transfer verbatim (:157-160); Pad is ignored (:213)
                m_op = op;
                m_len = len;
            }
        }
        piece = e != p;
        if (piece) {  // one piece of the open op against the block under p (update_ref2_cigar_segment)
            if (nk == p) {  // the walk reached the next key: it becomes the current block
                blk_k = nk;
                blk_v = nv;
                pgap = ngap;
                ++ti;
                if (ti < t1) {
                    const TabEntry b = tab[ti];
                    nk = b.key;
                    nv = b.val;
                    ngap = b.gap;
                } else {
                    nk = INF;
                }
            }
            const uint32_t seg_end = min(nk, e);  // > p: keys are strictly increasing
            plen = seg_end - p;
            p = seg_end;
        }
        return true;
    };
    // ---- until the start position is known (leading clips, pieces outside aligned blocks: a few events)
    while (!start_set) {
        uint32_t m_op = 0, m_len = 0, plen = 0, gap = 0;
        bool piece = false;
        if (!step_event(m_op, m_len, plen, piece)) break;
        if (piece) {
            if (blk_v >= 0) {
                if (is_match) {  // :84-88
                    start = blk_v + int32_t(p - plen - blk_k);
                    start_set = true;
                    if (some_seen) gap = pgap;  // :91-96 (a leading deletion: the sink turns it into a position shift)
                    m_op = OP_M;
                    m_len = plen;
                }
                pgap = 0;
                some_seen = true;  // ref2_end_pos is set by every Some piece (:98-100)
            } else if (is_match) {
                m_op = (blk_v == -1) ? uint32_t(OP_I) : uint32_t(OP_S);  // None block: insertion (:111-115); no block: clip (:117-123)
                m_len = plen;
            }
        }
        if (gap) sink.push(OP_D, gap);
        sink.push(m_op, m_len);
    }
    if (!start_set) {
        sink.close_unfinished();
        return false;
    }
    // ---- steady state (ncu s5b: the one-loop-for-everything version spent 150 instructions per event): the start is
    //      set, a Some piece and an alignment match have been seen, so the leading-edge rules of both the liftover and
    //      the sink are out of the way
    for (;;) {
        uint32_t m_op = 0, m_len = 0, plen = 0, gap = 0;
        bool piece = false;
        if (!step_event(m_op, m_len, plen, piece)) break;
        if (piece) {
            if (blk_v >= 0) {
                gap = pgap;  // :91-96, only the first piece of a block can see a positive distance
                pgap = 0;
                m_op = main_op;  // :102-109
                m_len = plen;
            } else if (is_match) {
                m_op = (blk_v == -1) ? uint32_t(OP_I) : uint32_t(OP_S);
                m_len = plen;
            }
        }
        if (gap) sink.push_steady(OP_D, gap);
        sink.push_steady(m_op, m_len);
    }
    sink.finish();
    *out_pos = int64_t(start) + int64_t(sink.lead_del_shift);
    return true;
}

// ------------------------------------------------------------------------------------------------------------------
// a9: simplify_alignment_indels (src/simplify_alignment_indels.rs:119-156) with CigarBlockInfo::end_indel (:35-111).
// Same three-phase shape as the left shift: (A) walk and record the MIXED clusters (both I and D, not 1/1) in `rec`
// (4 words each), (B) warp-converged base trimming of the k-th mixed cluster of every lane, (C) walk again and emit.
// Warp-collective.  Returns the (possibly shifted) position.
__device__ __forceinline__ int64_t run_simplify_warp(bool active, const OpSource& in, int64_t ref_pos, const uint8_t* __restrict__ ref_seq,
                                                     uint64_t ref_len, const ReadBases& read, uint32_t* __restrict__ rec, OpSink& sink,
                                                     PairCounters& cnt, int& err) {
    // ---- phase A
    uint32_t n_mixed = 0;
    if (active) {
        uint32_t ref_off = 0, read_head = 0, blk_ref = 0, blk_read = 0, del = 0, ins = 0;
        bool in_indel = false;
        for (uint32_t i = 0; i <= in.n; ++i) {
            const uint32_t c = (i < in.n) ? in.get(i) : 0u;  // a sentinel match op closes a trailing cluster
            const uint32_t op = c & 0xfu, len = c >> 4;
            if (i < in.n && (op == OP_D || op == OP_I)) {
                if (!in_indel) { in_indel = true; blk_ref = ref_off; blk_read = read_head; }
                if (op == OP_D) del += len; else ins += len;
            } else if (in_indel) {
                if (del > 0 && ins > 0 && !(del == 1 && ins == 1)) {
                    rec[4 * n_mixed] = blk_ref;
                    rec[4 * n_mixed + 1] = blk_read;
                    rec[4 * n_mixed + 2] = del;
                    rec[4 * n_mixed + 3] = ins;
                    ++n_mixed;
                }
                in_indel = false;
                del = 0;
                ins = 0;
            }
            ref_off += op_ref_adv(c);
            read_head += op_read_adv(c);
        }
    }
    // ---- phase B: trim equal bases from the right first, then from the left (:55-85)
    for (uint32_t k = 0; __any_sync(FULL, k < n_mixed); ++k) {
        const bool mine = k < n_mixed;
        int64_t blk_ref = 0;
        uint32_t blk_read = 0, d = 0, n = 0, pre = 0, post = 0;
        if (mine) {
            blk_ref = ref_pos + int64_t(rec[4 * k]);
            blk_read = rec[4 * k + 1];
            d = rec[4 * k + 2];
            n = rec[4 * k + 3];
        }
        uint32_t side = mine ? 1u : 0u;  // 1 right, 2 left, 0 done
        while (__any_sync(FULL, side != 0u)) {
            if (side != 0u) {
                if (d == 0 || n == 0) {
                    side = (side == 1u) ? 2u : 0u;
                } else {
                    const int64_t f_ref = (side == 1u) ? blk_ref + int64_t(d) - 1 : blk_ref + int64_t(pre);
                    const uint32_t f_read = (side == 1u) ? blk_read + n - 1u : blk_read + pre;
                    if (f_ref < 0 || uint64_t(f_ref) >= ref_len || f_read >= read.len) {
                        err = ST_ERR_BOUNDS;  // Rust slice index panic
                        side = 0u;
                    } else {
                        const uint8_t rb = ref_seq[f_ref];
                        const uint8_t qb = read.at(f_read);
                        cnt.base_bytes += 2;
                        if (rb == qb) {
                            --d; --n;
                            if (side == 1u) ++post; else ++pre;
                        } else {
                            side = (side == 1u) ? 2u : 0u;
                        }
                    }
                }
            }
        }
        if (mine) {
            if (d == 1 && n == 1) { d = 0; n = 0; ++post; }  // down to a SNP: 1 edit instead of 2 (:88-92)
            rec[4 * k] = pre;
            rec[4 * k + 1] = post;
            rec[4 * k + 2] = d;
            rec[4 * k + 3] = n;
        }
    }
    // ---- phase C: every iteration ends in the same push sequence M I D M op
    if (active) {
        uint32_t del = 0, ins = 0, k = 0;
        bool in_indel = false;
        for (uint32_t i = 0; i <= in.n; ++i) {
            const uint32_t c = (i < in.n) ? in.get(i) : 0u;
            const uint32_t op = c & 0xfu, len = c >> 4;
            uint32_t m1 = 0, e_ins = 0, e_del = 0, m2 = 0, o_len = 0;
            if (i < in.n && (op == OP_D || op == OP_I)) {
                in_indel = true;
                if (op == OP_D) del += len; else ins += len;
            } else {
                if (in_indel) {  // end_indel (:35-111)
                    if (del == 0 || ins == 0) {
                        e_ins = ins;  // (0,0) nothing, (0,len) Ins, (len,0) Del
                        e_del = del;
                    } else if (del == 1 && ins == 1) {
                        m1 = 1;  // do not even look at the bases (:45-48)
                    } else {
                        m1 = rec[4 * k];
                        e_ins = rec[4 * k + 3];
                        e_del = rec[4 * k + 2];
                        m2 = rec[4 * k + 1];
                        ++k;
                    }
                    in_indel = false;
                    del = 0;
                    ins = 0;
                }
                if (i < in.n) o_len = len;
            }
            sink.push(OP_M, m1);
            sink.push(OP_I, e_ins);
            sink.push(OP_D, e_del);
            sink.push(OP_M, m2);
            sink.push(op, o_len);
        }
        sink.finish();
    }
    return ref_pos + int64_t(sink.lead_del_shift);
}

}  // namespace ptl
