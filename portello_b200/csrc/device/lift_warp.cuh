// Warp-cooperative liftover (sm_100a): ONE WARP per (read segment x contig segment) pair, lanes over CIGAR ops.
//
//   a6  liftover_read_alignment + update_ref2_cigar_segment   (src/liftover_read_alignment.rs:35-223)
//   a8  clean_up_cigar_edge_indels + compress_cigar            (lib/rust-vc-utils/src/bam_utils/cigar/mod.rs:204-291)
//
// The per-thread walk of lift_device.cuh::run_liftover pays one dependent load per op and leaves a 100 kb read (hundreds
// of ops) to a single lane; here the same result is computed as data-parallel steps over chunks of 32 ops:
//
//   1. coalesced load of 32 ops; warp scan of their reference advance -> contig interval [p_i, e_i) of every op
//   2. the table entries inside the chunk's interval are loaded ONCE, one per lane (the "window"); every lane counts the
//      keys <= p_i and < e_i by register shuffles -> block under p_i and number of pieces of its op
//   3. warp scan of the piece counts; lanes are re-assigned to PIECES (load-balancing search over the scan), so an op
//      that crosses many table keys costs rounds, not a serial inner loop
//   4. every piece yields at most two ops (the deletion pushed when the walk enters an aligned run, then the piece
//      itself); the sequential state of the reference (ref2_start_pos set / ref2_end_pos set) is a prefix-OR = a ballot
//   5. WarpSink: edge clean-up + run compression of the 64 candidate ops of a round with ballots and one scan, carrying the
//      open run between rounds; complete runs go to memory with one coalesced store
//
// Only min() / compares / adds on integers; identical results to run_liftover by construction of the same piece
// sequence (tests: every golden liftover vector + randomised parity against the oracle through both paths).
#pragma once
#include <cstdint>

#include "cigar_ops.cuh"
#include "device_types.hpp"
#include "lift_device.cuh"  // ReadBases, walk_homology8, PairCounters

#ifndef PTL_ASSERT_UNIFORM
#define PTL_ASSERT_UNIFORM(v)  // (the lock-step CPU emulation of tests/emul checks warp-uniform state here)
#endif

namespace ptl {

__device__ __forceinline__ uint32_t lanes_lt(uint32_t lane) { return (1u << lane) - 1u; }
__device__ __forceinline__ uint32_t lanes_le(uint32_t lane) { return (2u << lane) - 1u; }  // lane 31: 0 - 1 = all

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, v, d);
        if (int(lane) >= d) v += o;
    }
    return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// ------------------------------------------------------------------------------------------------------------------
// Streaming clean_up_cigar_edge_indels + compress_cigar for a warp.  Every round each lane offers up to two ops in
// order: A = Del(a_len) then B = (b_op, b_len); empty ops (len 0) vanish.  Uniform state is identical in all lanes.
struct WarpSink {
    uint32_t* buf;
    uint32_t cap;
    uint32_t n = 0;               // stored ops (uniform)
    uint32_t pend = 0;            // open run (len << 4 | op), len 0 = none (uniform)
    int32_t last_match_idx = -1;  // index of the last alignment-match run, counting the open one (uniform)
    bool seen_match = false;      // (uniform)
    // per-lane partial sums, reduced once in finish()
    uint32_t acc_span = 0, acc_shift = 0;
    bool lane_mixed = false;
    // results of finish() (uniform)
    bool overflow = false, mixed_cluster = false;
    uint32_t ref_span = 0, lead_del_shift = 0;

    __device__ __forceinline__ WarpSink(uint32_t* b, uint32_t c) : buf(b), cap(c) {}

    __device__ __forceinline__ void store(uint32_t idx, uint32_t word) {
        if (idx < cap) buf[idx] = word;
    }

    __device__ __forceinline__ void push2(uint32_t lane, uint32_t a_len, uint32_t b_op, uint32_t b_len) {
        // leading edge (cigar/mod.rs:278-280): before the first alignment match I -> S, D -> dropped (+ position shift)
        const bool b_match = b_len != 0u && op_is_match(b_op);
        const uint32_t mmask = __ballot_sync(FULL, b_match);
        if (!seen_match && (mmask & lanes_lt(lane)) == 0u) {
            acc_shift += a_len;
            a_len = 0;
            if (!b_match) {
                if (b_op == OP_D) { acc_shift += b_len; b_len = 0; }
                else if (b_op == OP_I) b_op = OP_S;
            }
        }
        seen_match = seen_match || mmask != 0u;
        acc_span += a_len + (((kRefMask >> b_op) & 1u) ? b_len : 0u);
        // compress_cigar (cigar/mod.rs:204-228) over the 64 candidates + the open run
        const bool nzA = a_len != 0u, nzB = b_len != 0u;
        const uint32_t any_mask = __ballot_sync(FULL, nzA || nzB);
        if (any_mask == 0u) return;
        const uint32_t before = any_mask & lanes_lt(lane);
        const uint32_t last_op = nzB ? b_op : uint32_t(OP_D);
        uint32_t prev_op = __shfl_sync(FULL, last_op, before ? (31 - __clz(before)) : 0);
        if (!before) prev_op = (pend >> 4) ? (pend & 0xfu) : 0xffu;
        const bool headA = nzA && prev_op != OP_D;
        const uint32_t prev_b = nzA ? uint32_t(OP_D) : prev_op;
        const bool headB = nzB && b_op != prev_b;
        // a Pad after a Pad is dropped, not merged (cigar/mod.rs:208-215)
        const uint32_t cb = (b_op == OP_P && !headB) ? 0u : b_len;
        // an I/D run of the compressed CIGAR holds both kinds iff two adjacent runs are {I, D}
        lane_mixed = lane_mixed || (headA && prev_op == OP_I) || (headB && ((b_op == OP_I && prev_b == OP_D) || (b_op == OP_D && prev_b == OP_I)));
        const uint32_t t = a_len + cb;
        const uint32_t incl = warp_incl_scan(t, lane);
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        const uint32_t posA = incl - t, posB = posA + a_len;  // flat length prefix in front of A / B
        const uint32_t hA = __ballot_sync(FULL, headA), hB = __ballot_sync(FULL, headB), hAny = hA | hB;
        if (hAny == 0u) {  // everything merges into the open run
            pend += total << 4;
            return;
        }
        const uint32_t fh = headA ? posA : posB;  // prefix in front of this lane's first head
        const uint32_t after = hAny & ~lanes_le(lane);
        uint32_t nextpos = __shfl_sync(FULL, fh, after ? (__ffs(after) - 1) : 0);
        if (!after) nextpos = total;
        const uint32_t segA = headA ? ((headB ? posB : nextpos) - posA) : 0u;
        const uint32_t segB = headB ? (nextpos - posB) : 0u;
        const uint32_t n_heads = __popc(hA) + __popc(hB);
        const uint32_t idxA = __popc(hA & lanes_lt(lane)) + __popc(hB & lanes_lt(lane));
        const uint32_t idxB = idxA + (headA ? 1u : 0u);
        // the open run absorbs what precedes the first head, then is stored
        const uint32_t fpos = __shfl_sync(FULL, fh, __ffs(hAny) - 1);
        const uint32_t closed = pend + (fpos << 4);
        uint32_t base = n;
        if (closed >> 4) {
            if (lane == 0u) store(base, closed);
            ++base;
        }
        // every head but the last is a complete run; the last one stays open
        if (headA && idxA != n_heads - 1u) store(base + idxA, (segA << 4) | OP_D);
        if (headB && idxB != n_heads - 1u) store(base + idxB, (segB << 4) | b_op);
        const uint32_t my_last = headB ? ((segB << 4) | b_op) : ((segA << 4) | OP_D);
        pend = __shfl_sync(FULL, my_last, 31 - __clz(hAny));
        n = base + n_heads - 1u;
        const uint32_t mh = __ballot_sync(FULL, headB && op_is_match(b_op));
        if (mh) last_match_idx = int32_t(__shfl_sync(FULL, base + idxB, 31 - __clz(mh)));
        PTL_ASSERT_UNIFORM(n); PTL_ASSERT_UNIFORM(pend); PTL_ASSERT_UNIFORM(last_match_idx);
    }

    // flush + trailing edge (cigar/mod.rs:282-288) + re-merge of what the conversion made adjacent
    __device__ __forceinline__ void finish(uint32_t lane) {
        if (pend >> 4) {
            if (lane == 0u) store(n, pend);
            ++n;
        }
        pend = 0;
        overflow = n > cap;
        ref_span = warp_sum(acc_span);
        lead_del_shift = warp_sum(acc_shift);
        mixed_cluster = __any_sync(FULL, lane_mixed);
        __syncwarp();
        if (overflow || last_match_idx < 0) return;
        // the tail behind the last match is a handful of ops: lane 0 rewrites it, the others take its counts
        uint32_t w = n, dropped = 0;
        if (lane == 0u) {
            const uint32_t start = uint32_t(last_match_idx) + 1u;
            uint32_t prev = NO_OP;
            w = start;
            for (uint32_t i = start; i < n; ++i) {
                uint32_t c = buf[i];
                uint32_t op = c & 0xfu;
                if (op == OP_D) { dropped += c >> 4; continue; }
                if (op == OP_I) { op = OP_S; c = (c & ~0xfu) | OP_S; }
                if (prev != NO_OP && (prev & 0xfu) == op) {
                    if (op != OP_P) prev += c & ~0xfu;
                } else {
                    if (prev != NO_OP) buf[w++] = prev;
                    prev = c;
                }
            }
            if (prev != NO_OP) buf[w++] = prev;
        }
        n = __shfl_sync(FULL, w, 0);
        ref_span -= __shfl_sync(FULL, dropped, 0);
        PTL_ASSERT_UNIFORM(n); PTL_ASSERT_UNIFORM(ref_span);
        __syncwarp();
    }
};

// ------------------------------------------------------------------------------------------------------------------
// First index in [lo, hi) whose key is > pos (kUpper) or >= pos (!kUpper): 32-ary search, the lanes probe 32 evenly
// spaced entries per round (a 2^17-entry segment table takes 4 rounds instead of 17 dependent loads).  Uniform result.
template <bool kUpper>
__device__ __forceinline__ uint32_t warp_table_bound(const TabEntry* __restrict__ tab, uint32_t lo, uint32_t hi, uint32_t pos, uint32_t lane) {
    while (lo < hi) {
        const uint32_t step = (hi - lo + 31u) >> 5;
        const uint32_t idx = lo + (lane + 1u) * step - 1u;
        bool below = false;
        if (idx < hi) {
            const uint32_t k = tab[idx].key;
            below = kUpper ? (k <= pos) : (k < pos);
        }
        const uint32_t c = __popc(__ballot_sync(FULL, below));  // monotone: a prefix of the probes
        const uint32_t lo_old = lo;
        lo = lo_old + c * step;
        hi = min(hi, lo_old + (c + 1u) * step - 1u);
    }
    return lo;
}

__device__ __forceinline__ uint32_t lane_table_bound_upper(const TabEntry* __restrict__ tab, uint32_t lo, uint32_t hi, uint32_t pos) {
    while (lo < hi) {  // first index with key > pos
        const uint32_t mid = (lo + hi) >> 1;
        if (tab[mid].key <= pos) lo = mid + 1u; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ uint32_t lane_table_bound_lower(const TabEntry* __restrict__ tab, uint32_t lo, uint32_t hi, uint32_t pos) {
    while (lo < hi) {  // first index with key >= pos
        const uint32_t mid = (lo + hi) >> 1;
        if (tab[mid].key < pos) lo = mid + 1u; else hi = mid;
    }
    return lo;
}

// a6 for one pair, all 32 lanes.  `in`/`n`/`reversed`: the read->contig CIGAR (op i = in[reversed ? n-1-i : i]); `pos`: its
// start on the strand the segment's table [t0, t1) is written in.  Ops go through `sink`; the caller runs sink.finish().
// Returns true if ref2_start_pos was set (uniform); *start = that position (before the leading-deletion shift).
__device__ __forceinline__ bool warp_liftover(const uint32_t* __restrict__ in, uint32_t n, bool reversed, uint32_t pos,
                                              const TabEntry* __restrict__ tab, uint32_t t0, uint32_t t1, WarpSink& sink, uint32_t lane,
                                              int32_t* start_out) {
    constexpr uint32_t INF = 0xffffffffu;
    // cursor: tcur = first entry with key > p_base; (b0_*) = the entry in front of it = the block under p_base
    uint32_t tcur = warp_table_bound<true>(tab, t0, t1, pos, lane);
    uint32_t b0_key = 0, b0_gap = 0;
    int32_t b0_val = -2;  // -2: in front of the first key
    if (tcur > t0) {
        const TabEntry b = tab[tcur - 1u];
        b0_key = b.key; b0_val = b.val; b0_gap = b.gap;
    }
    uint32_t p_base = pos;
    bool start_set = false, some_seen = false;
    int32_t start = 0;
    for (uint32_t base = 0; base < n; base += 32u) {
        // ---- 1. ops of the chunk and their contig intervals
        const uint32_t i = base + lane;
        const uint32_t c = (i < n) ? in[reversed ? (n - 1u - i) : i] : 0u;
        const uint32_t op = c & 0xfu, len = c >> 4;
        const uint32_t radv = ((kRefMask >> op) & 1u) ? len : 0u;
        const uint32_t incl = warp_incl_scan(radv, lane);
        const uint32_t e_i = p_base + incl, p_i = e_i - radv;
        const uint32_t chunk_end = __shfl_sync(FULL, e_i, 31);
        // ---- 2. table window: entries with p_base < key < chunk_end, one per lane
        uint32_t wkey = INF, wgap = 0;
        int32_t wval = -1;
        if (tcur + lane < t1) {
            const TabEntry w = tab[tcur + lane];
            wkey = w.key; wval = w.val; wgap = w.gap;
        }
        const uint32_t nwin = __popc(__ballot_sync(FULL, wkey < chunk_end));  // keys increase: a prefix of the lanes
        const bool big = nwin == 32u;  // more keys may follow: per-lane look-ups in memory for this chunk
        uint32_t cnt_p = 0, cnt_e = 0;  // window keys <= p_i, < e_i
        if (!big) {
            for (uint32_t j = 0; j < nwin; ++j) {
                const uint32_t kj = __shfl_sync(FULL, wkey, j);
                cnt_p += (kj <= p_i) ? 1u : 0u;
                cnt_e += (kj < e_i) ? 1u : 0u;
            }
        } else if (radv) {
            cnt_p = lane_table_bound_upper(tab, tcur, t1, p_i) - tcur;
            cnt_e = lane_table_bound_lower(tab, tcur, t1, e_i) - tcur;
        }
        // ---- 3. pieces: a reference-consuming op is cut at every key inside it; I/S/H pass verbatim; Pad and empty ops vanish
        const bool verbatim = op == OP_I || op == OP_S || op == OP_H;
        const uint32_t pieces = radv ? (cnt_e - cnt_p + 1u) : (verbatim ? 1u : 0u);
        const uint32_t pincl = warp_incl_scan(pieces, lane);
        const uint32_t pex = pincl - pieces;
        const uint32_t n_pieces = __shfl_sync(FULL, pincl, 31);
        for (uint32_t q0 = 0; q0 < n_pieces; q0 += 32u) {
            const uint32_t q = q0 + lane;
            const bool act = q < n_pieces;
            // owner op of piece q = last lane whose exclusive piece prefix is <= q
            uint32_t lo = 0, hi = 31;
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const uint32_t mid = (lo + hi + 1u) >> 1;
                const uint32_t v = __shfl_sync(FULL, pex, mid);
                if (v <= q) lo = mid; else hi = mid - 1u;
            }
            const uint32_t oc = __shfl_sync(FULL, c, lo);
            const uint32_t o_p = __shfl_sync(FULL, p_i, lo), o_e = __shfl_sync(FULL, e_i, lo);
            const uint32_t o_cnt = __shfl_sync(FULL, cnt_p | (cnt_e << 16), lo);  // (window counts <= 32; in `big` chunks see below)
            const uint32_t o_pex = __shfl_sync(FULL, pex, lo);
            uint32_t o_cp = o_cnt & 0xffffu, o_ce = o_cnt >> 16;
            if (big) {  // counts may exceed 16 bits
                o_cp = __shfl_sync(FULL, cnt_p, lo);
                o_ce = __shfl_sync(FULL, cnt_e, lo);
            }
            const uint32_t o_op = oc & 0xfu, o_len = oc >> 4;
            const bool o_ref = ((kRefMask >> o_op) & 1u) && o_len != 0u;
            const uint32_t j = q - o_pex;
            // block of the piece: window index rel (-1 = the block under p_base), and the key that ends it
            const int32_t rel = (act && o_ref) ? int32_t(o_cp + j) - 1 : 0;
            uint32_t bkey, bgap, nkey;
            int32_t bval;
            if (!big) {
                bkey = __shfl_sync(FULL, wkey, rel & 31);
                bval = __shfl_sync(FULL, wval, rel & 31);
                bgap = __shfl_sync(FULL, wgap, rel & 31);
                nkey = __shfl_sync(FULL, wkey, (rel + 1) & 31);
            } else {
                bkey = 0; bval = -1; bgap = 0; nkey = INF;
                if (act && o_ref) {
                    if (rel >= 0) { const TabEntry b = tab[tcur + uint32_t(rel)]; bkey = b.key; bval = b.val; bgap = b.gap; }
                    if (tcur + uint32_t(rel + 1) < t1) nkey = tab[tcur + uint32_t(rel + 1)].key;
                }
            }
            if (rel < 0) { bkey = b0_key; bval = b0_val; bgap = b0_gap; }
            const bool is_m = op_is_match(o_op);
            const bool last = j == o_ce - o_cp;
            const uint32_t pstart = j ? bkey : o_p;
            const uint32_t plen = (last ? o_e : nkey) - pstart;
            const bool ref_piece = act && o_ref;
            const bool some = ref_piece && bval >= 0;
            // ---- 4. sequential state as prefix-ORs over the pieces (src/liftover_read_alignment.rs:84-100)
            const uint32_t cand_mask = __ballot_sync(FULL, some && is_m);  // pieces that can set ref2_start_pos
            const uint32_t some_mask = __ballot_sync(FULL, some);           // pieces that set ref2_end_pos
            const bool start_incl = start_set || (cand_mask & lanes_le(lane)) != 0u;
            const bool seen_excl = some_seen || (some_mask & lanes_lt(lane)) != 0u;
            if (!start_set && cand_mask) {
                start = __shfl_sync(FULL, bval + int32_t(pstart - bkey), __ffs(cand_mask) - 1);
                start_set = true;
            }
            some_seen = some_seen || some_mask != 0u;
            uint32_t a_len = 0, b_op = 0, b_len = 0;
            if (ref_piece) {
                if (some) {
                    if (start_incl) {
                        b_op = is_m ? uint32_t(OP_M) : o_op;  // D stays D, N stays N, M/=/X become M (:103-107)
                        b_len = plen;
                        // the deletion between the previous aligned run and this one (:91-96): only the first piece
                        // inside a block can see a positive distance
                        if (seen_excl && (j != 0u || o_p == bkey)) a_len = bgap;
                    }
                } else if (is_m) {
                    b_op = (bval == -1) ? uint32_t(OP_I) : uint32_t(OP_S);  // None block: insertion (:111-115); no block: clip (:117-123)
                    b_len = plen;
                }
            } else if (act) {  // I/S/H transfer verbatim (:157-160)
                b_op = o_op;
                b_len = o_len;
            }
            sink.push2(lane, a_len, b_op, b_len);
        }
        // ---- move the cursor behind the chunk: entries with key <= chunk_end are passed
        uint32_t n_le;
        if (!big) {
            const uint32_t k_next = __shfl_sync(FULL, wkey, nwin & 31u);  // (nwin < 32 here)
            n_le = nwin + ((k_next == chunk_end) ? 1u : 0u);
            if (n_le) {
                b0_key = __shfl_sync(FULL, wkey, n_le - 1u);
                b0_val = __shfl_sync(FULL, wval, n_le - 1u);
                b0_gap = __shfl_sync(FULL, wgap, n_le - 1u);
            }
        } else {
            n_le = warp_table_bound<true>(tab, tcur, t1, chunk_end, lane) - tcur;
            const TabEntry b = tab[tcur + n_le - 1u];
            b0_key = b.key; b0_val = b.val; b0_gap = b.gap;
        }
        tcur += n_le;
        p_base = chunk_end;
        PTL_ASSERT_UNIFORM(tcur); PTL_ASSERT_UNIFORM(p_base); PTL_ASSERT_UNIFORM(b0_key); PTL_ASSERT_UNIFORM(start_set); PTL_ASSERT_UNIFORM(some_seen);
    }
    *start_out = start;
    return start_set;
}

// ------------------------------------------------------------------------------------------------------------------
// a5 for one pair, all 32 lanes: left_shift_indels (lib/rust-vc-utils/.../shift_indels/left_shift_indels.rs:17-39) with
// CigarShiftBuilder in Left mode (cigar_indel_shifter.rs:45-164), op-parallel.
//
// The sequential builder is a chain of EVENTS: "close" (an I/D cluster ends: emit M(match_block - shift), I, D and carry
// `shift` matched bases over it) and "other" (a non-match, non-indel op or the end: flush the match block, emit the op).
//   pass A  lanes over ops: scans give the (ref, read) heads; ballots find cluster starts / ends; every event is written
//           to `ev` as 6 words {CLOSE | op word, matches since the previous event, ins, del, ref_end, read_end}
//   pass B  lanes over events: the walk bound limit_k = min(max_left_k, gap_k + limit_{k-1}) and the carried match block
//           carry_k = min(carry_{k-1} + gap_k, homology_k) are both scans of the functions x -> min(x + a, b), which
//           compose associatively; the homology walks of 32 clusters run together (walk_homology8); each event's
//           output ops replace its record in place
//   pass C  the <= 3 ops per event stream through the WarpSink (edge clean-up + compression)
// Returns the number of base bytes compared by this lane; *err is uniform.
struct MinPlus {
    uint32_t a, b;  // x -> min(x + a, b); identity = (0, kMpInf)
};
constexpr uint32_t kMpInf = 0x7fffffffu;
__device__ __forceinline__ MinPlus mp_then(MinPlus f, MinPlus g) { return MinPlus{f.a + g.a, min(f.b + g.a, g.b)}; }  // f first, then g
__device__ __forceinline__ MinPlus warp_incl_scan_mp(MinPlus v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        MinPlus o;
        o.a = __shfl_up_sync(FULL, v.a, d);
        o.b = __shfl_up_sync(FULL, v.b, d);
        if (int(lane) >= d) v = mp_then(o, v);
    }
    return v;
}
__device__ __forceinline__ int msb(uint32_t m) { return 31 - __clz(m); }

constexpr uint32_t kEvClose = 0xffffffffu;
constexpr uint32_t kEvWords = 6;

__device__ __forceinline__ void warp_left_shift(const uint32_t* __restrict__ in, uint32_t n, bool reversed, uint32_t ref_pos,
                                                const uint8_t* __restrict__ ref_seq, uint32_t ref_len, const ReadBases& read,
                                                const uint64_t* __restrict__ win, uint32_t n_win, uint32_t* __restrict__ ev, uint32_t ev_cap,
                                                WarpSink& sink, uint32_t lane, PairCounters& cnt, int* err) {
    const uint32_t lt = lanes_lt(lane);
    // ---- pass A
    uint32_t n_ev = 0, n_clus = 0;
    uint32_t ref_base = ref_pos, read_base = 0;
    bool c_open = false;
    uint32_t c_blk_ref = 0, c_blk_read = 0, c_del = 0, c_ins = 0, c_gap = 0;
    for (uint32_t base = 0; base <= n; base += 32u) {  // index n = the sentinel that closes a trailing cluster and flushes
        const uint32_t i = base + lane;
        const uint32_t c = (i < n) ? in[reversed ? (n - 1u - i) : i] : uint32_t(OP_S);
        const uint32_t op = c & 0xfu, len = c >> 4;
        const bool is_id = op == OP_I || op == OP_D;
        // 0 transparent (empty I/D, lanes past the sentinel), 1 indel, 2 match op, 3 other op
        const uint32_t kind = (i > n) ? 0u : is_id ? (len ? 1u : 0u) : (op_is_match(op) ? 2u : 3u);
        const uint32_t radv = (i < n) ? op_ref_adv(c) : 0u, qadv = (i < n) ? op_read_adv(c) : 0u;
        const uint32_t dlen = (kind == 1u && op == OP_D) ? len : 0u, ilen = (kind == 1u && op == OP_I) ? len : 0u;
        const uint32_t mlen = (kind == 2u) ? len : 0u;
        const uint32_t r_in = warp_incl_scan(radv, lane), q_in = warp_incl_scan(qadv, lane);
        const uint32_t d_in = warp_incl_scan(dlen, lane), i_in = warp_incl_scan(ilen, lane), m_in = warp_incl_scan(mlen, lane);
        const uint32_t ref_ex = ref_base + r_in - radv, read_ex = read_base + q_in - qadv;
        const uint32_t d_ex = d_in - dlen, i_ex = i_in - ilen, m_ex = m_in - mlen;
        const uint32_t nt_mask = __ballot_sync(FULL, kind != 0u), idl_mask = __ballot_sync(FULL, kind == 1u);
        const uint32_t closer_mask = nt_mask & ~idl_mask;
        const uint32_t pm = nt_mask & lt;
        const bool in_prev = pm ? ((idl_mask >> msb(pm)) & 1u) != 0u : c_open;  // is a cluster open in front of this op?
        const uint32_t start_mask = __ballot_sync(FULL, kind == 1u && !in_prev);
        const bool is_close = kind >= 2u && in_prev, is_other = kind == 3u;
        const uint32_t close_mask = __ballot_sync(FULL, is_close), other_mask = __ballot_sync(FULL, is_other);
        const uint32_t ev_mask = close_mask | other_mask;
        // the cluster a close lane ends: started in this chunk behind the last closer, or carried in
        const uint32_t sm = start_mask & lt, cm = closer_mask & lt;
        const bool started_here = sm != 0u && (cm == 0u || msb(sm) > msb(cm));
        const int s_lane = started_here ? msb(sm) : 0;
        const uint32_t s_ref = __shfl_sync(FULL, ref_ex, s_lane), s_read = __shfl_sync(FULL, read_ex, s_lane);
        const uint32_t s_dex = __shfl_sync(FULL, d_ex, s_lane), s_iex = __shfl_sync(FULL, i_ex, s_lane);
        const uint32_t blk_ref = started_here ? s_ref : c_blk_ref, blk_read = started_here ? s_read : c_blk_read;
        const uint32_t del = started_here ? d_ex - s_dex : c_del + d_ex, ins = started_here ? i_ex - s_iex : c_ins + i_ex;
        // matches since the previous event (a match op that closed a cluster counts itself towards the next one)
        const uint32_t em = ev_mask & lt;
        const uint32_t e_mex = __shfl_sync(FULL, m_ex, em ? msb(em) : 0);
        const uint32_t gap = em ? m_ex - e_mex : c_gap + m_ex;
        const uint32_t n_mine = (is_close ? 1u : 0u) + (is_other ? 1u : 0u);
        const uint32_t e_in = warp_incl_scan(n_mine, lane);
        uint32_t at = n_ev + e_in - n_mine;
        if (is_close) {
            if (at < ev_cap) {
                uint32_t* r = ev + size_t(at) * kEvWords;
                r[0] = kEvClose; r[1] = gap; r[2] = ins; r[3] = del; r[4] = blk_ref + del; r[5] = blk_read + ins;
            }
            ++at;
        }
        if (is_other && at < ev_cap) {
            uint32_t* r = ev + size_t(at) * kEvWords;
            r[0] = c; r[1] = is_close ? 0u : gap; r[2] = 0; r[3] = 0; r[4] = 0; r[5] = 0;
        }
        n_ev += __shfl_sync(FULL, e_in, 31);
        n_clus += __popc(close_mask);
        // carries into the next chunk
        const bool open_end = nt_mask ? ((idl_mask >> msb(nt_mask)) & 1u) != 0u : c_open;
        const uint32_t d_tot = __shfl_sync(FULL, d_in, 31), i_tot = __shfl_sync(FULL, i_in, 31), m_tot = __shfl_sync(FULL, m_in, 31);
        if (open_end) {
            const bool started = start_mask != 0u && (closer_mask == 0u || msb(start_mask) > msb(closer_mask));
            const int sl = started ? msb(start_mask) : 0;
            const uint32_t o_ref = __shfl_sync(FULL, ref_ex, sl), o_read = __shfl_sync(FULL, read_ex, sl);
            const uint32_t o_dex = __shfl_sync(FULL, d_ex, sl), o_iex = __shfl_sync(FULL, i_ex, sl);
            if (started) { c_blk_ref = o_ref; c_blk_read = o_read; c_del = d_tot - o_dex; c_ins = i_tot - o_iex; }
            else { c_del += d_tot; c_ins += i_tot; }
        }
        c_open = open_end;
        const uint32_t l_mex = __shfl_sync(FULL, m_ex, ev_mask ? msb(ev_mask) : 0);
        c_gap = ev_mask ? m_tot - l_mex : c_gap + m_tot;
        ref_base += __shfl_sync(FULL, r_in, 31);
        read_base += __shfl_sync(FULL, q_in, 31);
    }
    PTL_ASSERT_UNIFORM(n_ev); PTL_ASSERT_UNIFORM(n_clus); PTL_ASSERT_UNIFORM(c_gap);
    if (n_ev > ev_cap) { *err = ST_ERR_CAPACITY; return; }
    __syncwarp();
    // ---- pass B: walk bounds, homologies, carried match block; the output ops of an event replace its record
    const bool have_win = (n_win == n_clus) && n_clus > 0u;
    uint32_t lim_carry = 0, mb_carry = 0, k_base = 0;
    bool lane_err = false;
    for (uint32_t eb = 0; eb < n_ev; eb += 32u) {
        const uint32_t e = eb + lane;
        const bool act = e < n_ev;
        uint32_t w0 = 0, gap = 0, ins = 0, del = 0, ref_end = 0, read_end = 0;
        uint32_t* r = ev + size_t(act ? e : 0u) * kEvWords;
        if (act) { w0 = r[0]; gap = r[1]; ins = r[2]; del = r[3]; ref_end = r[4]; read_end = r[5]; }
        const bool is_close = act && w0 == kEvClose;
        uint32_t max_left = is_close ? min(ref_end - del, read_end - ins) : 0u;
        // first access is ref_seq[ref_end-1] / read[read_end-1]; later ones only move down and stay >= 0
        if (is_close && max_left > 0u && (ref_end > ref_len || read_end > read.len)) { lane_err = true; max_left = 0; }
        const MinPlus lf = warp_incl_scan_mp(act ? (is_close ? MinPlus{gap, max_left} : MinPlus{0u, 0u}) : MinPlus{0u, kMpInf}, lane);
        const uint32_t limit = min(lim_carry + lf.a, lf.b);  // (an "other" event flushes the match block: no carry across it)
        lim_carry = __shfl_sync(FULL, limit, 31);
        const uint32_t close_mask = __ballot_sync(FULL, is_close);
        uint32_t hom = 0;
        bool walking = is_close && limit > 0u;
        uint64_t w = 0;
        if (have_win && walking) w = win[k_base + __popc(close_mask & lt)];
        k_base += __popc(close_mask);
        while (__any_sync(FULL, walking)) {
            if (walking) walk_homology8(ref_seq, read, ref_end, read_end, limit, w, have_win, hom, walking, cnt);
        }
        const MinPlus cf = warp_incl_scan_mp(act ? (is_close ? MinPlus{gap, hom} : MinPlus{0u, 0u}) : MinPlus{0u, kMpInf}, lane);
        const uint32_t after = min(mb_carry + cf.a, cf.b);  // match block carried behind this event (:124-131)
        uint32_t before = __shfl_up_sync(FULL, after, 1);
        if (lane == 0u) before = mb_carry;
        mb_carry = __shfl_sync(FULL, after, 31);
        PTL_ASSERT_UNIFORM(lim_carry); PTL_ASSERT_UNIFORM(mb_carry); PTL_ASSERT_UNIFORM(k_base);
        if (act) {
            if (is_close) {  // end_indel (:101-148): M(match_block - shift) then nImD
                r[0] = ((before + gap - after) << 4) | OP_M; r[1] = (ins << 4) | OP_I; r[2] = (del << 4) | OP_D;
            } else {         // add_other (:155-164): the match block, then the op itself
                r[0] = ((before + gap) << 4) | OP_M; r[1] = w0; r[2] = 0;
            }
        }
    }
    if (__any_sync(FULL, lane_err)) { *err = ST_ERR_BOUNDS; return; }
    __syncwarp();
    // ---- pass C
    const uint32_t n_words = 3u * n_ev;
    for (uint32_t q0 = 0; q0 < n_words; q0 += 32u) {
        const uint32_t q = q0 + lane;
        const uint32_t word = (q < n_words) ? ev[size_t(q / 3u) * kEvWords + q % 3u] : 0u;
        sink.push2(lane, 0u, word & 0xfu, word >> 4);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// a9 for one pair, all 32 lanes: simplify_alignment_indels (src/simplify_alignment_indels.rs:119-156) with
// CigarBlockInfo::end_indel (:35-111), op-parallel.  `in`/`n`: a cleaned + compressed CIGAR starting at ref_pos.
//   pass A  lanes over ops: every I/D run (cluster) is summed by scans; the op that ends it emits the rewritten run in
//           front of itself: nothing, I, D, M1, or for a MIXED cluster (both kinds, not 1/1) the five slots
//           M(pre) I D M(post) op, whose numbers need bases.  All ops go to `x` (uncompressed, in order); mixed clusters
//           are listed in `rec` as {blk_ref, blk_read, slot}
//   pass B  lanes over mixed clusters: trim equal bases from the right first, then from the left (:55-85)
//   pass C  `x` streams through the sink, which compresses IN PLACE (a run is stored only after the ops it merges were
//           read, and stored runs never outnumber consumed ops): sink.buf == x
__device__ __forceinline__ int warp_simplify(const uint32_t* __restrict__ in, uint32_t n, int64_t ref_pos, const uint8_t* __restrict__ ref_seq,
                                             uint64_t ref_len, const ReadBases& read, uint32_t* __restrict__ x, uint32_t x_cap,
                                             uint32_t* __restrict__ rec, uint32_t rec_cap, WarpSink& sink, uint32_t lane, PairCounters& cnt) {
    const uint32_t lt = lanes_lt(lane);
    // ---- pass A
    uint32_t n_x = 0, n_mixed = 0;
    uint32_t ref_base = 0, read_base = 0;  // reference offset from ref_pos, read head
    bool c_open = false;
    uint32_t c_blk_ref = 0, c_blk_read = 0, c_del = 0, c_ins = 0;
    for (uint32_t base = 0; base <= n; base += 32u) {  // index n = a sentinel match op of length 0: closes a trailing cluster
        const uint32_t i = base + lane;
        const uint32_t c = (i < n) ? in[i] : 0u;
        const uint32_t op = c & 0xfu, len = c >> 4;
        const bool is_id = i < n && (op == OP_I || op == OP_D);
        const uint32_t kind = (i > n) ? 0u : (is_id ? 1u : 2u);  // 0 past the end, 1 indel, 2 any other op
        const uint32_t radv = (i < n) ? op_ref_adv(c) : 0u, qadv = (i < n) ? op_read_adv(c) : 0u;
        const uint32_t dlen = (is_id && op == OP_D) ? len : 0u, ilen = (is_id && op == OP_I) ? len : 0u;
        const uint32_t r_in = warp_incl_scan(radv, lane), q_in = warp_incl_scan(qadv, lane);
        const uint32_t d_in = warp_incl_scan(dlen, lane), i_in = warp_incl_scan(ilen, lane);
        const uint32_t ref_ex = ref_base + r_in - radv, read_ex = read_base + q_in - qadv;
        const uint32_t d_ex = d_in - dlen, i_ex = i_in - ilen;
        const uint32_t nt_mask = __ballot_sync(FULL, kind != 0u), idl_mask = __ballot_sync(FULL, kind == 1u);
        const uint32_t closer_mask = nt_mask & ~idl_mask;
        const uint32_t pm = nt_mask & lt;
        const bool in_prev = pm ? ((idl_mask >> msb(pm)) & 1u) != 0u : c_open;
        const uint32_t start_mask = __ballot_sync(FULL, kind == 1u && !in_prev);
        const bool is_close = kind == 2u && in_prev;
        const uint32_t sm = start_mask & lt, cm = closer_mask & lt;
        const bool started_here = sm != 0u && (cm == 0u || msb(sm) > msb(cm));
        const int s_lane = started_here ? msb(sm) : 0;
        const uint32_t s_ref = __shfl_sync(FULL, ref_ex, s_lane), s_read = __shfl_sync(FULL, read_ex, s_lane);
        const uint32_t s_dex = __shfl_sync(FULL, d_ex, s_lane), s_iex = __shfl_sync(FULL, i_ex, s_lane);
        const uint32_t blk_ref = started_here ? s_ref : c_blk_ref, blk_read = started_here ? s_read : c_blk_read;
        const uint32_t del = started_here ? d_ex - s_dex : c_del + d_ex, ins = started_here ? i_ex - s_iex : c_ins + i_ex;
        const bool mixed = is_close && del > 0u && ins > 0u && !(del == 1u && ins == 1u);
        const uint32_t mixed_mask = __ballot_sync(FULL, mixed);
        const uint32_t n_mine = (kind == 2u ? 1u : 0u) + (is_close ? (mixed ? 4u : 1u) : 0u);
        const uint32_t e_in = warp_incl_scan(n_mine, lane);
        uint32_t at = n_x + e_in - n_mine;
        if (kind == 2u && at + n_mine <= x_cap) {
            if (mixed) {
                const uint32_t m = n_mixed + __popc(mixed_mask & lt);
                if (m < rec_cap) { rec[3u * m] = blk_ref; rec[3u * m + 1u] = blk_read; rec[3u * m + 2u] = at; }
                x[at] = 0; x[at + 1u] = ins; x[at + 2u] = del; x[at + 3u] = 0;
                at += 4u;
            } else if (is_close) {
                // (0, i) -> Ins, (d, 0) -> Del, (1, 1) -> Match(1) without looking at the bases (:38-48)
                x[at] = (del == 0u) ? ((ins << 4) | OP_I) : (ins == 0u) ? ((del << 4) | OP_D) : ((1u << 4) | OP_M);
                ++at;
            }
            x[at] = c;
        }
        n_x += __shfl_sync(FULL, e_in, 31);
        n_mixed += __popc(mixed_mask);
        const bool open_end = nt_mask ? ((idl_mask >> msb(nt_mask)) & 1u) != 0u : c_open;
        const uint32_t d_tot = __shfl_sync(FULL, d_in, 31), i_tot = __shfl_sync(FULL, i_in, 31);
        if (open_end) {
            const bool started = start_mask != 0u && (closer_mask == 0u || msb(start_mask) > msb(closer_mask));
            const int sl = started ? msb(start_mask) : 0;
            const uint32_t o_ref = __shfl_sync(FULL, ref_ex, sl), o_read = __shfl_sync(FULL, read_ex, sl);
            const uint32_t o_dex = __shfl_sync(FULL, d_ex, sl), o_iex = __shfl_sync(FULL, i_ex, sl);
            if (started) { c_blk_ref = o_ref; c_blk_read = o_read; c_del = d_tot - o_dex; c_ins = i_tot - o_iex; }
            else { c_del += d_tot; c_ins += i_tot; }
        }
        c_open = open_end;
        ref_base += __shfl_sync(FULL, r_in, 31);
        read_base += __shfl_sync(FULL, q_in, 31);
    }
    PTL_ASSERT_UNIFORM(n_x); PTL_ASSERT_UNIFORM(n_mixed);
    if (n_x > x_cap || n_mixed > rec_cap) return ST_ERR_CAPACITY;
    __syncwarp();
    // ---- pass B
    bool lane_err = false;
    for (uint32_t kb = 0; kb < n_mixed; kb += 32u) {
        const uint32_t k = kb + lane;
        const bool mine = k < n_mixed;
        int64_t blk_ref = 0;
        uint32_t blk_read = 0, at = 0, d = 0, q = 0, pre = 0, post = 0;
        if (mine) {
            blk_ref = ref_pos + int64_t(rec[3u * k]);
            blk_read = rec[3u * k + 1u];
            at = rec[3u * k + 2u];
            q = x[at + 1u];
            d = x[at + 2u];
        }
        // The first probe of BOTH sides is issued up front (4 independent loads: almost every cluster ends each side at
        // its first or second base, and the dependent DRAM round trips of a few lanes were the whole cost of this pass);
        // the left probe is still valid after the right trimming because it reads blk_ref + 0 / blk_read + 0.
        const uint32_t d0 = d, q0 = q;
        uint32_t pr_r = 0x100u, pr_q = 0x200u, pl_r = 0x100u, pl_q = 0x200u;  // (unequal sentinels: an invalid probe is re-done below)
        if (mine && d0 > 0u && q0 > 0u) {
            const int64_t rr = blk_ref + int64_t(d0) - 1, rl = blk_ref;
            if (rr >= 0 && uint64_t(rr) < ref_len && blk_read + q0 - 1u < read.len) { pr_r = ref_seq[rr]; pr_q = read.at(blk_read + q0 - 1u); }
            if (rl >= 0 && uint64_t(rl) < ref_len && blk_read < read.len) { pl_r = ref_seq[rl]; pl_q = read.at(blk_read); }
        }
        uint32_t side = mine ? 1u : 0u;  // 1 right, 2 left, 0 done
        while (__any_sync(FULL, side != 0u)) {
            if (side != 0u) {
                if (d == 0u || q == 0u) {
                    side = (side == 1u) ? 2u : 0u;
                } else {
                    const int64_t f_ref = (side == 1u) ? blk_ref + int64_t(d) - 1 : blk_ref + int64_t(pre);
                    const uint32_t f_read = (side == 1u) ? blk_read + q - 1u : blk_read + pre;
                    if (f_ref < 0 || uint64_t(f_ref) >= ref_len || f_read >= read.len) {
                        lane_err = true;  // Rust slice index panic
                        side = 0u;
                    } else {
                        uint32_t rb, qb;
                        if (side == 1u && d == d0 && pr_r < 0x100u) { rb = pr_r; qb = pr_q; }
                        else if (side == 2u && pre == 0u && pl_r < 0x100u) { rb = pl_r; qb = pl_q; }
                        else { rb = ref_seq[f_ref]; qb = read.at(f_read); }
                        cnt.base_bytes += 2;
                        if (rb == qb) {
                            --d; --q;
                            if (side == 1u) ++post; else ++pre;
                        } else {
                            side = (side == 1u) ? 2u : 0u;
                        }
                    }
                }
            }
        }
        if (mine) {
            if (d == 1u && q == 1u) { d = 0; q = 0; ++post; }  // down to a SNP: 1 edit instead of 2 (:88-92)
            x[at] = (pre << 4) | OP_M; x[at + 1u] = (q << 4) | OP_I; x[at + 2u] = (d << 4) | OP_D; x[at + 3u] = (post << 4) | OP_M;
        }
    }
    if (__any_sync(FULL, lane_err)) return ST_ERR_BOUNDS;
    __syncwarp();
    // ---- pass C (in place: sink.buf == x)
    for (uint32_t q0 = 0; q0 < n_x; q0 += 32u) {
        const uint32_t q = q0 + lane;
        const uint32_t word = (q < n_x) ? x[q] : 0u;
        sink.push2(lane, 0u, word & 0xfu, word >> 4);
    }
    return 0;
}

// a4 + a5 + a6 + a8 for entry t of DevWork::long_list: lift_pair_body did the strand logic and parked the pair with
// ST_PENDING_LIFT (pair_pos = start on the table's strand, pair_n_out = ops of the segment CIGAR, pair_flip).
__device__ __forceinline__ void lift_long_pair_body(const DevStatic& S, const DevBatch& B, const DevWork& W, DevTotals* T, uint32_t p, uint32_t lane,
                                                    uint32_t stage_mask) {
    const uint32_t s = W.pair_rseg[p], g = W.pair_seg[p];
    const uint32_t r = W.rseg_read[s];
    const bool contig_fwd = S.seg_is_fwd[g] != 0;
    const uint64_t slot0 = W.pair_slot_begin[p];
    const uint32_t cap_b = W.pair_cap_b[p];
    const uint32_t cap_a = uint32_t(W.pair_slot_begin[p + 1] - slot0) - cap_b;
    uint32_t* buf_b = W.scratch + slot0;
    uint32_t* buf_a = buf_b + cap_b;
    uint32_t n = W.pair_n_out[p];
    const uint32_t* in = B.cigar + B.rseg_cigar_begin[s];
    bool reversed = !contig_fwd;  // the CIGAR reversal of :167 is just the read order
    uint32_t cpos = uint32_t(W.pair_pos[p]);
    int status = ST_LIFTED;
    uint32_t span = 0;
    PairCounters cnt;
    // ---- a5 on the contig's reverse strand (:168-175)
    if (!contig_fwd && (stage_mask & 1u)) {
        const uint32_t ctg = B.rseg_contig[s];
        const uint64_t rev_off = S.contig_rev_off[ctg];
        if (rev_off == ~0ull) {
            status = ST_ERR_BOUNDS;  // Option::unwrap on None (:174)
        } else {
            const ReadBases read{B.seq4 + B.read_seq_off[r], B.read_seq_len[r], W.pair_flip[p] != 0};
            const uint64_t* win = nullptr;
            uint32_t n_win = 0;
            if (B.rseg_win_begin) {
                const uint32_t w0 = B.rseg_win_begin[s];
                win = B.indel_win + w0;
                n_win = B.rseg_win_begin[s + 1] - w0;
            }
            WarpSink shifted(buf_a, cap_a);
            int err = 0;
            warp_left_shift(in, n, true, cpos, S.rev_pool + rev_off, uint32_t(S.contig_len[ctg]), read, win, n_win, buf_b, cap_b / kEvWords, shifted,
                            lane, cnt, &err);
            shifted.finish(lane);
            if (err) status = err;
            else if (shifted.overflow) status = ST_ERR_CAPACITY;
            in = buf_a;
            n = shifted.n;
            span = shifted.ref_span;
            reversed = false;
            cpos += shifted.lead_del_shift;
        }
    }
    if (!(stage_mask & 2u)) {  // stage test: left shift only, hand the shifted CIGAR back
        if (in != buf_a) {     // nothing to shift: the (possibly reversed) input, verbatim
            n = min(n, cap_a);
            for (uint32_t i = lane; i < n; i += 32u) {
                const uint32_t c = in[reversed ? (W.pair_n_out[p] - 1u - i) : i];
                buf_a[i] = c;
                span += op_ref_adv(c);
            }
            span = warp_sum(span);
            in = buf_a;
            __syncwarp();
        }
        if (lane == 0u) {
            const bool ok = status == ST_LIFTED;
            W.pair_status[p] = int8_t(status);
            W.pair_pos[p] = ok ? int64_t(cpos) : 0;
            W.pair_n_out[p] = ok ? n : 0u;
            W.pair_out_off[p] = uint64_t(in - W.scratch);
            W.pair_bin[p] = ok ? reg2bin(cpos, int64_t(cpos) + int64_t(span)) : uint16_t(0);
        }
        return;
    }
    // ---- a6 + a8
    WarpSink sink(buf_b, cap_b);
    int32_t start = 0;
    if (status == ST_LIFTED) {
        const bool some = warp_liftover(in, n, reversed, cpos, S.table, S.seg_tab_begin[g], S.seg_tab_begin[g + 1], sink, lane, &start);
        sink.finish(lane);
        if (sink.overflow) status = ST_ERR_CAPACITY;
        else if (!some) status = ST_NONE;
        else if (W.rseg_read_len[s] != B.read_seq_len[r]) status = ST_ERR_LENGTH;  // (:204-229, see lift_pair_body)
    }
    int64_t rpos = int64_t(start) + int64_t(sink.lead_del_shift);
    uint32_t n_out = sink.n;
    uint64_t out_off = slot0;
    span = sink.ref_span;
    // ---- a9 (:236-243): only when the lifted CIGAR holds a mixed I/D run (otherwise the identity, see lift_pair_body): B -> A
    if (status == ST_LIFTED && sink.mixed_cluster && (stage_mask & 4u)) {
        const ReadBases read{B.seq4 + B.read_seq_off[r], B.read_seq_len[r], W.pair_flip[p] != 0};
        const int32_t chrom = S.seg_chrom[g];
        // buffer A of a long pair: [ops, uncompressed then compressed in place | 3 words per mixed cluster]  (pair_fill_body)
        const uint32_t n_clus_cap = (cap_a - cap_b - 16u) / 9u;
        uint32_t* rec = buf_a + (cap_a - 3u * n_clus_cap);
        WarpSink simp(buf_a, cap_a - 3u * n_clus_cap);
        const int err = warp_simplify(buf_b, n_out, rpos, S.ref + S.chrom_off[chrom], S.chrom_off[chrom + 1] - S.chrom_off[chrom], read, buf_a,
                                      simp.cap, rec, n_clus_cap, simp, lane, cnt);
        simp.finish(lane);
        if (err) status = err;
        else if (simp.overflow) status = ST_ERR_CAPACITY;
        rpos += int64_t(simp.lead_del_shift);
        n_out = simp.n;
        out_off = slot0 + cap_b;
        span = simp.ref_span;
    }
    const bool ok = status == ST_LIFTED;
    const uint32_t bytes = warp_sum(cnt.base_bytes);
    if (lane == 0u) {
        W.pair_status[p] = int8_t(status);
        W.pair_pos[p] = ok ? rpos : 0;
        W.pair_n_out[p] = ok ? n_out : 0u;
        W.pair_out_off[p] = out_off;
        W.pair_bin[p] = ok ? reg2bin(rpos, rpos + int64_t(span)) : uint16_t(0);
        if (bytes) atomicAdd(&T->n_base_bytes, (unsigned long long)bytes);
    }
}

// a9 for a pair lifted by lift_pairs_kernel whose CIGAR holds a mixed I/D run (DevWork::simplify_list): B -> A.
__device__ __forceinline__ void simplify_warp_pair_body(const DevStatic& S, const DevBatch& B, const DevWork& W, DevTotals* T, uint32_t p, uint32_t lane) {
    const uint32_t s = W.pair_rseg[p], g = W.pair_seg[p];
    const uint32_t r = W.rseg_read[s];
    const uint64_t slot0 = W.pair_slot_begin[p];
    const uint32_t cap_b = W.pair_cap_b[p];
    const uint32_t cap_a = uint32_t(W.pair_slot_begin[p + 1] - slot0) - cap_b;
    uint32_t* buf_b = W.scratch + slot0;
    uint32_t* buf_a = buf_b + cap_b;
    // buffer A: [ops, uncompressed then compressed in place | 3 words per mixed cluster]; its size was chosen by
    // pair_fill_body from (n_id + n_keys) = the possible I/D clusters of the lifted CIGAR
    const bool long_layout = B.rseg_cigar_len[s] > W.long_ops;
    const uint32_t n_clus_cap = long_layout ? (cap_a - cap_b - 16u) / 9u : (cap_a - cap_b - 8u) / 6u;
    uint32_t* rec = buf_a + (cap_a - 3u * n_clus_cap);
    const ReadBases read{B.seq4 + B.read_seq_off[r], B.read_seq_len[r], W.pair_flip[p] != 0};
    const int32_t chrom = S.seg_chrom[g];
    const int64_t rpos = W.pair_pos[p];
    PairCounters cnt;
    WarpSink simp(buf_a, cap_a - 3u * n_clus_cap);
    // the lifted CIGAR: where lift_pairs_kernel left it (the pair's slot, or the dense region of its block)
    const uint32_t* lifted = W.scratch + W.pair_out_off[p];
    int status = warp_simplify(lifted, W.pair_n_out[p], rpos, S.ref + S.chrom_off[chrom], S.chrom_off[chrom + 1] - S.chrom_off[chrom], read, buf_a,
                               simp.cap, rec, n_clus_cap, simp, lane, cnt);
    simp.finish(lane);
    if (!status && simp.overflow) status = ST_ERR_CAPACITY;
    const int64_t out_pos = rpos + int64_t(simp.lead_del_shift);
    const uint32_t bytes = warp_sum(cnt.base_bytes);
    if (lane == 0u) {
        W.pair_status[p] = int8_t(status ? status : ST_LIFTED);
        W.pair_pos[p] = status ? 0 : out_pos;
        W.pair_n_out[p] = status ? 0u : simp.n;
        W.pair_out_off[p] = slot0 + cap_b;
        W.pair_bin[p] = status ? uint16_t(0) : reg2bin(out_pos, out_pos + int64_t(simp.ref_span));
        if (bytes) atomicAdd(&T->n_base_bytes, (unsigned long long)bytes);
    }
}

}  // namespace ptl
