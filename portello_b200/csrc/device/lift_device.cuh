// Per-pair device algorithms of the liftover path (sm_100a).  One thread owns one (read segment x contig segment)
// pair and runs the stages as streaming transducers over BAM-encoded CIGAR ops:
//
//     [left_shift]  ->  liftover  ->  [simplify]
//
// Every stage writes through an `OpSink`, which fuses the reference's two post-passes
// (clean_up_cigar_edge_indels + compress_cigar, lib/rust-vc-utils/src/bam_utils/cigar/mod.rs:204-291) into the
// write: leading-edge conversion and run merging happen as ops are pushed, the trailing edge is fixed in place at
// finish().  Integer arithmetic only; bases are compared as the exact bytes the reference compares.
#pragma once
#include <cstdint>

#include "device_types.hpp"

namespace ptl {

enum : uint32_t { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };
constexpr uint32_t kMatchMask = (1u << OP_M) | (1u << OP_EQ) | (1u << OP_X);
constexpr uint32_t kRefMask = kMatchMask | (1u << OP_D) | (1u << OP_N);
constexpr uint32_t kReadMask = kMatchMask | (1u << OP_I) | (1u << OP_S) | (1u << OP_H);
constexpr uint32_t NO_OP = 0xffffffffu;

constexpr int ST_LIFTED = 1, ST_NONE = 0, ST_ERR_LENGTH = -1, ST_ERR_BOUNDS = -2, ST_ERR_CAPACITY = -3;

__device__ __forceinline__ bool op_is_match(uint32_t op) { return (kMatchMask >> op) & 1u; }
__device__ __forceinline__ uint32_t op_ref_adv(uint32_t c) { return ((kRefMask >> (c & 0xf)) & 1u) ? (c >> 4) : 0u; }
__device__ __forceinline__ uint32_t op_read_adv(uint32_t c) { return ((kReadMask >> (c & 0xf)) & 1u) ? (c >> 4) : 0u; }

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
        uint8_t c = decode_nt16(nib);
        if (flip) c = comp(c);
        return c;
    }
    static __device__ __forceinline__ uint8_t decode_nt16(uint32_t nib) {
        // "=ACMGRSVTWYHKDBN" packed little-endian into two 64-bit immediates (no table load)
        const unsigned long long w0 = 0x565352474D43413DULL;  // V S R G M C A =
        const unsigned long long w1 = 0x4E42444B48595754ULL;  // N B D K H Y W T
        const unsigned long long w = (nib & 8u) ? w1 : w0;
        return uint8_t(w >> ((nib & 7u) * 8u));
    }
    static __device__ __forceinline__ uint8_t comp(uint8_t c) {
        switch (c) {
            case 'A': return 'T';
            case 'T': return 'A';
            case 'C': return 'G';
            case 'G': return 'C';
            default: return 'N';  // 'N' -> 'N'; every other 4-bit symbol -> 'N' (lower case never occurs in a BAM decode)
        }
    }
};

// ------------------------------------------------------------------------------------------------------------------
// Streaming clean_up_cigar_edge_indels + compress_cigar.
struct OpSink {
    uint32_t* buf;
    uint32_t cap;
    uint32_t n = 0;
    uint32_t pend = NO_OP;
    int32_t last_match_idx = -1;  // index (in buf) of the last alignment-match op, counting the pending one
    bool seen_match = false;
    bool overflow = false;
    uint64_t lead_del_shift = 0;  // return value of clean_up_cigar_edge_indels
    uint64_t read_len = 0;        // get_cigar_read_offset(result, ignore_hard_clip=false)

    __device__ __forceinline__ OpSink(uint32_t* b, uint32_t c) : buf(b), cap(c) {}

    __device__ __forceinline__ void flush() {
        if (pend != NO_OP) {
            if (n < cap) buf[n] = pend;
            else overflow = true;
            ++n;
            pend = NO_OP;
        }
    }
    __device__ __forceinline__ void push(uint32_t op, uint32_t len) {
        if (len == 0) return;  // compress_cigar filters empty elements (no stage emits an empty alignment match)
        if ((kReadMask >> op) & 1u) read_len += len;
        if (!seen_match) {  // leading edge (cigar/mod.rs:278-280)
            if (op_is_match(op)) seen_match = true;
            else if (op == OP_I) op = OP_S;
            else if (op == OP_D) { lead_del_shift += len; return; }  // -> SoftClip(0), later dropped
        }
        if (pend != NO_OP && (pend & 0xfu) == op) {
            if (op != OP_P) pend += len << 4;  // a Pad after a Pad is dropped, not merged (cigar/mod.rs:208-215)
            return;
        }
        flush();
        pend = (len << 4) | op;
        if (op_is_match(op)) last_match_idx = int32_t(n);
    }
    // trailing edge (cigar/mod.rs:282-288) + re-merge of what the conversion made adjacent
    __device__ __forceinline__ void finish() {
        flush();
        if (overflow || last_match_idx < 0) return;  // no match at all: the leading pass already converted everything
        const uint32_t start = uint32_t(last_match_idx) + 1u;
        uint32_t w = start, prev = NO_OP;
        for (uint32_t i = start; i < n; ++i) {
            uint32_t c = buf[i];
            uint32_t op = c & 0xfu;
            if (op == OP_D) continue;
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
// a5: left_shift_indels (lib/rust-vc-utils/.../shift_indels/left_shift_indels.rs:17-39) with CigarShiftBuilder in Left
// mode (cigar_indel_shifter.rs:45-164) and the left walk of get_indel_breakend_homology_info
// (indel_breakend_homology.rs:35-49).  Only min(match_block, homology) matters (:124), so the walk stops at
// match_block: same result, bounded work.
struct LeftShifter {
    const uint8_t* ref_seq;
    uint64_t ref_len;
    ReadBases read;
    OpSink& sink;
    PairCounters& cnt;
    int err = 0;
    uint32_t match_block = 0;
    bool in_indel = false;
    int64_t blk_ref = 0;
    uint64_t blk_read = 0;
    uint32_t del = 0, ins = 0;

    __device__ __forceinline__ LeftShifter(const uint8_t* r, uint64_t rl, ReadBases rd, OpSink& s, PairCounters& c)
        : ref_seq(r), ref_len(rl), read(rd), sink(s), cnt(c) {}

    __device__ __forceinline__ void end_indel() {
        if (!in_indel) return;
        in_indel = false;
        const int64_t ref_end = blk_ref + int64_t(del);
        const int64_t read_end = int64_t(blk_read) + int64_t(ins);
        const int64_t max_left = min(blk_ref, int64_t(blk_read));
        uint32_t hom = 0;
        if (max_left > 0) {
            // first access is ref_seq[ref_end-1] / read_seq[read_end-1]; later ones only move down and stay >= 0
            if (ref_end > int64_t(ref_len) || read_end > int64_t(read.len)) {
                err = ST_ERR_BOUNDS;
            } else {
                const uint32_t limit = uint32_t(min(max_left, int64_t(match_block)));
                while (hom < limit) {
                    const uint8_t rb = ref_seq[ref_end - 1 - hom];
                    const uint8_t qb = read.at(uint32_t(read_end - 1 - hom));
                    cnt.base_bytes += 2;
                    if (rb != qb) break;
                    ++hom;
                }
            }
        }
        const uint32_t actual = min(match_block, hom);
        sink.push(OP_M, match_block - actual);
        match_block = actual;
        sink.push(OP_I, ins);  // nImD order (:141-147)
        sink.push(OP_D, del);
        ins = 0;
        del = 0;
    }
    __device__ __forceinline__ void add(uint32_t c, int64_t ref_head, uint64_t read_head) {
        const uint32_t op = c & 0xfu, len = c >> 4;
        if (op == OP_D || op == OP_I) {
            if (len > 0) {
                if (!in_indel) { in_indel = true; blk_ref = ref_head; blk_read = read_head; }
                if (op == OP_D) del += len; else ins += len;
            }
        } else if (op_is_match(op)) {
            end_indel();
            match_block += len;
        } else {
            end_indel();
            sink.push(OP_M, match_block);
            match_block = 0;
            sink.push(op, len);
        }
    }
    __device__ __forceinline__ void end() {
        end_indel();
        sink.push(OP_M, match_block);
        match_block = 0;
    }
};

// returns the shifted position
__device__ __forceinline__ int64_t run_left_shift(const OpSource& in, int64_t ref_pos, const uint8_t* ref_seq, uint64_t ref_len,
                                                  const ReadBases& read, OpSink& sink, PairCounters& cnt, int& err) {
    LeftShifter ls(ref_seq, ref_len, read, sink, cnt);
    int64_t ref_head = ref_pos;
    uint64_t read_head = 0;
    for (uint32_t i = 0; i < in.n; ++i) {
        const uint32_t c = in.get(i);
        ls.add(c, ref_head, read_head);
        ref_head += op_ref_adv(c);
        read_head += op_read_adv(c);
    }
    ls.end();
    sink.finish();
    if (ls.err) err = ls.err;
    return ref_pos + int64_t(sink.lead_del_shift);
}

// ------------------------------------------------------------------------------------------------------------------
// a6: liftover_read_alignment (src/liftover_read_alignment.rs:137-223).  The reference re-searches its BTreeMap for
// every reference-consuming op; read-op boundaries and table keys both advance monotonically, so one binary search
// per pair plus a forward merge visits exactly the same (piece, block) sequence.
// Returns true if ref2_start_pos was set (Some); *out_pos = start + leading-deletion shift.
__device__ __forceinline__ bool run_liftover(const OpSource& in, int64_t pos, const int2* __restrict__ tab, uint32_t t0,
                                             uint32_t t1, OpSink& sink, int64_t* out_pos) {
    bool start_set = false, end2_set = false;
    int64_t start = 0, end2 = 0;
    // cursor = index of the greatest key <= pos, or t0-1 (as int64 to allow -1 when t0 == 0)
    int64_t cur;
    {
        uint32_t lo = t0, hi = t1;  // first index with key > pos
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (int64_t(uint32_t(tab[mid].x)) <= pos) lo = mid + 1;
            else hi = mid;
        }
        cur = int64_t(lo) - 1;
    }
    int64_t p = pos;
    for (uint32_t i = 0; i < in.n; ++i) {
        const uint32_t c = in.get(i);
        const uint32_t op = c & 0xfu, len = c >> 4;
        if (op == OP_I || op == OP_S || op == OP_H) {
            sink.push(op, len);  // read-only ops transfer verbatim (:157-160)
            continue;
        }
        if (!((kRefMask >> op) & 1u)) continue;  // Pad (:213)
        if (len == 0) continue;
        const bool is_match = op_is_match(op);
        const int64_t s = p, e = p + int64_t(len);
        while (cur + 1 < int64_t(t1) && int64_t(uint32_t(tab[cur + 1].x)) <= s) ++cur;
        int64_t bp = s;
        int64_t li = (cur >= int64_t(t0)) ? cur : -1;  // "last" block, -1 = None
        int64_t nx = cur + 1;
        for (;;) {
            const int64_t nk = (nx < int64_t(t1)) ? int64_t(uint32_t(tab[nx].x)) : INT64_MAX;
            const int64_t seg_end = min(nk, e);
            if (seg_end > bp) {  // update_ref2_cigar_segment (:35-133) for the piece [bp, seg_end)
                const uint32_t plen = uint32_t(seg_end - bp);
                if (li < 0) {
                    if (is_match) sink.push(OP_S, plen);
                } else {
                    const int2 blk = tab[li];
                    const int64_t k = int64_t(uint32_t(blk.x));
                    if (blk.y < 0) {
                        if (is_match) sink.push(OP_I, plen);
                    } else {
                        const int64_t r = int64_t(blk.y);
                        if (is_match && !start_set) { start = r + (bp - k); start_set = true; }
                        if (end2_set) {
                            const int64_t dlen = r - end2;
                            if (dlen > 0 && start_set) sink.push(OP_D, uint32_t(dlen));
                        }
                        end2 = r + (seg_end - k);
                        end2_set = true;
                        if (is_match || start_set) sink.push(op == OP_D ? OP_D : (op == OP_N ? OP_N : OP_M), plen);
                    }
                }
                bp = seg_end;
            }
            if (nk >= e) break;
            li = nx;
            ++nx;
        }
        cur = nx - 1;
        p = e;
    }
    if (!start_set) return false;
    sink.finish();
    *out_pos = start + int64_t(sink.lead_del_shift);
    return true;
}

// ------------------------------------------------------------------------------------------------------------------
// a9: simplify_alignment_indels (src/simplify_alignment_indels.rs:119-156) with CigarBlockInfo::end_indel (:35-111).
struct Simplifier {
    const uint8_t* ref_seq;
    uint64_t ref_len;
    ReadBases read;
    OpSink& sink;
    PairCounters& cnt;
    int err = 0;
    bool in_indel = false;
    int64_t blk_ref = 0;
    uint64_t blk_read = 0;
    uint32_t del = 0, ins = 0;

    __device__ __forceinline__ Simplifier(const uint8_t* r, uint64_t rl, ReadBases rd, OpSink& s, PairCounters& c)
        : ref_seq(r), ref_len(rl), read(rd), sink(s), cnt(c) {}

    __device__ __forceinline__ bool fetch(int64_t ref_idx, int64_t read_idx, uint8_t& rb, uint8_t& qb) {
        if (ref_idx < 0 || uint64_t(ref_idx) >= ref_len || read_idx < 0 || uint64_t(read_idx) >= read.len) {
            err = ST_ERR_BOUNDS;  // Rust slice index panic
            return false;
        }
        rb = ref_seq[ref_idx];
        qb = read.at(uint32_t(read_idx));
        cnt.base_bytes += 2;
        return true;
    }
    __device__ __forceinline__ void end_indel() {
        if (!in_indel) return;
        in_indel = false;
        uint32_t d = del, n = ins;
        del = 0;
        ins = 0;
        if (d == 0 || n == 0) {  // (0,0) nothing, (0,len) Ins, (len,0) Del
            sink.push(OP_I, n);
            sink.push(OP_D, d);
            return;
        }
        if (d == 1 && n == 1) { sink.push(OP_M, 1); return; }
        uint32_t pre = 0, post = 0;
        uint8_t rb, qb;
        while (d > 0 && n > 0) {  // right side first
            if (!fetch(blk_ref + int64_t(d) - 1, int64_t(blk_read) + int64_t(n) - 1, rb, qb)) return;
            if (rb != qb) break;
            --d; --n; ++post;
        }
        while (d > 0 && n > 0) {  // then left side
            if (!fetch(blk_ref + int64_t(pre), int64_t(blk_read) + int64_t(pre), rb, qb)) return;
            if (rb != qb) break;
            --d; --n; ++pre;
        }
        if (d == 1 && n == 1) { d = 0; n = 0; ++post; }
        sink.push(OP_M, pre);
        sink.push(OP_I, n);
        sink.push(OP_D, d);
        sink.push(OP_M, post);
    }
};

__device__ __forceinline__ int64_t run_simplify(const OpSource& in, int64_t ref_pos, const uint8_t* ref_seq, uint64_t ref_len,
                                                const ReadBases& read, OpSink& sink, PairCounters& cnt, int& err) {
    Simplifier sp(ref_seq, ref_len, read, sink, cnt);
    int64_t ref_head = ref_pos;
    uint64_t read_head = 0;
    for (uint32_t i = 0; i < in.n; ++i) {
        const uint32_t c = in.get(i);
        const uint32_t op = c & 0xfu, len = c >> 4;
        if (op == OP_D || op == OP_I) {
            if (!sp.in_indel) { sp.in_indel = true; sp.blk_ref = ref_head; sp.blk_read = read_head; }
            if (op == OP_D) sp.del += len; else sp.ins += len;
        } else {
            sp.end_indel();
            sink.push(op, len);
        }
        ref_head += op_ref_adv(c);
        read_head += op_read_adv(c);
    }
    sp.end_indel();
    sink.finish();
    if (sp.err) err = sp.err;
    return ref_pos + int64_t(sink.lead_del_shift);
}

// bam_reg2bin (lib/rust-vc-utils/src/bam_utils/util.rs:10-35)
__device__ __forceinline__ uint16_t reg2bin(int64_t begin, int64_t end) {
    const uint64_t b = uint64_t(begin), e = uint64_t(end) - 1ull;
    if ((b >> 14) == (e >> 14)) return uint16_t(4681u + (b >> 14));
    if ((b >> 17) == (e >> 17)) return uint16_t(585u + (b >> 17));
    if ((b >> 20) == (e >> 20)) return uint16_t(73u + (b >> 20));
    if ((b >> 23) == (e >> 23)) return uint16_t(9u + (b >> 23));
    if ((b >> 26) == (e >> 26)) return uint16_t(1u + (b >> 26));
    return 0;
}

}  // namespace ptl
