// ORACLE — TEST INFRASTRUCTURE ONLY (see cigar.hpp header).
//
// Restates:
//   lib/rust-vc-utils/src/bam_utils/read_to_ref_map.rs:59-137   (ReadToRefTreeMap + builder)
//   src/liftover_read_alignment.rs:35-223                       (update_ref2_cigar_segment, liftover_read_alignment)
//   src/simplify_alignment_indels.rs:5-156                      (CigarBlockInfo, simplify_alignment_indels)
//   lib/rust-vc-utils/src/indel_breakend_homology.rs:24-73      (get_indel_breakend_homology_info)
//   lib/rust-vc-utils/src/bam_utils/cigar/shift_indels/{cigar_indel_shifter.rs:10-165,left_shift_indels.rs:17-39,
//                                                      right_shift_indels.rs}
#pragma once
#include <map>
#include <optional>

#include "cigar.hpp"

namespace orc {

// ---------------------------------------------------------------------------------------------------------------
// read_to_ref_map.rs:59-90.  std::map<size_t, optional<int64>> stands in for BTreeMap<usize, Option<i64>>.
struct ReadToRefTreeMap {
    std::map<size_t, std::optional<int64_t>> map;
    using It = std::map<size_t, std::optional<int64_t>>::const_iterator;

    // :67-72
    std::optional<int64_t> get_ref_pos(size_t read_pos) const {
        auto it = map.upper_bound(read_pos);
        if (it == map.begin()) return std::nullopt;
        --it;
        if (!it->second) return std::nullopt;
        return *it->second + int64_t(read_pos - it->first);
    }
    // :74-85  half-open key range [k0, read_end_pos) where k0 = greatest key <= read_start_pos, else read_start_pos
    std::pair<It, It> get_ref_range(size_t read_start_pos, size_t read_end_pos) const {
        size_t start_block = read_start_pos;
        auto it = map.upper_bound(read_start_pos);
        if (it != map.begin()) {
            --it;
            start_block = it->first;
        }
        // Rust's BTreeMap::range panics if start > end; start_block <= read_start_pos <= read_end_pos always here.
        return {map.lower_bound(start_block), map.lower_bound(read_end_pos)};
    }
};

// read_to_ref_map.rs:101-137
inline ReadToRefTreeMap get_read_segment_to_ref_pos_tree_map(int64_t ref_pos, const CigarVec& cigar,
                                                             bool ignore_hard_clip) {
    ReadToRefTreeMap out;
    size_t read_pos = 0;
    size_t match_len = 0;
    auto update_map = [&](int64_t rp, size_t qp) {
        if (match_len > 0) {
            out.map[qp - match_len] = rp - int64_t(match_len);  // insert overwrites a previous None at that key
            out.map[qp] = std::nullopt;
            match_len = 0;
        }
    };
    for (const auto& c : cigar) {
        if (is_alignment_match(c)) match_len += c.len;
        else update_map(ref_pos, read_pos);
        update_ref_and_read_pos(c, ref_pos, read_pos, ignore_hard_clip);
    }
    update_map(ref_pos, read_pos);
    return out;
}

// ---------------------------------------------------------------------------------------------------------------
// src/liftover_read_alignment.rs:35-133
using MapBlock = std::optional<std::pair<size_t, std::optional<int64_t>>>;

inline void update_ref2_cigar_segment(const MapBlock& this_block, const MapBlock& last_block,
                                      int64_t ref1_seg_end_pos, const Cigar& ref1_seg, int64_t& block_ref1_pos,
                                      std::optional<int64_t>& ref2_start_pos, std::optional<int64_t>& ref2_end_pos,
                                      CigarVec& ref2_cigar) {
    const int64_t remapped_end =
        this_block ? std::min<int64_t>(int64_t(this_block->first), ref1_seg_end_pos) : ref1_seg_end_pos;

    if (remapped_end > block_ref1_pos) {
        const uint32_t remapped_len = uint32_t(remapped_end - block_ref1_pos);
        const bool is_match_segment = is_alignment_match(ref1_seg);

        if (last_block) {
            const int64_t last_ref1_start = int64_t(last_block->first);
            if (last_block->second) {
                const int64_t last_ref2_start = *last_block->second;
                if (is_match_segment && !ref2_start_pos) {
                    ref2_start_pos = last_ref2_start + (block_ref1_pos - last_ref1_start);
                }
                if (ref2_end_pos) {
                    const int64_t deletion_len = last_ref2_start - *ref2_end_pos;
                    if (deletion_len > 0 && ref2_start_pos) ref2_cigar.push_back(Cigar{D, uint32_t(deletion_len)});
                }
                ref2_end_pos = last_ref2_start + (remapped_end - last_ref1_start);
                if (is_match_segment || ref2_start_pos) {
                    const uint8_t op = (ref1_seg.op == D) ? D : (ref1_seg.op == N) ? N : M;
                    ref2_cigar.push_back(Cigar{op, remapped_len});
                }
            } else {
                if (is_match_segment) ref2_cigar.push_back(Cigar{I, remapped_len});
            }
        } else {
            if (is_match_segment) ref2_cigar.push_back(Cigar{S, remapped_len});
        }
        block_ref1_pos = remapped_end;
    }
}

struct PosCigar {
    int64_t pos;
    CigarVec cigar;
};

// src/liftover_read_alignment.rs:137-223
inline std::optional<PosCigar> liftover_read_alignment(const ReadToRefTreeMap& ref1_to_ref2_map,
                                                       int64_t ref1_seg_start_pos, const CigarVec& ref1_cigar) {
    std::optional<int64_t> ref2_start_pos, ref2_end_pos;
    CigarVec ref2_cigar;
    for (const auto& seg : ref1_cigar) {
        switch (seg.op) {
            case I: case S: case H:
                ref2_cigar.push_back(seg);
                break;
            case X: case EQ: case M: case D: case N: {
                MapBlock last_block;
                int64_t block_ref1_pos = ref1_seg_start_pos;
                const int64_t seg_end = ref1_seg_start_pos + int64_t(seg.len);
                auto range = ref1_to_ref2_map.get_ref_range(size_t(ref1_seg_start_pos), size_t(seg_end));
                for (auto it = range.first; it != range.second; ++it) {
                    MapBlock this_block = std::make_pair(it->first, it->second);
                    update_ref2_cigar_segment(this_block, last_block, seg_end, seg, block_ref1_pos, ref2_start_pos,
                                              ref2_end_pos, ref2_cigar);
                    last_block = this_block;
                }
                update_ref2_cigar_segment(std::nullopt, last_block, seg_end, seg, block_ref1_pos, ref2_start_pos,
                                          ref2_end_pos, ref2_cigar);
                break;
            }
            default:  // Pad
                break;
        }
        ref1_seg_start_pos += get_cigarseg_ref_offset(seg);
    }
    if (!ref2_start_pos) return std::nullopt;
    const size_t shift = clean_up_cigar_edge_indels(ref2_cigar);
    return PosCigar{*ref2_start_pos + int64_t(shift), compress_cigar(ref2_cigar)};
}

// ---------------------------------------------------------------------------------------------------------------
// Base-compare accounting for the roofline's s(p) term: bytes the algorithm NEEDS to compare
// (both operands): the left homology walk of a5 and the trim loops of a9.
struct CompareStats {
    uint64_t base_bytes = 0;
};

inline uint8_t seq_at(const uint8_t* seq, size_t len, int64_t idx) {
    if (idx < 0 || size_t(idx) >= len) throw Panic(-2, "sequence index out of bounds");
    return seq[idx];
}

// The read bases an algorithm indexes: either a decoded ASCII array (what the reference materialises with
// `record.seq().as_bytes()` + rev_comp_in_place) or a lazy view over the BAM 4-bit bases giving identical bytes.
struct ReadSeq {
    const uint8_t* ascii = nullptr;
    const uint8_t* seq4 = nullptr;
    size_t len = 0;
    bool flip = false;
    static ReadSeq from_ascii(const uint8_t* p, size_t n) { return ReadSeq{p, nullptr, n, false}; }
    static ReadSeq from_seq4(const uint8_t* p, size_t n, bool flip) { return ReadSeq{nullptr, p, n, flip}; }
    size_t size() const { return len; }
    uint8_t at(int64_t idx) const {
        if (idx < 0 || size_t(idx) >= len) throw Panic(-2, "read sequence index out of bounds");
        if (ascii) return ascii[idx];
        const size_t j = flip ? len - 1 - size_t(idx) : size_t(idx);
        const uint8_t byte = seq4[j >> 1];
        const uint8_t c = uint8_t(kSeqNt16[(j & 1) ? (byte & 0xf) : (byte >> 4)]);
        return flip ? comp_base(c) : c;
    }
};

// src/simplify_alignment_indels.rs:5-112
struct CigarBlockInfo {
    bool is_in_indel_block = false;
    int64_t block_ref_start = 0;
    size_t block_read_start = 0;
    uint32_t block_del_size = 0, block_ins_size = 0;

    void add_indel(int64_t ref_pos, size_t read_pos) {
        if (!is_in_indel_block) {
            is_in_indel_block = true;
            block_ref_start = ref_pos;
            block_read_start = read_pos;
        }
    }
    void end_indel(const uint8_t* ref_seq, size_t ref_len, const ReadSeq& read_seq, CigarVec& out,
                   CompareStats* stats) {
        if (!is_in_indel_block) return;
        is_in_indel_block = false;
        uint32_t del_len = block_del_size, ins_len = block_ins_size;
        if (del_len == 0 && ins_len == 0) {
        } else if (del_len == 0) {
            out.push_back(Cigar{I, ins_len});
        } else if (ins_len == 0) {
            out.push_back(Cigar{D, del_len});
        } else if (del_len == 1 && ins_len == 1) {
            out.push_back(Cigar{M, 1});
        } else {
            uint32_t pre = 0, post = 0;
            while (del_len > 0 && ins_len > 0) {  // right side first (:55-68)
                const uint8_t rb = seq_at(ref_seq, ref_len, block_ref_start + int64_t(del_len) - 1);
                const uint8_t qb = read_seq.at(int64_t(block_read_start + ins_len) - 1);
                if (stats) stats->base_bytes += 2;
                if (rb != qb) break;
                --del_len; --ins_len; ++post;
            }
            while (del_len > 0 && ins_len > 0) {  // then left side (:71-85)
                const uint8_t rb = seq_at(ref_seq, ref_len, block_ref_start + int64_t(pre));
                const uint8_t qb = read_seq.at(int64_t(block_read_start + pre));
                if (stats) stats->base_bytes += 2;
                if (rb != qb) break;
                --del_len; --ins_len; ++pre;
            }
            if (del_len == 1 && ins_len == 1) {  // :88-92
                del_len = 0; ins_len = 0; ++post;
            }
            if (pre) out.push_back(Cigar{M, pre});
            if (ins_len) out.push_back(Cigar{I, ins_len});
            if (del_len) out.push_back(Cigar{D, del_len});
            if (post) out.push_back(Cigar{M, post});
        }
        block_ins_size = 0;
        block_del_size = 0;
    }
};

// src/simplify_alignment_indels.rs:119-156
inline PosCigar simplify_alignment_indels(int64_t ref_pos, const CigarVec& cigar, const uint8_t* ref_seq,
                                          size_t ref_len, const ReadSeq& read_seq, CompareStats* stats = nullptr) {
    int64_t ref_head = ref_pos;
    size_t read_head = 0;
    CigarBlockInfo blk;
    CigarVec simple;
    for (const auto& c : cigar) {
        if (c.op == D) {
            blk.add_indel(ref_head, read_head);
            blk.block_del_size += c.len;
        } else if (c.op == I) {
            blk.add_indel(ref_head, read_head);
            blk.block_ins_size += c.len;
        } else {
            blk.end_indel(ref_seq, ref_len, read_seq, simple, stats);
            simple.push_back(c);
        }
        update_ref_and_read_pos(c, ref_head, read_head, false);
    }
    blk.end_indel(ref_seq, ref_len, read_seq, simple, stats);
    const size_t shift = clean_up_cigar_edge_indels(simple);
    return PosCigar{ref_pos + int64_t(shift), compress_cigar(simple)};
}

// ---------------------------------------------------------------------------------------------------------------
// indel_breakend_homology.rs:24-73.  Returns (range, hom_seq).
inline std::pair<IntRange, std::vector<uint8_t>> get_indel_breakend_homology_info(
    const uint8_t* ref_seq, size_t ref_len, const IntRange& ref_range, const ReadSeq& read_seq,
    const IntRange& read_range, CompareStats* stats = nullptr) {
    const size_t read_len = read_seq.size();
    std::vector<uint8_t> hom;
    const int64_t max_left = std::min(ref_range.start, read_range.start);
    int64_t left = 0;
    while (left < max_left) {
        const uint8_t rb = seq_at(ref_seq, ref_len, ref_range.end - left - 1);
        const uint8_t qb = read_seq.at(read_range.end - left - 1);
        if (stats) stats->base_bytes += 2;
        if (rb != qb) break;
        hom.push_back(rb);
        ++left;
    }
    std::reverse(hom.begin(), hom.end());
    const int64_t max_right = std::min(int64_t(ref_len) - ref_range.end, int64_t(read_len) - read_range.end);
    int64_t right = 0;
    while (right < max_right) {
        const uint8_t rb = seq_at(ref_seq, ref_len, ref_range.start + right);
        const uint8_t qb = read_seq.at(read_range.start + right);
        if (rb != qb) break;
        hom.push_back(rb);
        ++right;
    }
    return {IntRange{-left, right}, hom};
}

// cigar_indel_shifter.rs:10-165
enum class ShiftDirection { Left, Right };

struct CigarShiftBuilder {
    ShiftDirection dir;
    const uint8_t* ref_seq; size_t ref_len;
    ReadSeq read_seq;
    CompareStats* stats;
    uint32_t match_block_size = 0;
    bool is_in_indel_block = false;
    int64_t indel_block_ref_start = 0;
    size_t indel_block_read_start = 0;
    uint32_t indel_block_del_size = 0, indel_block_ins_size = 0;
    CigarVec shift_cigar;

    void add_element(const Cigar& c, int64_t ref_pos, size_t read_pos) {
        switch (c.op) {
            case D: if (c.len > 0) { add_indel(ref_pos, read_pos); indel_block_del_size += c.len; } break;
            case I: if (c.len > 0) { add_indel(ref_pos, read_pos); indel_block_ins_size += c.len; } break;
            case M: case EQ: case X: end_indel(); match_block_size += c.len; break;
            default: add_other(&c);
        }
    }
    CigarVec get_cigar() {
        add_other(nullptr);
        if (dir == ShiftDirection::Right) std::reverse(shift_cigar.begin(), shift_cigar.end());
        CigarVec out;
        out.swap(shift_cigar);
        return out;
    }
    void add_indel(int64_t ref_pos, size_t read_pos) {
        if (dir == ShiftDirection::Right || !is_in_indel_block) {
            indel_block_ref_start = ref_pos;
            indel_block_read_start = read_pos;
            is_in_indel_block = true;
        }
    }
    void push_del() {
        if (indel_block_del_size > 0) { shift_cigar.push_back(Cigar{D, indel_block_del_size}); indel_block_del_size = 0; }
    }
    void push_ins() {
        if (indel_block_ins_size > 0) { shift_cigar.push_back(Cigar{I, indel_block_ins_size}); indel_block_ins_size = 0; }
    }
    void end_indel() {
        if (!is_in_indel_block) return;
        is_in_indel_block = false;
        const IntRange ref_range{indel_block_ref_start, indel_block_ref_start + int64_t(indel_block_del_size)};
        const IntRange read_range{int64_t(indel_block_read_start),
                                  int64_t(indel_block_read_start) + int64_t(indel_block_ins_size)};
        const IntRange hom =
            get_indel_breakend_homology_info(ref_seq, ref_len, ref_range, read_seq, read_range, stats).first;
        const uint32_t shift_len =
            uint32_t(std::max<int64_t>(0, dir == ShiftDirection::Left ? -hom.start : hom.end));
        const uint32_t actual = std::min(match_block_size, shift_len);
        const uint32_t shifted_match = match_block_size - actual;
        if (shifted_match > 0) shift_cigar.push_back(Cigar{M, shifted_match});
        match_block_size = actual;
        if (dir == ShiftDirection::Left) push_ins();
        push_del();
        if (dir == ShiftDirection::Right) push_ins();
    }
    void add_other(const Cigar* c) {
        end_indel();
        if (match_block_size > 0) {
            shift_cigar.push_back(Cigar{M, match_block_size});
            match_block_size = 0;
        }
        if (c) shift_cigar.push_back(*c);
    }
};

// left_shift_indels.rs:17-39
inline PosCigar left_shift_indels(int64_t ref_pos, const CigarVec& cigar, const uint8_t* ref_seq, size_t ref_len,
                                  const ReadSeq& read_seq, CompareStats* stats = nullptr) {
    int64_t ref_head = ref_pos;
    size_t read_head = 0;
    CigarShiftBuilder b{ShiftDirection::Left, ref_seq, ref_len, read_seq, stats};
    for (const auto& c : cigar) {
        b.add_element(c, ref_head, read_head);
        update_ref_and_read_pos(c, ref_head, read_head, false);
    }
    CigarVec shifted = b.get_cigar();
    const size_t shift = clean_up_cigar_edge_indels(shifted);
    return PosCigar{ref_pos + int64_t(shift), compress_cigar(shifted)};
}

// right_shift_indels.rs:20-50 (NOT called by portello; restated because the reference's shift tests exercise the
// shared CigarShiftBuilder in both directions, which gives more vectors to pin the builder).
inline PosCigar right_shift_indels(int64_t ref_pos, const CigarVec& cigar, const uint8_t* ref_seq, size_t ref_len,
                                   const ReadSeq& read_seq) {
    std::vector<std::pair<int64_t, size_t>> heads;
    int64_t ref_head = ref_pos;
    size_t read_head = 0;
    for (const auto& c : cigar) {
        heads.push_back({ref_head, read_head});
        update_ref_and_read_pos(c, ref_head, read_head, false);
    }
    CigarShiftBuilder b{ShiftDirection::Right, ref_seq, ref_len, read_seq, nullptr};
    for (size_t k = cigar.size(); k-- > 0;) b.add_element(cigar[k], heads[k].first, heads[k].second);
    CigarVec shifted = b.get_cigar();
    const size_t shift = clean_up_cigar_edge_indels(shifted);
    return PosCigar{ref_pos + int64_t(shift), compress_cigar(shifted)};
}

}  // namespace orc
