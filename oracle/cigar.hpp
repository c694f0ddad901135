// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's CIGAR utilities.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use oracle/.
//
// Restates (paths relative to the reference repo root):
//   lib/rust-vc-utils/src/bam_utils/cigar/mod.rs:16-327
//   lib/rust-vc-utils/src/bam_utils/cigar/clip_alignment.rs:103-181
//   lib/rust-vc-utils/src/bam_utils/cigar/score_alignment.rs:68-74,138-165
//   lib/rust-vc-utils/src/int_range.rs:11-95, src/seq_util.rs:1-40, src/bam_utils/util.rs:10-35, src/util.rs:50-67
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

// BAM op codes == rust_htslib Cigar variants.
enum Op : uint8_t { M = 0, I = 1, D = 2, N = 3, S = 4, H = 5, P = 6, EQ = 7, X = 8 };

struct Cigar {
    uint8_t op;
    uint32_t len;
    bool operator==(const Cigar& o) const { return op == o.op && len == o.len; }
};
using CigarVec = std::vector<Cigar>;

// A reference `panic!` / failed assert / slice-index panic.
struct Panic : std::runtime_error {
    int code;
    Panic(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline uint32_t encode(const Cigar& c) { return (c.len << 4) | c.op; }
inline Cigar decode(uint32_t v) { return Cigar{uint8_t(v & 0xf), v >> 4}; }

// cigar/mod.rs:16-18
inline bool is_clip(const Cigar& c) { return c.op == S || c.op == H; }
// cigar/mod.rs:22-24
inline bool is_alignment_match(const Cigar& c) { return c.op == M || c.op == EQ || c.op == X; }

// cigar/mod.rs:26-39
inline size_t get_cigarseg_read_offset(const Cigar& c, bool ignore_hard_clip) {
    switch (c.op) {
        case I: case S: case X: case EQ: case M: return c.len;
        case H: return ignore_hard_clip ? 0 : c.len;
        default: return 0;
    }
}
// cigar/mod.rs:41-47
inline int64_t get_cigarseg_ref_offset(const Cigar& c) {
    switch (c.op) {
        case D: case N: case X: case EQ: case M: return c.len;
        default: return 0;
    }
}
// cigar/mod.rs:70-78
inline void update_ref_and_read_pos(const Cigar& c, int64_t& ref_pos, size_t& read_pos, bool ignore_hard_clip) {
    read_pos += get_cigarseg_read_offset(c, ignore_hard_clip);
    ref_pos += get_cigarseg_ref_offset(c);
}

struct ClipPositions { size_t left, right, size; };
// cigar/mod.rs:85-118
inline ClipPositions get_read_clip_positions(const CigarVec& cigar, bool ignore_hard_clip) {
    int64_t ref_pos = 0;
    size_t read_pos = 0, left_clip = 0, right_clip = 0;
    bool in_left = true;
    for (const auto& c : cigar) {
        if (c.op == S) {
            (in_left ? left_clip : right_clip) += c.len;
        } else if (c.op == H) {
            if (!ignore_hard_clip) (in_left ? left_clip : right_clip) += c.len;
        } else {
            in_left = false;
        }
        update_ref_and_read_pos(c, ref_pos, read_pos, ignore_hard_clip);
    }
    return {left_clip, read_pos - right_clip, read_pos};
}

// cigar/mod.rs:164-170
inline size_t get_cigar_read_offset(const CigarVec& cigar, bool ignore_hard_clip) {
    size_t r = 0;
    for (const auto& c : cigar) r += get_cigarseg_read_offset(c, ignore_hard_clip);
    return r;
}
// cigar/mod.rs:174-180
inline int64_t get_cigar_ref_offset(const CigarVec& cigar) {
    int64_t r = 0;
    for (const auto& c : cigar) r += get_cigarseg_ref_offset(c);
    return r;
}

// cigar/mod.rs:204-228.  Note the Pad quirk: a Pad following a Pad is dropped (neither merged nor pushed).
inline CigarVec compress_cigar(const CigarVec& in) {
    CigarVec out;
    Cigar last{M, 0};
    for (const auto& e : in) {
        if (e.len == 0) continue;
        if (e.op == last.op) {
            if (last.op != P) last.len += e.len;
        } else {
            if (last.len != 0) out.push_back(last);
            last = e;
        }
    }
    if (last.len != 0) out.push_back(last);
    return out;
}

// cigar/mod.rs:265-291.  Returns the leading-deletion shift.
inline size_t clean_up_cigar_edge_indels(CigarVec& cigar) {
    auto fix = [](Cigar& c) -> size_t {
        size_t ret = 0;
        if (c.op == D) {
            ret = c.len;
            c = Cigar{S, 0};
        } else if (c.op == I) {
            c = Cigar{S, c.len};
        }
        return ret;
    };
    size_t del_shift = 0;
    for (auto& c : cigar) {
        if (is_alignment_match(c)) break;
        del_shift += fix(c);
    }
    for (auto it = cigar.rbegin(); it != cigar.rend(); ++it) {
        if (is_alignment_match(*it)) break;
        fix(*it);
    }
    return del_shift;
}

// cigar/mod.rs:295-297
inline bool has_aligned_segments(const CigarVec& cigar) {
    return std::any_of(cigar.begin(), cigar.end(), is_alignment_match);
}
// cigar/mod.rs:300-312
inline void strip_leading_clip(CigarVec& cigar) {
    size_t k = 0;
    while (k < cigar.size() && is_clip(cigar[k])) ++k;
    cigar.erase(cigar.begin(), cigar.begin() + k);
}
// cigar/mod.rs:315-327: drops every clip found after the first non-clip element.
inline void strip_trailing_clip(CigarVec& cigar) {
    CigarVec out;
    bool non_clip_found = false;
    for (const auto& c : cigar) {
        if (non_clip_found) {
            if (!is_clip(c)) out.push_back(c);
        } else {
            if (!is_clip(c)) non_clip_found = true;
            out.push_back(c);
        }
    }
    cigar.swap(out);
}

// cigar/clip_alignment.rs:103-156
inline std::pair<CigarVec, int64_t> clip_alignment_read_start(const CigarVec& in, size_t min_left_clip) {
    int64_t ref_pos = 0;
    size_t read_pos = 0;
    CigarVec out;
    int64_t left_ref_clip_shift = 0;
    for (const auto& c : in) {
        switch (c.op) {
            case D: case N:
                if (read_pos <= min_left_clip) left_ref_clip_shift += c.len;
                else out.push_back(c);
                break;
            case I:
                if (read_pos < min_left_clip) out.push_back(Cigar{S, c.len});
                else out.push_back(c);
                break;
            case M: case X: case EQ:
                if (read_pos < min_left_clip) {
                    const int64_t remaining = int64_t(min_left_clip - read_pos);
                    const int64_t match_size = std::max<int64_t>(int64_t(c.len) - remaining, 0);
                    const int64_t clip_size = int64_t(c.len) - match_size;
                    out.push_back(Cigar{S, uint32_t(clip_size)});
                    if (match_size > 0) out.push_back(Cigar{c.op, uint32_t(match_size)});
                    left_ref_clip_shift += clip_size;
                } else {
                    out.push_back(c);
                }
                break;
            default:
                out.push_back(c);
        }
        update_ref_and_read_pos(c, ref_pos, read_pos, false);
    }
    return {out, left_ref_clip_shift};
}

// cigar/clip_alignment.rs:166-181
inline std::pair<CigarVec, int64_t> clip_alignment_read_edges(const CigarVec& in, size_t min_left_clip,
                                                              size_t min_right_clip) {
    CigarVec rev(in.rbegin(), in.rend());
    CigarVec right_clipped = clip_alignment_read_start(rev, min_right_clip).first;
    std::reverse(right_clipped.begin(), right_clipped.end());
    auto lr = clip_alignment_read_start(right_clipped, min_left_clip);
    return {compress_cigar(lr.first), lr.second};
}

// cigar/score_alignment.rs:68-74,138-165.  Throws Panic where the reference returns Err (the caller panics on it,
// contig_repeated_match_trimmer.rs:41-48).
inline double get_gap_compressed_identity_no_align_match(const CigarVec& cigar) {
    uint32_t mismatch_events = 0, match_bases = 0;
    for (const auto& c : cigar) {
        switch (c.op) {
            case I: case D: case N: mismatch_events += 1; break;
            case X: mismatch_events += c.len; break;
            case EQ: match_bases += c.len; break;
            case M: throw Panic(6, "GCI: CIGAR uses alignment match (M) instead of =/X");
            default: break;
        }
    }
    if (match_bases + mismatch_events == 0) return 1.0;
    return double(match_bases) / double(match_bases + mismatch_events);
}

// int_range.rs:11-95 (only what portello calls)
struct IntRange {
    int64_t start, end;
    bool intersect_pos(int64_t pos) const { return pos >= start && pos < end; }
    // NB asymmetric: adjacency on the low side counts (int_range.rs:56-58)
    bool intersect_range(const IntRange& o) const { return o.end >= start && o.start < end; }
    IntRange get_reverse_range(int64_t size) const { return {size - end, size - start}; }
    bool operator==(const IntRange& o) const { return start == o.start && end == o.end; }
};

// seq_util.rs:1-15
inline uint8_t comp_base(uint8_t x) {
    switch (x) {
        case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; case 'N': return 'N';
        case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c'; case 'n': return 'n';
        default: return 'N';
    }
}
// seq_util.rs:29-40 (same result as reversing then complementing every base)
inline void rev_comp_in_place(std::vector<uint8_t>& dna) {
    std::reverse(dna.begin(), dna.end());
    for (auto& b : dna) b = comp_base(b);
}

// rust-htslib `Seq::as_bytes()` decode of BAM 4-bit bases.
static const char kSeqNt16[] = "=ACMGRSVTWYHKDBN";
inline std::vector<uint8_t> decode_seq4(const uint8_t* seq4, size_t len) {
    std::vector<uint8_t> out(len);
    for (size_t i = 0; i < len; ++i) {
        const uint8_t byte = seq4[i >> 1];
        out[i] = uint8_t(kSeqNt16[(i & 1) ? (byte & 0xf) : (byte >> 4)]);
    }
    return out;
}

// bam_utils/util.rs:10-35
inline uint16_t bam_reg2bin(size_t begin, size_t end) {
    const unsigned min_shift = 14, depth = 5;
    end = end - 1;
    unsigned l = depth, s = min_shift;
    size_t t = ((size_t(1) << (depth * 3)) - 1) / 7;
    while (l > 0) {
        if ((begin >> s) == (end >> s)) return uint16_t(t + (begin >> s));
        l -= 1;
        s += 3;
        t -= size_t(1) << (l * 3);
    }
    return 0;
}

// util.rs:50-67
inline std::vector<std::pair<uint64_t, uint64_t>> get_region_segments(uint64_t size, uint64_t segment_size) {
    const uint64_t count = 1 + ((size - 1) / segment_size);
    const uint64_t base = size / count, n_plus_one = size % count;
    std::vector<std::pair<uint64_t, uint64_t>> out;
    uint64_t start = 0;
    for (uint64_t i = 0; i < count; ++i) {
        const uint64_t sz = base + (i < n_plus_one ? 1 : 0);
        const uint64_t end = std::min(start + sz, size);
        out.push_back({start, end});
        start = end;
    }
    return out;
}

// CIGAR text (rust_htslib CigarString Display / TryFrom<&[u8]>)
static const char kOpChars[] = "MIDNSHP=X";
inline std::string cigar_to_string(const CigarVec& c) {
    std::string s;
    for (const auto& e : c) {
        s += std::to_string(e.len);
        s += kOpChars[e.op];
    }
    return s;
}
inline CigarVec cigar_from_string(const std::string& s) {
    CigarVec out;
    uint64_t n = 0;
    bool have = false;
    for (char ch : s) {
        if (ch >= '0' && ch <= '9') {
            n = n * 10 + uint64_t(ch - '0');
            have = true;
        } else {
            const char* p = std::char_traits<char>::find(kOpChars, 9, ch);
            if (!p || !have) throw Panic(6, "bad CIGAR text: " + s);
            out.push_back(Cigar{uint8_t(p - kOpChars), uint32_t(n)});
            n = 0;
            have = false;
        }
    }
    if (have) throw Panic(6, "bad CIGAR text (trailing digits): " + s);
    return out;
}

}  // namespace orc
