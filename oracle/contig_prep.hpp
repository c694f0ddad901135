// ORACLE — TEST INFRASTRUCTURE ONLY (see cigar.hpp header).
//
// Restates the table-preparation side of the path:
//   lib/rust-vc-utils/src/bam_utils/aux/sa_tag_parser.rs:25-59         (parse_sa_segment / parse_sa_aux_val)
//   lib/rust-vc-utils/src/bam_utils/split_read.rs:15-155               (SeqOrderSplitReadSegment, get_seq_order_read_split_segments)
//   src/contig_alignment_scanner/mod.rs:25-76,91-183,360-456           (segment assembly, supplementary CIGAR fill-in)
//   src/contig_alignment_scanner/contig_repeated_match_trimmer.rs:18-303
//   src/contig_alignment_scanner/contig_colinear_segment_joiner.rs:15-186
#pragma once
#include <cstdlib>
#include <map>
#include <string>
#include <tuple>
#include <unordered_map>

#include "liftover.hpp"

namespace orc {

// split_read.rs:15-32
struct SeqOrderSplitReadSegment {
    size_t seq_order_read_start = 0, seq_order_read_end = 0;
    size_t chrom_index = 0;
    int64_t pos = 0;
    bool is_fwd_strand = true;
    CigarVec cigar;
    uint8_t mapq = 0;
    bool from_primary_bam_record = false;
};

// sa_tag_parser.rs:4-22
struct SplitReadSegment {
    std::string rname;
    int64_t pos;
    CigarVec cigar;
    bool is_fwd_strand;
    uint8_t mapq;
    int32_t nm;
};

inline std::vector<std::string> split_terminator(const std::string& s, char sep) {
    // Rust str::split_terminator: like split, but a trailing empty piece is dropped.
    std::vector<std::string> out;
    size_t start = 0;
    while (true) {
        size_t p = s.find(sep, start);
        if (p == std::string::npos) {
            if (start < s.size()) out.push_back(s.substr(start));
            break;
        }
        out.push_back(s.substr(start, p - start));
        start = p + 1;
    }
    return out;
}

inline int64_t parse_int_strict(const std::string& s, int64_t lo, int64_t hi) {
    // Rust `str::parse::<iN>()`: optional sign, digits only, range-checked; anything else -> unwrap panic.
    if (s.empty()) throw Panic(6, "SA tag: empty integer field");
    size_t k = 0;
    bool neg = false;
    if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; k = 1; }
    if (k == s.size()) throw Panic(6, "SA tag: bad integer '" + s + "'");
    int64_t v = 0;
    for (; k < s.size(); ++k) {
        if (s[k] < '0' || s[k] > '9') throw Panic(6, "SA tag: bad integer '" + s + "'");
        v = v * 10 + (s[k] - '0');
        if (v > (int64_t(1) << 40)) throw Panic(6, "SA tag: integer out of range '" + s + "'");
    }
    if (neg) v = -v;
    if (v < lo || v > hi) throw Panic(6, "SA tag: integer out of range '" + s + "'");
    return v;
}

// sa_tag_parser.rs:25-46
inline SplitReadSegment parse_sa_segment(const std::string& seg) {
    auto f = split_terminator(seg, ',');
    if (f.size() != 6) throw Panic(6, "Unexpected segment in bam SA tag: " + seg);
    SplitReadSegment out;
    out.rname = f[0];
    out.pos = parse_int_strict(f[1], INT64_MIN / 2, INT64_MAX / 2) - 1;
    out.is_fwd_strand = (f[2] == "+");
    out.cigar = cigar_from_string(f[3]);
    out.mapq = uint8_t(parse_int_strict(f[4], 0, 255));
    out.nm = int32_t(parse_int_strict(f[5], INT32_MIN, INT32_MAX));
    return out;
}
// sa_tag_parser.rs:54-59
inline std::vector<SplitReadSegment> parse_sa_aux_val(const std::string& v) {
    std::vector<SplitReadSegment> out;
    for (const auto& s : split_terminator(v, ';')) out.push_back(parse_sa_segment(s));
    return out;
}

// Minimal stand-in for the fields of a bam::Record the path reads.
struct AlnRecord {
    int32_t tid = 0;
    int64_t pos = 0;
    uint16_t flag = 0;
    uint8_t mapq = 0;
    CigarVec cigar;
    bool has_sa = false;
    std::string sa;
    bool is_reverse() const { return flag & 0x10; }
    bool is_supplementary() const { return flag & 0x800; }
    bool is_secondary() const { return flag & 0x100; }
    bool is_unmapped() const { return flag & 0x4; }
};

using NameIndex = std::unordered_map<std::string, size_t>;

// split_read.rs:56-155
inline std::vector<SeqOrderSplitReadSegment> get_seq_order_read_split_segments(const NameIndex& chrom_index,
                                                                               const AlnRecord& rec) {
    auto seq_order = [](size_t rs, size_t re, size_t size, bool fwd) -> std::pair<size_t, size_t> {
        return fwd ? std::make_pair(rs, re) : std::make_pair(size - re, size - rs);
    };
    std::vector<SeqOrderSplitReadSegment> out;
    const ClipPositions pc = get_read_clip_positions(rec.cigar, false);
    const size_t primary_read_size = pc.size;
    {
        auto so = seq_order(pc.left, pc.right, pc.size, !rec.is_reverse());
        SeqOrderSplitReadSegment s;
        s.seq_order_read_start = so.first;
        s.seq_order_read_end = so.second;
        s.chrom_index = size_t(rec.tid);
        s.pos = rec.pos;
        s.is_fwd_strand = !rec.is_reverse();
        s.cigar = rec.cigar;
        s.mapq = rec.mapq;
        s.from_primary_bam_record = true;
        out.push_back(std::move(s));
    }
    if (rec.has_sa) {
        for (const auto& sa : parse_sa_aux_val(rec.sa)) {
            if (!has_aligned_segments(sa.cigar)) throw Panic(6, "Bam record split segment is unaligned");
            const ClipPositions c = get_read_clip_positions(sa.cigar, false);
            if (c.size != primary_read_size) throw Panic(6, "SA segment read size differs from primary");
            auto so = seq_order(c.left, c.right, c.size, sa.is_fwd_strand);
            auto it = chrom_index.find(sa.rname);
            if (it == chrom_index.end()) throw Panic(6, "SA tag contig not found: " + sa.rname);
            SeqOrderSplitReadSegment s;
            s.seq_order_read_start = so.first;
            s.seq_order_read_end = so.second;
            s.chrom_index = it->second;
            s.pos = sa.pos;
            s.is_fwd_strand = sa.is_fwd_strand;
            s.cigar = sa.cigar;
            s.mapq = sa.mapq;
            s.from_primary_bam_record = false;
            out.push_back(std::move(s));
        }
        std::stable_sort(out.begin(), out.end(), [](const auto& a, const auto& b) {
            return a.seq_order_read_start < b.seq_order_read_start;
        });
    }
    for (const auto& s : out) {
        if (s.seq_order_read_start >= s.seq_order_read_end)
            throw Panic(6, "Can't parse consistent split read information from SA tag");
    }
    return out;
}

// mod.rs:25-47
struct ContigMappingSegmentInfo {
    SeqOrderSplitReadSegment seq_order_segment;
    ReadToRefTreeMap contig_to_ref_map;
};
struct ContigMappingInfo {
    std::string qname;
    std::vector<ContigMappingSegmentInfo> ordered_contig_segment_info;
    bool has_rev_contig_seq = false;
    std::vector<uint8_t> rev_contig_seq;
};
using AllContigMappingInfo = std::vector<ContigMappingInfo>;

// mod.rs:49-56
using SplitReadKey = std::tuple<size_t, int64_t, bool, uint32_t, uint32_t>;
inline SplitReadKey make_split_read_key(size_t chrom, int64_t pos, bool fwd, const CigarVec& cigar) {
    const ClipPositions c = get_read_clip_positions(cigar, false);
    return {chrom, pos, fwd, uint32_t(c.left), uint32_t(c.size - c.right)};
}

// A contig->ref BAM record as the scan sees it (mod.rs:186-240): `seq` is the ASCII decode of the stored bases,
// only needed on the primary record.
struct ContigRecord {
    AlnRecord aln;
    size_t contig_id = 0;
    std::vector<uint8_t> seq;
};

// mod.rs:91-133
inline ContigMappingInfo add_primary_read(const NameIndex& ref_chrom_index, const ContigRecord& rec,
                                          const std::string& qname) {
    ContigMappingInfo info;
    info.qname = qname;
    for (auto& seg : get_seq_order_read_split_segments(ref_chrom_index, rec.aln)) {
        ContigMappingSegmentInfo si;
        if (seg.from_primary_bam_record) si.contig_to_ref_map = get_read_segment_to_ref_pos_tree_map(seg.pos, seg.cigar, false);
        si.seq_order_segment = std::move(seg);
        info.ordered_contig_segment_info.push_back(std::move(si));
    }
    bool need_rev = false;
    for (const auto& s : info.ordered_contig_segment_info) need_rev |= !s.seq_order_segment.is_fwd_strand;
    if (need_rev) {
        info.has_rev_contig_seq = true;
        info.rev_contig_seq = rec.seq;
        if (!rec.aln.is_reverse()) rev_comp_in_place(info.rev_contig_seq);
    }
    return info;
}

// ---------------------------------------------------------------------------------------------------------------
// contig_repeated_match_trimmer.rs:18-49
inline double get_seg_gap_compressed_identity(const SeqOrderSplitReadSegment& seg, const IntRange& isec_seq_order) {
    const size_t read_len = get_cigar_read_offset(seg.cigar, false);
    const IntRange r = seg.is_fwd_strand ? isec_seq_order : isec_seq_order.get_reverse_range(int64_t(read_len));
    const CigarVec clipped = clip_alignment_read_edges(seg.cigar, size_t(r.start), read_len - size_t(r.end)).first;
    return get_gap_compressed_identity_no_align_match(clipped);
}

// contig_repeated_match_trimmer.rs:54-112.  Returns true if the segment is eliminated.
inline bool clip_seg_isec_range(SeqOrderSplitReadSegment& seg, const IntRange& isec_seq_order) {
    const bool clipping_seq_order_prefix = (isec_seq_order.start == int64_t(seg.seq_order_read_start));
    const bool clipping_prefix = clipping_seq_order_prefix ^ (!seg.is_fwd_strand);
    const size_t read_len = get_cigar_read_offset(seg.cigar, false);
    IntRange r = seg.is_fwd_strand ? isec_seq_order : isec_seq_order.get_reverse_range(int64_t(read_len));
    const size_t min_left = clipping_prefix ? size_t(r.end) : 0;
    const size_t min_right = clipping_prefix ? 0 : read_len - size_t(r.start);
    auto clipped = clip_alignment_read_edges(seg.cigar, min_left, min_right);
    seg.cigar = clipped.first;
    seg.pos += clipped.second;
    const ClipPositions c = get_read_clip_positions(seg.cigar, false);
    if (c.left >= c.right) return true;
    if (clipping_prefix) r.end = int64_t(c.left);
    else r.start = int64_t(c.right);
    const IntRange so = seg.is_fwd_strand ? r : r.get_reverse_range(int64_t(read_len));
    if (clipping_seq_order_prefix) seg.seq_order_read_start = size_t(so.end);
    else seg.seq_order_read_end = size_t(so.start);
    return false;
}

// contig_repeated_match_trimmer.rs:117-136
inline bool clip_seg_info_isec_range(ContigMappingSegmentInfo& si, const IntRange& isec) {
    if (clip_seg_isec_range(si.seq_order_segment, isec)) return true;
    si.contig_to_ref_map = get_read_segment_to_ref_pos_tree_map(si.seq_order_segment.pos, si.seq_order_segment.cigar, false);
    return false;
}

// contig_repeated_match_trimmer.rs:144-204.  Returns false if no intersection.
inline bool get_seg_clip_info(const ContigMappingInfo& info, size_t i1, size_t i2, IntRange& isec, size_t& clip_index) {
    const auto& seg1 = info.ordered_contig_segment_info[i1].seq_order_segment;
    const auto& seg2 = info.ordered_contig_segment_info[i2].seq_order_segment;
    if (seg1.seq_order_read_end <= seg2.seq_order_read_start) return false;
    isec = IntRange{int64_t(seg2.seq_order_read_start), int64_t(seg1.seq_order_read_end)};
    const double gci1 = get_seg_gap_compressed_identity(seg1, isec);
    const double gci2 = get_seg_gap_compressed_identity(seg2, isec);
    // seg2_gci.partial_cmp(&seg1_gci).unwrap().then(seg2.mapq.cmp(&seg1.mapq)) == Greater
    bool clip_seg1;
    if (gci2 > gci1) clip_seg1 = true;
    else if (gci2 < gci1) clip_seg1 = false;
    else clip_seg1 = seg2.mapq > seg1.mapq;
    clip_index = clip_seg1 ? i1 : i2;
    return true;
}

// contig_repeated_match_trimmer.rs:214-303.  Returns the number of clip operations.
inline size_t clip_repeated_contig_matches(AllContigMappingInfo& result) {
    size_t segments_clipped = 0;
    for (auto& info : result) {
        auto& segs = info.ordered_contig_segment_info;
        if (segs.empty()) continue;
        const size_t n = segs.size();
        std::vector<bool> eliminated(n, false);
        for (size_t i1 = 0; i1 < n; ++i1) {
            for (size_t i2 = i1 + 1; i2 < n; ++i2) {
                if (eliminated[i1] || eliminated[i2]) continue;
                IntRange isec{0, 0};
                size_t clip_index = 0;
                if (!get_seg_clip_info(info, i1, i2, isec, clip_index)) break;
                if (clip_seg_info_isec_range(segs[clip_index], isec)) eliminated[clip_index] = true;
                ++segments_clipped;
            }
        }
        std::vector<ContigMappingSegmentInfo> kept;
        for (size_t i = 0; i < n; ++i)
            if (!eliminated[i]) kept.push_back(std::move(segs[i]));
        segs.swap(kept);
    }
    return segments_clipped;
}

// ---------------------------------------------------------------------------------------------------------------
// contig_colinear_segment_joiner.rs:15-24
inline int64_t get_seg_ref_gap(const SeqOrderSplitReadSegment& s1, const SeqOrderSplitReadSegment& s2) {
    if (s1.is_fwd_strand) return s2.pos - (s1.pos + get_cigar_ref_offset(s1.cigar));
    return s1.pos - (s2.pos + get_cigar_ref_offset(s2.cigar));
}
// :27-49
inline bool are_segments_joinable(const SeqOrderSplitReadSegment& s1, const SeqOrderSplitReadSegment& s2) {
    const int64_t max_segment_ref_gap = 1000;
    if (s1.chrom_index != s2.chrom_index || s1.is_fwd_strand != s2.is_fwd_strand) return false;
    const int64_t gap = get_seg_ref_gap(s1, s2);
    if (gap < 0 || gap > max_segment_ref_gap) return false;
    return s1.mapq == s2.mapq;
}
// :57-122
inline void join_segments(ContigMappingSegmentInfo& si1, ContigMappingSegmentInfo si2) {
    auto& s1 = si1.seq_order_segment;
    auto& s2 = si2.seq_order_segment;
    const int64_t join_del = get_seg_ref_gap(s1, s2);
    if (join_del < 0) throw Panic(6, "join_segments: negative ref gap");
    if (s2.seq_order_read_start < s1.seq_order_read_end) throw Panic(6, "join_segments: overlapping segments");
    const size_t join_ins = s2.seq_order_read_start - s1.seq_order_read_end;
    auto join_cigars = [&](CigarVec& a, CigarVec& b) {
        strip_trailing_clip(a);
        if (join_ins > 0) a.push_back(Cigar{I, uint32_t(join_ins)});
        if (join_del > 0) a.push_back(Cigar{D, uint32_t(join_del)});
        strip_leading_clip(b);
        a.insert(a.end(), b.begin(), b.end());
        b.clear();
    };
    if (s1.is_fwd_strand) {
        join_cigars(s1.cigar, s2.cigar);
    } else {
        join_cigars(s2.cigar, s1.cigar);
        std::swap(s1.cigar, s2.cigar);
        s1.pos = s2.pos;
    }
    s1.seq_order_read_end = s2.seq_order_read_end;
    si1.contig_to_ref_map = get_read_segment_to_ref_pos_tree_map(s1.pos, s1.cigar, false);
}
// :124-186.  Returns the number of joins.
inline size_t join_colinear_contig_segments(AllContigMappingInfo& result) {
    size_t joined = 0;
    for (auto& info : result) {
        if (info.ordered_contig_segment_info.empty()) continue;
        std::vector<ContigMappingSegmentInfo> old;
        old.swap(info.ordered_contig_segment_info);
        auto& out = info.ordered_contig_segment_info;
        for (auto& seg : old) {
            if (out.empty()) {
                out.push_back(std::move(seg));
                continue;
            }
            auto& last = out.back();
            if (seg.seq_order_segment.seq_order_read_start < last.seq_order_segment.seq_order_read_end)
                throw Panic(6, "Incomplete repeat trimming on qname: " + info.qname);
            if (are_segments_joinable(last.seq_order_segment, seg.seq_order_segment)) {
                join_segments(last, std::move(seg));
                ++joined;
            } else {
                out.push_back(std::move(seg));
            }
        }
    }
    return joined;
}

// ---------------------------------------------------------------------------------------------------------------
// scan_contig_bam minus BAM I/O (mod.rs:186-240 record dispatch, :360-439 fill-in, :452-456 trim + join).
// `records` are all records of the contig->ref BAM in any order; windows/threads do not affect the result.
inline AllContigMappingInfo scan_contig_records(const std::vector<ContigRecord>& records, size_t n_contigs,
                                                const NameIndex& ref_chrom_index,
                                                const std::vector<std::string>& contig_names, bool do_trim = true,
                                                bool do_join = true) {
    struct Cell {
        bool has_info = false;
        ContigMappingInfo info;
        std::map<SplitReadKey, std::pair<CigarVec, ReadToRefTreeMap>> supp_cigars;
    };
    std::vector<Cell> cells(n_contigs);
    for (const auto& rec : records) {
        if (rec.aln.is_unmapped() || rec.aln.is_secondary()) continue;  // mod.rs:208-210
        Cell& cell = cells.at(rec.contig_id);
        if (!rec.aln.is_supplementary()) {
            cell.info = add_primary_read(ref_chrom_index, rec, contig_names[rec.contig_id]);
            cell.has_info = true;
        } else {  // mod.rs:135-183
            SplitReadKey key = make_split_read_key(size_t(rec.aln.tid), rec.aln.pos, !rec.aln.is_reverse(), rec.aln.cigar);
            auto map = get_read_segment_to_ref_pos_tree_map(rec.aln.pos, rec.aln.cigar, false);
            auto ins = cell.supp_cigars.emplace(key, std::make_pair(rec.aln.cigar, std::move(map)));
            if (!ins.second)
                throw Panic(6, "Can't uniquely identify split read alignment info in contig '" + contig_names[rec.contig_id] + "'");
        }
    }
    AllContigMappingInfo result;
    for (size_t ci = 0; ci < n_contigs; ++ci) {
        Cell& cell = cells[ci];
        ContigMappingInfo info = cell.has_info ? std::move(cell.info) : ContigMappingInfo{};  // unwrap_or_default
        for (auto& seg : info.ordered_contig_segment_info) {
            auto& s = seg.seq_order_segment;
            if (s.from_primary_bam_record) continue;
            SplitReadKey key = make_split_read_key(s.chrom_index, s.pos, s.is_fwd_strand, s.cigar);
            auto it = cell.supp_cigars.find(key);
            if (it == cell.supp_cigars.end())
                throw Panic(6, "Can't find supplementary alignment record corresponding to segment reported in SA tag");
            s.cigar = it->second.first;
            seg.contig_to_ref_map = it->second.second;
        }
        result.push_back(std::move(info));
    }
    if (do_trim) clip_repeated_contig_matches(result);
    if (do_join) join_colinear_contig_segments(result);
    return result;
}

}  // namespace orc
