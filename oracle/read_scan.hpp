// ORACLE — TEST INFRASTRUCTURE ONLY (see cigar.hpp header).
//
// Restates the per-read driver of src/read_alignment_scanner.rs minus BAM I/O:
//   :80-103   get_contig_split_segments_from_read_mapping
//   :136-288  get_liftover_alignment_for_read_and_contig_segment
//   :292-301  get_sa_tag_segment
//   :310-366  finish_remapped_alignment_set
//   :419-487  loop body of scan_chromosome_segment
#pragma once
#include "contig_prep.hpp"

namespace orc {

// A primary read->contig record as scan_chromosome_segment sees it.
struct ReadRecord {
    uint16_t flag = 0;
    uint8_t mapq = 0;
    uint16_t bin = 0;
    uint32_t seq_len = 0;
    const uint8_t* seq4 = nullptr;  // BAM 4-bit packed bases
    bool is_reverse() const { return flag & 0x10; }
};

// The fields of a remapped bam::Record that the path computes (everything else is copied from the input record).
struct LiftedRecord {
    int8_t status = 1;            // 1 lifted, 0 unmapped fallback
    uint32_t read_segment = 0;    // index into the read's ordered split segments
    uint32_t contig_segment = 0;  // PS split index (:260)
    uint32_t contig_index = 0;    // read_segment.chrom_index (PS contig name, :255-256)
    bool contig_is_fwd = true;    // PS strand char (:261)
    int32_t tid = -1;
    int64_t pos = -1;
    uint8_t mapq = 255;
    uint8_t zm = 0;               // ZM:C = original primary record MAPQ (:251,267-269)
    uint16_t flag = 0;
    uint16_t bin = 0;
    bool need_flip = false;
    CigarVec cigar;
};

struct ScanOptions {
    // The reference decodes the full read twice per pair and clones the record (:170,238,245). `faithful_decode`
    // reproduces those decodes (cost model of the reference); false decodes lazily only what is compared.
    bool faithful_decode = true;
    uint32_t stage_mask = 7;  // PTL_STAGE_* (testing only)
};

// :80-103
inline std::vector<size_t> get_contig_split_segments_from_read_mapping(
    const SeqOrderSplitReadSegment& read_segment, const std::vector<ContigMappingSegmentInfo>& contig_segments) {
    std::vector<size_t> ret;
    const IntRange read_range{read_segment.pos, read_segment.pos + get_cigar_ref_offset(read_segment.cigar)};
    for (size_t i = 0; i < contig_segments.size(); ++i) {
        const auto& s = contig_segments[i].seq_order_segment;
        const IntRange seg_range{int64_t(s.seq_order_read_start), int64_t(s.seq_order_read_end)};
        if (seg_range.intersect_range(read_range)) ret.push_back(i);
    }
    return ret;
}

struct PairStats {
    uint64_t n_pairs = 0, n_lifted = 0, n_in_ops = 0, n_out_ops = 0;
    CompareStats cmp;
};

// :136-288.  Returns false for None.  Throws Panic(-1) on the length mismatch (:204-229), Panic(-2) on
// out-of-bounds base access.
inline bool get_liftover_alignment_for_read_and_contig_segment(
    const std::vector<std::vector<uint8_t>>& reference, const std::vector<uint64_t>& contig_len,
    const ReadRecord& record, const SeqOrderSplitReadSegment& read_segment, size_t contig_segment_index,
    const ContigMappingSegmentInfo& cseg, const ContigMappingInfo& cinfo, const ScanOptions& opt, PairStats& stats,
    LiftedRecord& out) {
    const bool contig_is_fwd = cseg.seq_order_segment.is_fwd_strand;
    const bool read_segment_changes_strand = (record.is_reverse() == read_segment.is_fwd_strand);
    const bool need_flip = (!contig_is_fwd) ^ read_segment_changes_strand;

    // faithful: materialise the ASCII read (and its reverse complement) exactly where the reference does (:170,:238);
    // lean: index the packed bases lazily (same bytes, no decode cost).
    std::vector<uint8_t> read_ascii;
    ReadSeq read_seq = ReadSeq::from_seq4(record.seq4, record.seq_len, need_flip);
    auto decode_read = [&]() {
        if (!opt.faithful_decode) return;
        read_ascii = decode_seq4(record.seq4, record.seq_len);
        if (need_flip) rev_comp_in_place(read_ascii);
        read_seq = ReadSeq::from_ascii(read_ascii.data(), read_ascii.size());
    };

    int64_t pos_on_ref_strand = read_segment.pos;
    CigarVec cigar_on_ref_strand = read_segment.cigar;
    if (!contig_is_fwd) {
        const int64_t contig_length = int64_t(contig_len[read_segment.chrom_index]);
        const int64_t read_segment_end = read_segment.pos + get_cigar_ref_offset(read_segment.cigar);
        const int64_t rev_pos = contig_length - read_segment_end;
        // Invalid input (read runs past the contig end): the reference wraps `rev_pos as usize` and then either panics
        // inside BTreeMap::range (start > end) or looks up garbage.  Both implementations report a bounds panic.
        if (rev_pos < 0) throw Panic(-2, "read alignment extends past the end of a reverse-strand contig segment");
        CigarVec rev_cigar(read_segment.cigar.rbegin(), read_segment.cigar.rend());
        if (opt.stage_mask & 1) {
            decode_read();
            if (!cinfo.has_rev_contig_seq) throw Panic(6, "rev_contig_seq missing for reverse-strand contig segment");
            PosCigar shifted = left_shift_indels(rev_pos, rev_cigar, cinfo.rev_contig_seq.data(),
                                                 cinfo.rev_contig_seq.size(), read_seq, &stats.cmp);
            pos_on_ref_strand = shifted.pos;
            cigar_on_ref_strand = std::move(shifted.cigar);
        } else {
            pos_on_ref_strand = rev_pos;
            cigar_on_ref_strand = std::move(rev_cigar);
        }
    }

    out.need_flip = need_flip;
    out.contig_segment = uint32_t(contig_segment_index);
    out.contig_index = uint32_t(read_segment.chrom_index);
    out.contig_is_fwd = contig_is_fwd;
    out.zm = record.mapq;
    out.status = 1;

    int64_t ref2_pos = pos_on_ref_strand;
    CigarVec ref2_cigar = cigar_on_ref_strand;
    int32_t tid = -2;
    if (opt.stage_mask & 2) {
        auto lifted = liftover_read_alignment(cseg.contig_to_ref_map, pos_on_ref_strand, cigar_on_ref_strand);
        if (!lifted) return false;
        if (size_t(record.seq_len) != get_cigar_read_offset(lifted->cigar, false))
            throw Panic(-1, "lifted CIGAR read length != seq_len");
        ref2_pos = lifted->pos;
        ref2_cigar = std::move(lifted->cigar);
        tid = int32_t(cseg.seq_order_segment.chrom_index);
    }
    if (opt.stage_mask & 4) {
        const size_t chrom_index = cseg.seq_order_segment.chrom_index;
        const auto& chrom_ref = reference.at(chrom_index);
        decode_read();
        PosCigar simple = simplify_alignment_indels(ref2_pos, ref2_cigar, chrom_ref.data(), chrom_ref.size(), read_seq,
                                                    &stats.cmp);
        ref2_pos = simple.pos;
        ref2_cigar = std::move(simple.cigar);
    }

    out.tid = tid;
    out.mapq = cseg.seq_order_segment.mapq;
    out.pos = ref2_pos;
    out.cigar = std::move(ref2_cigar);
    uint16_t flag = record.flag;
    if (need_flip) flag ^= 0x10;
    const int64_t ref2_end = ref2_pos + get_cigar_ref_offset(out.cigar);  // get_alignment_end
    out.bin = bam_reg2bin(size_t(ref2_pos), size_t(ref2_end));
    flag |= 0x800;  // set_supplementary until the primary is chosen (:282)
    out.flag = flag;
    return true;
}

// :310-366 (is_target_region = false).  Field part only; SA text is format_sa_tags below.
inline std::vector<LiftedRecord> finish_remapped_alignment_set(const ReadRecord& orig, std::vector<LiftedRecord> recs) {
    if (recs.empty()) {
        LiftedRecord u;
        u.status = 0;
        uint16_t flag = orig.flag;
        flag |= 0x4;     // set_unmapped
        flag &= ~0x800;  // unset_supplementary
        u.mapq = 255;
        u.tid = -1;
        u.pos = -1;
        u.bin = orig.bin;  // untouched (:322-334)
        u.zm = orig.mapq;
        if (flag & 0x10) {  // reverse_alignment_seq_and_qual
            flag ^= 0x10;
            u.need_flip = true;
        }
        u.flag = flag;
        return {u};
    }
    size_t primary = 0;
    for (size_t i = 1; i < recs.size(); ++i)
        if (recs[primary].mapq < recs[i].mapq) primary = i;
    recs[primary].flag &= ~0x800;
    return recs;
}

// :292-301 + :349-363
inline std::vector<std::string> format_sa_tags(const std::vector<std::string>& chrom_names,
                                               const std::vector<LiftedRecord>& recs) {
    std::vector<std::string> out(recs.size());
    if (recs.size() == 1 && recs[0].status == 0) return out;
    auto seg = [&](const LiftedRecord& r) {
        return chrom_names.at(size_t(r.tid)) + "," + std::to_string(r.pos + 1) + "," + ((r.flag & 0x10) ? "-" : "+") +
               "," + cigar_to_string(r.cigar) + "," + std::to_string(unsigned(r.mapq)) + ",0;";
    };
    for (size_t i = 0; i < recs.size(); ++i)
        for (size_t j = 0; j < recs.size(); ++j)
            if (j != i) out[i] += seg(recs[j]);
    return out;
}

// :419-479 for one primary record whose split segments are already ordered.
inline std::vector<LiftedRecord> lift_read(const std::vector<std::vector<uint8_t>>& reference,
                                           const std::vector<uint64_t>& contig_len,
                                           const AllContigMappingInfo& all_contig_mapping_info, const ReadRecord& record,
                                           const std::vector<SeqOrderSplitReadSegment>& ordered_splits,
                                           const ScanOptions& opt, PairStats& stats) {
    std::vector<LiftedRecord> remapped;
    for (size_t rsi = 0; rsi < ordered_splits.size(); ++rsi) {
        const auto& read_segment = ordered_splits[rsi];
        const auto& cinfo = all_contig_mapping_info.at(read_segment.chrom_index);
        const auto& contig_segments = cinfo.ordered_contig_segment_info;
        for (size_t csi : get_contig_split_segments_from_read_mapping(read_segment, contig_segments)) {
            LiftedRecord rec;
            stats.n_pairs++;
            stats.n_in_ops += read_segment.cigar.size();
            if (get_liftover_alignment_for_read_and_contig_segment(reference, contig_len, record, read_segment, csi,
                                                                   contig_segments[csi], cinfo, opt, stats, rec)) {
                rec.read_segment = uint32_t(rsi);
                stats.n_lifted++;
                stats.n_out_ops += rec.cigar.size();
                remapped.push_back(std::move(rec));
            }
        }
    }
    if (opt.stage_mask != 7) return remapped;  // stage tests: raw per-pair outputs, no finish
    return finish_remapped_alignment_set(record, std::move(remapped));
}

}  // namespace orc
