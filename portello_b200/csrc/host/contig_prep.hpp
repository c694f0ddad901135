// Host-side contig segment model of the product library (see contig_prep.cpp).
#pragma once
#include <string>
#include <vector>

#include "../../../include/portello_b200.h"
#include "host_util.hpp"

namespace ptl {

// SeqOrderSplitReadSegment of a contig (lib/rust-vc-utils/src/bam_utils/split_read.rs:15-32)
struct HostSegment {
    uint32_t so_start = 0, so_end = 0;
    int32_t chrom = 0;
    int64_t pos = 0;
    bool is_fwd = true;
    uint8_t mapq = 0;
    Ops cigar;
};
// ContigMappingInfo (src/contig_alignment_scanner/mod.rs:38-47) without the tree maps (those live on the device)
struct HostContig {
    std::string name;
    uint64_t length = 0;
    std::vector<HostSegment> segments;
    bool has_rev_seq = false;
    std::vector<uint8_t> rev_seq;
};

std::vector<HostContig> assemble_from_records(const ptl_contig_records& r);
std::vector<HostContig> contigs_from_flat(const ptl_contig_segments& s);
size_t trim_repeated_matches(std::vector<HostContig>& contigs);
size_t join_colinear(std::vector<HostContig>& contigs);

// Flat SoA copy lent out through ptl_get_contig_segments and uploaded to the device.
struct FlatContigs {
    std::vector<uint64_t> contig_len;
    std::vector<uint32_t> seg_begin;
    std::vector<std::vector<uint8_t>> rev_seq;
    std::vector<char> has_rev;
    std::vector<const uint8_t*> rev_ptr;
    std::vector<uint32_t> so_start, so_end;
    std::vector<int32_t> chrom;
    std::vector<int64_t> pos;
    std::vector<uint8_t> is_fwd, mapq;
    std::vector<uint64_t> cigar_begin;
    std::vector<uint32_t> cigar;
    void build(const std::vector<HostContig>& contigs);
    void view(ptl_contig_segments* out) const;
};

}  // namespace ptl
