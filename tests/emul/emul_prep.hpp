// TEST INFRASTRUCTURE ONLY.  Bridge between the product's host contig preparation (contig_prep.cpp) and the emulated
// device pipeline (emul.cpp); a separate translation unit because the host and device CIGAR headers both define ptl::OP_*.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/portello_b200.h"

struct EmulFlat {
    std::vector<uint64_t> contig_len;
    std::vector<uint32_t> seg_begin;
    std::vector<std::vector<uint8_t>> rev_seq;
    std::vector<char> has_rev;
    std::vector<uint32_t> so_start, so_end;
    std::vector<int32_t> chrom;
    std::vector<int64_t> pos;
    std::vector<uint8_t> is_fwd, mapq;
    std::vector<uint64_t> cigar_begin;
    std::vector<uint32_t> cigar;
    std::vector<const uint8_t*> rev_ptr;
    void view(ptl_contig_segments* out) const;
};

// mode 0: already trimmed/joined segments, 1: raw segments (trim + join), 2: contig BAM records (assemble + trim + join).
// Returns 0 or PTL_ERR_INPUT / PTL_ERR_INVALID_ARG with *err set.
int emul_prepare(int mode, const void* in, EmulFlat* out, std::string* err);
