// TEST INFRASTRUCTURE ONLY (see emul_prep.hpp).
#include "emul_prep.hpp"

#include <cstdlib>

#include "../../portello_b200/csrc/host/contig_prep.hpp"

using namespace ptl;

void EmulFlat::view(ptl_contig_segments* out) const {
    out->n_contigs = uint32_t(contig_len.size());
    out->contig_len = contig_len.data();
    out->contig_seg_begin = seg_begin.data();
    out->rev_contig_seq = rev_ptr.data();
    out->n_segments = uint32_t(so_start.size());
    out->seg_seq_order_start = so_start.data();
    out->seg_seq_order_end = so_end.data();
    out->seg_chrom_index = chrom.data();
    out->seg_pos = pos.data();
    out->seg_is_fwd = is_fwd.data();
    out->seg_mapq = mapq.data();
    out->seg_cigar_begin = cigar_begin.data();
    out->cigar = cigar.data();
}

int emul_prepare(int mode, const void* in, EmulFlat* out, std::string* err) {
    try {
        std::vector<HostContig> contigs = (mode == 2) ? assemble_from_records(*static_cast<const ptl_contig_records*>(in))
                                                      : contigs_from_flat(*static_cast<const ptl_contig_segments*>(in));
        if (mode >= 1) {
            trim_repeated_matches(contigs);
            join_colinear(contigs);
        }
        FlatContigs f;
        f.build(contigs);
        out->contig_len = f.contig_len;
        out->seg_begin = f.seg_begin;
        out->rev_seq = f.rev_seq;
        out->has_rev = f.has_rev;
        out->so_start = f.so_start;
        out->so_end = f.so_end;
        out->chrom = f.chrom;
        out->pos = f.pos;
        out->is_fwd = f.is_fwd;
        out->mapq = f.mapq;
        out->cigar_begin = f.cigar_begin;
        out->cigar = f.cigar;
        out->rev_ptr.assign(out->contig_len.size(), nullptr);
        for (size_t c = 0; c < out->contig_len.size(); ++c)
            if (out->has_rev[c]) out->rev_ptr[c] = out->rev_seq[c].data();
        return PTL_OK;
    } catch (const InputError& e) {
        *err = e.what();
        return PTL_ERR_INPUT;
    } catch (const std::exception& e) {
        *err = e.what();
        return PTL_ERR_INVALID_ARG;
    }
}

// pack.cpp (linked for ptl_pack_split_segments, which the contig assembly uses) refers to the pinned allocator of the
// CUDA library; the emulation never asks for pinned batches.
extern "C" void* ptl_host_alloc(size_t bytes) { return std::malloc(bytes ? bytes : 1); }
extern "C" void ptl_host_free(void* p) { std::free(p); }
