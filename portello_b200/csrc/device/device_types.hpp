// Device-side data model shared by the kernels (kernels.cu) and the host orchestration (context.cu).
// Layout rationale is in DESIGN.md §3 ("Data layout in HBM").
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define PTL_HD __host__ __device__
#else
#define PTL_HD
#endif

namespace ptl {

// One entry of a flat ReadToRefTreeMap (lib/rust-vc-utils/src/bam_utils/read_to_ref_map.rs:59-137), 16 bytes so that a
// key crossing of the liftover walk is ONE 128-bit load.
//   key : contig read_pos where the block starts
//   val : reference position of the block start, or -1 for a None block (run end)
//   gap : for a Some block, max(0, val - reference end of the previous aligned run of the same segment) = the length
//         of the deletion update_ref2_cigar_segment (src/liftover_read_alignment.rs:91-96) pushes when the walk enters
//         this block from the previous one; 0 for the first run and for None blocks.  Precomputed because the walk
//         covers the contig interval contiguously: when it enters a block, `ref2_end_pos` is always the end of the
//         previous aligned run (if the walk touched it at all).
struct alignas(16) TabEntry {
    uint32_t key;
    int32_t val;
    uint32_t gap;
    uint32_t pad;
};

// ---- static, read-only state (replicated per GPU): reference, contigs, segment tables ------------------------
struct DevStatic {
    // reference genome, ASCII, chromosomes concatenated (ptl_set_reference)
    const uint8_t* ref = nullptr;
    const uint64_t* chrom_off = nullptr;  // [n_chrom+1]
    uint32_t n_chrom = 0;
    // contigs
    uint32_t n_contigs = 0;
    const uint32_t* contig_seg_begin = nullptr;  // [n_contigs+1]
    const uint64_t* contig_len = nullptr;        // [n_contigs]
    const uint64_t* contig_rev_off = nullptr;    // [n_contigs] offset into rev_pool, ~0 = None
    const uint8_t* rev_pool = nullptr;           // rev_contig_seq bytes, ASCII
    // contig->ref segments (post trim/join), SoA
    uint32_t n_segments = 0;
    const uint32_t* seg_so_start = nullptr;
    const uint32_t* seg_so_end = nullptr;
    const int32_t* seg_chrom = nullptr;
    const int64_t* seg_pos = nullptr;
    const uint8_t* seg_is_fwd = nullptr;
    const uint8_t* seg_mapq = nullptr;
    const uint64_t* seg_cigar_begin = nullptr;   // [n_segments+1]
    const uint32_t* seg_cigar = nullptr;
    // flat ReadToRefTreeMap: per segment a sorted run of TabEntry
    const uint32_t* seg_tab_begin = nullptr;     // [n_segments+1]
    const TabEntry* table = nullptr;
};

// ---- one batch on the device ------------------------------------------------------------------------------------
struct DevBatch {
    uint32_t n_reads = 0, n_rsegs = 0;
    uint64_t n_cigar = 0;
    const uint16_t* read_flag = nullptr;
    const uint8_t* read_mapq = nullptr;
    const uint16_t* read_bin = nullptr;
    const uint32_t* read_seq_len = nullptr;
    const uint64_t* read_seq_off = nullptr;
    const uint32_t* read_seg_begin = nullptr;
    const uint32_t* rseg_contig = nullptr;
    const int64_t* rseg_pos = nullptr;
    const uint8_t* rseg_is_fwd = nullptr;
    const uint64_t* rseg_cigar_begin = nullptr;
    const uint32_t* rseg_cigar_len = nullptr;
    const uint32_t* cigar = nullptr;
    const uint8_t* seq4 = nullptr;  // device copy, or a mapped pinned host pointer in zero-copy mode
    uint64_t seq4_bytes = 0;
    // optional indel windows (ptl_batch.indel_win): the first 16 read nibbles the homology walk of each I/D cluster touches
    const uint64_t* indel_win = nullptr;
    const uint32_t* rseg_win_begin = nullptr;  // [n_rsegs+1], nullptr = no windows
};


// What lift_pairs_kernel needs to know about a pair before it can walk it, packed by pair_fill_kernel into ONE 32-byte
// sector (the lanes of a tile own work-sorted, i.e. scattered, pairs: the same facts read from fifteen separate arrays cost
// fifteen sectors and a chain of four dependent loads per pair: 20 % of the kernel's stall samples on the whole-genome
// workload, ncu r06a).  Reverse-strand pairs read a second sector, PairDescRev.
struct alignas(16) PairDesc {
    uint64_t cigar_begin;  // first op of the read-segment CIGAR in the batch pool
    uint32_t n_ops;        // ops of that CIGAR
    uint32_t cpos;         // where the walk starts, on the strand the segment's table is written in
    uint32_t tab_lo;       // first table entry with key >= cpos (cursor hint)
    uint32_t t0, t1;       // the segment's table range
    uint32_t flags_cap_b;  // bits 0-7: kPd* flags; bits 8-31: capacity of the pair's buffer B (slot bound, ops)
};
enum : uint32_t {
    kPdContigFwd = 1,   // the contig segment is forward-strand
    kPdNeedFlip = 2,    // need_flipped_read_alignment (src/read_alignment_scanner.rs:153-157)
    kPdErrBounds = 4,   // the read runs past the end of a reverse-strand contig (negative rev_pos, :166)
    kPdErrLength = 8,   // read length of the segment CIGAR != seq_len (the reference's assertion, :204-229)
    kPdNoRevSeq = 16    // reverse-strand segment on a contig without rev_contig_seq (Option::unwrap on None, :174)
};
struct alignas(16) PairDescRev {
    uint64_t seq_off;      // read_seq_off of the read
    uint64_t rev_off;      // contig_rev_off of the contig (~0 = none)
    uint32_t contig_len;
    uint32_t seq_len;
    uint32_t win_begin;    // indel windows of the read segment: [win_begin, win_begin + n_win) (0, 0 without windows)
    uint32_t n_win;
};
static_assert(sizeof(PairDesc) == 32 && sizeof(PairDescRev) == 32, "one sector each");

// ---- per-batch work arrays ---------------------------------------------------------------------------------------
struct DevWork {
    // per read segment
    uint32_t* rseg_read = nullptr;       // owning read
    uint32_t* rseg_pair_begin = nullptr; // [n_rsegs+1] (count, then exclusive scan in place)
    int64_t* rseg_ref_len = nullptr;     // get_cigar_ref_offset of the segment
    uint32_t* rseg_n_id = nullptr;       // number of non-match ops (I, D, N, S, H, P) of the segment CIGAR (slot bounds)
    uint32_t* rseg_read_len = nullptr;   // read bases consumed by the segment CIGAR incl. hard clips (length check)
    // per pair
    uint32_t pair_cap = 0;
    uint32_t* pair_rseg = nullptr;
    uint32_t* pair_seg = nullptr;        // global contig-segment index
    uint64_t* pair_slot_begin = nullptr; // [pair_cap+1] (bound, then exclusive scan in place) ops of scratch
    uint32_t* pair_cap_b = nullptr;      // capacity of the pair's buffer B (the slot is [B | A])
    uint32_t* pair_tab_lo = nullptr;     // first table entry with key >= the start of the walked contig interval (cursor hint)
    int8_t* pair_status = nullptr;
    uint8_t* pair_flip = nullptr;
    int64_t* pair_pos = nullptr;
    uint32_t* pair_n_out = nullptr;
    uint16_t* pair_bin = nullptr;        // bam_reg2bin(pos, end) of the final record
    uint64_t* pair_out_off = nullptr;    // where the final ops of the pair start in scratch
    uint32_t* simplify_list = nullptr;   // [pair_cap] pairs whose lifted CIGAR needs simplify_alignment_indels
    uint32_t* long_list = nullptr;       // [pair_cap] pairs whose liftover runs warp-cooperatively (status ST_PENDING_LIFT)
    uint32_t* pair_order = nullptr;      // [pair_cap] work order of lift_pairs_kernel (pair_order_kernel)
    PairDesc* pair_desc = nullptr;       // [pair_cap] (pair_fill_kernel)
    PairDescRev* pair_desc_rev = nullptr;  // [pair_cap], written for reverse-strand pairs only
    uint32_t long_ops = 64;              // a pair with more CIGAR ops than this goes to long_list
    // scratch op slots, followed in the same allocation by the dense output region of the lift kernel: tile t (32 pairs)
    // of lift_pairs_kernel owns scratch[dense_off + t * kLiftTileOut, + kLiftTileOut) and packs the lifted CIGARs of its
    // pairs there back to back (pair_out_off is an offset into `scratch` either way)
    uint64_t scratch_cap = 0;            // in ops
    uint32_t* scratch = nullptr;
    uint64_t dense_off = 0;              // in ops
    // per read
    uint2* read_counts = nullptr;        // [n_reads+1] (.x records, .y ops) -> exclusive scan
    uint32_t* read_primary = nullptr;    // pair index chosen as primary, ~0 if none lifted
};

// lift_pairs_kernel: the dense output region of a tile of 32 pairs, in ops
#ifndef LIFT_TILE_OUT
#define LIFT_TILE_OUT 1536
#endif
constexpr uint32_t kLiftTileOut = LIFT_TILE_OUT;

// ---- results of one batch: ONE compact arena (device), copied to the host with a single DMA ------------------------
// [header: DevTotals, 128 B][read_rec_begin][rec_cigar_begin][rec_pos][rec_read_segment][rec_contig_segment][rec_tid]
// [rec_flag][rec_bin][rec_status][rec_mapq][rec_need_flip][cigar], every array sized EXACTLY by the batch's own record and
// op counts (known on the device after the scan of the per-read counts) and 16-byte aligned.  The host derives the same
// layout from the header, so no per-array copies and no size round trip before the copy are needed.
struct ResultLayout {
    uint64_t read_rec_begin, rec_cigar_begin, rec_pos, rec_rseg, rec_cseg, rec_tid, rec_flag, rec_bin, rec_status, rec_mapq, rec_flip, cigar, total;
};
constexpr uint64_t kResultHeaderBytes = 128;
PTL_HD inline ResultLayout result_layout(uint64_t n_reads, uint64_t n_rec, uint64_t n_cigar) {
    ResultLayout L;
    uint64_t off = kResultHeaderBytes;
    auto take = [&](uint64_t bytes) { const uint64_t o = off; off = (off + bytes + 15ull) & ~15ull; return o; };
    L.read_rec_begin = take((n_reads + 1) * 4);
    L.rec_cigar_begin = take((n_rec + 1) * 8);
    L.rec_pos = take(n_rec * 8);
    L.rec_rseg = take(n_rec * 4);
    L.rec_cseg = take(n_rec * 4);
    L.rec_tid = take(n_rec * 4);
    L.rec_flag = take(n_rec * 2);
    L.rec_bin = take(n_rec * 2);
    L.rec_status = take(n_rec);
    L.rec_mapq = take(n_rec);
    L.rec_flip = take(n_rec);
    L.cigar = take(n_cigar * 4);
    L.total = off;
    return L;
}

// typed view of an arena (mirrors ptl_result)
struct DevResult {
    uint32_t* read_rec_begin = nullptr;  // [n_reads+1]
    int8_t* rec_status = nullptr;
    uint32_t* rec_read_segment = nullptr;
    uint32_t* rec_contig_segment = nullptr;
    int32_t* rec_tid = nullptr;
    int64_t* rec_pos = nullptr;
    uint8_t* rec_mapq = nullptr;
    uint16_t* rec_flag = nullptr;
    uint16_t* rec_bin = nullptr;
    uint8_t* rec_need_flip = nullptr;
    uint64_t* rec_cigar_begin = nullptr; // [n_rec+1]
    uint32_t* cigar = nullptr;
    PTL_HD static DevResult view(char* base, const ResultLayout& L) {
        DevResult R;
        R.read_rec_begin = reinterpret_cast<uint32_t*>(base + L.read_rec_begin);
        R.rec_cigar_begin = reinterpret_cast<uint64_t*>(base + L.rec_cigar_begin);
        R.rec_pos = reinterpret_cast<int64_t*>(base + L.rec_pos);
        R.rec_read_segment = reinterpret_cast<uint32_t*>(base + L.rec_rseg);
        R.rec_contig_segment = reinterpret_cast<uint32_t*>(base + L.rec_cseg);
        R.rec_tid = reinterpret_cast<int32_t*>(base + L.rec_tid);
        R.rec_flag = reinterpret_cast<uint16_t*>(base + L.rec_flag);
        R.rec_bin = reinterpret_cast<uint16_t*>(base + L.rec_bin);
        R.rec_status = reinterpret_cast<int8_t*>(base + L.rec_status);
        R.rec_mapq = reinterpret_cast<uint8_t*>(base + L.rec_mapq);
        R.rec_need_flip = reinterpret_cast<uint8_t*>(base + L.rec_flip);
        R.cigar = reinterpret_cast<uint32_t*>(base + L.cigar);
        return R;
    }
};

// ---- totals / flags written by the kernels, read back once per batch ---------------------------------------------
struct DevTotals {
    unsigned long long n_pairs;
    unsigned long long scratch_needed;   // ops
    unsigned long long n_records;
    unsigned long long n_cigar_out;
    unsigned long long n_lifted;
    unsigned long long n_errors;
    long long first_error_read;          // min read index with an error, or INT64_MAX
    int first_error_status;
    unsigned int overflow;               // bit0 pairs, bit1 scratch, bit2 result arena, bit4 malformed batch
    unsigned long long n_in_ops;         // sum of input CIGAR ops over attempted pairs   (roofline arithmetic)
    unsigned long long n_base_bytes;     // base bytes compared (both operands)            (roofline arithmetic)
    unsigned int n_simplify;             // length of DevWork::simplify_list
    unsigned int n_long;                 // length of DevWork::long_list
};

// [off, off + n) inside a pool of `size` elements, without the sum (a caller's offset may be anything, 2^64 - 1 included:
// off + n would wrap and pass; found by tools/fuzz/fuzz_batch_through_device_code.py under AddressSanitizer)
PTL_HD inline bool range_in_pool(uint64_t off, uint64_t n, uint64_t size) { return off <= size && n <= size - off; }

// OVF_INVALID: the batch itself is malformed (an index outside its pool); reported by ptl_lift_wait as PTL_ERR_INVALID_ARG
enum : unsigned { OVF_PAIRS = 1, OVF_SCRATCH = 2, OVF_RESULT = 4, OVF_INVALID = 16 };
static_assert(sizeof(DevTotals) <= kResultHeaderBytes, "the totals are the header of the result arena");

}  // namespace ptl
