// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// C-ABI around the CPU restatement of the reference (cigar.hpp, liftover.hpp, contig_prep.hpp, read_scan.hpp).
// Exports the SAME structs as include/portello_b200.h under the `ptl_oracle_` prefix so one harness drives both the
// CUDA product and this checker, plus function-level entry points (`ptl_oracle_fn_*`) that mirror the reference's
// pure functions one-to-one so the reference's own unit-test vectors (tests/golden/) can be replayed.
//
// Parity status: PINNED by the reference's in-file unit tests, transcribed into tests/golden/*.json.
// The reference itself (Rust) cannot be built in this image (no cargo/rustc), so there is no oracle/_ref.
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg / --impl reference) may load this library.
#include <atomic>
#include <cstring>
#include <memory>
#include <thread>

#include "../include/portello_b200.h"
#include "read_scan.hpp"

using namespace orc;

namespace {

struct Slot {
    bool submitted = false;
    std::vector<uint32_t> read_rec_begin;
    std::vector<int8_t> rec_status;
    std::vector<uint32_t> rec_read_segment, rec_contig_segment;
    std::vector<int32_t> rec_tid;
    std::vector<int64_t> rec_pos;
    std::vector<uint8_t> rec_mapq, rec_need_flip;
    std::vector<uint16_t> rec_flag, rec_bin;
    std::vector<uint64_t> rec_cigar_begin;
    std::vector<uint32_t> cigar;
    ptl_result res{};
    PairStats stats;
    // record assembly (ptl_oracle_assemble_bases): the submitted batch (borrowed: the caller keeps it alive) and outputs
    ptl_batch batch{};
    std::vector<uint64_t> asm_seq_begin, asm_qual_begin;
    std::vector<uint8_t> asm_seq4, asm_qual;
    // ptl_oracle_assemble_records
    std::vector<uint64_t> bam_begin;
    std::vector<uint8_t> bam_bytes;
};

// Flat copy of the installed segments, lent out by ptl_oracle_get_contig_segments.
struct FlatSegments {
    std::vector<uint64_t> contig_len;
    std::vector<uint32_t> contig_seg_begin;
    std::vector<const uint8_t*> rev_ptr;
    std::vector<uint32_t> so_start, so_end;
    std::vector<int32_t> chrom;
    std::vector<int64_t> pos;
    std::vector<uint8_t> is_fwd, mapq;
    std::vector<uint64_t> cigar_begin;
    std::vector<uint32_t> cigar;
};

}  // namespace

struct ptl_ctx {
    std::string err;
    int n_slots = 1;
    int n_threads = 1;
    bool faithful_decode = true;
    std::vector<std::vector<uint8_t>> reference;
    std::vector<uint64_t> contig_len;
    AllContigMappingInfo contigs;
    FlatSegments flat;
    std::vector<Slot> slots;
    std::vector<std::string> contig_names, chrom_names;  // ptl_oracle_set_names
};

namespace {

int fail(ptl_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

CigarVec decode_cigar(const uint32_t* ops, size_t n) {
    CigarVec v(n);
    for (size_t i = 0; i < n; ++i) v[i] = decode(ops[i]);
    return v;
}

void flatten(ptl_ctx* ctx) {
    FlatSegments& f = ctx->flat;
    f = FlatSegments{};
    f.contig_len = ctx->contig_len;
    f.contig_seg_begin.push_back(0);
    f.cigar_begin.push_back(0);
    for (const auto& info : ctx->contigs) {
        f.rev_ptr.push_back(info.has_rev_contig_seq ? info.rev_contig_seq.data() : nullptr);
        for (const auto& si : info.ordered_contig_segment_info) {
            const auto& s = si.seq_order_segment;
            f.so_start.push_back(uint32_t(s.seq_order_read_start));
            f.so_end.push_back(uint32_t(s.seq_order_read_end));
            f.chrom.push_back(int32_t(s.chrom_index));
            f.pos.push_back(s.pos);
            f.is_fwd.push_back(s.is_fwd_strand);
            f.mapq.push_back(s.mapq);
            for (const auto& c : s.cigar) f.cigar.push_back(encode(c));
            f.cigar_begin.push_back(f.cigar.size());
        }
        f.contig_seg_begin.push_back(uint32_t(f.so_start.size()));
    }
}

// Build AllContigMappingInfo from the flat struct (tables via get_read_segment_to_ref_pos_tree_map, as the reference
// does after every trim/join, contig_repeated_match_trimmer.rs:132-133, contig_colinear_segment_joiner.rs:117).
void load_segments(ptl_ctx* ctx, const ptl_contig_segments* s) {
    ctx->contigs.clear();
    ctx->contig_len.assign(s->contig_len, s->contig_len + s->n_contigs);
    for (uint32_t ci = 0; ci < s->n_contigs; ++ci) {
        ContigMappingInfo info;
        info.qname = "contig" + std::to_string(ci);
        if (s->rev_contig_seq && s->rev_contig_seq[ci]) {
            info.has_rev_contig_seq = true;
            info.rev_contig_seq.assign(s->rev_contig_seq[ci], s->rev_contig_seq[ci] + s->contig_len[ci]);
        }
        for (uint32_t k = s->contig_seg_begin[ci]; k < s->contig_seg_begin[ci + 1]; ++k) {
            ContigMappingSegmentInfo si;
            auto& g = si.seq_order_segment;
            g.seq_order_read_start = s->seg_seq_order_start[k];
            g.seq_order_read_end = s->seg_seq_order_end[k];
            g.chrom_index = size_t(s->seg_chrom_index[k]);
            g.pos = s->seg_pos[k];
            g.is_fwd_strand = s->seg_is_fwd[k] != 0;
            g.mapq = s->seg_mapq[k];
            g.cigar = decode_cigar(s->cigar + s->seg_cigar_begin[k], s->seg_cigar_begin[k + 1] - s->seg_cigar_begin[k]);
            si.contig_to_ref_map = get_read_segment_to_ref_pos_tree_map(g.pos, g.cigar, false);
            info.ordered_contig_segment_info.push_back(std::move(si));
        }
        ctx->contigs.push_back(std::move(info));
    }
}

struct ChunkOut {
    std::vector<uint32_t> n_rec_per_read;
    std::vector<LiftedRecord> recs;
    std::vector<uint32_t> rec_rseg;  // batch-global read-segment index
    PairStats stats;
    int64_t first_error_read = -1;
    int first_error_status = 0;
    uint64_t n_errors = 0;
};

void lift_chunk(const ptl_ctx* ctx, const ptl_batch* b, uint32_t r0, uint32_t r1, uint32_t stage_mask, ChunkOut& out) {
    ScanOptions opt;
    opt.faithful_decode = ctx->faithful_decode;
    opt.stage_mask = stage_mask;
    std::vector<SeqOrderSplitReadSegment> splits;
    for (uint32_t r = r0; r < r1; ++r) {
        ReadRecord rec;
        rec.flag = b->read_flag[r];
        rec.mapq = b->read_mapq[r];
        rec.bin = b->read_bin[r];
        rec.seq_len = b->read_seq_len[r];
        rec.seq4 = b->seq4 + b->read_seq_off[r];
        splits.clear();
        const uint32_t s0 = b->read_seg_begin[r], s1 = b->read_seg_begin[r + 1];
        for (uint32_t s = s0; s < s1; ++s) {
            SeqOrderSplitReadSegment g;
            g.chrom_index = b->rseg_contig[s];
            g.pos = b->rseg_pos[s];
            g.is_fwd_strand = b->rseg_is_fwd[s] != 0;
            g.cigar = decode_cigar(b->cigar + b->rseg_cigar_begin[s], b->rseg_cigar_len[s]);
            splits.push_back(std::move(g));
        }
        std::vector<LiftedRecord> recs;
        const PairStats before = out.stats;
        try {
            recs = lift_read(ctx->reference, ctx->contig_len, ctx->contigs, rec, splits, opt, out.stats);
        } catch (const Panic& p) {
            // The reference aborts the run here. Record the error and emit the unmapped fallback so shapes stay defined.
            // Counters follow the ABI's definition: n_pairs = all enumerated pairs, n_lifted = lifted pairs of reads
            // that produced records.
            out.stats.n_lifted = before.n_lifted;
            out.stats.n_out_ops = before.n_out_ops;
            out.stats.n_pairs = before.n_pairs;
            for (const auto& g : splits)
                out.stats.n_pairs += get_contig_split_segments_from_read_mapping(g, ctx->contigs.at(g.chrom_index).ordered_contig_segment_info).size();
            out.n_errors++;
            if (out.first_error_read < 0) {
                out.first_error_read = r;
                out.first_error_status = p.code;
            }
            recs = finish_remapped_alignment_set(rec, {});
        }
        out.n_rec_per_read.push_back(uint32_t(recs.size()));
        for (auto& x : recs) {
            out.rec_rseg.push_back(x.status == 0 ? s0 : s0 + x.read_segment);
            out.recs.push_back(std::move(x));
        }
    }
}

}  // namespace

extern "C" {

int ptl_oracle_create(int /*device*/, int n_slots, ptl_ctx** out) {
    if (!out || n_slots < 1) return PTL_ERR_INVALID_ARG;
    auto* ctx = new ptl_ctx();
    ctx->n_slots = n_slots;
    ctx->slots.resize(size_t(n_slots));
    *out = ctx;
    return PTL_OK;
}
void ptl_oracle_destroy(ptl_ctx* ctx) { delete ctx; }
const char* ptl_oracle_last_error(const ptl_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }
const char* ptl_oracle_version(void) { return "portello_b200 oracle 0.1 (CPU restatement of portello v0.6.1)"; }

// CPU-baseline knobs (no counterpart in the product ABI).
int ptl_oracle_set_threads(ptl_ctx* ctx, int n) {
    if (!ctx || n < 1) return PTL_ERR_INVALID_ARG;
    ctx->n_threads = n;
    return PTL_OK;
}
int ptl_oracle_set_faithful_decode(ptl_ctx* ctx, int on) {
    if (!ctx) return PTL_ERR_INVALID_ARG;
    ctx->faithful_decode = on != 0;
    return PTL_OK;
}

int ptl_oracle_set_reference(ptl_ctx* ctx, uint32_t n_chrom, const uint64_t* chrom_len, const uint8_t* const* chrom_seq) {
    if (!ctx) return PTL_ERR_INVALID_ARG;
    ctx->reference.clear();
    for (uint32_t i = 0; i < n_chrom; ++i) ctx->reference.emplace_back(chrom_seq[i], chrom_seq[i] + chrom_len[i]);
    return PTL_OK;
}

int ptl_oracle_set_contig_segments(ptl_ctx* ctx, const ptl_contig_segments* s) {
    if (!ctx || !s) return PTL_ERR_INVALID_ARG;
    load_segments(ctx, s);
    flatten(ctx);
    return PTL_OK;
}

int ptl_oracle_set_raw_contig_segments(ptl_ctx* ctx, const ptl_contig_segments* s) {
    if (!ctx || !s) return PTL_ERR_INVALID_ARG;
    try {
        load_segments(ctx, s);
        clip_repeated_contig_matches(ctx->contigs);
        join_colinear_contig_segments(ctx->contigs);
    } catch (const Panic& p) {
        return fail(ctx, PTL_ERR_INPUT, p.what());
    }
    flatten(ctx);
    return PTL_OK;
}

int ptl_oracle_set_contig_records(ptl_ctx* ctx, const ptl_contig_records* r) {
    if (!ctx || !r) return PTL_ERR_INVALID_ARG;
    try {
        NameIndex ref_index;
        for (uint32_t i = 0; i < r->n_ref_chrom; ++i) ref_index[r->ref_chrom_names[i]] = i;
        std::vector<std::string> names;
        for (uint32_t i = 0; i < r->n_contigs; ++i) names.push_back(r->contig_names ? r->contig_names[i] : "contig" + std::to_string(i));
        std::vector<ContigRecord> recs(r->n_records);
        for (uint32_t i = 0; i < r->n_records; ++i) {
            auto& c = recs[i];
            c.contig_id = r->contig_id[i];
            c.aln.tid = r->tid[i];
            c.aln.pos = r->pos[i];
            c.aln.flag = r->flag[i];
            c.aln.mapq = r->mapq[i];
            c.aln.cigar = decode_cigar(r->cigar + r->cigar_begin[i], r->cigar_begin[i + 1] - r->cigar_begin[i]);
            if (r->sa_tag && r->sa_tag[i]) {
                c.aln.has_sa = true;
                c.aln.sa = r->sa_tag[i];
            }
            if (r->seq && r->seq[i]) c.seq.assign(r->seq[i], r->seq[i] + r->contig_len[c.contig_id]);
        }
        ctx->contig_len.assign(r->contig_len, r->contig_len + r->n_contigs);
        ctx->contigs = scan_contig_records(recs, r->n_contigs, ref_index, names);
    } catch (const Panic& p) {
        return fail(ctx, PTL_ERR_INPUT, p.what());
    }
    flatten(ctx);
    return PTL_OK;
}

int ptl_oracle_get_contig_segments(const ptl_ctx* ctx, ptl_contig_segments* out) {
    if (!ctx || !out) return PTL_ERR_INVALID_ARG;
    const FlatSegments& f = ctx->flat;
    out->n_contigs = uint32_t(f.contig_len.size());
    out->contig_len = f.contig_len.data();
    out->contig_seg_begin = f.contig_seg_begin.data();
    out->rev_contig_seq = f.rev_ptr.data();
    out->n_segments = uint32_t(f.so_start.size());
    out->seg_seq_order_start = f.so_start.data();
    out->seg_seq_order_end = f.so_end.data();
    out->seg_chrom_index = f.chrom.data();
    out->seg_pos = f.pos.data();
    out->seg_is_fwd = f.is_fwd.data();
    out->seg_mapq = f.mapq.data();
    out->seg_cigar_begin = f.cigar_begin.data();
    out->cigar = f.cigar.data();
    return PTL_OK;
}

int ptl_oracle_get_segment_table(ptl_ctx* ctx, uint32_t segment, uint32_t cap, uint32_t* keys, int32_t* vals, uint32_t* n) {
    if (!ctx || !n) return PTL_ERR_INVALID_ARG;
    uint32_t k = 0;
    for (const auto& info : ctx->contigs) {
        for (const auto& si : info.ordered_contig_segment_info) {
            if (k++ != segment) continue;
            uint32_t i = 0;
            for (const auto& kv : si.contig_to_ref_map.map) {
                if (i < cap) {
                    keys[i] = uint32_t(kv.first);
                    vals[i] = kv.second ? int32_t(*kv.second) : -1;
                }
                ++i;
            }
            *n = i;
            return i <= cap ? PTL_OK : PTL_ERR_INVALID_ARG;
        }
    }
    return fail(ctx, PTL_ERR_INVALID_ARG, "segment index out of range");
}

int ptl_oracle_lift_submit_ex(ptl_ctx* ctx, int slot, const ptl_batch* b, uint32_t stage_mask) {
    if (!ctx || !b || slot < 0 || slot >= ctx->n_slots) return PTL_ERR_INVALID_ARG;
    Slot& sl = ctx->slots[size_t(slot)];
    const uint32_t chunk = 512;
    const uint32_t n_chunks = (b->n_reads + chunk - 1) / chunk;
    std::vector<ChunkOut> outs(n_chunks);
    std::atomic<uint32_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const uint32_t c = next.fetch_add(1);
            if (c >= n_chunks) break;
            lift_chunk(ctx, b, c * chunk, std::min(b->n_reads, (c + 1) * chunk), stage_mask, outs[c]);
        }
    };
    const int nt = std::max(1, std::min<int>(ctx->n_threads, int(n_chunks)));
    if (nt == 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    // gather in submission order
    size_t n_rec = 0, n_ops = 0;
    for (const auto& o : outs) {
        n_rec += o.recs.size();
        for (const auto& r : o.recs) n_ops += r.cigar.size();
    }
    sl.read_rec_begin.assign(1, 0);
    sl.read_rec_begin.reserve(b->n_reads + 1);
    sl.rec_status.resize(n_rec); sl.rec_read_segment.resize(n_rec); sl.rec_contig_segment.resize(n_rec);
    sl.rec_tid.resize(n_rec); sl.rec_pos.resize(n_rec); sl.rec_mapq.resize(n_rec); sl.rec_need_flip.resize(n_rec);
    sl.rec_flag.resize(n_rec); sl.rec_bin.resize(n_rec);
    sl.rec_cigar_begin.assign(1, 0);
    sl.rec_cigar_begin.reserve(n_rec + 1);
    sl.cigar.clear();
    sl.cigar.reserve(n_ops);
    sl.stats = PairStats{};
    ptl_result& res = sl.res;
    res = ptl_result{};
    res.first_error_read = -1;
    size_t k = 0;
    for (const auto& o : outs) {
        for (uint32_t n : o.n_rec_per_read) sl.read_rec_begin.push_back(sl.read_rec_begin.back() + n);
        for (size_t i = 0; i < o.recs.size(); ++i, ++k) {
            const LiftedRecord& r = o.recs[i];
            sl.rec_status[k] = r.status;
            sl.rec_read_segment[k] = o.rec_rseg[i];
            sl.rec_contig_segment[k] = r.status == 0 ? 0xffffffffu : r.contig_segment;
            sl.rec_tid[k] = r.tid;
            sl.rec_pos[k] = r.pos;
            sl.rec_mapq[k] = r.mapq;
            sl.rec_need_flip[k] = r.need_flip;
            sl.rec_flag[k] = r.flag;
            sl.rec_bin[k] = r.bin;
            for (const auto& c : r.cigar) sl.cigar.push_back(encode(c));
            sl.rec_cigar_begin.push_back(sl.cigar.size());
        }
        sl.stats.n_pairs += o.stats.n_pairs;
        sl.stats.n_lifted += o.stats.n_lifted;
        sl.stats.n_in_ops += o.stats.n_in_ops;
        sl.stats.n_out_ops += o.stats.n_out_ops;
        sl.stats.cmp.base_bytes += o.stats.cmp.base_bytes;
        res.n_errors += o.n_errors;
        if (res.first_error_read < 0 && o.first_error_read >= 0) {
            res.first_error_read = o.first_error_read;
            res.first_error_status = o.first_error_status;
        }
    }
    res.n_reads = b->n_reads;
    res.read_rec_begin = sl.read_rec_begin.data();
    res.n_records = uint32_t(n_rec);
    res.rec_status = sl.rec_status.data();
    res.rec_read_segment = sl.rec_read_segment.data();
    res.rec_contig_segment = sl.rec_contig_segment.data();
    res.rec_tid = sl.rec_tid.data();
    res.rec_pos = sl.rec_pos.data();
    res.rec_mapq = sl.rec_mapq.data();
    res.rec_flag = sl.rec_flag.data();
    res.rec_bin = sl.rec_bin.data();
    res.rec_need_flip = sl.rec_need_flip.data();
    res.rec_cigar_begin = sl.rec_cigar_begin.data();
    res.cigar = sl.cigar.data();
    res.n_cigar = sl.cigar.size();
    res.n_pairs = sl.stats.n_pairs;
    res.n_lifted = sl.stats.n_lifted;
    sl.submitted = true;
    sl.batch = *b;  // (pointers borrowed: ptl_oracle_assemble_bases reads the batch's bases again)
    return PTL_OK;
}
int ptl_oracle_lift_submit(ptl_ctx* ctx, int slot, const ptl_batch* b) {
    return ptl_oracle_lift_submit_ex(ctx, slot, b, PTL_STAGE_ALL);
}

// Record assembly, bases: what reverse_alignment_seq_and_qual (src/read_alignment_scanner.rs:125-133) does to a flipped
// record, byte for byte: seq().as_bytes() (4-bit decode) -> rev_comp_in_place (seq_util.rs:29-40) -> Record::set, which
// re-encodes the ASCII bases with htslib's seq_nt16_table (only A, C, G, T, N can come out of rev_comp); qualities reversed.
// Records that were not flipped keep the read's bytes (clone_record, :105-117).
int ptl_oracle_assemble_bases(ptl_ctx* ctx, int slot, const ptl_read_quals* quals, uint32_t /*flags*/, ptl_record_bases* out) {
    if (!ctx || !out || !quals || slot < 0 || slot >= ctx->n_slots) return PTL_ERR_INVALID_ARG;
    Slot& sl = ctx->slots[size_t(slot)];
    if (!sl.submitted) return fail(ctx, PTL_ERR_STATE, "assemble without a lifted batch");
    const ptl_batch& b = sl.batch;
    const uint32_t n_rec = sl.res.n_records;
    // read of every record (every record of a read is a clone of the same input record, :105-117; a read without any
    // split segment still yields its unmapped fallback, so go through the record CSR, not through the segment index)
    std::vector<uint32_t> rec_read(n_rec);
    for (uint32_t r = 0; r < b.n_reads; ++r)
        for (uint32_t k = sl.res.read_rec_begin[r]; k < sl.res.read_rec_begin[r + 1]; ++k) rec_read[k] = r;
    auto encode = [](uint8_t c) -> uint8_t {  // htslib seq_nt16_table
        switch (c) {
            case '=': return 0; case 'A': case 'a': return 1; case 'C': case 'c': return 2; case 'M': case 'm': return 3;
            case 'G': case 'g': return 4; case 'R': case 'r': return 5; case 'S': case 's': return 6; case 'V': case 'v': return 7;
            case 'T': case 't': return 8; case 'W': case 'w': return 9; case 'Y': case 'y': return 10; case 'H': case 'h': return 11;
            case 'K': case 'k': return 12; case 'D': case 'd': return 13; case 'B': case 'b': return 14; default: return 15;
        }
    };
    sl.asm_seq_begin.assign(size_t(n_rec) + 1, 0);
    sl.asm_qual_begin.assign(size_t(n_rec) + 1, 0);
    for (uint32_t k = 0; k < n_rec; ++k) {
        const uint64_t len = b.read_seq_len[rec_read[k]];
        sl.asm_seq_begin[k + 1] = sl.asm_seq_begin[k] + ((((len + 1) >> 1) + 15) & ~15ull);
        sl.asm_qual_begin[k + 1] = sl.asm_qual_begin[k] + ((len + 15) & ~15ull);
    }
    sl.asm_seq4.assign(sl.asm_seq_begin[n_rec], 0);
    sl.asm_qual.assign(sl.asm_qual_begin[n_rec], 0);
    for (uint32_t k = 0; k < n_rec; ++k) {
        const uint32_t r = rec_read[k];
        const uint64_t len = b.read_seq_len[r];
        const uint8_t* seq4 = b.seq4 + b.read_seq_off[r];
        const uint8_t* q = quals->qual + quals->read_qual_off[r];
        uint8_t* o_seq = sl.asm_seq4.data() + sl.asm_seq_begin[k];
        uint8_t* o_q = sl.asm_qual.data() + sl.asm_qual_begin[k];
        if (!sl.rec_need_flip[k]) {
            std::memcpy(o_seq, seq4, (len + 1) >> 1);
            std::memcpy(o_q, q, len);
            continue;
        }
        std::vector<uint8_t> ascii = decode_seq4(seq4, len);  // record.seq().as_bytes()
        rev_comp_in_place(ascii);
        for (uint64_t i = 0; i < len; ++i) o_seq[i >> 1] |= uint8_t(encode(ascii[i]) << ((i & 1) ? 0 : 4));  // Record::set
        for (uint64_t i = 0; i < len; ++i) o_q[i] = q[len - 1 - i];                                          // qual().iter().rev()
    }
    *out = ptl_record_bases{};
    out->n_records = n_rec;
    out->rec_seq_begin = sl.asm_seq_begin.data();
    out->seq4 = sl.asm_seq4.data();
    out->rec_qual_begin = sl.asm_qual_begin.data();
    out->qual = sl.asm_qual.data();
    out->bytes_read = out->bytes_written = sl.asm_seq_begin[n_rec] + sl.asm_qual_begin[n_rec];
    return PTL_OK;
}
int ptl_oracle_set_names(ptl_ctx* ctx, uint32_t n_contigs, const char* const* contig_names, uint32_t n_chrom, const char* const* chrom_names) {
    if (!ctx || (n_contigs && !contig_names) || (n_chrom && !chrom_names)) return PTL_ERR_INVALID_ARG;
    ctx->contig_names.assign(contig_names, contig_names + n_contigs);
    ctx->chrom_names.assign(chrom_names, chrom_names + n_chrom);
    return PTL_OK;
}

namespace {
// [begin, end) of the FIRST aux field named `tag` (htslib bam_aux_get + bam_aux_del as used by remove_aux), or false.
// Field sizes per SAM spec 4.2.4; a malformed tail ends the search (nothing behind it can be found).
bool find_aux(const uint8_t* aux, size_t n, const char* tag, size_t* begin, size_t* end) {
    size_t i = 0;
    while (i + 3 <= n) {
        const size_t start = i;
        const uint8_t type = aux[i + 2];
        i += 3;
        size_t sz = 0;
        switch (type) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'd': sz = 8; break;
            case 'Z': case 'H': {
                size_t j = i;
                while (j < n && aux[j] != 0) ++j;
                if (j >= n) return false;
                sz = j - i + 1;
                break;
            }
            case 'B': {
                if (i + 5 > n) return false;
                const uint8_t sub = aux[i];
                uint32_t cnt;
                std::memcpy(&cnt, aux + i + 1, 4);
                size_t es = 0;
                switch (sub) {
                    case 'c': case 'C': es = 1; break;
                    case 's': case 'S': es = 2; break;
                    case 'i': case 'I': case 'f': es = 4; break;
                    default: return false;
                }
                sz = 5 + size_t(cnt) * es;
                break;
            }
            default: return false;
        }
        if (i + sz > n) return false;
        i += sz;
        if (aux[start] == uint8_t(tag[0]) && aux[start + 1] == uint8_t(tag[1])) { *begin = start; *end = i; return true; }
    }
    return false;
}
void put_u32(std::vector<uint8_t>& v, uint32_t x) { for (int b = 0; b < 4; ++b) v.push_back(uint8_t(x >> (8 * b))); }
void put_u16(std::vector<uint8_t>& v, uint16_t x) { v.push_back(uint8_t(x)); v.push_back(uint8_t(x >> 8)); }
}  // namespace

// Record assembly, whole records: clone_record (src/read_alignment_scanner.rs:105-117) + the field updates and tag
// pushes of :245-282 + finish_remapped_alignment_set (:310-366), serialised the way bam_write1 lays a record out
// (SAM spec 4.2).  Tag payloads follow rust-htslib push_aux: Aux::String -> 'Z' + bytes + NUL, Aux::U8 -> 'C' + byte.
int ptl_oracle_assemble_records(ptl_ctx* ctx, int slot, const ptl_read_extras* x, uint32_t /*flags*/, ptl_bam_records* out) {
    if (!ctx || !out || !x || slot < 0 || slot >= ctx->n_slots) return PTL_ERR_INVALID_ARG;
    Slot& sl = ctx->slots[size_t(slot)];
    if (!sl.submitted) return fail(ctx, PTL_ERR_STATE, "assemble without a lifted batch");
    const ptl_batch& b = sl.batch;
    const uint32_t n_rec = sl.res.n_records;
    static const char kOps[] = "MIDNSHP=X";
    auto encode = [](uint8_t c) -> uint8_t {  // htslib seq_nt16_table restricted to what rev_comp can produce
        switch (c) { case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8; default: return 15; }
    };
    sl.bam_begin.assign(1, 0);
    sl.bam_bytes.clear();
    for (uint32_t r = 0; r < b.n_reads; ++r) {
        const uint32_t k0 = sl.read_rec_begin[r], k1 = sl.read_rec_begin[r + 1];
        if (k0 == k1) continue;
        const uint8_t* aux = x->aux + x->aux_off[r];
        const size_t aux_n = size_t(x->aux_off[r + 1] - x->aux_off[r]);
        // clone_record: remove NM, then SA, PS, ZM (each: first occurrence, on the already shortened block)
        std::vector<uint8_t> kept(aux, aux + aux_n);
        for (const char* tag : {"NM", "SA", "PS", "ZM"}) {
            size_t a = 0, e = 0;
            if (find_aux(kept.data(), kept.size(), tag, &a, &e)) kept.erase(kept.begin() + long(a), kept.begin() + long(e));
        }
        // get_sa_tag_segment of every record of the read (:292-301)
        std::vector<std::string> sa(k1 - k0);
        for (uint32_t k = k0; k < k1; ++k) {
            if (sl.rec_status[k] != 1) continue;
            const int32_t tid = sl.rec_tid[k];
            if (tid < 0 || size_t(tid) >= ctx->chrom_names.size()) return fail(ctx, PTL_ERR_STATE, "ptl_set_names: no name for a reference chromosome");
            std::string t = ctx->chrom_names[size_t(tid)] + "," + std::to_string(sl.rec_pos[k] + 1) + "," + ((sl.rec_flag[k] & 0x10) ? "-" : "+") + ",";
            for (uint64_t i = sl.rec_cigar_begin[k]; i < sl.rec_cigar_begin[k + 1]; ++i) {
                t += std::to_string(sl.cigar[i] >> 4);
                t += kOps[std::min<uint32_t>(sl.cigar[i] & 0xfu, 8)];
            }
            t += "," + std::to_string(unsigned(sl.rec_mapq[k])) + ",0;";
            sa[k - k0] = t;
        }
        const uint64_t len = b.read_seq_len[r];
        const uint8_t* seq4 = b.seq4 + b.read_seq_off[r];
        const uint8_t* q = x->quals.qual + x->quals.read_qual_off[r];
        const uint8_t* name = x->names + x->name_off[r];
        const size_t name_n = size_t(x->name_off[r + 1] - x->name_off[r]);
        if (name_n > 254) return fail(ctx, PTL_ERR_STATE, "a read name longer than 254 bytes (BAM l_read_name is a u8)");
        for (uint32_t k = k0; k < k1; ++k) {
            const bool lifted = sl.rec_status[k] == 1;
            std::vector<uint8_t> a = kept;
            const uint64_t c0 = sl.rec_cigar_begin[k], c1 = sl.rec_cigar_begin[k + 1];
            if (lifted) {
                const uint32_t s = sl.rec_read_segment[k];
                const uint32_t ctg = b.rseg_contig[s];
                if (ctg >= ctx->contig_names.size()) return fail(ctx, PTL_ERR_STATE, "ptl_set_names: no name for a contig");
                const uint32_t g = ctx->flat.contig_seg_begin[ctg] + sl.rec_contig_segment[k];
                const std::string ps = ctx->contig_names[ctg] + "_split" + std::to_string(sl.rec_contig_segment[k]) + (ctx->flat.is_fwd[g] ? "+" : "-");
                a.insert(a.end(), {'P', 'S', 'Z'});
                a.insert(a.end(), ps.begin(), ps.end());
                a.push_back(0);
                a.insert(a.end(), {'Z', 'M', 'C', b.read_mapq[r]});
                std::string text;
                for (uint32_t j = k0; j < k1; ++j)
                    if (j != k) text += sa[j - k0];
                if (!text.empty()) {
                    a.insert(a.end(), {'S', 'A', 'Z'});
                    a.insert(a.end(), text.begin(), text.end());
                    a.push_back(0);
                }
            }
            // bam_write1 (htslib sam.c; hts-sys 2.2.0 is the reference's pin, Cargo.lock:740): a record with more than 65535
            // CIGAR ops is written with the placeholder CIGAR <l_seq>S<ref_len>N, n_cigar_op = 2, and the real CIGAR in a
            // CG:B,I tag behind every other tag (SAM spec 4.2.2); it refuses a reference length the N op cannot hold.
            const bool long_cigar = lifted && c1 - c0 > 65535;
            uint64_t cig_ref_len = 0;
            if (long_cigar) {
                for (uint64_t i = c0; i < c1; ++i)
                    if ((0x18du >> (sl.cigar[i] & 0xfu)) & 1u) cig_ref_len += sl.cigar[i] >> 4;  // M D N = X (bam_cigar2rlen)
                if (cig_ref_len >= (1ull << 28)) return fail(ctx, PTL_ERR_STATE, "a lifted CIGAR with more than 65535 ops spans 2^28 reference bases or more: bam_write1 cannot write it");
                a.insert(a.end(), {'C', 'G', 'B', 'I'});
                put_u32(a, uint32_t(c1 - c0));
                for (uint64_t i = c0; i < c1; ++i) put_u32(a, sl.cigar[i]);
            }
            std::vector<uint8_t> rec;
            put_u32(rec, uint32_t(sl.rec_tid[k]));
            put_u32(rec, uint32_t(int32_t(sl.rec_pos[k])));
            rec.push_back(uint8_t(name_n + 1));
            rec.push_back(sl.rec_mapq[k]);
            put_u16(rec, sl.rec_bin[k]);
            put_u16(rec, uint16_t(long_cigar ? 2 : lifted ? c1 - c0 : 0));
            put_u16(rec, sl.rec_flag[k]);
            put_u32(rec, uint32_t(len));
            put_u32(rec, uint32_t(x->mate_tid[r]));
            put_u32(rec, uint32_t(x->mate_pos[r]));
            put_u32(rec, uint32_t(x->tlen[r]));
            rec.insert(rec.end(), name, name + name_n);
            rec.push_back(0);
            if (long_cigar) {
                put_u32(rec, (uint32_t(len) << 4) | 4u);
                put_u32(rec, (uint32_t(cig_ref_len) << 4) | 3u);
            } else if (lifted) {
                for (uint64_t i = c0; i < c1; ++i) put_u32(rec, sl.cigar[i]);
            }
            if (!sl.rec_need_flip[k]) {
                rec.insert(rec.end(), seq4, seq4 + ((len + 1) >> 1));
                rec.insert(rec.end(), q, q + len);
            } else {  // reverse_alignment_seq_and_qual (:125-133)
                std::vector<uint8_t> ascii = decode_seq4(seq4, len);
                rev_comp_in_place(ascii);
                std::vector<uint8_t> packed((len + 1) >> 1, 0);
                for (uint64_t i = 0; i < len; ++i) packed[i >> 1] |= uint8_t(encode(ascii[i]) << ((i & 1) ? 0 : 4));
                rec.insert(rec.end(), packed.begin(), packed.end());
                for (uint64_t i = 0; i < len; ++i) rec.push_back(q[len - 1 - i]);
            }
            rec.insert(rec.end(), a.begin(), a.end());
            put_u32(sl.bam_bytes, uint32_t(rec.size()));
            sl.bam_bytes.insert(sl.bam_bytes.end(), rec.begin(), rec.end());
            sl.bam_begin.push_back(sl.bam_bytes.size());
        }
    }
    if (sl.bam_begin.size() != size_t(n_rec) + 1) return fail(ctx, PTL_ERR_STATE, "record count mismatch (internal)");
    *out = ptl_bam_records{};
    out->n_records = n_rec;
    out->rec_begin = sl.bam_begin.data();
    sl.bam_bytes.resize(sl.bam_bytes.size() + 16, 0);
    out->bytes = sl.bam_bytes.data();
    return PTL_OK;
}
int ptl_oracle_lift_wait(ptl_ctx* ctx, int slot, ptl_result* out) {
    if (!ctx || !out || slot < 0 || slot >= ctx->n_slots) return PTL_ERR_INVALID_ARG;
    Slot& sl = ctx->slots[size_t(slot)];
    if (!sl.submitted) return fail(ctx, PTL_ERR_STATE, "wait without submit");
    *out = sl.res;
    if (sl.res.n_errors) return fail(ctx, PTL_ERR_LIFT_PANIC, "reference would panic on read " + std::to_string(sl.res.first_error_read));
    return PTL_OK;
}
// out[0..5) = n_pairs, n_lifted, n_in_ops, n_out_ops, compared base bytes  (roofline arithmetic)
int ptl_oracle_last_stats(ptl_ctx* ctx, int slot, uint64_t* out) {
    if (!ctx || !out || slot < 0 || slot >= ctx->n_slots) return PTL_ERR_INVALID_ARG;
    const PairStats& s = ctx->slots[size_t(slot)].stats;
    out[0] = s.n_pairs; out[1] = s.n_lifted; out[2] = s.n_in_ops; out[3] = s.n_out_ops; out[4] = s.cmp.base_bytes;
    return PTL_OK;
}

// ---------------------------------------------------------------- function-level mirrors (golden vectors)
// All return the number of output ops (or -1 for None / -100-code on a reference panic); out arrays need `cap` entries.

static int64_t emit(const CigarVec& c, uint32_t* out, uint32_t cap) {
    for (size_t i = 0; i < c.size() && i < cap; ++i) out[i] = encode(c[i]);
    return int64_t(c.size());
}

int64_t ptl_oracle_fn_liftover(int64_t c2r_pos, const uint32_t* c2r_cigar, uint32_t n_c2r, int has_map, int64_t pos,
                               const uint32_t* cigar, uint32_t n, int64_t* out_pos, uint32_t* out, uint32_t cap) {
    ReadToRefTreeMap map;
    if (has_map) map = get_read_segment_to_ref_pos_tree_map(c2r_pos, decode_cigar(c2r_cigar, n_c2r), false);
    auto r = liftover_read_alignment(map, pos, decode_cigar(cigar, n));
    if (!r) return -1;
    *out_pos = r->pos;
    return emit(r->cigar, out, cap);
}
int64_t ptl_oracle_fn_simplify(int64_t pos, const uint32_t* cigar, uint32_t n, const uint8_t* ref, uint64_t ref_len,
                               const uint8_t* read, uint64_t read_len, int64_t* out_pos, uint32_t* out, uint32_t cap) {
    try {
        auto r = simplify_alignment_indels(pos, decode_cigar(cigar, n), ref, ref_len, ReadSeq::from_ascii(read, read_len));
        *out_pos = r.pos;
        return emit(r.cigar, out, cap);
    } catch (const Panic& p) { return -100 + p.code; }
}
int64_t ptl_oracle_fn_shift(int right, int64_t pos, const uint32_t* cigar, uint32_t n, const uint8_t* ref, uint64_t ref_len,
                            const uint8_t* read, uint64_t read_len, int64_t* out_pos, uint32_t* out, uint32_t cap) {
    try {
        auto rs = ReadSeq::from_ascii(read, read_len);
        auto r = right ? right_shift_indels(pos, decode_cigar(cigar, n), ref, ref_len, rs)
                       : left_shift_indels(pos, decode_cigar(cigar, n), ref, ref_len, rs);
        *out_pos = r.pos;
        return emit(r.cigar, out, cap);
    } catch (const Panic& p) { return -100 + p.code; }
}
int64_t ptl_oracle_fn_compress(const uint32_t* cigar, uint32_t n, uint32_t* out, uint32_t cap) {
    return emit(compress_cigar(decode_cigar(cigar, n)), out, cap);
}
// in-place; returns the shift
int64_t ptl_oracle_fn_cleanup(uint32_t* cigar, uint32_t n) {
    CigarVec c = decode_cigar(cigar, n);
    const size_t shift = clean_up_cigar_edge_indels(c);
    for (uint32_t i = 0; i < n; ++i) cigar[i] = encode(c[i]);
    return int64_t(shift);
}
int64_t ptl_oracle_fn_clip_read_edges(const uint32_t* cigar, uint32_t n, uint64_t left, uint64_t right, int64_t* ref_shift,
                                      uint32_t* out, uint32_t cap) {
    auto r = clip_alignment_read_edges(decode_cigar(cigar, n), left, right);
    *ref_shift = r.second;
    return emit(r.first, out, cap);
}
void ptl_oracle_fn_read_clip_positions(const uint32_t* cigar, uint32_t n, int ignore_hard_clip, uint64_t* out3) {
    auto c = get_read_clip_positions(decode_cigar(cigar, n), ignore_hard_clip != 0);
    out3[0] = c.left; out3[1] = c.right; out3[2] = c.size;
}
void ptl_oracle_fn_offsets(const uint32_t* cigar, uint32_t n, int ignore_hard_clip, int64_t* ref_pos, uint64_t* read_pos) {
    // running (ref_pos, read_pos) after each op, update_ref_and_read_pos
    int64_t rp = ref_pos[0];
    size_t qp = size_t(read_pos[0]);
    for (uint32_t i = 0; i < n; ++i) {
        update_ref_and_read_pos(decode(cigar[i]), rp, qp, ignore_hard_clip != 0);
        ref_pos[i] = rp;
        read_pos[i] = qp;
    }
}
void ptl_oracle_fn_homology(const uint8_t* ref, uint64_t ref_len, int64_t ref_s, int64_t ref_e, const uint8_t* read,
                            uint64_t read_len, int64_t read_s, int64_t read_e, int64_t* out_range2, uint8_t* hom,
                            uint32_t cap, uint32_t* hom_len) {
    auto r = get_indel_breakend_homology_info(ref, ref_len, IntRange{ref_s, ref_e}, ReadSeq::from_ascii(read, read_len),
                                              IntRange{read_s, read_e});
    out_range2[0] = r.first.start;
    out_range2[1] = r.first.end;
    *hom_len = uint32_t(r.second.size());
    for (size_t i = 0; i < r.second.size() && i < cap; ++i) hom[i] = r.second[i];
}
// keys/vals of the tree map; vals -1 = None.  Returns count.
int64_t ptl_oracle_fn_tree_map(int64_t ref_pos, const uint32_t* cigar, uint32_t n, int ignore_hard_clip, uint64_t* keys,
                               int64_t* vals, uint32_t cap) {
    auto m = get_read_segment_to_ref_pos_tree_map(ref_pos, decode_cigar(cigar, n), ignore_hard_clip != 0);
    uint32_t i = 0;
    for (const auto& kv : m.map) {
        if (i < cap) {
            keys[i] = kv.first;
            vals[i] = kv.second ? *kv.second : -1;
        }
        ++i;
    }
    return i;
}
// get_ref_pos for read_pos in [0, n_pos); INT64_MIN = None
void ptl_oracle_fn_tree_map_ref_pos(int64_t ref_pos, const uint32_t* cigar, uint32_t n, int ignore_hard_clip,
                                    uint32_t n_pos, int64_t* out) {
    auto m = get_read_segment_to_ref_pos_tree_map(ref_pos, decode_cigar(cigar, n), ignore_hard_clip != 0);
    for (uint32_t i = 0; i < n_pos; ++i) {
        auto v = m.get_ref_pos(i);
        out[i] = v ? *v : INT64_MIN;
    }
}
int64_t ptl_oracle_fn_tree_map_ref_range(int64_t ref_pos, const uint32_t* cigar, uint32_t n, int ignore_hard_clip,
                                         uint64_t start, uint64_t end, uint64_t* keys, int64_t* vals, uint32_t cap) {
    auto m = get_read_segment_to_ref_pos_tree_map(ref_pos, decode_cigar(cigar, n), ignore_hard_clip != 0);
    auto r = m.get_ref_range(start, end);
    uint32_t i = 0;
    for (auto it = r.first; it != r.second; ++it, ++i) {
        if (i < cap) {
            keys[i] = it->first;
            vals[i] = it->second ? *it->second : -1;
        }
    }
    return i;
}
// clip_seg_isec_range on one segment; io arrays: so[2] (seq-order start,end), pos; returns new cigar len or -1 if eliminated
int64_t ptl_oracle_fn_clip_seg_isec_range(uint64_t* so, int64_t* pos, int is_fwd, const uint32_t* cigar, uint32_t n,
                                          int64_t isec_start, int64_t isec_end, uint32_t* out, uint32_t cap) {
    SeqOrderSplitReadSegment s;
    s.seq_order_read_start = so[0];
    s.seq_order_read_end = so[1];
    s.pos = *pos;
    s.is_fwd_strand = is_fwd != 0;
    s.cigar = decode_cigar(cigar, n);
    const bool eliminated = clip_seg_isec_range(s, IntRange{isec_start, isec_end});
    so[0] = s.seq_order_read_start;
    so[1] = s.seq_order_read_end;
    *pos = s.pos;
    const int64_t len = emit(s.cigar, out, cap);
    return eliminated ? -1 : len;
}
double ptl_oracle_fn_gci(const uint32_t* cigar, uint32_t n) {
    try {
        return get_gap_compressed_identity_no_align_match(decode_cigar(cigar, n));
    } catch (const Panic&) { return -1.0; }
}
uint16_t ptl_oracle_fn_reg2bin(int64_t begin, int64_t end) { return bam_reg2bin(size_t(begin), size_t(end)); }
uint32_t ptl_oracle_fn_region_segments(uint64_t size, uint64_t segment_size, uint64_t* begin, uint64_t* end, uint32_t cap) {
    auto v = get_region_segments(size, segment_size);
    for (size_t i = 0; i < v.size() && i < cap; ++i) {
        begin[i] = v[i].first;
        end[i] = v[i].second;
    }
    return uint32_t(v.size());
}
void ptl_oracle_fn_rev_comp(uint8_t* seq, uint64_t n) {
    std::vector<uint8_t> v(seq, seq + n);
    rev_comp_in_place(v);
    std::memcpy(seq, v.data(), n);
}
int64_t ptl_oracle_fn_strip_clip(int trailing, const uint32_t* cigar, uint32_t n, uint32_t* out, uint32_t cap) {
    CigarVec c = decode_cigar(cigar, n);
    if (trailing) strip_trailing_clip(c); else strip_leading_clip(c);
    return emit(c, out, cap);
}

// parse_sa_aux_val: returns segment count (or -1 on a reference panic); rnames are written '\n'-joined.
int64_t ptl_oracle_fn_parse_sa(const char* sa, uint32_t cap, int64_t* pos, uint8_t* is_fwd, uint8_t* mapq, int32_t* nm,
                               uint32_t* n_cigar, char* rnames, uint32_t rnames_cap) {
    try {
        auto v = parse_sa_aux_val(sa);
        std::string names;
        for (size_t i = 0; i < v.size(); ++i) {
            if (i < cap) {
                pos[i] = v[i].pos; is_fwd[i] = v[i].is_fwd_strand; mapq[i] = v[i].mapq; nm[i] = v[i].nm;
                n_cigar[i] = uint32_t(v[i].cigar.size());
            }
            names += v[i].rname;
            names.push_back('\n');
        }
        if (names.size() + 1 <= rnames_cap) std::memcpy(rnames, names.c_str(), names.size() + 1);
        return int64_t(v.size());
    } catch (const Panic&) { return -1; }
}

// get_seq_order_read_split_segments for one record.  Same outputs as ptl_pack_split_segments.
static thread_local std::string g_split_err;
const char* ptl_oracle_split_last_error(void) { return g_split_err.c_str(); }
int ptl_oracle_split_segments(uint32_t n_names, const char* const* names, int32_t tid, int64_t pos, uint16_t flag,
                              uint8_t mapq, const uint32_t* cigar, uint32_t n_cigar, const char* sa_tag,
                              uint32_t cap_segments, uint32_t cap_cigar, ptl_split_segments* out, uint32_t* n_segments,
                              uint32_t* n_cigar_out) {
    try {
        NameIndex idx;
        for (uint32_t i = 0; i < n_names; ++i) idx[names[i]] = i;
        AlnRecord rec;
        rec.tid = tid; rec.pos = pos; rec.flag = flag; rec.mapq = mapq;
        rec.cigar = decode_cigar(cigar, n_cigar);
        if (sa_tag) { rec.has_sa = true; rec.sa = sa_tag; }
        auto segs = get_seq_order_read_split_segments(idx, rec);
        size_t total = 0;
        for (const auto& s : segs) total += s.cigar.size();
        *n_segments = uint32_t(segs.size());
        *n_cigar_out = uint32_t(total);
        if (segs.size() > cap_segments || total > cap_cigar) return PTL_ERR_INVALID_ARG;
        uint32_t w = 0;
        for (size_t i = 0; i < segs.size(); ++i) {
            const auto& s = segs[i];
            out->seq_order_start[i] = uint32_t(s.seq_order_read_start);
            out->seq_order_end[i] = uint32_t(s.seq_order_read_end);
            out->contig[i] = uint32_t(s.chrom_index);
            out->pos[i] = s.pos;
            out->is_fwd[i] = s.is_fwd_strand;
            out->mapq[i] = s.mapq;
            out->from_primary[i] = s.from_primary_bam_record;
            out->cigar_begin[i] = w;
            for (const auto& c : s.cigar) out->cigar[w++] = encode(c);
        }
        out->cigar_begin[segs.size()] = w;
        return PTL_OK;
    } catch (const Panic& p) {
        g_split_err = p.what();
        return PTL_ERR_INPUT;
    }
}

// The head of the reference's record loop (src/read_alignment_scanner.rs:393-421) for records [first, first + count):
// skip supplementary records, get_seq_order_read_split_segments for the others -> one ptl_batch, so that the CPU arm of
// the bench (bench.py --impl reference) needs nothing of the product library.  Same output layout as ptl_pack_batch.
struct ptl_oracle_packed {
    std::vector<uint16_t> read_flag, read_bin;
    std::vector<uint8_t> read_mapq, rseg_is_fwd;
    std::vector<uint32_t> read_seq_len, read_seg_begin, rseg_contig, rseg_cigar_len, cigar, record_index;
    std::vector<uint64_t> read_seq_off, rseg_cigar_begin;
    std::vector<int64_t> rseg_pos;
    ptl_batch view{};
};
static thread_local std::string g_pack_err;
const char* ptl_oracle_pack_last_error(void) { return g_pack_err.c_str(); }
int ptl_oracle_pack_batch(const ptl_read_records* recs, uint32_t first, uint32_t count, uint32_t n_names, const char* const* names, int n_threads,
                          ptl_oracle_packed** out) {
    if (!recs || !out || first > recs->n_reads || count > recs->n_reads - first) return PTL_ERR_INVALID_ARG;
    NameIndex idx;
    for (uint32_t i = 0; i < n_names; ++i) idx[names[i]] = i;
    struct Part { std::vector<std::vector<SeqOrderSplitReadSegment>> segs; std::vector<uint32_t> rec; std::string err; };
    const uint32_t chunk = 4096, n_chunks = (count + chunk - 1) / chunk;
    std::vector<Part> parts(n_chunks);
    std::atomic<uint32_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const uint32_t c = next.fetch_add(1);
            if (c >= n_chunks) break;
            Part& P = parts[c];
            try {
                for (uint32_t r = first + c * chunk; r < std::min(first + count, first + (c + 1) * chunk); ++r) {
                    if (recs->flag[r] & 0x800) continue;  // supplementary records are skipped (:404)
                    AlnRecord rec;
                    rec.tid = recs->tid[r]; rec.pos = recs->pos[r]; rec.flag = recs->flag[r]; rec.mapq = recs->mapq[r];
                    rec.cigar = decode_cigar(recs->cigar + recs->cigar_begin[r], uint32_t(recs->cigar_begin[r + 1] - recs->cigar_begin[r]));
                    if (recs->sa_tag && recs->sa_tag[r]) { rec.has_sa = true; rec.sa = recs->sa_tag[r]; }
                    P.segs.push_back(get_seq_order_read_split_segments(idx, rec));
                    P.rec.push_back(r);
                }
            } catch (const Panic& p) { P.err = p.what(); }
        }
    };
    const int nt = std::max(1, std::min<int>(n_threads, int(n_chunks)));
    if (nt == 1) worker();
    else { std::vector<std::thread> th; for (int t = 0; t < nt; ++t) th.emplace_back(worker); for (auto& t : th) t.join(); }
    for (const auto& P : parts) if (!P.err.empty()) { g_pack_err = P.err; return PTL_ERR_INPUT; }
    auto* pk = new ptl_oracle_packed();
    pk->read_seg_begin.push_back(0);
    for (const auto& P : parts)
        for (size_t i = 0; i < P.rec.size(); ++i) {
            const uint32_t r = P.rec[i];
            pk->record_index.push_back(r);
            pk->read_flag.push_back(recs->flag[r]); pk->read_mapq.push_back(recs->mapq[r]); pk->read_bin.push_back(recs->bin[r]);
            pk->read_seq_len.push_back(recs->seq_len[r]); pk->read_seq_off.push_back(recs->seq_off[r]);
            for (const auto& sg : P.segs[i]) {
                pk->rseg_contig.push_back(uint32_t(sg.chrom_index)); pk->rseg_pos.push_back(sg.pos); pk->rseg_is_fwd.push_back(sg.is_fwd_strand);
                pk->rseg_cigar_begin.push_back(pk->cigar.size()); pk->rseg_cigar_len.push_back(uint32_t(sg.cigar.size()));
                for (const auto& c : sg.cigar) pk->cigar.push_back(encode(c));
            }
            pk->read_seg_begin.push_back(uint32_t(pk->rseg_contig.size()));
        }
    ptl_batch& b = pk->view;
    b.n_reads = uint32_t(pk->read_flag.size());
    b.read_flag = pk->read_flag.data(); b.read_mapq = pk->read_mapq.data(); b.read_bin = pk->read_bin.data();
    b.read_seq_len = pk->read_seq_len.data(); b.read_seq_off = pk->read_seq_off.data(); b.read_seg_begin = pk->read_seg_begin.data();
    b.n_read_segments = uint32_t(pk->rseg_contig.size());
    b.rseg_contig = pk->rseg_contig.data(); b.rseg_pos = pk->rseg_pos.data(); b.rseg_is_fwd = pk->rseg_is_fwd.data();
    b.rseg_cigar_begin = pk->rseg_cigar_begin.data(); b.rseg_cigar_len = pk->rseg_cigar_len.data();
    b.cigar = pk->cigar.data(); b.n_cigar = pk->cigar.size();
    b.seq4 = recs->seq4; b.seq4_bytes = recs->seq4_bytes;
    *out = pk;
    return PTL_OK;
}
void ptl_oracle_packed_view(const ptl_oracle_packed* p, ptl_batch* out) { *out = p->view; }
const uint32_t* ptl_oracle_packed_record_index(const ptl_oracle_packed* p) { return p->record_index.data(); }
void ptl_oracle_packed_free(ptl_oracle_packed* p) { delete p; }

// SA:Z text, same contract as ptl_format_sa_tags.
int ptl_oracle_format_sa_tags(const ptl_result* res, uint32_t n_chrom, const char* const* chrom_names, char* buf,
                              uint64_t cap, uint64_t* sa_begin, uint64_t* need) {
    std::vector<std::string> names(chrom_names, chrom_names + n_chrom);
    std::string all;
    std::vector<uint64_t> begin;
    for (uint32_t r = 0; r < res->n_reads; ++r) {
        std::vector<LiftedRecord> recs;
        for (uint32_t k = res->read_rec_begin[r]; k < res->read_rec_begin[r + 1]; ++k) {
            LiftedRecord x;
            x.status = res->rec_status[k];
            x.tid = res->rec_tid[k];
            x.pos = res->rec_pos[k];
            x.flag = res->rec_flag[k];
            x.mapq = res->rec_mapq[k];
            x.cigar = decode_cigar(res->cigar + res->rec_cigar_begin[k], res->rec_cigar_begin[k + 1] - res->rec_cigar_begin[k]);
            recs.push_back(std::move(x));
        }
        for (const auto& s : format_sa_tags(names, recs)) {
            begin.push_back(all.size());
            all += s;
            all.push_back('\0');
        }
    }
    begin.push_back(all.size());
    *need = all.size();
    if (all.size() > cap) return PTL_ERR_INVALID_ARG;
    std::memcpy(buf, all.data(), all.size());
    std::memcpy(sa_begin, begin.data(), begin.size() * sizeof(uint64_t));
    return PTL_OK;
}

}  // extern "C"
