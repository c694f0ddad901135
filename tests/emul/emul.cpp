// TEST INFRASTRUCTURE ONLY.  Host emulation of the CUDA liftover pipeline: the product's per-index device bodies
// (portello_b200/csrc/device/pair_bodies.cuh + lift_device.cuh) compiled under a one-lane SIMT shim and driven in the
// same order as launch_lift (kernels.cu), behind the same C-ABI as the product under the prefix ptl_emul_.  The CPU
// test-suite compares it bit-exactly with the oracle, so kernel logic regressions are caught without a GPU.  The warp
// cooperation itself (32 lanes in lock step, record emission, scans) is only exercised by the -m gpu tests.
// Nothing under portello_b200/ links or loads this library.
#include "cuda_shim.hpp"

#include <numeric>
#include <string>
#include <vector>

#include "../../portello_b200/csrc/device/pair_bodies.cuh"
#include "../../portello_b200/csrc/device/assemble_bam.cuh"
#include "emul_prep.hpp"

using namespace ptl;

// emul_warp.cpp: pair_count_warp_kernel under the 32-lane lock-step shim
void emul_pair_count_warp(const ptl::DevStatic& S, const ptl::DevBatch& B, const ptl::DevWork& W, ptl::DevTotals* T);
// emul_warp.cpp: table_build_kernel under the 32-lane lock-step shim
void emul_table_build_warp(const ptl::DevStatic& S, uint32_t* counts, ptl::TabEntry* out);
// emul_warp.cpp: lift_long_pairs_kernel under the 32-lane lock-step shim
void emul_lift_long_pairs(const ptl::DevStatic& S, const ptl::DevBatch& B, const ptl::DevWork& W, ptl::DevTotals* T, uint32_t stage_mask);

namespace {

template <class T>
std::vector<T> padded_copy(const T* p, size_t n, size_t pad = 16) {
    std::vector<T> v(n + pad, T{});
    if (n) std::memcpy(v.data(), p, n * sizeof(T));
    return v;
}

struct EmulSlot {
    // batch copies
    std::vector<uint16_t> read_flag, read_bin;
    std::vector<uint8_t> read_mapq, rseg_is_fwd, seq4;
    std::vector<uint32_t> read_seq_len, read_seg_begin, rseg_contig, rseg_cigar_len, cigar;
    std::vector<uint64_t> read_seq_off, rseg_cigar_begin, indel_win;
    std::vector<uint32_t> rseg_win_begin;
    std::vector<int64_t> rseg_pos;
    // results: the compact arena of the product (device_types.hpp: result_layout)
    std::vector<char> arena;
    ptl_result res{};
    DevTotals totals{};
    bool ran = false;
    // ptl_emul_assemble_records
    std::vector<uint32_t> rseg_read;
    std::vector<uint64_t> bam_begin;
    std::vector<uint8_t> bam_out;
};

}  // namespace

struct ptl_ctx {
    std::string err;
    std::vector<EmulSlot> slots;
    uint32_t long_pair_ops = 64;
    bool check_warp_count = false;  // ptl_emul_check_warp_pair_count
    // static state
    std::vector<uint8_t> ref;
    std::vector<uint64_t> chrom_off;
    EmulFlat flat;
    std::vector<uint64_t> contig_rev_off;
    std::vector<uint8_t> rev_pool;
    std::vector<uint32_t> tab_begin;
    std::vector<TabEntry> table;
    DevStatic S;
    bool have_reference = false, have_segments = false;
    std::vector<uint64_t> contig_name_off{0}, chrom_name_off{0};
    std::vector<uint8_t> contig_names, chrom_names;
    bool have_names = false;
};

namespace {

int fail(ptl_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

int install(ptl_ctx* ctx, int mode, const void* in) {
    EmulFlat f;
    std::string err;
    const int rc = emul_prepare(mode, in, &f, &err);
    if (rc != PTL_OK) return fail(ctx, rc, err);
    const uint32_t nc = uint32_t(f.contig_len.size()), ns = uint32_t(f.so_start.size());
    for (uint32_t g = 0; g < ns; ++g) {
        if (f.chrom[g] < 0 || (ctx->have_reference && uint32_t(f.chrom[g]) >= ctx->S.n_chrom))
            return fail(ctx, PTL_ERR_INPUT, "contig segment with a chromosome index outside the reference");
        if (f.pos[g] < 0 || f.pos[g] > 0x7fffffffLL) return fail(ctx, PTL_ERR_INPUT, "contig segment position outside the BAM int32 range");
    }
    for (uint32_t c = 0; c < nc; ++c)
        if (f.contig_len[c] > 0x7fffffffULL) return fail(ctx, PTL_ERR_INPUT, "contig longer than the BAM int32 sequence length limit");
    ctx->flat = std::move(f);
    const EmulFlat& F = ctx->flat;
    DevStatic& S = ctx->S;
    S.n_contigs = nc;
    S.n_segments = ns;
    S.contig_seg_begin = F.seg_begin.data();
    S.contig_len = F.contig_len.data();
    ctx->contig_rev_off.assign(nc, ~0ull);
    uint64_t pool = 0;
    for (uint32_t c = 0; c < nc; ++c)
        if (F.has_rev[c]) { ctx->contig_rev_off[c] = pool; pool += (F.contig_len[c] + 15) & ~15ull; }
    ctx->rev_pool.assign(pool + 32, 0);
    for (uint32_t c = 0; c < nc; ++c)
        if (F.has_rev[c] && F.contig_len[c]) std::memcpy(ctx->rev_pool.data() + ctx->contig_rev_off[c], F.rev_seq[c].data(), F.contig_len[c]);
    S.rev_pool = ctx->rev_pool.data();
    S.contig_rev_off = ctx->contig_rev_off.data();
    S.seg_so_start = F.so_start.data();
    S.seg_so_end = F.so_end.data();
    S.seg_chrom = F.chrom.data();
    S.seg_pos = F.pos.data();
    S.seg_is_fwd = F.is_fwd.data();
    S.seg_mapq = F.mapq.data();
    S.seg_cigar_begin = F.cigar_begin.data();
    S.seg_cigar = F.cigar.data();
    // tables: count -> scan -> fill (launch_table_build x2 + exclusive scan)
    ctx->tab_begin.assign(size_t(ns) + 1, 0);
    S.seg_tab_begin = ctx->tab_begin.data();
    S.table = nullptr;
    for (uint32_t g = 0; g < ns; ++g) table_build_body(S, g, ctx->tab_begin.data(), nullptr);
    uint32_t run = 0;
    for (uint32_t g = 0; g <= ns; ++g) { const uint32_t c = ctx->tab_begin[g]; ctx->tab_begin[g] = run; run += (g < ns) ? c : 0; }
    ctx->table.assign(size_t(run) + 1, TabEntry{0, 0, 0, 0});
    S.table = ctx->table.data();
    for (uint32_t g = 0; g < ns; ++g) table_build_body(S, g, nullptr, ctx->table.data());
    ctx->have_segments = true;
    return PTL_OK;
}

template <class T>
void exclusive_scan(T* a, size_t n) {
    T run{};
    for (size_t i = 0; i < n; ++i) { const T v = a[i]; a[i] = run; run += v; }
}

int run_batch(ptl_ctx* ctx, EmulSlot& sl, const ptl_batch* b, uint32_t stage_mask) {
    if (!ctx->have_segments) return fail(ctx, PTL_ERR_INVALID_ARG, "ptl_set_contig_segments / ptl_set_contig_records has not been called");
    if ((stage_mask & PTL_STAGE_SIMPLIFY) && !ctx->have_reference) return fail(ctx, PTL_ERR_INVALID_ARG, "ptl_set_reference has not been called");
    const uint32_t n = b->n_reads, ns = b->n_read_segments;
    if (n && (b->read_seg_begin[0] != 0 || b->read_seg_begin[n] != ns)) return fail(ctx, PTL_ERR_INVALID_ARG, "read_seg_begin must run from 0 to n_read_segments");
    sl.read_flag = padded_copy(b->read_flag, n);
    sl.read_mapq = padded_copy(b->read_mapq, n);
    sl.read_bin = padded_copy(b->read_bin, n);
    sl.read_seq_len = padded_copy(b->read_seq_len, n);
    sl.read_seq_off = padded_copy(b->read_seq_off, n);
    sl.read_seg_begin = padded_copy(b->read_seg_begin, n ? size_t(n) + 1 : 0);
    sl.rseg_contig = padded_copy(b->rseg_contig, ns);
    sl.rseg_pos = padded_copy(b->rseg_pos, ns);
    sl.rseg_is_fwd = padded_copy(b->rseg_is_fwd, ns);
    sl.rseg_cigar_begin = padded_copy(b->rseg_cigar_begin, ns);
    sl.rseg_cigar_len = padded_copy(b->rseg_cigar_len, ns);
    sl.cigar = padded_copy(b->cigar, b->n_cigar);
    sl.seq4 = padded_copy(b->seq4, b->seq4_bytes, 64);
    DevBatch B;
    B.n_reads = n; B.n_rsegs = ns; B.n_cigar = b->n_cigar;
    B.read_flag = sl.read_flag.data(); B.read_mapq = sl.read_mapq.data(); B.read_bin = sl.read_bin.data();
    B.read_seq_len = sl.read_seq_len.data(); B.read_seq_off = sl.read_seq_off.data(); B.read_seg_begin = sl.read_seg_begin.data();
    B.rseg_contig = sl.rseg_contig.data(); B.rseg_pos = sl.rseg_pos.data(); B.rseg_is_fwd = sl.rseg_is_fwd.data();
    B.rseg_cigar_begin = sl.rseg_cigar_begin.data(); B.rseg_cigar_len = sl.rseg_cigar_len.data(); B.cigar = sl.cigar.data();
    B.seq4 = sl.seq4.data();
    B.seq4_bytes = b->seq4_bytes;
    if (b->indel_win && b->rseg_win_begin) {
        sl.rseg_win_begin = padded_copy(b->rseg_win_begin, size_t(ns) + 1);
        sl.indel_win = padded_copy(b->indel_win, b->n_indel_win);
        B.rseg_win_begin = sl.rseg_win_begin.data();
        B.indel_win = sl.indel_win.data();
    }
    const DevStatic& S = ctx->S;
    DevTotals& T = sl.totals;
    totals_reset(&T);
    const int do_finish = (stage_mask == 7u);

    // ---- pair_count -> scan -> pair_fill -> scan
    std::vector<uint32_t> rseg_read(ns + 1), rseg_pair_begin(size_t(ns) + 1, 0), rseg_n_id(ns + 1), rseg_read_len(ns + 1);
    std::vector<int64_t> rseg_ref_len(ns + 1);
    DevWork W;
    W.rseg_read = rseg_read.data(); W.rseg_pair_begin = rseg_pair_begin.data(); W.rseg_ref_len = rseg_ref_len.data();
    W.rseg_n_id = rseg_n_id.data(); W.rseg_read_len = rseg_read_len.data();
    for (uint32_t r = 0; r < n; ++r) pair_count_body(S, B, W, &T, r);
    if (ctx->check_warp_count) {  // the warp-cooperative count (long-read batches on the device) must write the same arrays
        std::vector<uint32_t> r2(ns + 1), pb2(size_t(ns) + 1, 0), id2(ns + 1), rl2(ns + 1);
        std::vector<int64_t> ref2(ns + 1);
        DevWork W2 = W;
        W2.rseg_read = r2.data(); W2.rseg_pair_begin = pb2.data(); W2.rseg_ref_len = ref2.data(); W2.rseg_n_id = id2.data(); W2.rseg_read_len = rl2.data();
        DevTotals T2;
        totals_reset(&T2);
        emul_pair_count_warp(S, B, W2, &T2);
        for (uint32_t s2 = 0; s2 < ns; ++s2)
            if (r2[s2] != rseg_read[s2] || pb2[s2] != rseg_pair_begin[s2] || id2[s2] != rseg_n_id[s2] || rl2[s2] != rseg_read_len[s2] || ref2[s2] != rseg_ref_len[s2])
                return fail(ctx, PTL_ERR_STATE, "pair_count_warp_body differs from pair_count_body at read segment " + std::to_string(s2));
        if ((T2.overflow ^ T.overflow) & OVF_INVALID) return fail(ctx, PTL_ERR_STATE, "pair_count_warp_body: validity flag differs");
    }
    rseg_pair_begin[ns] = 0;
    exclusive_scan(rseg_pair_begin.data(), size_t(ns) + 1);
    const uint32_t np = rseg_pair_begin[ns];
    T.n_pairs = np;
    std::vector<uint32_t> pair_rseg(np + 1), pair_seg(np + 1), pair_cap_b(np + 1), pair_tab_lo(np + 1), pair_n_out(np + 1), simplify_list(np + 1);
    std::vector<uint64_t> pair_slot_begin(size_t(np) + 1, 0), pair_out_off(np + 1);
    std::vector<int8_t> pair_status(np + 1);
    std::vector<uint8_t> pair_flip(np + 1);
    std::vector<int64_t> pair_pos(np + 1);
    std::vector<uint16_t> pair_bin(np + 1);
    W.pair_cap = np;
    W.pair_rseg = pair_rseg.data(); W.pair_seg = pair_seg.data(); W.pair_slot_begin = pair_slot_begin.data(); W.pair_cap_b = pair_cap_b.data(); W.pair_tab_lo = pair_tab_lo.data();
    W.pair_status = pair_status.data(); W.pair_flip = pair_flip.data(); W.pair_pos = pair_pos.data(); W.pair_n_out = pair_n_out.data();
    W.pair_bin = pair_bin.data(); W.pair_out_off = pair_out_off.data(); W.simplify_list = simplify_list.data();
    std::vector<uint32_t> long_list(np + 1);
    W.long_list = long_list.data();
    std::vector<PairDesc> pair_desc(np + 1);
    std::vector<PairDescRev> pair_desc_rev(np + 1);
    W.pair_desc = pair_desc.data();
    W.pair_desc_rev = pair_desc_rev.data();
    W.long_ops = ctx->long_pair_ops;
    for (uint32_t s = 0; s < ns; ++s) pair_fill_body(S, B, W, s);
    pair_slot_begin[np] = 0;
    exclusive_scan(pair_slot_begin.data(), size_t(np) + 1);
    std::vector<uint32_t> scratch(pair_slot_begin[np] + 16, 0xdeadbeefu);
    W.scratch_cap = pair_slot_begin[np];
    W.scratch = scratch.data();
    T.scratch_needed = pair_slot_begin[np];

    // ---- lift_pairs (+ simplify worklist)
    uint32_t in_ops = 0, base_bytes = 0;
    for (uint32_t p = 0; p < np; ++p) {  // the kernel launches the production instantiation for the full stage mask
        if (stage_mask == 7u) lift_pair_body<true>(S, B, W, &T, p, true, stage_mask, in_ops, base_bytes);
        else lift_pair_body<false>(S, B, W, &T, p, true, stage_mask, in_ops, base_bytes);
    }
    T.n_in_ops = in_ops;
    T.n_base_bytes = base_bytes;
    if ((stage_mask & 2u) || stage_mask == 1u) emul_lift_long_pairs(S, B, W, &T, stage_mask);  // warp_pairs_kernel
    if ((stage_mask & 6u) == 6u) {  // simplify_pairs_kernel: one thread per listed pair (even entries; see emul_warp.cpp)
        uint32_t bb = 0;
        for (uint32_t i = 0; i < std::min<uint32_t>(T.n_simplify, W.pair_cap); i += 2) simplify_thread_pair_body(S, B, W, &T, i, true, bb);
        T.n_base_bytes += bb;
    }

    // ---- read_finalize -> scan -> emit_records
    std::vector<uint2> read_counts(size_t(n) + 1, make_uint2(0, 0));
    std::vector<uint32_t> read_primary(n + 1);
    W.read_counts = read_counts.data();
    W.read_primary = read_primary.data();
    for (uint32_t r = 0; r < n; ++r) T.n_lifted += read_finalize_body(S, B, W, &T, r, do_finish);
    uint2 run = make_uint2(0, 0);
    for (uint32_t r = 0; r <= n; ++r) { const uint2 v = read_counts[r]; read_counts[r] = run; run.x += v.x; run.y += v.y; }
    const uint32_t nr = read_counts[n].x;
    const uint64_t nc = read_counts[n].y;
    T.n_records = nr;
    T.n_cigar_out = nc;
    const ResultLayout L = result_layout(n, nr, nc);
    sl.arena.assign(L.total + 64, char(0x5a));
    const DevResult R = DevResult::view(sl.arena.data(), L);
    for (uint32_t r = 0; r < n; ++r) {  // what emit_records_kernel does for read r (lane per read on the device)
        const uint2 base = read_counts[r], next = read_counts[r + 1];
        R.read_rec_begin[r] = base.x;
        uint32_t k = base.x;
        uint64_t op_at = base.y;
        const uint32_t s0 = min(B.read_seg_begin[r], B.n_rsegs), s1 = min(max(B.read_seg_begin[r + 1], s0), B.n_rsegs);
        if (next.x == base.x) continue;
        if (read_primary[r] == 0xffffffffu) {
            emit_unmapped_record(B, R, r, k, s0, op_at);
            continue;
        }
        for (uint32_t p = W.rseg_pair_begin[s0]; p < W.rseg_pair_begin[s1]; ++p) {
            if (W.pair_status[p] != ST_LIFTED) continue;
            emit_lifted_record(S, B, W, R, p, k, B.read_flag[r], read_primary[r], op_at, stage_mask);
            const uint32_t no = W.pair_n_out[p];
            for (uint32_t i = 0; i < no; ++i) R.cigar[op_at + i] = W.scratch[W.pair_out_off[p] + i];
            ++k;
            op_at += no;
        }
    }
    R.read_rec_begin[n] = nr;
    R.rec_cigar_begin[nr] = nc;

    ptl_result& res = sl.res;
    res = ptl_result{};
    res.n_reads = n;
    res.read_rec_begin = R.read_rec_begin;
    res.n_records = nr;
    res.rec_status = R.rec_status; res.rec_read_segment = R.rec_read_segment; res.rec_contig_segment = R.rec_contig_segment;
    res.rec_tid = R.rec_tid; res.rec_pos = R.rec_pos; res.rec_mapq = R.rec_mapq; res.rec_flag = R.rec_flag;
    res.rec_bin = R.rec_bin; res.rec_need_flip = R.rec_need_flip; res.rec_cigar_begin = R.rec_cigar_begin;
    res.cigar = R.cigar;
    res.n_cigar = nc;
    res.n_pairs = T.n_pairs;
    res.n_lifted = T.n_lifted;
    res.n_errors = T.n_errors;
    if (T.n_errors) {
        res.first_error_read = T.first_error_read >> 8;
        res.first_error_status = int32_t(int8_t(T.first_error_read & 0xff));
    } else {
        res.first_error_read = -1;
        res.first_error_status = 0;
    }
    sl.rseg_read = rseg_read;
    sl.ran = true;
    return PTL_OK;
}

EmulSlot* get_slot(ptl_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || size_t(slot) >= ctx->slots.size()) return nullptr;
    return &ctx->slots[size_t(slot)];
}

}  // namespace

extern "C" {

const char* ptl_emul_version(void) { return "portello_b200 device-body emulation (test infrastructure)"; }
int ptl_emul_create(int /*device*/, int n_slots, ptl_ctx** out) {
    if (!out || n_slots < 1 || n_slots > 64) return PTL_ERR_INVALID_ARG;
    auto* ctx = new ptl_ctx();
    ctx->slots.resize(size_t(n_slots));
    *out = ctx;
    return PTL_OK;
}
void ptl_emul_destroy(ptl_ctx* ctx) { delete ctx; }
const char* ptl_emul_last_error(const ptl_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }

int ptl_emul_set_reference(ptl_ctx* ctx, uint32_t n_chrom, const uint64_t* chrom_len, const uint8_t* const* chrom_seq) {
    if (!ctx || (n_chrom && (!chrom_len || !chrom_seq))) return PTL_ERR_INVALID_ARG;
    if (ctx->have_segments)  // (as ptl_set_reference: segments installed first must fit the reference)
        for (int32_t c : ctx->flat.chrom)
            if (c < 0 || uint32_t(c) >= n_chrom) return fail(ctx, PTL_ERR_INPUT, "the reference has fewer chromosomes than the installed contig segments use");
    ctx->chrom_off.assign(size_t(n_chrom) + 1, 0);
    for (uint32_t c = 0; c < n_chrom; ++c) ctx->chrom_off[c + 1] = ctx->chrom_off[c] + chrom_len[c];
    ctx->ref.assign(ctx->chrom_off[n_chrom] + 32, 0);
    for (uint32_t c = 0; c < n_chrom; ++c)
        if (chrom_len[c]) std::memcpy(ctx->ref.data() + ctx->chrom_off[c], chrom_seq[c], chrom_len[c]);
    ctx->S.ref = ctx->ref.data();
    ctx->S.chrom_off = ctx->chrom_off.data();
    ctx->S.n_chrom = n_chrom;
    ctx->have_reference = true;
    return PTL_OK;
}
int ptl_emul_set_contig_segments(ptl_ctx* ctx, const ptl_contig_segments* s) { return (ctx && s) ? install(ctx, 0, s) : PTL_ERR_INVALID_ARG; }
int ptl_emul_set_raw_contig_segments(ptl_ctx* ctx, const ptl_contig_segments* s) { return (ctx && s) ? install(ctx, 1, s) : PTL_ERR_INVALID_ARG; }
int ptl_emul_set_contig_records(ptl_ctx* ctx, const ptl_contig_records* r) { return (ctx && r) ? install(ctx, 2, r) : PTL_ERR_INVALID_ARG; }
int ptl_emul_get_contig_segments(const ptl_ctx* ctx, ptl_contig_segments* out) {
    if (!ctx || !out || !ctx->have_segments) return PTL_ERR_INVALID_ARG;
    ctx->flat.view(out);
    return PTL_OK;
}
int ptl_emul_get_segment_table(ptl_ctx* ctx, uint32_t segment, uint32_t cap, uint32_t* keys, int32_t* vals, uint32_t* n) {
    if (!ctx || !n || !ctx->have_segments || segment >= ctx->S.n_segments) return PTL_ERR_INVALID_ARG;
    const uint32_t t0 = ctx->tab_begin[segment], t1 = ctx->tab_begin[segment + 1];
    *n = t1 - t0;
    if (*n > cap) return PTL_ERR_INVALID_ARG;
    for (uint32_t i = 0; i < *n; ++i) {
        keys[i] = ctx->table[t0 + i].key;
        vals[i] = ctx->table[t0 + i].val;
    }
    return PTL_OK;
}
// gaps of a segment table (TabEntry::gap), for the table tests
int ptl_emul_get_segment_table_gaps(ptl_ctx* ctx, uint32_t segment, uint32_t cap, uint32_t* gaps) {
    if (!ctx || !ctx->have_segments || segment >= ctx->S.n_segments) return PTL_ERR_INVALID_ARG;
    const uint32_t t0 = ctx->tab_begin[segment], t1 = ctx->tab_begin[segment + 1];
    if (t1 - t0 > cap) return PTL_ERR_INVALID_ARG;
    for (uint32_t i = 0; i < t1 - t0; ++i) gaps[i] = ctx->table[t0 + i].gap;
    return PTL_OK;
}

int ptl_emul_lift_submit_ex(ptl_ctx* ctx, int slot, const ptl_batch* b, uint32_t stage_mask) {
    EmulSlot* sl = get_slot(ctx, slot);
    if (!sl || !b) return PTL_ERR_INVALID_ARG;
    return run_batch(ctx, *sl, b, stage_mask);
}
int ptl_emul_lift_submit(ptl_ctx* ctx, int slot, const ptl_batch* b) { return ptl_emul_lift_submit_ex(ctx, slot, b, PTL_STAGE_ALL); }
int ptl_emul_lift_wait(ptl_ctx* ctx, int slot, ptl_result* out) {
    EmulSlot* sl = get_slot(ctx, slot);
    if (!sl || !out) return PTL_ERR_INVALID_ARG;
    if (!sl->ran) return fail(ctx, PTL_ERR_STATE, "ptl_lift_wait without a submitted batch");
    if (sl->totals.overflow & OVF_INVALID) return fail(ctx, PTL_ERR_INVALID_ARG, "malformed batch");
    *out = sl->res;
    if (sl->res.n_errors)
        return fail(ctx, PTL_ERR_LIFT_PANIC, "the reference would panic on read " + std::to_string(sl->res.first_error_read) + " (status " +
                                                 std::to_string(sl->res.first_error_status) + ")");
    return PTL_OK;
}
int ptl_emul_set_long_pair_ops(ptl_ctx* ctx, uint32_t n_ops) {
    if (!ctx) return PTL_ERR_INVALID_ARG;
    ctx->long_pair_ops = n_ops;
    return PTL_OK;
}
// pairs of the last batch that went through the warp-cooperative liftover
int ptl_emul_slot_long_pairs(ptl_ctx* ctx, int slot, uint64_t* out) {
    EmulSlot* sl = get_slot(ctx, slot);
    if (!sl || !out || !sl->ran) return PTL_ERR_INVALID_ARG;
    *out = sl->totals.n_long;
    return PTL_OK;
}
// counters of the last batch: n_pairs, n_lifted, n_in_ops, n_out_ops, base bytes compared, scratch ops
int ptl_emul_slot_counters(ptl_ctx* ctx, int slot, uint64_t* out) {
    EmulSlot* sl = get_slot(ctx, slot);
    if (!sl || !out || !sl->ran) return PTL_ERR_INVALID_ARG;
    const DevTotals& t = sl->totals;
    out[0] = t.n_pairs; out[1] = t.n_lifted; out[2] = t.n_in_ops; out[3] = t.n_cigar_out; out[4] = t.n_base_bytes; out[5] = t.scratch_needed;
    return PTL_OK;
}

// Every following batch also runs the warp-cooperative pair count and compares it with the scalar one (submit fails on a difference).
int ptl_emul_check_warp_pair_count(ptl_ctx* ctx, int enable) {
    if (!ctx) return PTL_ERR_INVALID_ARG;
    ctx->check_warp_count = enable != 0;
    return PTL_OK;
}

// The installed segment tables rebuilt by the warp-cooperative kernel body (32 lanes in lock step) and compared with the scalar
// build: returns the number of differing words (counts and entries), 0 = identical.
int64_t ptl_emul_table_build_warp_mismatches(ptl_ctx* ctx) {
    if (!ctx || !ctx->have_segments) return -1;
    const uint32_t ns = ctx->S.n_segments;
    std::vector<uint32_t> counts(size_t(ns) + 1, 0);
    emul_table_build_warp(ctx->S, counts.data(), nullptr);
    int64_t bad = 0;
    for (uint32_t g = 0; g < ns; ++g) bad += counts[g] != ctx->tab_begin[g + 1] - ctx->tab_begin[g];
    if (bad) return bad;
    std::vector<TabEntry> t(ctx->table.size(), TabEntry{0xdeadbeefu, 0, 0, 0});
    emul_table_build_warp(ctx->S, nullptr, t.data());
    for (uint32_t i = 0; i < ctx->tab_begin[ns]; ++i)
        bad += t[i].key != ctx->table[i].key || t[i].val != ctx->table[i].val || t[i].gap != ctx->table[i].gap;
    return bad;
}

// Record assembly through the device bodies of assemble_bam.cuh (every "thread" of a block runs in turn: the bodies need no
// cooperation between threads), byte for byte what ptl_assemble_records launches.
int ptl_emul_set_names(ptl_ctx* ctx, uint32_t n_contigs, const char* const* contig_names, uint32_t n_chrom, const char* const* chrom_names) {
    if (!ctx || (n_contigs && !contig_names) || (n_chrom && !chrom_names)) return PTL_ERR_INVALID_ARG;
    auto pool = [](uint32_t n, const char* const* names, std::vector<uint64_t>& off, std::vector<uint8_t>& bytes) {
        off.assign(1, 0);
        bytes.clear();
        for (uint32_t i = 0; i < n; ++i) {
            bytes.insert(bytes.end(), names[i], names[i] + std::strlen(names[i]));
            off.push_back(bytes.size());
        }
        bytes.resize(bytes.size() + 16, 0);
    };
    pool(n_contigs, contig_names, ctx->contig_name_off, ctx->contig_names);
    pool(n_chrom, chrom_names, ctx->chrom_name_off, ctx->chrom_names);
    ctx->have_names = true;
    return PTL_OK;
}
int ptl_emul_assemble_records(ptl_ctx* ctx, int slot, const ptl_read_extras* x, uint32_t /*flags*/, ptl_bam_records* out) {
    EmulSlot* sl = get_slot(ctx, slot);
    if (!sl || !x || !out) return PTL_ERR_INVALID_ARG;
    if (!sl->ran) return fail(ctx, PTL_ERR_STATE, "assemble without a lifted batch");
    if (!ctx->have_names) return fail(ctx, PTL_ERR_STATE, "ptl_set_names has not been called");
    const uint32_t n = sl->res.n_reads, n_rec = sl->res.n_records;
    constexpr size_t kPad = 64;  // the copy loops read a few bytes around a field (inside the device pools)
    auto padded = [&](const uint8_t* p, size_t bytes) {
        std::vector<uint8_t> v(bytes + 2 * kPad, 0);
        if (bytes) std::memcpy(v.data() + kPad, p, bytes);
        return v;
    };
    const std::vector<uint8_t> names = padded(x->names, n ? x->name_off[n] : 0), aux = padded(x->aux, n ? x->aux_off[n] : 0),
                               qual = padded(x->quals.qual, x->quals.qual_bytes), seq4 = padded(sl->seq4.data(), sl->seq4.size());
    const std::vector<uint8_t> cig = padded(reinterpret_cast<const uint8_t*>(sl->res.cigar), size_t(sl->res.n_cigar) * 4);
    std::vector<uint32_t> keep(size_t(n) * 10 + 10), sa_len(n_rec + 1);
    std::vector<uint4> desc(size_t(n_rec) * 2 + 2);
    sl->bam_begin.assign(size_t(n_rec) + 1, 0);
    unsigned int err = 0;
    BamAsmArgs A{};
    A.n_reads = n; A.n_records = n_rec;
    A.seg_is_fwd = ctx->S.seg_is_fwd; A.contig_seg_begin = ctx->S.contig_seg_begin;
    A.contig_name_off = ctx->contig_name_off.data(); A.contig_names = ctx->contig_names.data(); A.n_contig_names = uint32_t(ctx->contig_name_off.size() - 1);
    A.chrom_name_off = ctx->chrom_name_off.data(); A.chrom_names = ctx->chrom_names.data(); A.n_chrom_names = uint32_t(ctx->chrom_name_off.size() - 1);
    A.read_mapq = sl->read_mapq.data(); A.read_seq_len = sl->read_seq_len.data(); A.read_seq_off = sl->read_seq_off.data();
    A.seq4 = seq4.data() + kPad; A.rseg_contig = sl->rseg_contig.data(); A.rseg_read = sl->rseg_read.data();
    A.name_off = x->name_off; A.names = names.data() + kPad; A.aux_off = x->aux_off; A.aux = aux.data() + kPad;
    A.mate_tid = x->mate_tid; A.mate_pos = x->mate_pos; A.tlen = x->tlen; A.qual_off = x->quals.read_qual_off; A.qual = qual.data() + kPad;
    A.names_bytes = n ? x->name_off[n] : 0; A.aux_bytes = n ? x->aux_off[n] : 0; A.qual_bytes = x->quals.qual_bytes;
    A.read_rec_begin = sl->res.read_rec_begin; A.rec_status = sl->res.rec_status; A.rec_read_segment = sl->res.rec_read_segment;
    A.rec_contig_segment = sl->res.rec_contig_segment; A.rec_tid = sl->res.rec_tid; A.rec_pos = sl->res.rec_pos; A.rec_mapq = sl->res.rec_mapq;
    A.rec_flag = sl->res.rec_flag; A.rec_bin = sl->res.rec_bin; A.rec_need_flip = sl->res.rec_need_flip; A.rec_cigar_begin = sl->res.rec_cigar_begin;
    A.cigar = reinterpret_cast<const uint32_t*>(cig.data() + kPad);
    A.read_keep = keep.data(); A.rec_sa_len = sa_len.data(); A.rec_begin = sl->bam_begin.data(); A.rec_desc = desc.data(); A.error = &err;
    for (uint32_t r = 0; r < n; ++r) bam_read_prep_body(A, r);
    for (uint32_t k = 0; k < n_rec; ++k) bam_rec_prep_body(A, k);
    uint64_t run = 0;
    for (uint32_t k = 0; k < n_rec; ++k) {  // bam_rec_size_kernel + the exclusive scan
        const BamRecLayout L = bam_rec_layout(A, k);
        if (L.ps_n > 0xffu || L.sa_n > 0xffffffu) err |= 4u;
        if (L.name_n > 254u) err |= 8u;
        L.pack(desc.data() + 2 * size_t(k));
        sl->bam_begin[k] = run;
        run += L.total;
    }
    sl->bam_begin[n_rec] = run;
    if (err & 16u) return fail(ctx, PTL_ERR_INVALID_ARG, "ptl_read_extras: an offset lies outside its pool");
    if (err & 1u) return fail(ctx, PTL_ERR_INVALID_ARG, "ptl_set_names: a contig or reference chromosome of this batch has no name");
    if (err & 2u) return fail(ctx, PTL_ERR_INVALID_ARG, "a lifted CIGAR has more than 65535 ops");
    if (err & 4u) return fail(ctx, PTL_ERR_INVALID_ARG, "descriptor field overflow");
    if (err & 8u) return fail(ctx, PTL_ERR_INVALID_ARG, "a read name longer than 254 bytes");
    sl->bam_out.assign(run + 2 * kPad, 0x5a);
    // (the device pool is 256-byte aligned: keep the same alignment so that the 16-byte chunking takes the same paths)
    uint8_t* base = sl->bam_out.data();
    base += (256 - (reinterpret_cast<uintptr_t>(base) & 255u)) & 255u;
    if (size_t(base - sl->bam_out.data()) + run > sl->bam_out.size()) { sl->bam_out.resize(run + 512, 0x5a); base = sl->bam_out.data(); base += (256 - (reinterpret_cast<uintptr_t>(base) & 255u)) & 255u; }
    A.out = base;
    for (uint32_t k = 0; k < n_rec; ++k) {
        for (uint32_t tid = 0; tid < 32; ++tid) bam_write_meta_body(A, k, tid, 32);
        for (uint32_t tid = 0; tid < 256; ++tid) bam_write_bases_body(A, k, tid, 256);
    }
    *out = ptl_bam_records{};
    out->n_records = n_rec;
    out->rec_begin = sl->bam_begin.data();
    out->bytes = base;
    out->bytes_written = run;
    return PTL_OK;
}

}  // extern "C"
