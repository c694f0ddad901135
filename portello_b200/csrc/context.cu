// Host orchestration + C-ABI of libportello_b200.so (include/portello_b200.h).
//
// One ptl_ctx per GPU.  Static state (reference, rev_contig_seq, segment tables) is uploaded once; each of the
// n_slots batch slots owns a CUDA stream, grow-only device buffers and pinned host result buffers, which replaces
// the reference's per-thread reader/worker pool (src/worker_thread_data.rs:8-30) with a multi-buffered stream
// pipeline.  There is NO CPU fallback: without a usable device ptl_create fails with PTL_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <cstdlib>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/portello_b200.h"
#include "device/assemble.cuh"
#include "device/assemble_bam.cuh"
#include "device/bgzf_store.cuh"
#include "host/bgzf_tables.hpp"
#include "device/kernels.hpp"
#include "host/contig_prep.hpp"

using namespace ptl;

namespace {

struct CudaError {
    std::string msg;
};
#define CK(call)                                                                                                  \
    do {                                                                                                          \
        cudaError_t e_ = (call);                                                                                  \
        if (e_ != cudaSuccess)                                                                                    \
            throw CudaError{std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"}; \
    } while (0)

// grow-only device buffer
struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes, cudaStream_t st) {
        if (bytes <= cap) return;
        if (p) {
            CK(cudaStreamSynchronize(st));
            CK(cudaFree(p));
            p = nullptr;
            cap = 0;
        }
        const size_t want = (std::max<size_t>(bytes + bytes / 8, 256) + 255) & ~size_t(255);  // (whole granules, as ptl_host_alloc)
        CK(cudaMalloc(&p, want));
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};
// grow-only pinned host buffer
struct HBuf {
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) CK(cudaFreeHost(p));
        p = nullptr;
        cap = 0;
        const size_t want = (std::max<size_t>(bytes + bytes / 8, 256) + 255) & ~size_t(255);
        CK(cudaHostAlloc(&p, want, cudaHostAllocPortable));
        cap = want;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct Slot {
    cudaStream_t stream = nullptr;
    StageEvents ev{};
    bool have_events = false;
    // device copies of the batch
    DBuf d_read_flag, d_read_mapq, d_read_bin, d_read_seq_len, d_read_seq_off, d_read_seg_begin, d_rseg_contig, d_rseg_pos,
        d_rseg_is_fwd, d_rseg_cigar_begin, d_rseg_cigar_len, d_cigar, d_seq4, d_arena, d_win_begin, d_win;
    // work
    DBuf w_rseg_read, w_rseg_pair_begin, w_rseg_ref_len, w_rseg_n_id, w_rseg_read_len, w_pair_cap_b, w_pair_tab_lo, w_pair_rseg, w_pair_seg, w_pair_slot_begin, w_pair_status, w_pair_flip,
        w_pair_pos, w_pair_n_out, w_pair_bin, w_pair_out_off, w_simplify_list, w_long_list, w_pair_order, w_pair_desc, w_pair_desc_rev, w_scratch, w_read_counts, w_read_primary, w_scan_tmp, w_totals;
    // record assembly (ptl_assemble_bases): uploaded qualities, per-record offsets, output pools and their pinned twins
    DBuf a_qual, a_qual_off, a_rec_read, a_seq_begin, a_qual_begin, a_out_seq, a_out_qual;
    HBuf ha_seq_begin, ha_qual_begin, ha_out_seq, ha_out_qual;
    cudaEvent_t a_ev[2] = {nullptr, nullptr};
    uint64_t a_qual_bytes = 0;
    uint32_t a_qual_n_reads = 0;      // reads the resident qualities were uploaded for (0: none; cleared by a new batch)
    // record assembly, whole BAM records (ptl_assemble_records): uploaded names / aux / mate fields, work arrays, output
    DBuf b_name_off, b_names, b_aux_off, b_aux, b_mate_tid, b_mate_pos, b_tlen, b_keep, b_sa_len, b_rec_begin, b_rec_desc, b_out, b_err;
    HBuf hb_rec_begin, hb_out, hb_err;
    bool b_resident = false;
    uint64_t b_in_bytes = 0, b_name_bytes = 0, b_aux_bytes = 0;
    uint64_t b_total = 0;             // bytes of the records assembled last (they start kBamFront bytes into b_out)
    DBuf z_out, z_prefix;             // ptl_bgzf_store_records / ptl_frame_records
    HBuf hz_out;
    cudaStream_t b_stream = nullptr;  // the meta kernel of the record assembly runs beside the streaming kernel
    cudaEvent_t b_fork = nullptr, b_join = nullptr;
    // results: one compact arena on the device (device_types.hpp: result_layout) and its pinned host twin
    DBuf r_arena;
    HBuf h_arena;
    uint64_t arena_cap = 0;     // bytes usable by the kernels (<= both buffers)
    uint64_t copied_bytes = 0;  // prefix of the arena already enqueued for D2H behind the kernels
    double est_rec_per_read = 0, est_out_per_in = 0;  // shape of the previous batch on this slot (sizes the speculative copy)
    DevBatch B;
    DevWork W;
    uint32_t stage_mask = PTL_STAGE_ALL;
    bool uploaded = false, ran = false, with_results = true;
    uint64_t n_cigar_in = 0;
    ptl_result res{};
    DevTotals totals{};
};

}  // namespace

struct ptl_ctx {
    int device = 0;
    std::string err;
    std::vector<Slot> slots;
    cudaStream_t setup_stream = nullptr;
    std::atomic<uint64_t> launches{0};  // bumped from any slot's driving thread
    bool zero_copy_seq = false;
    uint32_t long_pair_ops = 64;  // ptl_set_long_pair_ops
    // static state
    DBuf s_ref, s_chrom_off, s_contig_seg_begin, s_contig_len, s_contig_rev_off, s_rev_pool, s_so_start, s_so_end, s_chrom, s_pos,
        s_is_fwd, s_mapq, s_cigar_begin, s_cigar, s_tab_begin, s_table, s_contig_name_off, s_contig_names, s_chrom_name_off, s_chrom_names,
        s_bgzf_tables;
    uint32_t n_contig_names = 0, n_chrom_names = 0;  // ptl_set_names
    bool have_names = false;
    DevStatic S;
    bool have_reference = false, have_segments = false;
    FlatContigs flat;
    std::vector<uint32_t> h_tab_begin;
};

namespace {

// Launch wrappers count into a plain local; the total is added to the ctx counter atomically when the scope ends
// (distinct slots may be driven from distinct host threads).
struct LaunchTally {
    std::atomic<uint64_t>& total;
    uint64_t n = 0;
    explicit LaunchTally(ptl_ctx* ctx) : total(ctx->launches) {}
    ~LaunchTally() { total.fetch_add(n, std::memory_order_relaxed); }
    operator uint64_t*() { return &n; }
};

int fail(ptl_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

template <class F>
int guarded(ptl_ctx* ctx, F f) {
    try {
        if (ctx) CK(cudaSetDevice(ctx->device));
        return f();
    } catch (const CudaError& e) {
        return fail(ctx, PTL_ERR_CUDA, e.msg);
    } catch (const InputError& e) {
        return fail(ctx, PTL_ERR_INPUT, e.what());
    } catch (const std::exception& e) {
        return fail(ctx, PTL_ERR_INVALID_ARG, e.what());
    }
}

template <class T>
const T* upload(DBuf& b, const T* src, size_t n, cudaStream_t st) {
    b.ensure(std::max<size_t>(n, 1) * sizeof(T), st);
    if (n) CK(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
    return b.as<T>();
}

void install_segments(ptl_ctx* ctx, std::vector<HostContig>& contigs) {
    cudaStream_t st = ctx->setup_stream;
    ctx->flat.build(contigs);
    const FlatContigs& f = ctx->flat;
    const uint32_t nc = uint32_t(f.contig_len.size()), ns = uint32_t(f.so_start.size());
    for (uint32_t g = 0; g < ns; ++g) {
        if (f.chrom[g] < 0 || (ctx->have_reference && uint32_t(f.chrom[g]) >= ctx->S.n_chrom))
            throw InputError("contig segment with a chromosome index outside the reference");
        if (f.pos[g] < 0 || f.pos[g] > 0x7fffffffLL) throw InputError("contig segment position outside the BAM int32 range");
    }
    for (uint32_t c = 0; c < nc; ++c)
        if (f.contig_len[c] > 0x7fffffffULL) throw InputError("contig longer than the BAM int32 sequence length limit");
    DevStatic& S = ctx->S;
    S.n_contigs = nc;
    S.n_segments = ns;
    S.contig_seg_begin = upload(ctx->s_contig_seg_begin, f.seg_begin.data(), nc + 1, st);
    S.contig_len = upload(ctx->s_contig_len, f.contig_len.data(), nc, st);
    // rev_contig_seq pool
    std::vector<uint64_t> rev_off(nc, ~0ull);
    uint64_t pool = 0;
    for (uint32_t c = 0; c < nc; ++c)
        if (f.has_rev[c]) { rev_off[c] = pool; pool += (f.contig_len[c] + 15) & ~15ull; }
    ctx->s_rev_pool.ensure(std::max<uint64_t>(pool, 16), st);
    for (uint32_t c = 0; c < nc; ++c)
        if (f.has_rev[c] && f.contig_len[c])
            CK(cudaMemcpyAsync(ctx->s_rev_pool.as<uint8_t>() + rev_off[c], f.rev_seq[c].data(), f.contig_len[c], cudaMemcpyHostToDevice, st));
    S.rev_pool = ctx->s_rev_pool.as<uint8_t>();
    S.contig_rev_off = upload(ctx->s_contig_rev_off, rev_off.data(), nc, st);
    S.seg_so_start = upload(ctx->s_so_start, f.so_start.data(), ns, st);
    S.seg_so_end = upload(ctx->s_so_end, f.so_end.data(), ns, st);
    S.seg_chrom = upload(ctx->s_chrom, f.chrom.data(), ns, st);
    S.seg_pos = upload(ctx->s_pos, f.pos.data(), ns, st);
    S.seg_is_fwd = upload(ctx->s_is_fwd, f.is_fwd.data(), ns, st);
    S.seg_mapq = upload(ctx->s_mapq, f.mapq.data(), ns, st);
    S.seg_cigar_begin = upload(ctx->s_cigar_begin, f.cigar_begin.data(), ns + 1, st);
    S.seg_cigar = upload(ctx->s_cigar, f.cigar.data(), f.cigar.size(), st);
    // device tables: count -> scan -> fill
    ctx->s_tab_begin.ensure((size_t(ns) + 1) * 4, st);
    CK(cudaMemsetAsync(ctx->s_tab_begin.p, 0, (size_t(ns) + 1) * 4, st));
    S.seg_tab_begin = ctx->s_tab_begin.as<uint32_t>();
    S.table = nullptr;
    ctx->h_tab_begin.assign(size_t(ns) + 1, 0);
    if (ns) {
        launch_table_build(S, ctx->s_tab_begin.as<uint32_t>(), nullptr, st);
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        DBuf tmp;
        tmp.ensure(scan_tmp_bytes(ns + 1), st);
        exclusive_scan_inplace<uint32_t>(ctx->s_tab_begin.as<uint32_t>(), uint64_t(ns) + 1, tmp.p, tmp.cap, st, LaunchTally(ctx));
        CK(cudaMemcpyAsync(ctx->h_tab_begin.data(), ctx->s_tab_begin.p, (size_t(ns) + 1) * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        tmp.release();
        const uint32_t total = ctx->h_tab_begin[ns];
        ctx->s_table.ensure(std::max<size_t>(total, 1) * sizeof(TabEntry), st);
        S.table = ctx->s_table.as<TabEntry>();
        launch_table_build(S, nullptr, ctx->s_table.as<TabEntry>(), st);
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    ctx->have_segments = true;
}

void upload_batch(ptl_ctx* ctx, Slot& sl, const ptl_batch* b) {
    cudaStream_t st = sl.stream;
    const uint32_t n = b->n_reads, ns = b->n_read_segments;
    // O(1) checks here; per-segment index validation happens on the device (pair_count_body -> OVF_INVALID), so that
    // submit costs the host no pass over the batch
    if (n && (b->read_seg_begin[0] != 0 || b->read_seg_begin[n] != ns)) throw std::runtime_error("read_seg_begin must run from 0 to n_read_segments");
    DevBatch& B = sl.B;
    B.n_reads = n;
    B.n_rsegs = ns;
    B.n_cigar = b->n_cigar;
    // The packer (ptl_pack_batch) lays every small array out in ONE contiguous arena: move it with a single DMA instead
    // of twelve (31 -> ~50 GB/s effective on the 136 MB of a 1 M-read batch).  Arbitrary caller layouts take the
    // per-array path.
    const bool have_win = b->indel_win != nullptr && b->rseg_win_begin != nullptr;
    if (have_win && ns && b->rseg_win_begin[ns] > b->n_indel_win) throw std::runtime_error("rseg_win_begin outside the indel window pool");
    struct Arr { const void* p; size_t bytes; };
    const Arr arrs[14] = {{b->rseg_win_begin, have_win ? (size_t(ns) + 1) * 4 : 0}, {b->indel_win, have_win ? size_t(b->n_indel_win) * 8 : 0},{b->read_flag, size_t(n) * 2}, {b->read_mapq, size_t(n)}, {b->read_bin, size_t(n) * 2}, {b->read_seq_len, size_t(n) * 4},
                          {b->read_seq_off, size_t(n) * 8}, {b->read_seg_begin, (size_t(n) + 1) * 4}, {b->rseg_contig, size_t(ns) * 4},
                          {b->rseg_pos, size_t(ns) * 8}, {b->rseg_is_fwd, size_t(ns)}, {b->rseg_cigar_begin, size_t(ns) * 8},
                          {b->rseg_cigar_len, size_t(ns) * 4}, {b->cigar, size_t(b->n_cigar) * 4}};
    uintptr_t lo = UINTPTR_MAX, hi = 0;
    size_t sum = 0;
    for (const Arr& a : arrs) {
        if (!a.bytes) continue;
        lo = std::min(lo, reinterpret_cast<uintptr_t>(a.p));
        hi = std::max(hi, reinterpret_cast<uintptr_t>(a.p) + a.bytes);
        sum += a.bytes;
    }
    const bool one_arena = n > 0 && hi > lo && (hi - lo) <= sum + 16 * 256 && (lo % 8) == 0;
    if (one_arena) {
        sl.d_arena.ensure(hi - lo, st);
        CK(cudaMemcpyAsync(sl.d_arena.p, reinterpret_cast<const void*>(lo), hi - lo, cudaMemcpyHostToDevice, st));
        const char* dbase = sl.d_arena.as<char>();
        auto at = [&](const void* hp) { return dbase + (reinterpret_cast<uintptr_t>(hp) - lo); };
        B.read_flag = reinterpret_cast<const uint16_t*>(at(b->read_flag));
        B.read_mapq = reinterpret_cast<const uint8_t*>(at(b->read_mapq));
        B.read_bin = reinterpret_cast<const uint16_t*>(at(b->read_bin));
        B.read_seq_len = reinterpret_cast<const uint32_t*>(at(b->read_seq_len));
        B.read_seq_off = reinterpret_cast<const uint64_t*>(at(b->read_seq_off));
        B.read_seg_begin = reinterpret_cast<const uint32_t*>(at(b->read_seg_begin));
        B.rseg_contig = reinterpret_cast<const uint32_t*>(at(b->rseg_contig));
        B.rseg_pos = reinterpret_cast<const int64_t*>(at(b->rseg_pos));
        B.rseg_is_fwd = reinterpret_cast<const uint8_t*>(at(b->rseg_is_fwd));
        B.rseg_cigar_begin = reinterpret_cast<const uint64_t*>(at(b->rseg_cigar_begin));
        B.rseg_cigar_len = reinterpret_cast<const uint32_t*>(at(b->rseg_cigar_len));
        B.cigar = reinterpret_cast<const uint32_t*>(at(b->cigar));
        B.rseg_win_begin = have_win ? reinterpret_cast<const uint32_t*>(at(b->rseg_win_begin)) : nullptr;
        B.indel_win = have_win ? reinterpret_cast<const uint64_t*>(at(b->indel_win)) : nullptr;
    } else {
        B.read_flag = upload(sl.d_read_flag, b->read_flag, n, st);
        B.read_mapq = upload(sl.d_read_mapq, b->read_mapq, n, st);
        B.read_bin = upload(sl.d_read_bin, b->read_bin, n, st);
        B.read_seq_len = upload(sl.d_read_seq_len, b->read_seq_len, n, st);
        B.read_seq_off = upload(sl.d_read_seq_off, b->read_seq_off, n, st);
        B.read_seg_begin = upload(sl.d_read_seg_begin, b->read_seg_begin, size_t(n) + 1, st);
        B.rseg_contig = upload(sl.d_rseg_contig, b->rseg_contig, ns, st);
        B.rseg_pos = upload(sl.d_rseg_pos, b->rseg_pos, ns, st);
        B.rseg_is_fwd = upload(sl.d_rseg_is_fwd, b->rseg_is_fwd, ns, st);
        B.rseg_cigar_begin = upload(sl.d_rseg_cigar_begin, b->rseg_cigar_begin, ns, st);
        B.rseg_cigar_len = upload(sl.d_rseg_cigar_len, b->rseg_cigar_len, ns, st);
        B.cigar = upload(sl.d_cigar, b->cigar, b->n_cigar, st);
        B.rseg_win_begin = have_win ? upload(sl.d_win_begin, b->rseg_win_begin, size_t(ns) + 1, st) : nullptr;
        B.indel_win = have_win ? upload(sl.d_win, b->indel_win, size_t(b->n_indel_win), st) : nullptr;
    }
    if (ctx->zero_copy_seq) {
        void* dptr = nullptr;
        CK(cudaHostGetDevicePointer(&dptr, const_cast<uint8_t*>(b->seq4), 0));  // must come from ptl_host_alloc
        B.seq4 = static_cast<const uint8_t*>(dptr);
    } else {
        B.seq4 = upload(sl.d_seq4, b->seq4, b->seq4_bytes, st);
    }
    B.seq4_bytes = b->seq4_bytes;
    sl.n_cigar_in = b->n_cigar;
    sl.uploaded = true;
    sl.ran = false;
    // names / aux / qualities / assembled records on the slot belong to the previous batch
    sl.b_resident = false;
    sl.b_total = 0;
    sl.a_qual_n_reads = 0;
}

void size_work(Slot& sl, uint32_t want_pairs, uint64_t want_scratch, uint64_t want_arena) {
    cudaStream_t st = sl.stream;
    const uint32_t n = sl.B.n_reads, ns = sl.B.n_rsegs;
    DevWork& W = sl.W;
    sl.w_rseg_read.ensure(size_t(ns) * 4 + 4, st);
    sl.w_rseg_pair_begin.ensure((size_t(ns) + 1) * 4, st);
    sl.w_rseg_ref_len.ensure(size_t(ns) * 8 + 8, st);
    sl.w_rseg_n_id.ensure(size_t(ns) * 4 + 4, st);
    sl.w_rseg_read_len.ensure(size_t(ns) * 4 + 4, st);
    W.rseg_read = sl.w_rseg_read.as<uint32_t>();
    W.rseg_pair_begin = sl.w_rseg_pair_begin.as<uint32_t>();
    W.rseg_ref_len = sl.w_rseg_ref_len.as<int64_t>();
    W.rseg_n_id = sl.w_rseg_n_id.as<uint32_t>();
    W.rseg_read_len = sl.w_rseg_read_len.as<uint32_t>();
    const uint32_t pc = std::max(W.pair_cap, want_pairs);
    sl.w_pair_rseg.ensure(size_t(pc) * 4, st);
    sl.w_pair_seg.ensure(size_t(pc) * 4, st);
    sl.w_pair_slot_begin.ensure((size_t(pc) + 1) * 8, st);
    sl.w_pair_cap_b.ensure(size_t(pc) * 4, st);
    sl.w_pair_tab_lo.ensure(size_t(pc) * 4, st);
    sl.w_pair_status.ensure(size_t(pc), st);
    sl.w_pair_flip.ensure(size_t(pc), st);
    sl.w_pair_pos.ensure(size_t(pc) * 8, st);
    sl.w_pair_n_out.ensure(size_t(pc) * 4, st);
    sl.w_pair_bin.ensure(size_t(pc) * 2, st);
    sl.w_pair_out_off.ensure(size_t(pc) * 8, st);
    sl.w_simplify_list.ensure(size_t(pc) * 4, st);
    sl.w_long_list.ensure(size_t(pc) * 4, st);
    sl.w_pair_order.ensure(size_t(pc) * 4, st);
    sl.w_pair_desc.ensure(size_t(pc) * sizeof(PairDesc), st);
    sl.w_pair_desc_rev.ensure(size_t(pc) * sizeof(PairDescRev), st);
    W.pair_cap = pc;
    W.pair_rseg = sl.w_pair_rseg.as<uint32_t>();
    W.pair_seg = sl.w_pair_seg.as<uint32_t>();
    W.pair_slot_begin = sl.w_pair_slot_begin.as<uint64_t>();
    W.pair_cap_b = sl.w_pair_cap_b.as<uint32_t>();
    W.pair_tab_lo = sl.w_pair_tab_lo.as<uint32_t>();
    W.pair_status = sl.w_pair_status.as<int8_t>();
    W.pair_flip = sl.w_pair_flip.as<uint8_t>();
    W.pair_pos = sl.w_pair_pos.as<int64_t>();
    W.pair_n_out = sl.w_pair_n_out.as<uint32_t>();
    W.pair_bin = sl.w_pair_bin.as<uint16_t>();
    W.pair_out_off = sl.w_pair_out_off.as<uint64_t>();
    W.simplify_list = sl.w_simplify_list.as<uint32_t>();
    W.long_list = sl.w_long_list.as<uint32_t>();
    W.pair_order = sl.w_pair_order.as<uint32_t>();
    W.pair_desc = sl.w_pair_desc.as<PairDesc>();
    W.pair_desc_rev = sl.w_pair_desc_rev.as<PairDescRev>();
    const uint64_t sc = std::max(W.scratch_cap, want_scratch);
    const uint64_t dense = (uint64_t(pc) + 31) / 32 * kLiftTileOut;  // dense output regions of the lift kernel's tiles
    sl.w_scratch.ensure(size_t(sc + dense) * 4, st);
    W.scratch_cap = sc;
    W.dense_off = sc;
    W.scratch = sl.w_scratch.as<uint32_t>();
    sl.w_read_counts.ensure((size_t(n) + 1) * sizeof(uint2), st);
    sl.w_read_primary.ensure(size_t(n) * 4 + 4, st);
    W.read_counts = sl.w_read_counts.as<uint2>();
    W.read_primary = sl.w_read_primary.as<uint32_t>();
    const uint64_t biggest = std::max<uint64_t>({uint64_t(ns) + 1, uint64_t(pc) + 1, uint64_t(n) + 1});
    sl.w_scan_tmp.ensure(scan_tmp_bytes(biggest), st);
    if (!sl.w_totals.p) {
        sl.w_totals.ensure(sizeof(DevTotals), st);
        CK(cudaMemsetAsync(sl.w_totals.p, 0, sl.w_totals.cap, st));  // padding bytes too (they travel in the D2H of the totals)
    }
    const uint64_t ac = std::max<uint64_t>({sl.arena_cap, want_arena, kResultHeaderBytes});
    sl.r_arena.ensure(ac, st);
    sl.h_arena.ensure(ac);
    sl.arena_cap = ac;
}

// Bytes of the arena a batch of this shape is expected to fill (from the previous batch on the slot), or the whole
// capacity when nothing is known yet.
uint64_t predicted_result_bytes(const Slot& sl) {
    if (sl.est_rec_per_read <= 0) return sl.arena_cap;
    const uint64_t n = sl.B.n_reads;
    const uint64_t nr = uint64_t(double(n) * sl.est_rec_per_read * 1.02) + 256, nc = uint64_t(double(sl.n_cigar_in) * sl.est_out_per_in * 1.02) + 4096;
    return std::min<uint64_t>(sl.arena_cap, result_layout(n, nr, nc).total);
}

// Kernels + the D2H copy of the results behind them on the slot's stream.  `with_results`: copy the predicted extent of
// the arena right away (ptl_lift_submit: one DMA, no host round trip in between); otherwise only the 128-byte header
// (ptl_lift_run: the device-resident path).
void enqueue_run(ptl_ctx* ctx, Slot& sl, bool with_results) {
    launch_lift(ctx->S, sl.B, sl.W, sl.r_arena.as<char>(), sl.arena_cap, sl.w_totals.as<DevTotals>(), sl.stage_mask, sl.w_scan_tmp.p,
                sl.w_scan_tmp.cap, sl.stream, LaunchTally(ctx), sl.have_events ? &sl.ev : nullptr);
    sl.copied_bytes = with_results ? predicted_result_bytes(sl) : kResultHeaderBytes;
    CK(cudaMemcpyAsync(sl.h_arena.p, sl.r_arena.p, sl.copied_bytes, cudaMemcpyDeviceToHost, sl.stream));
    sl.with_results = with_results;
    sl.ran = true;
}

void run_batch(ptl_ctx* ctx, Slot& sl, uint32_t stage_mask, bool with_results) {
    if (!ctx->have_segments) throw std::runtime_error("ptl_set_contig_segments / ptl_set_contig_records has not been called");
    if ((stage_mask & PTL_STAGE_SIMPLIFY) && !ctx->have_reference) throw std::runtime_error("ptl_set_reference has not been called");
    sl.stage_mask = stage_mask;
    // optimistic capacities (no host round trip before the kernels); overflow is detected on the device and the
    // batch is re-run with exact sizes (finish_batch).  Steady-state batches of similar shape never re-run.
    const uint32_t ns = sl.B.n_rsegs, n = sl.B.n_reads;
    const uint32_t want_pairs = ns + ns / 8 + 1024;
    const uint64_t want_scratch = 8ull * sl.n_cigar_in + 64ull * want_pairs;
    const uint64_t want_recs = uint64_t(want_pairs) + n;
    const uint64_t want_cigar = sl.n_cigar_in + sl.n_cigar_in / 2 + 16ull * n + 1024;
    sl.W.long_ops = ctx->long_pair_ops;
    size_work(sl, want_pairs, want_scratch, result_layout(n, want_recs, want_cigar).total);
    enqueue_run(ctx, sl, with_results);
}

// Wait for the kernels, re-run on capacity overflow, leave totals in sl.totals.
void finish_batch(ptl_ctx* ctx, Slot& sl) {
    for (int attempt = 0; attempt < 6; ++attempt) {
        CK(cudaStreamSynchronize(sl.stream));
        CK(cudaGetLastError());
        std::memcpy(&sl.totals, sl.h_arena.p, sizeof(DevTotals));
        const DevTotals& t = sl.totals;
        if (t.overflow & OVF_INVALID)
            throw std::runtime_error("malformed batch: a read segment's contig index, CIGAR range, position (BAM int32) or base range lies outside its pool");
        if (!t.overflow) {
            if (sl.B.n_reads) {
                sl.est_rec_per_read = double(t.n_records) / double(sl.B.n_reads);
                sl.est_out_per_in = double(t.n_cigar_out) / double(std::max<uint64_t>(sl.n_cigar_in, 1));
            }
            return;
        }
        const uint32_t want_pairs = uint32_t(std::max<uint64_t>(t.n_pairs, sl.W.pair_cap));
        uint64_t want_scratch = sl.W.scratch_cap, want_arena = sl.arena_cap;
        if (!(t.overflow & OVF_PAIRS)) {
            want_scratch = std::max<uint64_t>(want_scratch, t.scratch_needed);
            if (!(t.overflow & OVF_SCRATCH)) want_arena = std::max<uint64_t>(want_arena, result_layout(sl.B.n_reads, t.n_records, t.n_cigar_out).total);
        }
        want_arena = std::max<uint64_t>(want_arena, result_layout(sl.B.n_reads, uint64_t(want_pairs) + sl.B.n_reads, 0).total);
        size_work(sl, want_pairs, want_scratch, want_arena);
        enqueue_run(ctx, sl, sl.with_results);
    }
    throw std::runtime_error("work buffers still overflow after resizing (internal error)");
}

int download(ptl_ctx* ctx, Slot& sl, ptl_result* out) {
    finish_batch(ctx, sl);
    const DevTotals& t = sl.totals;
    const size_t n = sl.B.n_reads;
    const ResultLayout L = result_layout(n, t.n_records, t.n_cigar_out);
    if (n && L.total > sl.copied_bytes) {  // the speculative copy was short (first batch of a new shape, or ptl_lift_run): fetch the rest
        CK(cudaMemcpyAsync(sl.h_arena.as<char>() + sl.copied_bytes, sl.r_arena.as<char>() + sl.copied_bytes, L.total - sl.copied_bytes,
                           cudaMemcpyDeviceToHost, sl.stream));
        CK(cudaStreamSynchronize(sl.stream));
        sl.copied_bytes = L.total;
    }
    ptl_result& r = sl.res;
    r = ptl_result{};
    r.n_reads = uint32_t(n);
    r.n_records = uint32_t(t.n_records);
    r.n_cigar = t.n_cigar_out;
    if (n) {
        const DevResult R = DevResult::view(sl.h_arena.as<char>(), L);
        r.read_rec_begin = R.read_rec_begin;
        r.rec_status = R.rec_status;
        r.rec_read_segment = R.rec_read_segment;
        r.rec_contig_segment = R.rec_contig_segment;
        r.rec_tid = R.rec_tid;
        r.rec_pos = R.rec_pos;
        r.rec_mapq = R.rec_mapq;
        r.rec_flag = R.rec_flag;
        r.rec_bin = R.rec_bin;
        r.rec_need_flip = R.rec_need_flip;
        r.rec_cigar_begin = R.rec_cigar_begin;
        r.cigar = R.cigar;
    } else {
        // an empty batch: the kernels only wrote the header; lend out the two one-element CSR arrays from behind it
        const ResultLayout L0 = result_layout(0, 0, 0);
        sl.h_arena.ensure(L0.total);
        const DevResult R = DevResult::view(sl.h_arena.as<char>(), L0);
        R.read_rec_begin[0] = 0;
        R.rec_cigar_begin[0] = 0;
        r.read_rec_begin = R.read_rec_begin;
        r.rec_cigar_begin = R.rec_cigar_begin;
    }
    r.n_pairs = t.n_pairs;
    r.n_lifted = t.n_lifted;
    r.n_errors = t.n_errors;
    if (t.n_errors) {
        r.first_error_read = t.first_error_read >> 8;
        r.first_error_status = int32_t(int8_t(t.first_error_read & 0xff));
    } else {
        r.first_error_read = -1;
        r.first_error_status = 0;
    }
    *out = r;
    if (t.n_errors)
        return fail(ctx, PTL_ERR_LIFT_PANIC,
                    "the reference would panic on read " + std::to_string(r.first_error_read) + " (status " + std::to_string(r.first_error_status) + ")");
    return PTL_OK;
}

Slot* get_slot(ptl_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || size_t(slot) >= ctx->slots.size()) return nullptr;
    return &ctx->slots[size_t(slot)];
}

}  // namespace

extern "C" {

const char* ptl_version(void) { return "portello_b200 0.1 (sm_100a CUDA kernels, C-ABI 1)"; }

int ptl_create(int device, int n_slots, ptl_ctx** out) {
    if (!out || n_slots < 1 || n_slots > 64) return PTL_ERR_INVALID_ARG;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
        cudaGetLastError();
        return PTL_ERR_NO_DEVICE;  // no CPU fallback by design
    }
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
        cudaGetLastError();
        return PTL_ERR_NO_DEVICE;  // the kernels are compiled for sm_100a only
    }
    auto* ctx = new ptl_ctx();
    ctx->device = device;
    const int rc = guarded(ctx, [&]() {
        CK(cudaStreamCreateWithFlags(&ctx->setup_stream, cudaStreamNonBlocking));
        ctx->slots.resize(size_t(n_slots));
        for (auto& sl : ctx->slots) {
            CK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
            for (auto& e : sl.ev.e) CK(cudaEventCreate(&e));
            sl.have_events = true;
        }
        return PTL_OK;
    });
    if (rc != PTL_OK) {
        std::fprintf(stderr, "ptl_create: %s\n", ctx->err.c_str());
        delete ctx;
        return rc;
    }
    *out = ctx;
    return PTL_OK;
}

void ptl_destroy(ptl_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (auto& sl : ctx->slots) {
        if (sl.stream) cudaStreamSynchronize(sl.stream);
        for (DBuf* b : {&sl.d_read_flag, &sl.d_read_mapq, &sl.d_read_bin, &sl.d_read_seq_len, &sl.d_read_seq_off, &sl.d_read_seg_begin,
                        &sl.d_rseg_contig, &sl.d_rseg_pos, &sl.d_rseg_is_fwd, &sl.d_rseg_cigar_begin, &sl.d_rseg_cigar_len, &sl.d_cigar,
                        &sl.d_seq4, &sl.d_arena, &sl.d_win_begin, &sl.d_win, &sl.w_rseg_read, &sl.w_rseg_pair_begin, &sl.w_rseg_ref_len, &sl.w_rseg_n_id, &sl.w_rseg_read_len, &sl.w_pair_cap_b, &sl.w_pair_tab_lo, &sl.w_pair_rseg, &sl.w_pair_seg,
                        &sl.w_pair_slot_begin, &sl.w_pair_status, &sl.w_pair_flip, &sl.w_pair_pos, &sl.w_pair_n_out, &sl.w_pair_bin, &sl.w_pair_out_off, &sl.w_simplify_list, &sl.w_long_list, &sl.w_pair_order, &sl.w_pair_desc, &sl.w_pair_desc_rev,
                        &sl.w_scratch, &sl.w_read_counts, &sl.w_read_primary, &sl.w_scan_tmp, &sl.w_totals, &sl.r_arena, &sl.a_qual, &sl.a_qual_off,
                        &sl.a_rec_read, &sl.a_seq_begin, &sl.a_qual_begin, &sl.a_out_seq, &sl.a_out_qual, &sl.b_name_off, &sl.b_names, &sl.b_aux_off,
                        &sl.b_aux, &sl.b_mate_tid, &sl.b_mate_pos, &sl.b_tlen, &sl.b_keep, &sl.b_sa_len, &sl.b_rec_begin, &sl.b_rec_desc, &sl.b_out, &sl.b_err, &sl.z_out, &sl.z_prefix})
            b->release();
        sl.h_arena.release();
        for (HBuf* b : {&sl.ha_seq_begin, &sl.ha_qual_begin, &sl.ha_out_seq, &sl.ha_out_qual, &sl.hb_rec_begin, &sl.hb_out, &sl.hb_err, &sl.hz_out}) b->release();
        for (auto& e : sl.a_ev) if (e) cudaEventDestroy(e);
        if (sl.b_fork) cudaEventDestroy(sl.b_fork);
        if (sl.b_join) cudaEventDestroy(sl.b_join);
        if (sl.b_stream) cudaStreamDestroy(sl.b_stream);
        if (sl.have_events)
            for (auto& e : sl.ev.e) cudaEventDestroy(e);
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    for (DBuf* b : {&ctx->s_ref, &ctx->s_chrom_off, &ctx->s_contig_seg_begin, &ctx->s_contig_len, &ctx->s_contig_rev_off, &ctx->s_rev_pool,
                    &ctx->s_so_start, &ctx->s_so_end, &ctx->s_chrom, &ctx->s_pos, &ctx->s_is_fwd, &ctx->s_mapq, &ctx->s_cigar_begin,
                    &ctx->s_cigar, &ctx->s_tab_begin, &ctx->s_table, &ctx->s_contig_name_off, &ctx->s_contig_names, &ctx->s_chrom_name_off,
                    &ctx->s_chrom_names, &ctx->s_bgzf_tables})
        b->release();
    if (ctx->setup_stream) cudaStreamDestroy(ctx->setup_stream);
    delete ctx;
}

const char* ptl_last_error(const ptl_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }

int ptl_set_reference(ptl_ctx* ctx, uint32_t n_chrom, const uint64_t* chrom_len, const uint8_t* const* chrom_seq) {
    if (!ctx || (n_chrom && (!chrom_len || !chrom_seq))) return PTL_ERR_INVALID_ARG;
    return guarded(ctx, [&]() {
        cudaStream_t st = ctx->setup_stream;
        // (segments may be installed first: they must fit the reference whichever call comes second)
        if (ctx->have_segments)
            for (int32_t c : ctx->flat.chrom)
                if (c < 0 || uint32_t(c) >= n_chrom) throw InputError("the reference has fewer chromosomes than the installed contig segments use");
        // chromosomes are packed back to back (byte loads need no alignment); chrom_off[c+1]-chrom_off[c] is the exact length
        std::vector<uint64_t> off(size_t(n_chrom) + 1, 0);
        for (uint32_t c = 0; c < n_chrom; ++c) off[c + 1] = off[c] + chrom_len[c];
        ctx->s_ref.ensure(std::max<uint64_t>(off[n_chrom], 16), st);
        for (uint32_t c = 0; c < n_chrom; ++c)
            if (chrom_len[c]) CK(cudaMemcpyAsync(ctx->s_ref.as<uint8_t>() + off[c], chrom_seq[c], chrom_len[c], cudaMemcpyHostToDevice, st));
        ctx->S.ref = ctx->s_ref.as<uint8_t>();
        ctx->S.chrom_off = upload(ctx->s_chrom_off, off.data(), size_t(n_chrom) + 1, st);
        ctx->S.n_chrom = n_chrom;
        CK(cudaStreamSynchronize(st));
        ctx->have_reference = true;
        return PTL_OK;
    });
}

int ptl_set_contig_segments(ptl_ctx* ctx, const ptl_contig_segments* segs) {
    if (!ctx || !segs) return PTL_ERR_INVALID_ARG;
    return guarded(ctx, [&]() {
        auto contigs = contigs_from_flat(*segs);
        install_segments(ctx, contigs);
        return PTL_OK;
    });
}
int ptl_set_raw_contig_segments(ptl_ctx* ctx, const ptl_contig_segments* raw) {
    if (!ctx || !raw) return PTL_ERR_INVALID_ARG;
    return guarded(ctx, [&]() {
        auto contigs = contigs_from_flat(*raw);
        trim_repeated_matches(contigs);
        join_colinear(contigs);
        install_segments(ctx, contigs);
        return PTL_OK;
    });
}
int ptl_set_contig_records(ptl_ctx* ctx, const ptl_contig_records* recs) {
    if (!ctx || !recs) return PTL_ERR_INVALID_ARG;
    return guarded(ctx, [&]() {
        auto contigs = assemble_from_records(*recs);
        trim_repeated_matches(contigs);
        join_colinear(contigs);
        install_segments(ctx, contigs);
        return PTL_OK;
    });
}
int ptl_get_contig_segments(const ptl_ctx* ctx, ptl_contig_segments* out) {
    if (!ctx || !out || !ctx->have_segments) return PTL_ERR_INVALID_ARG;
    ctx->flat.view(out);
    return PTL_OK;
}
int ptl_get_segment_table(ptl_ctx* ctx, uint32_t segment, uint32_t cap, uint32_t* keys, int32_t* vals, uint32_t* n) {
    if (!ctx || !n || !ctx->have_segments || segment >= ctx->S.n_segments) return PTL_ERR_INVALID_ARG;
    return guarded(ctx, [&]() {
        const uint32_t t0 = ctx->h_tab_begin[segment], t1 = ctx->h_tab_begin[segment + 1];
        *n = t1 - t0;
        if (*n > cap) return int(PTL_ERR_INVALID_ARG);
        std::vector<TabEntry> tmp(*n);
        if (*n) CK(cudaMemcpy(tmp.data(), ctx->s_table.as<TabEntry>() + t0, size_t(*n) * sizeof(TabEntry), cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < *n; ++i) {
            keys[i] = tmp[i].key;
            vals[i] = tmp[i].val;
        }
        return int(PTL_OK);
    });
}

int ptl_lift_upload(ptl_ctx* ctx, int slot, const ptl_batch* batch) {
    Slot* sl = get_slot(ctx, slot);
    if (!sl || !batch) return PTL_ERR_INVALID_ARG;
    return guarded(ctx, [&]() {
        if (!ctx->have_segments) throw std::runtime_error("ptl_set_contig_segments / ptl_set_contig_records has not been called");
        upload_batch(ctx, *sl, batch);
        return PTL_OK;
    });
}
int ptl_lift_run(ptl_ctx* ctx, int slot, uint32_t stage_mask) {
    Slot* sl = get_slot(ctx, slot);
    if (!sl) return PTL_ERR_INVALID_ARG;
    if (!sl->uploaded) return fail(ctx, PTL_ERR_STATE, "ptl_lift_run without ptl_lift_upload");
    return guarded(ctx, [&]() {
        run_batch(ctx, *sl, stage_mask, /*with_results=*/false);
        return PTL_OK;
    });
}
int ptl_lift_download(ptl_ctx* ctx, int slot, ptl_result* out) {
    Slot* sl = get_slot(ctx, slot);
    if (!sl || !out) return PTL_ERR_INVALID_ARG;
    if (!sl->ran) return fail(ctx, PTL_ERR_STATE, "ptl_lift_download / ptl_lift_wait without a submitted batch");
    return guarded(ctx, [&]() { return download(ctx, *sl, out); });
}
int ptl_lift_submit_ex(ptl_ctx* ctx, int slot, const ptl_batch* batch, uint32_t stage_mask) {
    Slot* sl = get_slot(ctx, slot);
    if (!sl || !batch) return PTL_ERR_INVALID_ARG;
    return guarded(ctx, [&]() {
        if (!ctx->have_segments) throw std::runtime_error("ptl_set_contig_segments / ptl_set_contig_records has not been called");
        upload_batch(ctx, *sl, batch);
        run_batch(ctx, *sl, stage_mask, /*with_results=*/true);
        return PTL_OK;
    });
}
int ptl_lift_submit(ptl_ctx* ctx, int slot, const ptl_batch* batch) { return ptl_lift_submit_ex(ctx, slot, batch, PTL_STAGE_ALL); }
int ptl_lift_wait(ptl_ctx* ctx, int slot, ptl_result* out) { return ptl_lift_download(ctx, slot, out); }

void* ptl_slot_stream(ptl_ctx* ctx, int slot) {
    Slot* sl = get_slot(ctx, slot);
    return sl ? static_cast<void*>(sl->stream) : nullptr;
}
int ptl_slot_kernel_times(ptl_ctx* ctx, int slot, int cap, const char** names, float* ms) {
    Slot* sl = get_slot(ctx, slot);
    if (!sl || !sl->ran || !sl->have_events) return 0;
    static const char* kNames[] = {"enumerate_pairs", "lift_pairs", "finalize_emit"};
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(sl->stream);
    int n = 0;
    for (int i = 0; i + 1 < StageEvents::N && n < cap; ++i, ++n) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, sl->ev.e[i], sl->ev.e[i + 1]) != cudaSuccess) { cudaGetLastError(); t = -1.f; }
        names[n] = kNames[i];
        ms[n] = t;
    }
    return n;
}
uint64_t ptl_launch_count(const ptl_ctx* ctx) { return ctx ? ctx->launches.load(std::memory_order_relaxed) : 0; }
// Record assembly, bases: see include/portello_b200.h.
int ptl_assemble_bases(ptl_ctx* ctx, int slot, const ptl_read_quals* quals, uint32_t flags, ptl_record_bases* out) {
    Slot* sl = get_slot(ctx, slot);
    if (!sl || !out || (!quals && !(flags & PTL_ASM_RESIDENT_QUAL))) return PTL_ERR_INVALID_ARG;
    if (!sl->ran) return fail(ctx, PTL_ERR_STATE, "ptl_assemble_bases without a lifted batch on the slot");
    return guarded(ctx, [&]() {
        finish_batch(ctx, *sl);
        cudaStream_t st = sl->stream;
        const DevTotals& t = sl->totals;
        const uint32_t n = sl->B.n_reads, n_rec = uint32_t(t.n_records);
        *out = ptl_record_bases{};
        out->n_records = n_rec;
        if (!(flags & PTL_ASM_RESIDENT_QUAL)) {
            sl->a_qual.ensure(std::max<uint64_t>(quals->qual_bytes, 4) + 8, st);
            if (quals->qual_bytes) CK(cudaMemcpyAsync(sl->a_qual.p, quals->qual, quals->qual_bytes, cudaMemcpyHostToDevice, st));
            upload(sl->a_qual_off, quals->read_qual_off, n, st);
            sl->a_qual_bytes = quals->qual_bytes;
            sl->a_qual_n_reads = n;
        } else if (!sl->a_qual.p || sl->a_qual_n_reads != n) {
            throw std::runtime_error("PTL_ASM_RESIDENT_QUAL without a previous upload for this batch on this slot");
        }
        const ResultLayout L = result_layout(n, t.n_records, t.n_cigar_out);
        const DevResult R = DevResult::view(sl->r_arena.as<char>(), L);
        sl->a_rec_read.ensure(size_t(n_rec) * 4 + 4, st);
        sl->a_seq_begin.ensure((size_t(n_rec) + 1) * 8, st);
        sl->a_qual_begin.ensure((size_t(n_rec) + 1) * 8, st);
        sl->w_scan_tmp.ensure(scan_tmp_bytes(uint64_t(n_rec) + 1), st);
        sl->b_err.ensure(4, st);
        CK(cudaMemsetAsync(sl->b_err.p, 0, 4, st));
        launch_assemble_sizes(n_rec, n, R.read_rec_begin, sl->B.read_seq_len, sl->a_qual_off.as<uint64_t>(), sl->a_qual_bytes, sl->a_rec_read.as<uint32_t>(),
                              sl->a_seq_begin.as<uint64_t>(), sl->a_qual_begin.as<uint64_t>(), sl->b_err.as<unsigned int>(), sl->w_scan_tmp.p,
                              sl->w_scan_tmp.cap, st, LaunchTally(ctx));
        sl->ha_seq_begin.ensure((size_t(n_rec) + 1) * 8);
        sl->ha_qual_begin.ensure((size_t(n_rec) + 1) * 8);
        sl->hb_err.ensure(4);
        CK(cudaMemcpyAsync(sl->ha_seq_begin.p, sl->a_seq_begin.p, (size_t(n_rec) + 1) * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(sl->ha_qual_begin.p, sl->a_qual_begin.p, (size_t(n_rec) + 1) * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(sl->hb_err.p, sl->b_err.p, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));  // the output sizes are only known now
        if (*sl->hb_err.as<unsigned int>() & 16u) throw std::runtime_error("ptl_read_quals: a quality range lies outside the pool");
        const uint64_t seq_total = sl->ha_seq_begin.as<uint64_t>()[n_rec], qual_total = sl->ha_qual_begin.as<uint64_t>()[n_rec];
        sl->a_out_seq.ensure(std::max<uint64_t>(seq_total, 16), st);
        sl->a_out_qual.ensure(std::max<uint64_t>(qual_total, 16), st);
        AsmArgs A{};
        A.n_records = n_rec;
        A.rec_read = sl->a_rec_read.as<uint32_t>();
        A.rec_flip = R.rec_need_flip;
        A.read_seq_len = sl->B.read_seq_len;
        A.read_seq_off = sl->B.read_seq_off;
        A.seq4 = sl->B.seq4;
        A.read_qual_off = sl->a_qual_off.as<uint64_t>();
        A.qual = sl->a_qual.as<uint8_t>();
        A.rec_seq_begin = sl->a_seq_begin.as<uint64_t>();
        A.rec_qual_begin = sl->a_qual_begin.as<uint64_t>();
        A.out_seq4 = sl->a_out_seq.as<uint8_t>();
        A.out_qual = sl->a_out_qual.as<uint8_t>();
        if (!sl->a_ev[0]) { CK(cudaEventCreate(&sl->a_ev[0])); CK(cudaEventCreate(&sl->a_ev[1])); }
        CK(cudaEventRecord(sl->a_ev[0], st));
        launch_assemble_records(A, st, LaunchTally(ctx));
        CK(cudaEventRecord(sl->a_ev[1], st));
        if (!(flags & PTL_ASM_NO_DOWNLOAD)) {
            sl->ha_out_seq.ensure(std::max<uint64_t>(seq_total, 4));
            sl->ha_out_qual.ensure(std::max<uint64_t>(qual_total, 4));
            if (seq_total) CK(cudaMemcpyAsync(sl->ha_out_seq.p, sl->a_out_seq.p, seq_total, cudaMemcpyDeviceToHost, st));
            if (qual_total) CK(cudaMemcpyAsync(sl->ha_out_qual.p, sl->a_out_qual.p, qual_total, cudaMemcpyDeviceToHost, st));
            out->seq4 = sl->ha_out_seq.as<uint8_t>();
            out->qual = sl->ha_out_qual.as<uint8_t>();
        }
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&out->kernel_ms, sl->a_ev[0], sl->a_ev[1]));
        out->rec_seq_begin = sl->ha_seq_begin.as<uint64_t>();
        out->rec_qual_begin = sl->ha_qual_begin.as<uint64_t>();
        // algorithmic bytes: every record reads its read's bases + qualities once and writes them once (unrounded)
        uint64_t bytes = 0;
        {
            const uint64_t* sb = out->rec_seq_begin;
            const uint64_t* qb = out->rec_qual_begin;
            bytes = sb[n_rec] + qb[n_rec];  // (rounded up to 16 per record: < 0.1 % for 15 kb reads)
        }
        out->bytes_read = bytes;
        out->bytes_written = bytes;
        return PTL_OK;
    });
}
// Records are assembled kBamFront bytes into their buffer, so that a stream prefix (the BAM header) can sit right in front.
constexpr uint64_t kBamFront = 1ull << 16;

// BGZF framing on the device, level 0: see include/portello_b200.h.
int ptl_bgzf_store_records(ptl_ctx* ctx, int slot, const uint8_t* prefix, uint64_t prefix_bytes, uint32_t flags, ptl_bgzf_stream* out) {
    Slot* sl = get_slot(ctx, slot);
    if (!sl || !out || (prefix_bytes && !prefix) || prefix_bytes > kBamFront / 2) return PTL_ERR_INVALID_ARG;
    if (!sl->b_resident || !sl->b_out.p) return fail(ctx, PTL_ERR_STATE, "ptl_bgzf_store_records without ptl_assemble_records on the slot");
    return guarded(ctx, [&]() {
        cudaStream_t st = sl->stream;
        static const std::vector<uint32_t> tables = bgzf_tables();
        if (!ctx->s_bgzf_tables.p) upload(ctx->s_bgzf_tables, tables.data(), tables.size(), st);
        uint8_t* stream = sl->b_out.as<uint8_t>() + kBamFront - prefix_bytes;
        if (prefix_bytes) CK(cudaMemcpyAsync(stream, prefix, prefix_bytes, cudaMemcpyHostToDevice, st));
        const uint64_t n = prefix_bytes + sl->b_total;
        const uint64_t n_blocks = (n + kBgzfIn - 1) / kBgzfIn;
        const uint64_t framed = n + uint64_t(kBgzfOverhead) * n_blocks;
        static const uint8_t kEof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const uint64_t total = framed + ((flags & PTL_BGZF_EOF) ? sizeof(kEof) : 0);
        sl->z_out.ensure(total + 64, st);
        BgzfArgs A{};
        A.in = stream;
        A.n = n;
        A.out = sl->z_out.as<uint8_t>();
        A.tables = ctx->s_bgzf_tables.as<uint32_t>();
        A.n_blocks = n_blocks;
        A.init_full = crc_zero_bytes(tables.data(), 0xffffffffu, kBgzfIn);
        A.init_last = crc_zero_bytes(tables.data(), 0xffffffffu, n_blocks ? n - (n_blocks - 1) * kBgzfIn : 0);
        if (!sl->a_ev[0]) { CK(cudaEventCreate(&sl->a_ev[0])); CK(cudaEventCreate(&sl->a_ev[1])); }
        CK(cudaEventRecord(sl->a_ev[0], st));
        launch_bgzf_store(A, st, LaunchTally(ctx));
        CK(cudaEventRecord(sl->a_ev[1], st));
        if (flags & PTL_BGZF_EOF) CK(cudaMemcpyAsync(sl->z_out.as<uint8_t>() + framed, kEof, sizeof(kEof), cudaMemcpyHostToDevice, st));
        *out = ptl_bgzf_stream{};
        if (!(flags & PTL_ASM_NO_DOWNLOAD)) {
            sl->hz_out.ensure(total + 64);
            if (total) CK(cudaMemcpyAsync(sl->hz_out.p, sl->z_out.p, total, cudaMemcpyDeviceToHost, st));
            out->bytes = sl->hz_out.as<uint8_t>();
        }
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&out->kernel_ms, sl->a_ev[0], sl->a_ev[1]));
        out->n_bytes = total;
        out->n_blocks = n_blocks;
        out->bytes_read = n;
        out->bytes_written = framed;
        return PTL_OK;
    });
}

// Record assembly, whole BAM records: see include/portello_b200.h.
int ptl_set_names(ptl_ctx* ctx, uint32_t n_contigs, const char* const* contig_names, uint32_t n_chrom, const char* const* chrom_names) {
    if (!ctx || (n_contigs && !contig_names) || (n_chrom && !chrom_names)) return PTL_ERR_INVALID_ARG;
    return guarded(ctx, [&]() {
        cudaStream_t st = ctx->setup_stream;
        auto pool = [&](uint32_t n, const char* const* names, DBuf& d_off, DBuf& d_bytes) {
            std::vector<uint64_t> off(size_t(n) + 1, 0);
            std::string bytes;
            for (uint32_t i = 0; i < n; ++i) {
                if (!names[i]) throw std::runtime_error("ptl_set_names: NULL name");
                bytes += names[i];
                off[i + 1] = bytes.size();
            }
            bytes.resize(bytes.size() + 16, '\0');
            upload(d_off, off.data(), off.size(), st);
            upload(d_bytes, reinterpret_cast<const uint8_t*>(bytes.data()), bytes.size(), st);
            CK(cudaStreamSynchronize(st));  // (the host vectors die here)
        };
        pool(n_contigs, contig_names, ctx->s_contig_name_off, ctx->s_contig_names);
        pool(n_chrom, chrom_names, ctx->s_chrom_name_off, ctx->s_chrom_names);
        ctx->n_contig_names = n_contigs;
        ctx->n_chrom_names = n_chrom;
        ctx->have_names = true;
        return PTL_OK;
    });
}
// ptl_assemble_records (zout == nullptr: the records as a byte stream) and ptl_frame_records (zout != nullptr: the same
// records, never materialised, straight into level-0 BGZF blocks behind `prefix`).
static int assemble_impl(ptl_ctx* ctx, int slot, const ptl_read_extras* x, uint32_t flags, ptl_bam_records* out, const uint8_t* prefix,
                         uint64_t prefix_bytes, ptl_bgzf_stream* zout) {
    Slot* sl = get_slot(ctx, slot);
    ptl_bam_records out_local{};
    if (zout && !out) out = &out_local;
    if (!sl || !out || (!x && !(flags & PTL_ASM_RESIDENT_QUAL)) || (prefix_bytes && !prefix) || prefix_bytes > kBamFront / 2) return PTL_ERR_INVALID_ARG;
    if (!sl->ran) return fail(ctx, PTL_ERR_STATE, "record assembly without a lifted batch on the slot");
    if (!ctx->have_names) return fail(ctx, PTL_ERR_STATE, "ptl_set_names has not been called");
    return guarded(ctx, [&]() {
        finish_batch(ctx, *sl);
        cudaStream_t st = sl->stream;
        const DevTotals& t = sl->totals;
        const uint32_t n = sl->B.n_reads, n_rec = uint32_t(t.n_records);
        *out = ptl_bam_records{};
        out->n_records = n_rec;
        if (!(flags & PTL_ASM_RESIDENT_QUAL)) {
            const uint64_t name_bytes = n ? x->name_off[n] : 0, aux_bytes = n ? x->aux_off[n] : 0;
            upload(sl->b_name_off, x->name_off, size_t(n) + 1, st);
            upload(sl->b_aux_off, x->aux_off, size_t(n) + 1, st);
            sl->b_names.ensure(name_bytes + 32, st);
            sl->b_aux.ensure(aux_bytes + 32, st);
            if (name_bytes) CK(cudaMemcpyAsync(sl->b_names.p, x->names, name_bytes, cudaMemcpyHostToDevice, st));
            if (aux_bytes) CK(cudaMemcpyAsync(sl->b_aux.p, x->aux, aux_bytes, cudaMemcpyHostToDevice, st));
            upload(sl->b_mate_tid, x->mate_tid, n, st);
            upload(sl->b_mate_pos, x->mate_pos, n, st);
            upload(sl->b_tlen, x->tlen, n, st);
            sl->a_qual.ensure(std::max<uint64_t>(x->quals.qual_bytes, 4) + 32, st);
            if (x->quals.qual_bytes) CK(cudaMemcpyAsync(sl->a_qual.p, x->quals.qual, x->quals.qual_bytes, cudaMemcpyHostToDevice, st));
            upload(sl->a_qual_off, x->quals.read_qual_off, n, st);
            sl->a_qual_bytes = x->quals.qual_bytes;
            sl->a_qual_n_reads = n;
            sl->b_resident = true;
            sl->b_in_bytes = name_bytes + aux_bytes;
            sl->b_name_bytes = name_bytes;
            sl->b_aux_bytes = aux_bytes;
        } else if (!sl->b_resident) {
            throw std::runtime_error("PTL_ASM_RESIDENT_QUAL without a previous ptl_assemble_records upload on this slot");
        }
        const ResultLayout L = result_layout(n, t.n_records, t.n_cigar_out);
        const DevResult R = DevResult::view(sl->r_arena.as<char>(), L);
        sl->b_keep.ensure(size_t(n) * 40 + 40, st);
        sl->b_sa_len.ensure(size_t(n_rec) * 4 + 4, st);
        sl->b_rec_begin.ensure((size_t(n_rec) + 1) * 8, st);
        sl->b_rec_desc.ensure(size_t(n_rec) * 32 + 32, st);
        sl->b_err.ensure(4, st);
        CK(cudaMemsetAsync(sl->b_err.p, 0, 4, st));
        sl->w_scan_tmp.ensure(scan_tmp_bytes(uint64_t(n_rec) + 1), st);
        BamAsmArgs A{};
        A.n_reads = n;
        A.n_records = n_rec;
        A.seg_is_fwd = ctx->S.seg_is_fwd;
        A.contig_seg_begin = ctx->S.contig_seg_begin;
        A.contig_name_off = ctx->s_contig_name_off.as<uint64_t>();
        A.contig_names = ctx->s_contig_names.as<uint8_t>();
        A.n_contig_names = ctx->n_contig_names;
        A.chrom_name_off = ctx->s_chrom_name_off.as<uint64_t>();
        A.chrom_names = ctx->s_chrom_names.as<uint8_t>();
        A.n_chrom_names = ctx->n_chrom_names;
        A.read_mapq = sl->B.read_mapq;
        A.read_seq_len = sl->B.read_seq_len;
        A.read_seq_off = sl->B.read_seq_off;
        A.seq4 = sl->B.seq4;
        A.rseg_contig = sl->B.rseg_contig;
        A.rseg_read = sl->W.rseg_read;
        A.name_off = sl->b_name_off.as<uint64_t>();
        A.names = sl->b_names.as<uint8_t>();
        A.aux_off = sl->b_aux_off.as<uint64_t>();
        A.aux = sl->b_aux.as<uint8_t>();
        A.mate_tid = sl->b_mate_tid.as<int32_t>();
        A.mate_pos = sl->b_mate_pos.as<int32_t>();
        A.tlen = sl->b_tlen.as<int32_t>();
        A.qual_off = sl->a_qual_off.as<uint64_t>();
        A.qual = sl->a_qual.as<uint8_t>();
        A.names_bytes = sl->b_name_bytes;
        A.aux_bytes = sl->b_aux_bytes;
        A.qual_bytes = sl->a_qual_bytes;
        A.read_rec_begin = R.read_rec_begin;
        A.rec_status = R.rec_status;
        A.rec_read_segment = R.rec_read_segment;
        A.rec_contig_segment = R.rec_contig_segment;
        A.rec_tid = R.rec_tid;
        A.rec_pos = R.rec_pos;
        A.rec_mapq = R.rec_mapq;
        A.rec_flag = R.rec_flag;
        A.rec_bin = R.rec_bin;
        A.rec_need_flip = R.rec_need_flip;
        A.rec_cigar_begin = R.rec_cigar_begin;
        A.cigar = R.cigar;
        A.read_keep = sl->b_keep.as<uint32_t>();
        A.rec_sa_len = sl->b_sa_len.as<uint32_t>();
        A.rec_begin = sl->b_rec_begin.as<uint64_t>();
        A.rec_desc = sl->b_rec_desc.as<uint4>();
        A.error = sl->b_err.as<unsigned int>();
        launch_bam_sizes(A, sl->w_scan_tmp.p, sl->w_scan_tmp.cap, st, LaunchTally(ctx));
        sl->hb_rec_begin.ensure((size_t(n_rec) + 1) * 8);
        sl->hb_err.ensure(4);
        CK(cudaMemcpyAsync(sl->hb_rec_begin.p, sl->b_rec_begin.p, (size_t(n_rec) + 1) * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(sl->hb_err.p, sl->b_err.p, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));  // the output size is only known now
        const unsigned err = *sl->hb_err.as<unsigned int>();
        if (err & 16u) throw std::runtime_error("ptl_read_extras: a name / aux / quality offset is not monotonic or lies outside its pool");
        if (err & 1u) throw std::runtime_error("ptl_set_names: a contig or reference chromosome of this batch has no name");
        if (err & 2u) throw std::runtime_error("a lifted CIGAR with more than 65535 ops spans 2^28 reference bases or more: BAM cannot hold it (bam_write1 refuses it too)");
        if (err & 4u) throw std::runtime_error("a contig name longer than 248 bytes or more than 16 MB of SA text in one record");
        if (err & 8u) throw std::runtime_error("a read name longer than 254 bytes (BAM l_read_name is a u8)");
        const uint64_t total = sl->hb_rec_begin.as<uint64_t>()[n_rec];
        sl->b_out.ensure(kBamFront + total + 64, st);  // (headroom in front: ptl_bgzf_store_records puts its prefix there)
        A.out = sl->b_out.as<uint8_t>() + kBamFront;
        sl->b_total = total;
        if (!sl->a_ev[0]) { CK(cudaEventCreate(&sl->a_ev[0])); CK(cudaEventCreate(&sl->a_ev[1])); }
        if (!sl->b_stream) {
            CK(cudaStreamCreateWithFlags(&sl->b_stream, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&sl->b_fork, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&sl->b_join, cudaEventDisableTiming));
        }
        if (zout) {  // ---- fused: small fields into the sparse scratch, then the payload of every BGZF block straight from the sources
            static const std::vector<uint32_t> tables = bgzf_tables();
            if (!ctx->s_bgzf_tables.p) upload(ctx->s_bgzf_tables, tables.data(), tables.size(), st);
            sl->z_prefix.ensure(std::max<uint64_t>(prefix_bytes, 16) + 64, st);
            if (prefix_bytes) CK(cudaMemcpyAsync(sl->z_prefix.p, prefix, prefix_bytes, cudaMemcpyHostToDevice, st));
            const uint64_t zn = prefix_bytes + total;
            const uint64_t n_blocks = (zn + kBgzfIn - 1) / kBgzfIn;
            const uint64_t framed = zn + uint64_t(kBgzfOverhead) * n_blocks;
            static const uint8_t kEof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            const uint64_t ztotal = framed + ((flags & PTL_BGZF_EOF) ? sizeof(kEof) : 0);
            sl->z_out.ensure(ztotal + 64, st);
            FrameArgs F{};
            F.A = A;
            F.prefix = sl->z_prefix.as<uint8_t>();
            F.prefix_bytes = prefix_bytes;
            F.Z.in = nullptr;
            F.Z.n = zn;
            F.Z.out = sl->z_out.as<uint8_t>();
            F.Z.tables = ctx->s_bgzf_tables.as<uint32_t>();
            F.Z.n_blocks = n_blocks;
            F.Z.init_full = crc_zero_bytes(tables.data(), 0xffffffffu, kBgzfIn);
            F.Z.init_last = crc_zero_bytes(tables.data(), 0xffffffffu, n_blocks ? zn - (n_blocks - 1) * kBgzfIn : 0);
            CK(cudaEventRecord(sl->a_ev[0], st));
            launch_bam_frame(F, st, LaunchTally(ctx));
            CK(cudaEventRecord(sl->a_ev[1], st));
            if (flags & PTL_BGZF_EOF) CK(cudaMemcpyAsync(sl->z_out.as<uint8_t>() + framed, kEof, sizeof(kEof), cudaMemcpyHostToDevice, st));
            *zout = ptl_bgzf_stream{};
            if (!(flags & PTL_ASM_NO_DOWNLOAD)) {
                sl->hz_out.ensure(ztotal + 64);
                if (ztotal) CK(cudaMemcpyAsync(sl->hz_out.p, sl->z_out.p, ztotal, cudaMemcpyDeviceToHost, st));
                zout->bytes = sl->hz_out.as<uint8_t>();
            }
            CK(cudaStreamSynchronize(st));
            CK(cudaGetLastError());
            CK(cudaEventElapsedTime(&zout->kernel_ms, sl->a_ev[0], sl->a_ev[1]));
            zout->n_bytes = ztotal;
            zout->n_blocks = n_blocks;
            // algorithmic bytes: every input byte of a record once (+ the prefix), the framed stream once
            zout->bytes_read = zn - std::min<uint64_t>(total, 36ull * n_rec);
            zout->bytes_written = framed;
            out->rec_begin = sl->hb_rec_begin.as<uint64_t>();
            sl->b_total = 0;  // (no assembled record stream exists on the slot: ptl_bgzf_store_records has nothing to frame)
            return PTL_OK;
        }
        CK(cudaEventRecord(sl->a_ev[0], st));
        CK(cudaEventRecord(sl->b_fork, st));
        CK(cudaStreamWaitEvent(sl->b_stream, sl->b_fork, 0));
        launch_bam_write(A, st, sl->b_stream, LaunchTally(ctx));
        CK(cudaEventRecord(sl->b_join, sl->b_stream));
        CK(cudaStreamWaitEvent(st, sl->b_join, 0));
        CK(cudaEventRecord(sl->a_ev[1], st));
        if (!(flags & PTL_ASM_NO_DOWNLOAD)) {
            sl->hb_out.ensure(total + 32);
            if (total) CK(cudaMemcpyAsync(sl->hb_out.p, sl->b_out.as<uint8_t>() + kBamFront, total, cudaMemcpyDeviceToHost, st));
            out->bytes = sl->hb_out.as<uint8_t>();
        }
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&out->kernel_ms, sl->a_ev[0], sl->a_ev[1]));
        out->rec_begin = sl->hb_rec_begin.as<uint64_t>();
        // algorithmic bytes: the output once; the input of a record = its output minus what the kernel generates itself
        // (block_size + core 36 B, PS / ZM / SA text), i.e. name, CIGAR, bases, qualities and surviving aux, read once
        out->bytes_written = total;
        out->bytes_read = total - std::min<uint64_t>(total, 36ull * n_rec);
        return PTL_OK;
    });
}
int ptl_assemble_records(ptl_ctx* ctx, int slot, const ptl_read_extras* x, uint32_t flags, ptl_bam_records* out) {
    if (!out) return PTL_ERR_INVALID_ARG;
    return assemble_impl(ctx, slot, x, flags, out, nullptr, 0, nullptr);
}
int ptl_frame_records(ptl_ctx* ctx, int slot, const ptl_read_extras* x, const uint8_t* prefix, uint64_t prefix_bytes, uint32_t flags, ptl_bgzf_stream* out) {
    if (!out) return PTL_ERR_INVALID_ARG;
    return assemble_impl(ctx, slot, x, flags, nullptr, prefix, prefix_bytes, out);
}
int ptl_set_long_pair_ops(ptl_ctx* ctx, uint32_t n_ops) {
    if (!ctx) return PTL_ERR_INVALID_ARG;
    ctx->long_pair_ops = n_ops;
    return PTL_OK;
}
int ptl_set_seq_zero_copy(ptl_ctx* ctx, int enable) {
    if (!ctx) return PTL_ERR_INVALID_ARG;
    ctx->zero_copy_seq = enable != 0;
    return PTL_OK;
}
// Counters of the last finished batch on a slot: out[0..6) = n_pairs, n_lifted, n_in_ops, n_out_ops, base bytes compared,
// scratch ops needed (roofline arithmetic).
int ptl_slot_counters(ptl_ctx* ctx, int slot, uint64_t* out) {
    Slot* sl = get_slot(ctx, slot);
    if (!sl || !out || !sl->ran) return PTL_ERR_INVALID_ARG;
    return guarded(ctx, [&]() {
        finish_batch(ctx, *sl);
        const DevTotals& t = sl->totals;
        out[0] = t.n_pairs; out[1] = t.n_lifted; out[2] = t.n_in_ops; out[3] = t.n_cigar_out; out[4] = t.n_base_bytes; out[5] = t.scratch_needed;
        return PTL_OK;
    });
}

void* ptl_host_alloc(size_t bytes) {
    void* p = nullptr;
    // whole 256-byte granules: the kernels fetch aligned 4 / 8 / 16-byte words around the bytes they need, and the word around
    // the last byte of a pool must still be inside its allocation
    if (cudaHostAlloc(&p, ((bytes ? bytes : 1) + 255) & ~size_t(255), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void ptl_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
