// Contig->reference segment preparation of the product library (host C++, runs once per run, O(#contig records)).
//
// Reference behaviour reproduced (paths relative to the reference repo):
//   a11  src/contig_alignment_scanner/mod.rs:91-183,360-439   add_primary_read, supplementary exact-CIGAR fill-in
//   a12  src/contig_alignment_scanner/contig_repeated_match_trimmer.rs:18-303  clip_repeated_contig_matches
//        src/contig_alignment_scanner/contig_colinear_segment_joiner.rs:15-186 join_colinear_contig_segments
//        lib/rust-vc-utils/src/bam_utils/cigar/clip_alignment.rs:103-181        clip_alignment_read_edges
//        lib/rust-vc-utils/src/bam_utils/cigar/score_alignment.rs:68-74,138-165 gap-compressed identity (the only f64)
// The device tables themselves are built by a CUDA kernel from the segment CIGARs this file produces (tables.cu).
#include "contig_prep.hpp"

#include <algorithm>
#include <cstring>
#include <map>
#include <tuple>
#include <unordered_map>

namespace ptl {

namespace {

// clip_alignment_read_start (clip_alignment.rs:103-156) applied to ops visited in `fwd` or reverse order.
// Emits in visiting order; returns the reference shift.
int64_t clip_from_start(const Ops& in, bool reversed, uint64_t min_clip, Ops& out) {
    uint64_t read_pos = 0;
    int64_t ref_shift = 0;
    const size_t n = in.size();
    for (size_t i = 0; i < n; ++i) {
        const uint32_t c = in[reversed ? n - 1 - i : i];
        const uint32_t op = op_of(c), len = len_of(c);
        if (op == OP_D || op == OP_N) {
            if (read_pos <= min_clip) ref_shift += len;
            else out.push_back(c);
        } else if (op == OP_I) {
            out.push_back(read_pos < min_clip ? mk(OP_S, len) : c);
        } else if (op == OP_M || op == OP_X || op == OP_EQ) {
            if (read_pos < min_clip) {
                const uint64_t remaining = min_clip - read_pos;
                const uint64_t keep = len > remaining ? len - remaining : 0;
                const uint64_t clip = len - keep;
                out.push_back(mk(OP_S, clip));  // may be zero-length; compress removes it
                if (keep) out.push_back(mk(op, keep));
                ref_shift += int64_t(clip);
            } else {
                out.push_back(c);
            }
        } else {
            out.push_back(c);
        }
        read_pos += read_adv(c);
    }
    return ref_shift;
}

// clip_alignment_read_edges (clip_alignment.rs:166-181)
int64_t clip_read_edges(const Ops& in, uint64_t min_left, uint64_t min_right, Ops& out) {
    Ops right_rev, right, left;
    clip_from_start(in, true, min_right, right_rev);
    right.assign(right_rev.rbegin(), right_rev.rend());
    const int64_t shift = clip_from_start(right, false, min_left, left);
    out.clear();
    compress_into(left.data(), left.size(), out);
    return shift;
}

// get_gap_compressed_identity_no_align_match + get_final_gci
double gap_compressed_identity(const Ops& c) {
    uint32_t mismatch = 0, match = 0;
    for (uint32_t o : c) {
        switch (op_of(o)) {
            case OP_I: case OP_D: case OP_N: mismatch += 1; break;
            case OP_X: mismatch += len_of(o); break;
            case OP_EQ: match += len_of(o); break;
            case OP_M: throw InputError("gap-compressed identity needs =/X CIGARs, found alignment match (M)");
            default: break;
        }
    }
    if (match + mismatch == 0) return 1.0;
    return double(match) / double(match + mismatch);
}

struct Range { int64_t start, end; };
inline Range mirrored(const Range& r, int64_t size) { return {size - r.end, size - r.start}; }

// get_seg_gap_compressed_identity (trimmer :18-49)
double seg_identity(const HostSegment& seg, const Range& isec_so) {
    const int64_t read_len = int64_t(read_span(seg.cigar.data(), seg.cigar.size()));
    const Range r = seg.is_fwd ? isec_so : mirrored(isec_so, read_len);
    Ops clipped;
    clip_read_edges(seg.cigar, uint64_t(r.start), uint64_t(read_len - r.end), clipped);
    return gap_compressed_identity(clipped);
}

// clip_seg_isec_range (trimmer :54-112); true = segment eliminated
bool clip_overlap(HostSegment& seg, const Range& isec_so) {
    const bool so_prefix = (isec_so.start == int64_t(seg.so_start));
    const bool aln_prefix = so_prefix ^ (!seg.is_fwd);
    const int64_t read_len = int64_t(read_span(seg.cigar.data(), seg.cigar.size()));
    Range r = seg.is_fwd ? isec_so : mirrored(isec_so, read_len);
    const uint64_t min_left = aln_prefix ? uint64_t(r.end) : 0;
    const uint64_t min_right = aln_prefix ? 0 : uint64_t(read_len - r.start);
    Ops clipped;
    seg.pos += clip_read_edges(seg.cigar, min_left, min_right, clipped);
    seg.cigar.swap(clipped);
    const ClipPos c = clip_positions(seg.cigar.data(), seg.cigar.size());
    if (c.left >= c.right) return true;
    if (aln_prefix) r.end = int64_t(c.left);
    else r.start = int64_t(c.right);
    const Range so = seg.is_fwd ? r : mirrored(r, read_len);
    if (so_prefix) seg.so_start = uint32_t(so.end);
    else seg.so_end = uint32_t(so.start);
    return false;
}

// get_seg_ref_gap (joiner :15-24)
int64_t ref_gap(const HostSegment& a, const HostSegment& b) {
    if (a.is_fwd) return b.pos - (a.pos + ref_span(a.cigar.data(), a.cigar.size()));
    return a.pos - (b.pos + ref_span(b.cigar.data(), b.cigar.size()));
}

void strip_trailing_clip(Ops& c) {  // cigar/mod.rs:315-327 (removes every clip after the first non-clip op)
    Ops o;
    bool seen = false;
    for (uint32_t x : c) {
        if (seen) { if (!is_clip(x)) o.push_back(x); }
        else { if (!is_clip(x)) seen = true; o.push_back(x); }
    }
    c.swap(o);
}
void strip_leading_clip(Ops& c) {  // cigar/mod.rs:300-312
    size_t k = 0;
    while (k < c.size() && is_clip(c[k])) ++k;
    c.erase(c.begin(), c.begin() + long(k));
}

}  // namespace

// clip_repeated_contig_matches (trimmer :214-303)
size_t trim_repeated_matches(std::vector<HostContig>& contigs) {
    size_t clipped = 0;
    for (auto& ct : contigs) {
        auto& segs = ct.segments;
        const size_t n = segs.size();
        if (!n) continue;
        std::vector<char> gone(n, 0);
        for (size_t i = 0; i < n; ++i) {
            for (size_t j = i + 1; j < n; ++j) {
                if (gone[i] || gone[j]) continue;
                if (segs[i].so_end <= segs[j].so_start) break;  // get_seg_clip_info -> None -> break (:236-241)
                const Range isec{int64_t(segs[j].so_start), int64_t(segs[i].so_end)};
                const double gi = seg_identity(segs[i], isec), gj = seg_identity(segs[j], isec);
                // loser: lower identity, then lower MAPQ, full tie -> the later segment (:183-201)
                const bool clip_i = (gj > gi) || (gj == gi && segs[j].mapq > segs[i].mapq);
                const size_t victim = clip_i ? i : j;
                if (clip_overlap(segs[victim], isec)) gone[victim] = 1;
                ++clipped;
            }
        }
        std::vector<HostSegment> kept;
        for (size_t i = 0; i < n; ++i)
            if (!gone[i]) kept.push_back(std::move(segs[i]));
        segs.swap(kept);
    }
    return clipped;
}

// join_colinear_contig_segments (joiner :124-186)
size_t join_colinear(std::vector<HostContig>& contigs) {
    size_t joined = 0;
    for (auto& ct : contigs) {
        if (ct.segments.empty()) continue;
        std::vector<HostSegment> old;
        old.swap(ct.segments);
        auto& out = ct.segments;
        for (auto& seg : old) {
            if (out.empty()) { out.push_back(std::move(seg)); continue; }
            HostSegment& last = out.back();
            if (seg.so_start < last.so_end) throw InputError("Incomplete repeat trimming on contig " + ct.name);  // :150-157
            bool joinable = last.chrom == seg.chrom && last.is_fwd == seg.is_fwd;
            int64_t gap = 0;
            if (joinable) {
                gap = ref_gap(last, seg);
                joinable = gap >= 0 && gap <= 1000 && last.mapq == seg.mapq;  // :37-47
            }
            if (!joinable) { out.push_back(std::move(seg)); continue; }
            const uint32_t ins = seg.so_start - last.so_end;
            // forward: last ++ I ++ D ++ seg ; reverse: seg ++ I ++ D ++ last, and the joined segment starts at seg.pos
            Ops& a = last.is_fwd ? last.cigar : seg.cigar;
            Ops& b = last.is_fwd ? seg.cigar : last.cigar;
            strip_trailing_clip(a);
            if (ins) a.push_back(mk(OP_I, ins));
            if (gap) a.push_back(mk(OP_D, uint64_t(gap)));
            strip_leading_clip(b);
            a.insert(a.end(), b.begin(), b.end());
            if (!last.is_fwd) {
                last.cigar.swap(seg.cigar);
                last.pos = seg.pos;
            }
            last.so_end = seg.so_end;
            ++joined;
        }
    }
    return joined;
}

static inline uint8_t comp_base(uint8_t x) {  // seq_util.rs:1-15
    switch (x) {
        case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; case 'N': return 'N';
        case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c'; case 'n': return 'n';
        default: return 'N';
    }
}

// scan_contig_bam minus BAM I/O (mod.rs:186-240 dispatch, :360-439 fill-in); trim/join are separate calls.
std::vector<HostContig> assemble_from_records(const ptl_contig_records& r) {
    // a public entry point (ptl_set_contig_records / ptl_prepare_contig_records): the arrays are checked before they index
    // anything (tools/fuzz/fuzz_contig_records.py).  cigar_begin[n_records] is the size of the caller's CIGAR pool.
    if (r.n_records && (!r.contig_id || !r.flag || !r.tid || !r.pos || !r.mapq || !r.cigar_begin)) throw InputError("contig records: missing arrays");
    if (r.n_contigs && !r.contig_len) throw InputError("contig records: missing contig lengths");
    if (r.n_ref_chrom && !r.ref_chrom_names) throw InputError("contig records: missing reference names");
    for (uint32_t i = 0; i < r.n_records; ++i) {
        if (r.cigar_begin[i] > r.cigar_begin[i + 1] || r.cigar_begin[i + 1] > r.cigar_begin[r.n_records]) throw InputError("contig records: cigar_begin is not monotone");
        if (r.cigar_begin[i + 1] > r.cigar_begin[i] && !r.cigar) throw InputError("contig records: missing CIGAR pool");
        if (r.flag[i] & (0x4 | 0x100)) continue;  // (skipped below: their other fields are never looked at)
        if (r.contig_id[i] >= r.n_contigs) throw InputError("contig record with an out-of-range contig id");
        if (r.tid[i] < 0 || uint32_t(r.tid[i]) >= r.n_ref_chrom) throw InputError("contig record with a reference index outside the header");
        if (r.pos[i] < 0 || r.pos[i] > 0x7fffffffLL) throw InputError("contig record position outside the BAM range");
    }
    std::vector<HostContig> contigs(r.n_contigs);
    for (uint32_t i = 0; i < r.n_contigs; ++i) {
        contigs[i].length = r.contig_len[i];
        contigs[i].name = r.contig_names ? r.contig_names[i] : ("contig" + std::to_string(i));
    }
    std::unordered_map<std::string, uint32_t> chrom_index;
    for (uint32_t i = 0; i < r.n_ref_chrom; ++i) chrom_index.emplace(r.ref_chrom_names[i], i);
    std::vector<const char*> chrom_names(r.ref_chrom_names, r.ref_chrom_names + r.n_ref_chrom);

    using Key = std::tuple<uint32_t, int64_t, bool, uint32_t, uint32_t>;  // SplitReadKey (mod.rs:49-56)
    auto make_key = [](uint32_t chrom, int64_t pos, bool fwd, const uint32_t* c, size_t n) {
        const ClipPos cp = clip_positions(c, n);
        return Key{chrom, pos, fwd, uint32_t(cp.left), uint32_t(cp.size - cp.right)};
    };
    std::vector<std::map<Key, uint32_t>> supp(r.n_contigs);  // -> record index
    std::vector<int64_t> primary_rec(r.n_contigs, -1);
    for (uint32_t i = 0; i < r.n_records; ++i) {
        const uint16_t flag = r.flag[i];
        if (flag & (0x4 | 0x100)) continue;  // unmapped / secondary skipped (mod.rs:208-210)
        const uint32_t cid = r.contig_id[i];
        if (cid >= r.n_contigs) throw InputError("contig record with an out-of-range contig id");
        const uint32_t* cg = r.cigar + r.cigar_begin[i];
        const size_t ncg = size_t(r.cigar_begin[i + 1] - r.cigar_begin[i]);
        if (!(flag & 0x800)) {
            primary_rec[cid] = i;  // a later primary record replaces an earlier one, as the assignment at :224-227 does
        } else {
            const Key k = make_key(uint32_t(r.tid[i]), r.pos[i], !(flag & 0x10), cg, ncg);
            if (!supp[cid].emplace(k, i).second)
                throw InputError("Can't uniquely identify split read alignment info in contig '" + contigs[cid].name + "'");
        }
    }
    for (uint32_t cid = 0; cid < r.n_contigs; ++cid) {
        if (primary_rec[cid] < 0) continue;  // unwrap_or_default: no segments (mod.rs:364-367)
        const uint32_t i = uint32_t(primary_rec[cid]);
        HostContig& ct = contigs[cid];
        const uint32_t* cg = r.cigar + r.cigar_begin[i];
        const uint32_t ncg = uint32_t(r.cigar_begin[i + 1] - r.cigar_begin[i]);
        // get_seq_order_read_split_segments through the shared packer entry point
        uint32_t ns = 0, nc = 0;
        ptl_split_segments probe{};
        int rc = ptl_pack_split_segments(r.n_ref_chrom, chrom_names.data(), r.tid[i], r.pos[i], r.flag[i], r.mapq[i], cg, ncg,
                                         r.sa_tag ? r.sa_tag[i] : nullptr, 0, 0, &probe, &ns, &nc);
        if (rc == PTL_ERR_INPUT) throw InputError(std::string(ptl_pack_last_error()) + " (contig '" + ct.name + "')");
        std::vector<uint32_t> so_s(ns), so_e(ns), ctg(ns), cb(ns + 1), cig(std::max<uint32_t>(nc, 1));
        std::vector<int64_t> pos(ns);
        std::vector<uint8_t> fwd(ns), mq(ns), prim(ns);
        ptl_split_segments o{so_s.data(), so_e.data(), ctg.data(), pos.data(), fwd.data(), mq.data(), prim.data(), cb.data(), cig.data()};
        rc = ptl_pack_split_segments(r.n_ref_chrom, chrom_names.data(), r.tid[i], r.pos[i], r.flag[i], r.mapq[i], cg, ncg,
                                     r.sa_tag ? r.sa_tag[i] : nullptr, ns, std::max<uint32_t>(nc, 1), &o, &ns, &nc);
        if (rc != PTL_OK) throw InputError(std::string(ptl_pack_last_error()) + " (contig '" + ct.name + "')");
        bool need_rev = false;
        for (uint32_t k = 0; k < ns; ++k) {
            HostSegment s;
            s.so_start = so_s[k]; s.so_end = so_e[k];
            s.chrom = int32_t(ctg[k]); s.pos = pos[k]; s.is_fwd = fwd[k]; s.mapq = mq[k];
            s.cigar.assign(cig.begin() + cb[k], cig.begin() + cb[k + 1]);
            if (!prim[k]) {  // replace the approximate SA CIGAR with the supplementary record's exact one (:371-416)
                const Key key = make_key(uint32_t(s.chrom), s.pos, s.is_fwd, s.cigar.data(), s.cigar.size());
                auto it = supp[cid].find(key);
                if (it == supp[cid].end())
                    throw InputError("Can't find supplementary alignment record corresponding to segment reported in SA tag for contig '" + ct.name + "'");
                const uint32_t j = it->second;
                s.cigar.assign(r.cigar + r.cigar_begin[j], r.cigar + r.cigar_begin[j + 1]);
            }
            need_rev |= !s.is_fwd;
            ct.segments.push_back(std::move(s));
        }
        if (need_rev) {  // mod.rs:113-125
            if (!r.seq || !r.seq[i]) throw InputError("primary contig record without bases but a reverse-strand segment needs rev_contig_seq");
            const uint64_t L = ct.length;
            ct.has_rev_seq = true;
            ct.rev_seq.resize(L);
            if (r.flag[i] & 0x10) std::memcpy(ct.rev_seq.data(), r.seq[i], L);
            else for (uint64_t k = 0; k < L; ++k) ct.rev_seq[k] = comp_base(r.seq[i][L - 1 - k]);
        }
    }
    return contigs;
}

std::vector<HostContig> contigs_from_flat(const ptl_contig_segments& s) {
    // a public entry point (ptl_set_contig_segments / ptl_prepare_raw_contig_segments): the two CSR arrays and the per-segment
    // fields are checked before anything is indexed with them (tools/fuzz/fuzz_segments_through_device_code.py)
    if (s.n_contigs && (!s.contig_len || !s.contig_seg_begin)) throw InputError("contig segments: missing contig arrays");
    if (s.n_segments && (!s.seg_seq_order_start || !s.seg_seq_order_end || !s.seg_chrom_index || !s.seg_pos || !s.seg_is_fwd || !s.seg_mapq || !s.seg_cigar_begin || !s.cigar))
        throw InputError("contig segments: missing segment arrays");
    if (s.n_contigs) {
        if (s.contig_seg_begin[0] != 0 || s.contig_seg_begin[s.n_contigs] != s.n_segments) throw InputError("contig segments: contig_seg_begin does not span the segments");
        for (uint32_t c = 0; c < s.n_contigs; ++c)
            if (s.contig_seg_begin[c] > s.contig_seg_begin[c + 1]) throw InputError("contig segments: contig_seg_begin is not monotone");
    } else if (s.n_segments) {
        throw InputError("contig segments: segments without contigs");
    }
    for (uint32_t k = 0; k < s.n_segments; ++k) {
        if (s.seg_cigar_begin[k] > s.seg_cigar_begin[k + 1]) throw InputError("contig segments: seg_cigar_begin is not monotone");
        if (s.seg_chrom_index[k] < 0) throw InputError("contig segments: negative chromosome index");
        if (s.seg_pos[k] < 0 || s.seg_pos[k] > 0x7fffffffLL) throw InputError("contig segments: position outside the BAM range");
        if (s.seg_seq_order_start[k] > s.seg_seq_order_end[k]) throw InputError("contig segments: sequencing-order interval is reversed");
    }
    std::vector<HostContig> contigs(s.n_contigs);
    for (uint32_t c = 0; c < s.n_contigs; ++c) {
        HostContig& ct = contigs[c];
        ct.length = s.contig_len[c];
        ct.name = "contig" + std::to_string(c);
        if (s.rev_contig_seq && s.rev_contig_seq[c]) {
            ct.has_rev_seq = true;
            ct.rev_seq.assign(s.rev_contig_seq[c], s.rev_contig_seq[c] + ct.length);
        }
        for (uint32_t k = s.contig_seg_begin[c]; k < s.contig_seg_begin[c + 1]; ++k) {
            HostSegment g;
            g.so_start = s.seg_seq_order_start[k];
            g.so_end = s.seg_seq_order_end[k];
            g.chrom = s.seg_chrom_index[k];
            g.pos = s.seg_pos[k];
            g.is_fwd = s.seg_is_fwd[k] != 0;
            g.mapq = s.seg_mapq[k];
            g.cigar.assign(s.cigar + s.seg_cigar_begin[k], s.cigar + s.seg_cigar_begin[k + 1]);
            ct.segments.push_back(std::move(g));
        }
    }
    return contigs;
}

void FlatContigs::build(const std::vector<HostContig>& contigs) {
    *this = FlatContigs{};
    seg_begin.push_back(0);
    cigar_begin.push_back(0);
    for (const auto& ct : contigs) {
        contig_len.push_back(ct.length);
        rev_seq.push_back(ct.has_rev_seq ? ct.rev_seq : std::vector<uint8_t>());
        has_rev.push_back(ct.has_rev_seq);
        for (const auto& g : ct.segments) {
            so_start.push_back(g.so_start);
            so_end.push_back(g.so_end);
            chrom.push_back(g.chrom);
            pos.push_back(g.pos);
            is_fwd.push_back(g.is_fwd);
            mapq.push_back(g.mapq);
            cigar.insert(cigar.end(), g.cigar.begin(), g.cigar.end());
            cigar_begin.push_back(cigar.size());
        }
        seg_begin.push_back(uint32_t(so_start.size()));
    }
    rev_ptr.clear();
    for (size_t i = 0; i < rev_seq.size(); ++i) rev_ptr.push_back(has_rev[i] ? rev_seq[i].data() : nullptr);
}

void FlatContigs::view(ptl_contig_segments* o) const {
    o->n_contigs = uint32_t(contig_len.size());
    o->contig_len = contig_len.data();
    o->contig_seg_begin = seg_begin.data();
    o->rev_contig_seq = rev_ptr.data();
    o->n_segments = uint32_t(so_start.size());
    o->seg_seq_order_start = so_start.data();
    o->seg_seq_order_end = so_end.data();
    o->seg_chrom_index = chrom.data();
    o->seg_pos = pos.data();
    o->seg_is_fwd = is_fwd.data();
    o->seg_mapq = mapq.data();
    o->seg_cigar_begin = cigar_begin.data();
    o->cigar = cigar.data();
}

}  // namespace ptl

// ---------------------------------------------------------------------------------------------- host-only C-ABI
struct ptl_prepared_contigs {
    ptl::FlatContigs flat;
};
namespace {
thread_local std::string g_prepare_err;
template <class F>
int prepare_guarded(ptl_prepared_contigs** out, F make) {
    if (!out) return PTL_ERR_INVALID_ARG;
    *out = nullptr;
    try {
        std::vector<ptl::HostContig> contigs = make();
        ptl::trim_repeated_matches(contigs);
        ptl::join_colinear(contigs);
        auto* p = new ptl_prepared_contigs();
        p->flat.build(contigs);
        *out = p;
        return PTL_OK;
    } catch (const ptl::InputError& e) {
        g_prepare_err = e.what();
        return PTL_ERR_INPUT;
    } catch (const std::exception& e) {
        g_prepare_err = e.what();
        return PTL_ERR_INVALID_ARG;
    }
}
}  // namespace
extern "C" {
int ptl_prepare_contig_records(const ptl_contig_records* recs, ptl_prepared_contigs** out) {
    if (!recs) return PTL_ERR_INVALID_ARG;
    return prepare_guarded(out, [&]() { return ptl::assemble_from_records(*recs); });
}
int ptl_prepare_raw_contig_segments(const ptl_contig_segments* raw, ptl_prepared_contigs** out) {
    if (!raw) return PTL_ERR_INVALID_ARG;
    return prepare_guarded(out, [&]() { return ptl::contigs_from_flat(*raw); });
}
void ptl_prepared_contigs_view(const ptl_prepared_contigs* p, ptl_contig_segments* out) { p->flat.view(out); }
void ptl_prepared_contigs_free(ptl_prepared_contigs* p) { delete p; }
const char* ptl_prepare_last_error(void) { return g_prepare_err.c_str(); }
}
