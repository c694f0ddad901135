// C-ABI of the BAM / FASTA input side (include/portello_b200.h: ptl_bam_*, ptl_decoded_*, ptl_fasta_*, ptl_scan_contig_bam):
// thin wrappers over bam_io.{hpp,cpp}, plus the record loop of scan_contig_bam (src/contig_alignment_scanner/mod.rs:186-354).
#include <algorithm>
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "bam_io.hpp"
#include "host_util.hpp"

using namespace ptl;

struct ptl_bam_file {
    BamFile* f = nullptr;
    std::vector<const char*> names;
};
struct ptl_decoded_batch {
    DecodedBatch d;
};
struct ptl_fasta {
    std::vector<FastaRecord> recs;
};
struct ptl_contig_scan {
    std::vector<uint32_t> contig_id;
    std::vector<uint16_t> flag;
    std::vector<int32_t> tid;
    std::vector<int64_t> pos;
    std::vector<uint8_t> mapq;
    std::vector<uint64_t> cigar_begin{0};
    std::vector<uint32_t> cigar;
    std::vector<std::string> sa_s;
    std::vector<std::vector<uint8_t>> seq_s;
    std::vector<const char*> sa;
    std::vector<const uint8_t*> seq;
    std::vector<uint64_t> contig_len;
    std::vector<std::string> contig_name_s, ref_name_s;
    std::vector<const char*> contig_names, ref_names;
};

namespace {
thread_local std::string g_err;
template <class F>
int guarded(F f) {
    try {
        return f();
    } catch (const BamIoError& e) {
        g_err = e.msg;
        return PTL_ERR_INPUT;
    } catch (const InputError& e) {
        g_err = e.what();
        return PTL_ERR_INPUT;
    } catch (const std::exception& e) {
        g_err = e.what();
        return PTL_ERR_INVALID_ARG;
    }
}
// region windows of get_region_segments (lib/rust-vc-utils/src/util.rs:50-67)
std::vector<std::pair<uint64_t, uint64_t>> windows(uint64_t size, uint64_t seg) {
    const uint32_t n = ptl_region_segment_count(size, seg);
    std::vector<uint64_t> b(n), e(n);
    if (n) ptl_region_segments(size, seg, b.data(), e.data());
    std::vector<std::pair<uint64_t, uint64_t>> out;
    for (uint32_t i = 0; i < n; ++i) out.emplace_back(b[i], e[i]);
    return out;
}
}  // namespace

extern "C" {

const char* ptl_bam_last_error(void) { return g_err.c_str(); }

int ptl_bam_open(const char* path, ptl_bam_file** out) {
    if (!path || !out) return PTL_ERR_INVALID_ARG;
    *out = nullptr;
    return guarded([&]() {
        auto* h = new ptl_bam_file();
        try {
            h->f = BamFile::open(path);
        } catch (...) {
            delete h;
            throw;
        }
        for (const auto& n : h->f->ref_names()) h->names.push_back(n.c_str());
        *out = h;
        return PTL_OK;
    });
}
void ptl_bam_close(ptl_bam_file* f) {
    if (!f) return;
    delete f->f;
    delete f;
}
uint32_t ptl_bam_n_ref(const ptl_bam_file* f) { return f ? uint32_t(f->f->ref_names().size()) : 0; }
const char* ptl_bam_ref_name(const ptl_bam_file* f, uint32_t i) { return (f && i < f->names.size()) ? f->names[i] : nullptr; }
uint64_t ptl_bam_ref_len(const ptl_bam_file* f, uint32_t i) { return (f && i < f->f->ref_len().size()) ? f->f->ref_len()[i] : 0; }
const char* ptl_bam_header_text(const ptl_bam_file* f) { return f ? f->f->header_text().c_str() : ""; }
int ptl_bam_has_index(const ptl_bam_file* f) { return f && f->f->has_index(); }
int ptl_bam_has_eof_marker(const ptl_bam_file* f) { return f && f->f->has_eof_marker(); }

int ptl_bam_fetch(const ptl_bam_file* f, int32_t tid, int64_t begin, int64_t end, uint32_t filter, ptl_decoded_batch** out) {
    if (!f || !out) return PTL_ERR_INVALID_ARG;
    *out = nullptr;
    return guarded([&]() {
        auto* d = new ptl_decoded_batch();
        try {
            f->f->fetch(tid, begin, end, filter, d->d);
        } catch (...) {
            delete d;
            throw;
        }
        *out = d;
        return PTL_OK;
    });
}
void ptl_decoded_view(const ptl_decoded_batch* d, ptl_read_records* recs, ptl_read_extras* extras) {
    if (d) d->d.view(recs, extras);
}
const uint8_t* ptl_decoded_raw(const ptl_decoded_batch* d, const uint64_t** rec_off, uint64_t* n_bytes) {
    if (!d) return nullptr;
    if (rec_off) *rec_off = d->d.raw_off.data();
    if (n_bytes) *n_bytes = d->d.raw.size();
    return d->d.raw.data();
}
void ptl_decoded_free(ptl_decoded_batch* d) { delete d; }

int ptl_bam_index_build(const char* bam_path, const char* bai_path) {
    if (!bam_path) return PTL_ERR_INVALID_ARG;
    return guarded([&]() {
        build_bai(bam_path, bai_path ? std::string(bai_path) : std::string(bam_path) + ".bai");
        return PTL_OK;
    });
}
int ptl_bam_index_build_csi(const char* bam_path, const char* csi_path, int min_shift, int depth) {
    if (!bam_path) return PTL_ERR_INVALID_ARG;
    return guarded([&]() {
        build_csi(bam_path, csi_path ? std::string(csi_path) : std::string(bam_path) + ".csi", min_shift, depth);
        return PTL_OK;
    });
}

int ptl_fasta_load(const char* path, int n_threads, ptl_fasta** out) {
    if (!path || !out) return PTL_ERR_INVALID_ARG;
    *out = nullptr;
    return guarded([&]() {
        auto* h = new ptl_fasta();
        try {
            h->recs = read_fasta(path, n_threads);
        } catch (...) {
            delete h;
            throw;
        }
        *out = h;
        return PTL_OK;
    });
}
uint32_t ptl_fasta_n(const ptl_fasta* f) { return f ? uint32_t(f->recs.size()) : 0; }
const char* ptl_fasta_name(const ptl_fasta* f, uint32_t i) { return (f && i < f->recs.size()) ? f->recs[i].name.c_str() : nullptr; }
const uint8_t* ptl_fasta_seq(const ptl_fasta* f, uint32_t i, uint64_t* len) {
    if (!f || i >= f->recs.size()) return nullptr;
    if (len) *len = f->recs[i].seq.size();
    return f->recs[i].seq.data();
}
void ptl_fasta_free(ptl_fasta* f) { delete f; }

// scan_contig_bam's record loop: one task per (reference chromosome x <= 20 Mb window), the records that start in the
// window (mod.rs:213-217), unmapped / secondary skipped (:208), gathered in (chromosome, window, file) order.
int ptl_scan_contig_bam(const ptl_bam_file* bam, uint32_t n_contigs, const char* const* contig_names, const uint64_t* contig_len, int n_threads,
                        ptl_contig_scan** out) {
    if (!bam || !out || (n_contigs && (!contig_names || !contig_len))) return PTL_ERR_INVALID_ARG;
    *out = nullptr;
    return guarded([&]() {
        const BamFile& f = *bam->f;
        if (!f.has_index()) throw BamIoError{"alignment file is not indexed"};
        std::unordered_map<std::string, uint32_t> name_to_contig;
        for (uint32_t i = 0; i < n_contigs; ++i) name_to_contig.emplace(contig_names[i], i);
        struct Task { int32_t tid; int64_t b, e; DecodedBatch d; std::string err; };
        std::vector<Task> tasks;
        for (size_t c = 0; c < f.ref_len().size(); ++c)
            for (const auto& w : windows(f.ref_len()[c], 20000000ull)) tasks.push_back(Task{int32_t(c), int64_t(w.first), int64_t(w.second), {}, {}});
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= tasks.size()) return;
                try {
                    f.fetch(tasks[k].tid, tasks[k].b, tasks[k].e, kKeepStartInRegion | kSkipUnmappedSecondary, tasks[k].d);
                } catch (const BamIoError& e) {
                    tasks[k].err = e.msg;
                }
            }
        };
        const int nt = std::max(1, std::min<int>(n_threads, int(tasks.size())));
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        auto s = std::make_unique<ptl_contig_scan>();
        static const char kDecode[] = "=ACMGRSVTWYHKDBN";
        for (Task& t : tasks) {
            if (!t.err.empty()) throw BamIoError{t.err};
            const DecodedBatch& d = t.d;
            for (uint32_t r = 0; r < d.size(); ++r) {
                const std::string qname(reinterpret_cast<const char*>(d.names.data() + d.name_off[r]), size_t(d.name_off[r + 1] - d.name_off[r]));
                auto it = name_to_contig.find(qname);
                // assembly_contig_list.label_to_index[qname] panics on an unknown contig (mod.rs:219-220)
                if (it == name_to_contig.end()) throw InputError("contig '" + qname + "' of the assembly-to-ref alignment file is not in the read-to-assembly header");
                s->contig_id.push_back(it->second);
                s->flag.push_back(d.flag[r]);
                s->tid.push_back(d.tid[r]);
                s->pos.push_back(d.pos[r]);
                s->mapq.push_back(d.mapq[r]);
                s->cigar.insert(s->cigar.end(), d.cigar.begin() + long(d.cigar_begin[r]), d.cigar.begin() + long(d.cigar_begin[r + 1]));
                s->cigar_begin.push_back(s->cigar.size());
                s->sa_s.push_back(d.sa_tag[r] ? std::string(d.sa_tag[r]) : std::string());
                std::vector<uint8_t> ascii;
                if (!(d.flag[r] & 0x800)) {  // add_primary_read decodes record.seq() (mod.rs:113-125)
                    const uint32_t L = d.seq_len[r];
                    // the contig preparation reads contig_len bases of a primary record (rev_contig_seq): a record that holds
                    // another number of bases (and is not simply without them) cannot be used
                    if (L != 0 && uint64_t(L) != contig_len[it->second])
                        throw InputError("contig '" + qname + "': the primary alignment record holds " + std::to_string(L) + " bases, the read-to-assembly header gives the contig " +
                                         std::to_string(contig_len[it->second]));
                    ascii.resize(L);
                    const uint8_t* p = d.seq4.data() + d.seq_off[r];
                    for (uint32_t i = 0; i < L; ++i) ascii[i] = uint8_t(kDecode[(i & 1u) ? (p[i >> 1] & 0xf) : (p[i >> 1] >> 4)]);
                }
                s->seq_s.push_back(std::move(ascii));
            }
            t.d = DecodedBatch{};  // (a contig record carries megabases: release as we go)
        }
        const size_t n = s->contig_id.size();
        for (size_t i = 0; i < n; ++i) {
            const bool has_sa = !s->sa_s[i].empty();
            s->sa.push_back(has_sa ? s->sa_s[i].c_str() : nullptr);
            s->seq.push_back(((s->flag[i] & 0x800) || s->seq_s[i].empty()) ? nullptr : s->seq_s[i].data());
        }
        for (uint32_t i = 0; i < n_contigs; ++i) {
            s->contig_len.push_back(contig_len[i]);
            s->contig_name_s.emplace_back(contig_names[i]);
        }
        for (const auto& nm : f.ref_names()) s->ref_name_s.push_back(nm);
        for (const auto& nm : s->contig_name_s) s->contig_names.push_back(nm.c_str());
        for (const auto& nm : s->ref_name_s) s->ref_names.push_back(nm.c_str());
        *out = s.release();
        return PTL_OK;
    });
}
void ptl_contig_scan_view(const ptl_contig_scan* s, ptl_contig_records* o) {
    if (!s || !o) return;
    *o = ptl_contig_records{};
    o->n_records = uint32_t(s->contig_id.size());
    o->contig_id = s->contig_id.data(); o->flag = s->flag.data(); o->tid = s->tid.data(); o->pos = s->pos.data(); o->mapq = s->mapq.data();
    o->cigar_begin = s->cigar_begin.data(); o->cigar = s->cigar.data(); o->sa_tag = s->sa.data(); o->seq = s->seq.data();
    o->n_contigs = uint32_t(s->contig_len.size()); o->contig_len = s->contig_len.data(); o->contig_names = s->contig_names.data();
    o->n_ref_chrom = uint32_t(s->ref_names.size()); o->ref_chrom_names = s->ref_names.data();
}
void ptl_contig_scan_free(ptl_contig_scan* s) { delete s; }

}  // extern "C"
