// BAM / BGZF / BAI / CSI input and FASTA loading without htslib: see bam_io.hpp.  Written from the SAM specification.
#include "bam_io.hpp"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

namespace ptl {
namespace {

constexpr uint32_t kMetaBin = 37450;  // = meta_bin_of(5)
const uint8_t kEofMarker[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

[[noreturn]] void fail(const std::string& m) { throw BamIoError{m}; }

uint32_t le32(const uint8_t* p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }
uint64_t le64(const uint8_t* p) { return uint64_t(le32(p)) | (uint64_t(le32(p + 4)) << 32); }
uint16_t le16(const uint8_t* p) { return uint16_t(p[0] | (p[1] << 8)); }

size_t pread_full(int fd, void* dst, size_t n, uint64_t off) {
    size_t got = 0;
    while (got < n) {
        const ssize_t k = ::pread(fd, static_cast<char*>(dst) + got, n - got, off_t(off + got));
        if (k < 0) fail("read error");
        if (k == 0) break;
        got += size_t(k);
    }
    return got;
}

// reg2bins (SAM spec 5.3; CSIv1 for other min_shift / depth): the bins that may hold records overlapping [beg, end).
// Level l (0 = the root) has 8^l bins of 2^(min_shift + 3 (depth - l)) bases, numbered from (8^l - 1) / 7.
void reg2bins(int64_t beg, int64_t end, int min_shift, int depth, std::vector<uint32_t>& out) {
    out.clear();
    if (beg < 0) beg = 0;
    const int64_t max_end = int64_t(1) << (min_shift + 3 * depth);
    if (end > max_end) end = max_end;
    if (end <= beg) end = beg + 1;
    --end;
    uint32_t first = 0;
    for (int l = 0, sh = min_shift + 3 * depth; l <= depth; ++l, sh -= 3) {
        for (uint32_t k = first + uint32_t(beg >> sh); k <= first + uint32_t(end >> sh); ++k) out.push_back(k);
        first += 1u << (3 * l);
    }
}
// the smallest bin that holds [beg, end) whole
uint32_t reg2bin_levels(int64_t beg, int64_t end, int min_shift, int depth) {
    --end;
    uint32_t first = ((1u << (3 * depth)) - 1u) / 7u;
    for (int l = depth, sh = min_shift; l > 0; --l, sh += 3) {
        if ((beg >> sh) == (end >> sh)) return first + uint32_t(beg >> sh);
        first -= 1u << (3 * (l - 1));
    }
    return 0;
}
inline uint32_t meta_bin_of(int depth) { return ((1u << (3 * depth + 3)) - 1u) / 7u + 1u; }  // 37450 for the BAI's five levels


// reference bases consumed by a BAM CIGAR (M, D, N, =, X)
int64_t cigar_ref_len(const uint8_t* cigar, uint32_t n) {
    int64_t len = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t c = le32(cigar + 4 * size_t(i));
        if ((0x18dU >> (c & 0xf)) & 1u) len += c >> 4;  // ops 0, 2, 3, 7, 8
    }
    return len;
}

// size of the aux field starting at aux[i] (tag, type, value), 0 if malformed
size_t aux_field_size(const uint8_t* aux, size_t i, size_t n) {
    if (i + 3 > n) return 0;
    size_t sz;
    switch (aux[i + 2]) {
        case 'A': case 'c': case 'C': sz = 1; break;
        case 's': case 'S': sz = 2; break;
        case 'i': case 'I': case 'f': sz = 4; break;
        case 'd': sz = 8; break;
        case 'Z': case 'H': {
            size_t j = i + 3;
            while (j < n && aux[j] != 0) ++j;
            if (j >= n) return 0;
            sz = j - (i + 3) + 1;
            break;
        }
        case 'B': {
            if (i + 8 > n) return 0;
            size_t es;
            switch (aux[i + 3]) {
                case 'c': case 'C': es = 1; break;
                case 's': case 'S': es = 2; break;
                case 'i': case 'I': case 'f': es = 4; break;
                default: return 0;
            }
            sz = 5 + size_t(le32(aux + i + 4)) * es;
            break;
        }
        default: return 0;
    }
    return (i + 3 + sz > n) ? 0 : 3 + sz;
}

struct RecordView {
    int32_t tid;
    int64_t pos, end;
    uint16_t flag;
};
RecordView peek(const uint8_t* rec, uint32_t block_size) {
    if (block_size < 32) fail("BAM record shorter than its fixed fields");
    RecordView v;
    v.tid = int32_t(le32(rec));
    v.pos = int32_t(le32(rec + 4));
    const uint32_t l_name = rec[8], n_cigar = le16(rec + 12);
    v.flag = le16(rec + 14);
    if (32ull + l_name + 4ull * n_cigar > block_size) fail("BAM record: name / CIGAR run past the record");
    int64_t rl = (v.flag & 0x4) ? 0 : cigar_ref_len(rec + 32 + l_name, n_cigar);
    v.end = v.pos + (rl > 0 ? rl : 1);  // bam_endpos
    return v;
}

}  // namespace

// ================================================================================================== BGZF
bool BgzfReader::load(uint64_t coffset) {
    uint8_t h[18];
    const size_t got = pread_full(fd_, h, 12, coffset);
    if (got == 0) return false;
    if (got < 12 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) fail("not a BGZF block at offset " + std::to_string(coffset));
    const uint32_t xlen = le16(h + 10);
    std::vector<uint8_t> extra(xlen);
    if (pread_full(fd_, extra.data(), xlen, coffset + 12) != xlen) fail("truncated BGZF header");
    int64_t bsize = -1;
    for (size_t i = 0; i + 4 <= xlen;) {
        const uint32_t slen = le16(extra.data() + i + 2);
        if (extra[i] == 66 && extra[i + 1] == 67 && slen == 2 && i + 6 <= xlen) bsize = int64_t(le16(extra.data() + i + 4)) + 1;
        i += 4 + slen;
    }
    if (bsize < 0 || bsize < int64_t(12 + xlen + 8)) fail("BGZF block without a BC subfield");
    const size_t clen = size_t(bsize) - 12 - xlen - 8;
    raw_.resize(clen + 8);
    if (pread_full(fd_, raw_.data(), clen + 8, coffset + 12 + xlen) != clen + 8) fail("truncated BGZF block");
    const uint32_t crc = le32(raw_.data() + clen), isize = le32(raw_.data() + clen + 4);
    if (isize > 0x10000) fail("BGZF block with ISIZE > 64 KB");
    buf_.resize(isize);
    z_stream zs{};
    if (inflateInit2(&zs, -15) != Z_OK) fail("inflateInit2 failed");
    zs.next_in = raw_.data();
    zs.avail_in = uInt(clen);
    zs.next_out = buf_.data();
    zs.avail_out = isize;
    const int rc = inflate(&zs, Z_FINISH);
    const bool ok = (rc == Z_STREAM_END) && zs.total_out == isize;
    inflateEnd(&zs);
    if (!ok) fail("BGZF block does not inflate to ISIZE bytes");
    if (uint32_t(crc32(crc32(0L, Z_NULL, 0), buf_.data(), isize)) != crc) fail("BGZF block CRC mismatch");
    block_off_ = coffset;
    next_off_ = coffset + uint64_t(bsize);
    pos_ = 0;
    loaded_ = true;
    return true;
}

void BgzfReader::seek(uint64_t voffset) {
    if (!load(voffset >> 16)) {
        block_off_ = next_off_ = voffset >> 16;
        buf_.clear();
        pos_ = 0;
        loaded_ = true;
        return;
    }
    pos_ = uint32_t(voffset & 0xffff);
    if (pos_ > buf_.size()) fail("virtual offset past the end of its block");
}

bool BgzfReader::read(void* dst, size_t n) {
    if (!loaded_) seek(0);
    uint8_t* d = static_cast<uint8_t*>(dst);
    size_t done = 0;
    while (done < n) {
        if (pos_ == buf_.size()) {
            if (!load(next_off_)) {
                if (done == 0) return false;
                fail("BAM stream ends inside a record");
            }
            continue;
        }
        const size_t k = std::min(n - done, buf_.size() - size_t(pos_));
        std::memcpy(d + done, buf_.data() + pos_, k);
        pos_ += uint32_t(k);
        done += k;
    }
    return true;
}

size_t BgzfReader::read_some(void* dst, size_t n) {
    if (at_eof()) return 0;
    const size_t k = std::min(n, buf_.size() - size_t(pos_));
    std::memcpy(dst, buf_.data() + pos_, k);
    pos_ += uint32_t(k);
    return k;
}

bool BgzfReader::at_eof() {
    if (!loaded_) seek(0);
    while (pos_ == buf_.size())
        if (!load(next_off_)) return true;
    return false;
}

// ================================================================================================== decoded records
void DecodedBatch::append(const uint8_t* rec, uint32_t block_size) {
    if (block_size < 32) fail("BAM record shorter than its fixed fields");
    const uint32_t l_name = rec[8], n_cig = le16(rec + 12);
    const uint32_t l_seq = le32(rec + 16);
    const size_t o_name = 32, o_cigar = o_name + l_name, o_seq = o_cigar + 4ull * n_cig, o_qual = o_seq + (size_t(l_seq) + 1) / 2,
                 o_aux = o_qual + l_seq;
    if (o_aux > block_size || l_name == 0) fail("BAM record: variable-length fields run past the record");
    tid.push_back(int32_t(le32(rec)));
    pos.push_back(int32_t(le32(rec + 4)));
    mapq.push_back(rec[9]);
    bin.push_back(le16(rec + 10));
    flag.push_back(le16(rec + 14));
    seq_len.push_back(l_seq);
    mate_tid.push_back(int32_t(le32(rec + 20)));
    mate_pos.push_back(int32_t(le32(rec + 24)));
    tlen.push_back(int32_t(le32(rec + 28)));
    names.insert(names.end(), rec + o_name, rec + o_name + l_name - 1);  // without the NUL
    name_off.push_back(names.size());
    seq_off.push_back(seq4.size());
    seq4.insert(seq4.end(), rec + o_seq, rec + o_qual);
    qual_off.push_back(qual.size());
    qual.insert(qual.end(), rec + o_qual, rec + o_aux);
    // aux block; the real CIGAR of a record with > 65535 ops travels in CG:B,I behind a `<l_seq>S<ref_len>N` placeholder
    // (SAM spec 4.2.2): restore it and drop the tag, as htslib does when it reads the record (bam_tag2cigar)
    const uint8_t* a = rec + o_aux;
    const size_t an = block_size - o_aux;
    size_t cg_at = an, cg_n = 0;
    int64_t sa = -1;
    const size_t aux_base = aux.size();
    const bool placeholder = n_cig == 2 && (le32(rec + o_cigar) & 0xf) == 4 && (le32(rec + o_cigar) >> 4) == l_seq && (le32(rec + o_cigar + 4) & 0xf) == 3;
    for (size_t i = 0; i < an;) {
        const size_t sz = aux_field_size(a, i, an);
        if (sz == 0) break;  // malformed tail: kept verbatim, never interpreted (clone_record copies it too)
        if (a[i] == 'S' && a[i + 1] == 'A' && a[i + 2] == 'Z' && sa < 0) sa = int64_t(i + 3);
        if (placeholder && a[i] == 'C' && a[i + 1] == 'G' && a[i + 2] == 'B' && a[i + 3] == 'I' && cg_at == an) { cg_at = i; cg_n = sz; }
        i += sz;
    }
    if (cg_at < an) {
        const uint32_t n_real = le32(a + cg_at + 4);
        for (uint32_t k = 0; k < n_real; ++k) cigar.push_back(le32(a + cg_at + 8 + 4 * size_t(k)));
        aux.insert(aux.end(), a, a + cg_at);
        aux.insert(aux.end(), a + cg_at + cg_n, a + an);
        if (sa >= 0 && size_t(sa) > cg_at) sa -= int64_t(cg_n);
    } else {
        for (uint32_t k = 0; k < n_cig; ++k) cigar.push_back(le32(rec + o_cigar + 4 * size_t(k)));
        aux.insert(aux.end(), a, a + an);
    }
    cigar_begin.push_back(cigar.size());
    aux_off.push_back(aux.size());
    sa_at.push_back(sa < 0 ? -1 : int64_t(aux_base) + sa);
    if (keep_raw) {
        uint8_t bs[4] = {uint8_t(block_size), uint8_t(block_size >> 8), uint8_t(block_size >> 16), uint8_t(block_size >> 24)};
        raw.insert(raw.end(), bs, bs + 4);
        raw.insert(raw.end(), rec, rec + block_size);
        raw_off.push_back(raw.size());
    }
}

void DecodedBatch::finish() {
    const size_t n_names = names.size(), n_aux = aux.size(), n_qual = qual.size(), n_seq = seq4.size();
    names.resize(n_names + 32, 0);
    aux.resize(n_aux + 32, 0);
    qual.resize(n_qual + 32, 0);
    seq4.resize(n_seq + 32, 0);
    names.resize(n_names);  // (capacity keeps the zero padding addressable; sizes stay exact)
    aux.resize(n_aux);
    qual.resize(n_qual);
    seq4.resize(n_seq);
    sa_tag.assign(sa_at.size(), nullptr);
    for (size_t i = 0; i < sa_at.size(); ++i)
        if (sa_at[i] >= 0) sa_tag[i] = reinterpret_cast<const char*>(aux.data() + sa_at[i]);
}

void DecodedBatch::view(ptl_read_records* r, ptl_read_extras* x) const {
    if (r) {
        *r = ptl_read_records{};
        r->n_reads = size();
        r->tid = tid.data(); r->pos = pos.data(); r->flag = flag.data(); r->mapq = mapq.data(); r->bin = bin.data();
        r->seq_len = seq_len.data(); r->seq_off = seq_off.data(); r->seq4 = seq4.data(); r->seq4_bytes = seq4.size();
        r->cigar_begin = cigar_begin.data(); r->cigar = cigar.data(); r->sa_tag = sa_tag.data();
    }
    if (x) {
        *x = ptl_read_extras{};
        x->name_off = name_off.data(); x->names = names.data(); x->aux_off = aux_off.data(); x->aux = aux.data();
        x->mate_tid = mate_tid.data(); x->mate_pos = mate_pos.data(); x->tlen = tlen.data();
        x->quals.qual = qual.data(); x->quals.read_qual_off = qual_off.data(); x->quals.qual_bytes = qual.size();
    }
}

// ================================================================================================== BAM file + BAI
namespace {
bool load_bai(const std::string& path, size_t n_ref, BaiIndex& idx) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st {};
    ::fstat(fd, &st);
    std::vector<uint8_t> b(size_t(st.st_size));
    const size_t got = pread_full(fd, b.data(), b.size(), 0);
    ::close(fd);
    if (got != b.size() || b.size() < 8 || std::memcmp(b.data(), "BAI\1", 4) != 0) fail("not a BAI index: " + path);
    size_t at = 4;
    auto need = [&](size_t n) { if (at + n > b.size()) fail("truncated BAI index: " + path); };
    need(4);
    const uint32_t nr = le32(b.data() + at); at += 4;
    if (nr != n_ref) fail("BAI index does not match the BAM header (reference count): " + path);
    idx.refs.assign(nr, BaiRef{});
    for (uint32_t r = 0; r < nr; ++r) {
        BaiRef& R = idx.refs[r];
        need(4);
        const uint32_t n_bin = le32(b.data() + at); at += 4;
        for (uint32_t k = 0; k < n_bin; ++k) {
            need(8);
            const uint32_t bin = le32(b.data() + at), n_chunk = le32(b.data() + at + 4); at += 8;
            need(16ull * n_chunk);
            if (bin == kMetaBin && n_chunk == 2) {
                R.has_meta = true;
                R.meta_beg = le64(b.data() + at); R.meta_end = le64(b.data() + at + 8);
                R.n_mapped = le64(b.data() + at + 16); R.n_unmapped = le64(b.data() + at + 24);
            } else {
                auto& v = R.bins[bin];
                for (uint32_t c = 0; c < n_chunk; ++c) v.emplace_back(le64(b.data() + at + 16ull * c), le64(b.data() + at + 16ull * c + 8));
            }
            at += 16ull * n_chunk;
        }
        need(4);
        const uint32_t n_intv = le32(b.data() + at); at += 4;
        need(8ull * n_intv);
        R.ioffset.resize(n_intv);
        for (uint32_t k = 0; k < n_intv; ++k) R.ioffset[k] = le64(b.data() + at + 8ull * k);
        at += 8ull * n_intv;
    }
    if (at + 8 <= b.size()) { idx.has_no_coor = true; idx.n_no_coor = le64(b.data() + at); }
    return true;
}

// CSIv1 (the index htslib writes for references longer than 2^29 bases, or on request): a BGZF stream of
// magic, min_shift, depth, l_aux, aux, n_ref, then per reference its bins { bin, loffset, n_chunk, chunks }.
bool load_csi(const std::string& path, size_t n_ref, BaiIndex& idx) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct FdGuard { int fd; ~FdGuard() { ::close(fd); } } g{fd};
    std::vector<uint8_t> b;
    {
        BgzfReader r(fd);
        uint8_t buf[65536];
        for (size_t n; (n = r.read_some(buf, sizeof buf)) != 0;) b.insert(b.end(), buf, buf + n);
    }
    if (b.size() < 16 || std::memcmp(b.data(), "CSI\1", 4) != 0) fail("not a CSI index: " + path);
    size_t at = 4;
    auto need = [&](size_t n) { if (at + n > b.size()) fail("truncated CSI index: " + path); };
    idx.csi = true;
    idx.min_shift = int(le32(b.data() + 4));
    idx.depth = int(le32(b.data() + 8));
    const uint32_t l_aux = le32(b.data() + 12);
    at = 16;
    if (idx.min_shift < 1 || idx.min_shift > 30 || idx.depth < 1 || idx.depth > 9 || idx.min_shift + 3 * idx.depth > 62) fail("unsupported CSI geometry: " + path);
    need(size_t(l_aux) + 4);
    at += l_aux;
    const uint32_t nr = le32(b.data() + at); at += 4;
    if (nr != n_ref) fail("CSI index does not match the BAM header (reference count): " + path);
    const uint32_t meta = meta_bin_of(idx.depth);
    idx.refs.assign(nr, BaiRef{});
    for (uint32_t r = 0; r < nr; ++r) {
        BaiRef& R = idx.refs[r];
        need(4);
        const uint32_t n_bin = le32(b.data() + at); at += 4;
        for (uint32_t k = 0; k < n_bin; ++k) {
            need(16);
            const uint32_t bin = le32(b.data() + at);
            const uint64_t loff = le64(b.data() + at + 4);
            const uint32_t n_chunk = le32(b.data() + at + 12);
            at += 16;
            need(16ull * n_chunk);
            if (bin == meta && n_chunk == 2) {
                R.has_meta = true;
                R.meta_beg = le64(b.data() + at); R.meta_end = le64(b.data() + at + 8);
                R.n_mapped = le64(b.data() + at + 16); R.n_unmapped = le64(b.data() + at + 24);
            } else {
                R.loffset[bin] = loff;
                auto& v = R.bins[bin];
                for (uint32_t c = 0; c < n_chunk; ++c) v.emplace_back(le64(b.data() + at + 16ull * c), le64(b.data() + at + 16ull * c + 8));
            }
            at += 16ull * n_chunk;
        }
    }
    if (at + 8 <= b.size()) { idx.has_no_coor = true; idx.n_no_coor = le64(b.data() + at); }
    return true;
}
}  // namespace

BamFile* BamFile::open(const std::string& path) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) fail("cannot open alignment file: " + path);
    auto* f = new BamFile();
    f->fd_ = fd;
    f->path_ = path;
    try {
        BgzfReader r(fd);
        uint8_t m[8];
        if (!r.read(m, 8) || std::memcmp(m, "BAM\1", 4) != 0) fail("not a BAM file (CRAM is not supported by this build): " + path);
        const uint32_t l_text = le32(m + 4);
        f->text_.resize(l_text);
        if (l_text && !r.read(&f->text_[0], l_text)) fail("truncated BAM header: " + path);
        while (!f->text_.empty() && f->text_.back() == '\0') f->text_.pop_back();
        uint8_t w[4];
        if (!r.read(w, 4)) fail("truncated BAM header: " + path);
        const uint32_t n_ref = le32(w);
        for (uint32_t i = 0; i < n_ref; ++i) {
            if (!r.read(w, 4)) fail("truncated BAM header: " + path);
            const uint32_t l_name = le32(w);
            std::string nm(l_name, '\0');
            if (l_name && !r.read(&nm[0], l_name)) fail("truncated BAM header: " + path);
            while (!nm.empty() && nm.back() == '\0') nm.pop_back();
            if (!r.read(w, 4)) fail("truncated BAM header: " + path);
            f->ref_names_.push_back(nm);
            f->ref_len_.push_back(le32(w));
        }
        f->first_record_voff_ = r.tell();
        struct stat st {};
        ::fstat(fd, &st);
        if (st.st_size >= 28) {
            uint8_t tail[28];
            pread_full(fd, tail, 28, uint64_t(st.st_size) - 28);
            f->has_eof_ = std::memcmp(tail, kEofMarker, 28) == 0;
        }
        f->has_index_ = load_bai(path + ".bai", n_ref, f->index_);
        const bool dot_bam = path.size() > 4 && path.compare(path.size() - 4, 4, ".bam") == 0;
        if (!f->has_index_ && dot_bam) f->has_index_ = load_bai(path.substr(0, path.size() - 4) + ".bai", n_ref, f->index_);
        if (!f->has_index_) f->has_index_ = load_csi(path + ".csi", n_ref, f->index_);
        if (!f->has_index_ && dot_bam) f->has_index_ = load_csi(path.substr(0, path.size() - 4) + ".csi", n_ref, f->index_);
    } catch (...) {
        delete f;
        throw;
    }
    return f;
}

BamFile::~BamFile() {
    if (fd_ >= 0) ::close(fd_);
}

void BamFile::fetch(int32_t tid, int64_t begin, int64_t end, uint32_t filter, DecodedBatch& out) const {
    out.keep_raw = (filter & kKeepRaw) != 0;
    BgzfReader r(fd_);
    std::vector<uint8_t> rec;
    auto keep = [&](const RecordView& v) {
        if ((filter & kSkipSupplementary) && (v.flag & 0x800)) return false;
        if ((filter & kSkipUnmappedSecondary) && (v.flag & (0x4 | 0x100))) return false;
        if ((filter & kOnlyUnmapped) && !(v.flag & 0x4)) return false;
        return true;
    };
    auto next = [&](uint32_t& bs) {
        uint8_t w[4];
        if (!r.read(w, 4)) return false;
        bs = le32(w);
        rec.resize(bs);
        if (bs && !r.read(rec.data(), bs)) fail("BAM stream ends inside a record: " + path_);
        return true;
    };
    if (tid == kFetchAll || tid == kFetchUnmapped) {
        uint64_t start = first_record_voff_;
        if (tid == kFetchUnmapped) {
            if (!has_index_) fail("alignment file is not indexed: " + path_);
            // the unplaced reads sort behind every mapped read: the end offset of the last reference that has records
            uint64_t off0 = ~0ull, max_end = 0;
            for (size_t i = index_.refs.size(); i-- > 0 && off0 == ~0ull;)
                if (index_.refs[i].has_meta) off0 = index_.refs[i].meta_end;
            if (off0 == ~0ull) {
                for (const auto& R : index_.refs)
                    for (const auto& kv : R.bins)
                        for (const auto& c : kv.second) max_end = std::max(max_end, c.second);
                off0 = max_end ? max_end : first_record_voff_;
            }
            start = off0;
        }
        r.seek(start);
        uint32_t bs = 0;
        while (next(bs)) {
            const RecordView v = peek(rec.data(), bs);
            if (tid == kFetchUnmapped && v.tid >= 0) continue;  // (only the no-coordinate tail)
            if (keep(v)) out.append(rec.data(), bs);
        }
        out.finish();
        return;
    }
    if (!has_index_) fail("alignment file is not indexed: " + path_);
    if (tid < 0 || size_t(tid) >= index_.refs.size()) fail("fetch: reference index out of range");
    const BaiRef& R = index_.refs[size_t(tid)];
    std::vector<uint32_t> bins;
    reg2bins(begin, end, index_.min_shift, index_.depth, bins);
    uint64_t min_off = 0;
    if (index_.csi) {
        // the lower bound of the finest bin around `begin` that the index has, or of its nearest ancestor (each bin's
        // loffset is the offset of the first record that overlaps the bin's first window)
        const int64_t b0 = std::min<int64_t>(std::max<int64_t>(begin, 0), (int64_t(1) << (index_.min_shift + 3 * index_.depth)) - 1);
        uint32_t bin = ((1u << (3 * index_.depth)) - 1u) / 7u + uint32_t(b0 >> index_.min_shift);
        for (;;) {
            auto it = R.loffset.find(bin);
            if (it != R.loffset.end()) { min_off = it->second; break; }
            if (bin == 0) break;
            bin = (bin - 1u) >> 3;
        }
    } else if (!R.ioffset.empty()) {
        const size_t w = size_t(std::max<int64_t>(begin, 0) >> 14);
        min_off = R.ioffset[std::min(w, R.ioffset.size() - 1)];
    }
    std::vector<std::pair<uint64_t, uint64_t>> chunks;
    for (uint32_t b : bins) {
        auto it = R.bins.find(b);
        if (it == R.bins.end()) continue;
        for (const auto& c : it->second)
            if (c.second > min_off) chunks.push_back(c);
    }
    std::sort(chunks.begin(), chunks.end());
    std::vector<std::pair<uint64_t, uint64_t>> merged;
    for (const auto& c : chunks) {
        if (!merged.empty() && c.first <= merged.back().second) merged.back().second = std::max(merged.back().second, c.second);
        else merged.push_back(c);
    }
    bool done = false;
    for (const auto& c : merged) {
        if (done) break;
        r.seek(c.first);
        uint32_t bs = 0;
        while (r.tell() < c.second && next(bs)) {
            const RecordView v = peek(rec.data(), bs);
            if (v.tid != tid || v.pos >= end) { done = true; break; }
            if (v.end <= begin) continue;
            if ((filter & kKeepStartInRegion) && !(v.pos >= begin && v.pos < end)) continue;
            if (keep(v)) out.append(rec.data(), bs);
        }
    }
    out.finish();
}

// ================================================================================================== BAI builder
namespace {
void build_index(const std::string& bam_path, const std::string& out_path, bool csi, int min_shift, int depth);
}
void build_bai(const std::string& bam_path, const std::string& bai_path) { build_index(bam_path, bai_path, false, 14, 5); }
void build_csi(const std::string& bam_path, const std::string& csi_path, int min_shift, int depth) {
    if (min_shift < 1 || min_shift > 30 || depth > 9) fail("build_csi: min_shift must be 1..30 and depth at most 9");
    build_index(bam_path, csi_path, true, min_shift, depth);
}
namespace {
void build_index(const std::string& bam_path, const std::string& bai_path, bool csi, int min_shift, int depth) {
    BamFile* f = BamFile::open(bam_path);
    struct Guard { BamFile* f; ~Guard() { delete f; } } g{f};
    const size_t n_ref = f->ref_names().size();
    struct Ref {
        std::unordered_map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
        std::vector<uint64_t> lin;
        uint64_t beg = ~0ull, end = 0, n_mapped = 0, n_unmapped = 0;
    };
    std::vector<Ref> refs(n_ref);
    uint64_t n_no_coor = 0;
    if (csi && depth <= 0) {
        uint64_t longest = 0;
        for (uint64_t l : f->ref_len()) longest = std::max(longest, l);
        for (depth = 1; depth < 9 && (uint64_t(1) << (min_shift + 3 * depth)) < longest; ++depth) {}
    }
    if (min_shift + 3 * depth > 62) fail("index geometry too large");
    const int64_t max_pos = int64_t(1) << (min_shift + 3 * depth);
    // walk the records with their virtual offsets
    const int fd = ::open(bam_path.c_str(), O_RDONLY);
    if (fd < 0) fail("cannot open " + bam_path);
    struct FdGuard { int fd; ~FdGuard() { ::close(fd); } } fg{fd};
    BgzfReader r(fd);
    {   // skip the header
        uint8_t m[8];
        r.read(m, 8);
        std::vector<uint8_t> skip(le32(m + 4));
        if (!skip.empty()) r.read(skip.data(), skip.size());
        uint8_t w[4];
        r.read(w, 4);
        for (uint32_t i = 0, n = le32(w); i < n; ++i) {
            r.read(w, 4);
            skip.resize(le32(w) + 4);
            r.read(skip.data(), skip.size());
        }
    }
    std::vector<uint8_t> rec;
    int32_t last_tid = 0;
    int64_t last_pos = -1;
    uint32_t last_bin = ~0u;
    int32_t last_bin_tid = -2;
    for (;;) {
        // (a record that starts exactly at a block end is addressed at the start of the next block, as htslib does)
        if (r.at_eof()) break;
        const uint64_t v0 = r.tell();
        uint8_t w[4];
        if (!r.read(w, 4)) break;
        const uint32_t bs = le32(w);
        rec.resize(bs);
        if (bs && !r.read(rec.data(), bs)) fail("BAM stream ends inside a record: " + bam_path);
        const uint64_t v1 = r.tell();
        const RecordView v = peek(rec.data(), bs);
        if (v.tid < 0) { ++n_no_coor; continue; }
        if (size_t(v.tid) >= n_ref) fail("record with a reference index outside the header");
        if (v.tid < last_tid || (v.tid == last_tid && v.pos < last_pos)) fail("alignment file is not coordinate-sorted: " + bam_path);
        last_tid = v.tid;
        last_pos = v.pos;
        Ref& R = refs[size_t(v.tid)];
        if (v.end > max_pos) fail(csi ? "a record lies beyond what this CSI geometry can index: raise depth" : "a record lies beyond 2^29: the BAI cannot index it, write a .csi");
        const uint32_t bin = reg2bin_levels(v.pos, v.end, min_shift, depth);
        auto& chunks = R.bins[bin];
        if (last_bin == bin && last_bin_tid == v.tid && !chunks.empty()) chunks.back().second = v1;
        else chunks.emplace_back(v0, v1);
        last_bin = bin;
        last_bin_tid = v.tid;
        const size_t w0 = size_t(v.pos >> min_shift), w1 = size_t((v.end - 1) >> min_shift);
        if (R.lin.size() <= w1) R.lin.resize(w1 + 1, 0);
        for (size_t k = w0; k <= w1; ++k)
            if (R.lin[k] == 0) R.lin[k] = v0;
        R.beg = std::min(R.beg, v0);
        R.end = std::max(R.end, v1);
        if (v.flag & 0x4) ++R.n_unmapped; else ++R.n_mapped;
    }
    std::vector<uint8_t> out;
    auto p32 = [&](uint32_t x) { for (int b = 0; b < 4; ++b) out.push_back(uint8_t(x >> (8 * b))); };
    auto p64 = [&](uint64_t x) { for (int b = 0; b < 8; ++b) out.push_back(uint8_t(x >> (8 * b))); };
    if (csi) {
        out.insert(out.end(), {'C', 'S', 'I', 1});
        p32(uint32_t(min_shift));
        p32(uint32_t(depth));
        p32(0);  // l_aux
    } else {
        out.insert(out.end(), {'B', 'A', 'I', 1});
    }
    p32(uint32_t(n_ref));
    const uint32_t meta_bin = meta_bin_of(depth);
    for (Ref& R : refs) {
        const bool any = R.beg != ~0ull;
        p32(uint32_t(R.bins.size() + (any ? 1 : 0)));
        std::vector<uint32_t> ids;
        for (const auto& kv : R.bins) ids.push_back(kv.first);
        std::sort(ids.begin(), ids.end());
        for (size_t k = 1; k < R.lin.size(); ++k)
            if (R.lin[k] == 0) R.lin[k] = R.lin[k - 1];
        for (uint32_t b : ids) {
            const auto& ch = R.bins[b];
            p32(b);
            if (csi) {  // loffset: the linear-index entry of the bin's first window
                int l = 0;
                uint32_t first = 0;
                while (l < depth && b >= first + (1u << (3 * l))) { first += 1u << (3 * l); ++l; }
                const uint64_t win = uint64_t(b - first) << (3 * (depth - l));
                p64(win < R.lin.size() ? R.lin[size_t(win)] : (R.lin.empty() ? 0 : R.lin.back()));
            }
            p32(uint32_t(ch.size()));
            for (const auto& c : ch) { p64(c.first); p64(c.second); }
        }
        if (any) {
            p32(meta_bin);
            if (csi) p64(0);
            p32(2); p64(R.beg); p64(R.end); p64(R.n_mapped); p64(R.n_unmapped);
        }
        if (!csi) {
            p32(uint32_t(R.lin.size()));
            for (uint64_t o : R.lin) p64(o);
        }
    }
    p64(n_no_coor);
    if (csi) {  // a .csi is a BGZF file
        std::vector<uint8_t> z(ptl_bgzf_bound(out.size()));
        const int64_t k = ptl_bgzf_compress(out.data(), out.size(), 6, 1, 1, z.data(), z.size());
        if (k < 0) fail("cannot compress the CSI index");
        z.resize(size_t(k));
        out.swap(z);
    }
    const int ofd = ::open(bai_path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (ofd < 0) fail("cannot write " + bai_path);
    size_t done = 0;
    while (done < out.size()) {
        const ssize_t k = ::write(ofd, out.data() + done, out.size() - done);
        if (k <= 0) { ::close(ofd); fail("write error: " + bai_path); }
        done += size_t(k);
    }
    ::close(ofd);
}
}  // namespace

// ================================================================================================== FASTA
std::vector<FastaRecord> read_fasta(const std::string& path, int n_threads) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) fail("Unable to open reference fasta file: '" + path + "'");
    struct stat st {};
    ::fstat(fd, &st);
    std::vector<uint8_t> b(size_t(st.st_size));
    const size_t got = pread_full(fd, b.data(), b.size(), 0);
    ::close(fd);
    if (got != b.size()) fail("read error: " + path);
    if (b.size() >= 2 && b[0] == 0x1f && b[1] == 0x8b) fail("compressed FASTA is not supported: " + path);
    // record boundaries: '>' at the start of a line
    std::vector<size_t> starts;
    for (size_t i = 0; i < b.size(); ++i)
        if (b[i] == '>' && (i == 0 || b[i - 1] == '\n')) starts.push_back(i);
    if (starts.empty() && !b.empty()) fail("Error during fasta record parsing: no '>' header in " + path);
    std::vector<FastaRecord> out(starts.size());
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= starts.size()) return;
            const size_t a = starts[k], e = (k + 1 < starts.size()) ? starts[k + 1] : b.size();
            size_t nl = a;
            while (nl < e && b[nl] != '\n') ++nl;
            size_t id_end = a + 1;
            while (id_end < nl && b[id_end] != ' ' && b[id_end] != '\t' && b[id_end] != '\r') ++id_end;
            FastaRecord& R = out[k];
            R.name.assign(reinterpret_cast<const char*>(b.data() + a + 1), id_end - a - 1);
            R.seq.reserve(e - nl);
            for (size_t i = nl + 1; i < e; ++i) {
                const uint8_t c = b[i];
                if (c == '\n' || c == '\r') continue;
                R.seq.push_back((c >= 'a' && c <= 'z') ? uint8_t(c - 32) : c);  // to_ascii_uppercase
            }
        }
    };
    const int nt = std::max(1, std::min<int>(n_threads, int(starts.size())));
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return out;
}

}  // namespace ptl
