// BAM container around the records of ptl_assemble_records (the "deflate out" half): the BAM header
// block and BGZF framing, host C++ with zlib.  The reference gets both from htslib (bam::Writer, src/read_alignment_scanner.rs:
// 537-559; `--threads` compression threads); this is the part of the output path that stays on the host (north star:
// "BGZF writing stays on the host").  No htslib: the format is written from the SAM specification (sections 4.1, 4.2).
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/portello_b200.h"

namespace {
constexpr uint64_t kBlockIn = 0xff00;  // payload bytes per BGZF block (htslib's BGZF_BLOCK_SIZE)
constexpr uint32_t kHeader = 18, kFooter = 8;
const uint8_t kEof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// one BGZF block: gzip member with the BC extra subfield carrying the block size (SAM spec 4.1)
int compress_block(const uint8_t* in, uint32_t n, int level, std::vector<uint8_t>& out) {
    out.resize(kHeader + compressBound(n) + kFooter + 64);
    z_stream zs{};
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
    zs.next_in = const_cast<uint8_t*>(in);
    zs.avail_in = n;
    zs.next_out = out.data() + kHeader;
    zs.avail_out = uInt(out.size() - kHeader - kFooter);
    const int rc = deflate(&zs, Z_FINISH);
    const uint32_t clen = uint32_t(zs.total_out);
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return -1;
    const uint32_t bsize = kHeader + clen + kFooter;
    if (bsize > 0x10000) return -2;  // (cannot happen for n <= 0xff00: deflate expands by < 0.1 % + 13 bytes)
    const uint8_t h[kHeader] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, uint8_t((bsize - 1) & 0xff), uint8_t((bsize - 1) >> 8)};
    std::memcpy(out.data(), h, kHeader);
    const uint32_t crc = uint32_t(crc32(crc32(0L, Z_NULL, 0), in, n));
    uint8_t* f = out.data() + kHeader + clen;
    for (int b = 0; b < 4; ++b) { f[b] = uint8_t(crc >> (8 * b)); f[4 + b] = uint8_t(n >> (8 * b)); }
    out.resize(bsize);
    return 0;
}
}  // namespace

extern "C" {

uint64_t ptl_bgzf_bound(uint64_t n) {
    const uint64_t blocks = (n + kBlockIn - 1) / kBlockIn;
    return blocks * 0x10000ull + sizeof(kEof);
}

int64_t ptl_bgzf_compress(const uint8_t* in, uint64_t n, int level, int n_threads, int append_eof, uint8_t* out, uint64_t cap) {
    if ((n && !in) || !out || level < 0 || level > 9) return -int64_t(PTL_ERR_INVALID_ARG);
    const uint64_t blocks = (n + kBlockIn - 1) / kBlockIn;
    std::vector<std::vector<uint8_t>> z(blocks);
    std::atomic<uint64_t> next{0};
    std::atomic<int> err{0};
    auto work = [&]() {
        for (;;) {
            const uint64_t b = next.fetch_add(1);
            if (b >= blocks) return;
            const uint64_t off = b * kBlockIn;
            if (compress_block(in + off, uint32_t(std::min<uint64_t>(kBlockIn, n - off)), level, z[b]) != 0) err = 1;
        }
    };
    const int nt = int(std::max<uint64_t>(1, std::min<uint64_t>(uint64_t(std::max(n_threads, 1)), blocks)));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (err) return -int64_t(PTL_ERR_INPUT);  // zlib failure
    uint64_t at = 0;
    for (const auto& b : z) {
        if (at + b.size() > cap) return -int64_t(PTL_ERR_INVALID_ARG);
        std::memcpy(out + at, b.data(), b.size());
        at += b.size();
    }
    if (append_eof) {
        if (at + sizeof(kEof) > cap) return -int64_t(PTL_ERR_INVALID_ARG);
        std::memcpy(out + at, kEof, sizeof(kEof));
        at += sizeof(kEof);
    }
    return int64_t(at);
}

int64_t ptl_bam_header(const char* sam_text, uint32_t n_ref, const char* const* ref_names, const uint64_t* ref_len, uint8_t* out, uint64_t cap) {
    if ((n_ref && (!ref_names || !ref_len)) || !out) return -int64_t(PTL_ERR_INVALID_ARG);
    std::vector<uint8_t> v;
    auto u32 = [&](uint32_t x) { for (int b = 0; b < 4; ++b) v.push_back(uint8_t(x >> (8 * b))); };
    const std::string text = sam_text ? sam_text : "";
    v.insert(v.end(), {'B', 'A', 'M', 1});
    u32(uint32_t(text.size()));
    v.insert(v.end(), text.begin(), text.end());
    u32(n_ref);
    for (uint32_t i = 0; i < n_ref; ++i) {
        if (!ref_names[i] || ref_len[i] > 0x7fffffffull) return -int64_t(PTL_ERR_INVALID_ARG);
        const std::string nm = ref_names[i];
        u32(uint32_t(nm.size() + 1));
        v.insert(v.end(), nm.begin(), nm.end());
        v.push_back(0);
        u32(uint32_t(ref_len[i]));
    }
    if (v.size() > cap) return -int64_t(PTL_ERR_INVALID_ARG);
    std::memcpy(out, v.data(), v.size());
    return int64_t(v.size());
}

}  // extern "C"
