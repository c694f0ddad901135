// BAM / BGZF / BAI / CSI input without htslib (input half), host C++ + zlib.
//
// What the reference gets from rust-htslib on its input side:
//   bam::IndexedReader::from_path + fetch(Region | Unmapped) + read   src/read_alignment_scanner.rs:382-393,537-559
//                                                                     src/contig_alignment_scanner/mod.rs:196-203
//   one reader per worker thread                                       src/worker_thread_data.rs:8-30
//   header -> ChromList                                                lib/rust-vc-utils/src/chrom_list.rs:26-44
//   EOF-marker check                                                   lib/rust-vc-utils/src/bam_utils/bam_reader_utils.rs:29
// Written from the SAM specification (sections 4.1 BGZF, 4.2 BAM, 5.2 BAI).  A BamFile is immutable after open(): any
// number of threads may fetch from it concurrently (every fetch uses its own BgzfReader over pread()).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../../include/portello_b200.h"

namespace ptl {

struct BamIoError {
    std::string msg;
};

// ---- BGZF (SAM spec 4.1): random-access reader over virtual offsets (compressed block offset << 16 | offset in block)
class BgzfReader {
public:
    explicit BgzfReader(int fd) : fd_(fd) {}
    void seek(uint64_t voffset);
    uint64_t tell() const { return (block_off_ << 16) | uint64_t(pos_); }
    // false at a clean end of file (before the first byte); throws on truncation / corruption
    bool read(void* dst, size_t n);
    bool at_eof();
    size_t read_some(void* dst, size_t n);  // up to n bytes of the current block (0 at the end of the file)

private:
    bool load(uint64_t coffset);  // inflate the block at coffset; false at end of file
    int fd_;
    uint64_t block_off_ = 0, next_off_ = 0;
    uint32_t pos_ = 0;
    bool loaded_ = false;
    std::vector<uint8_t> raw_, buf_;
};

// ---- BAI (SAM spec 5.2)
struct BaiRef {
    std::unordered_map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;  // bin -> chunks (virtual offsets)
    std::vector<uint64_t> ioffset;                                                  // 16 kb linear index (BAI)
    std::unordered_map<uint32_t, uint64_t> loffset;                                 // per-bin lower bound (CSI)
    bool has_meta = false;
    uint64_t meta_beg = 0, meta_end = 0, n_mapped = 0, n_unmapped = 0;              // pseudo-bin 37450
};
struct BaiIndex {
    int min_shift = 14, depth = 5;  // BAI: fixed; CSI: from the file (bins of 2^min_shift .. 2^(min_shift + 3 depth) bases)
    bool csi = false;
    std::vector<BaiRef> refs;
    bool has_no_coor = false;
    uint64_t n_no_coor = 0;
};

// ---- decoded records: the SoA the packer (ptl_read_records) and the record assembly (ptl_read_extras) take
struct DecodedBatch {
    std::vector<int32_t> tid, mate_tid, mate_pos, tlen;
    std::vector<int64_t> pos;
    std::vector<uint16_t> flag, bin;
    std::vector<uint8_t> mapq;
    std::vector<uint32_t> seq_len;
    std::vector<uint64_t> seq_off, cigar_begin{0}, name_off{0}, aux_off{0}, qual_off, raw_off{0};
    std::vector<uint32_t> cigar;
    std::vector<uint8_t> seq4, names, aux, qual, raw;  // raw: block_size + record bytes, only when keep_raw
    std::vector<int64_t> sa_at;                         // offset of the SA:Z value in `aux`, -1 if absent
    std::vector<const char*> sa_tag;                    // resolved by finish()
    bool keep_raw = false;
    // `rec`: the record WITHOUT its block_size prefix (block_size bytes)
    void append(const uint8_t* rec, uint32_t block_size);
    void finish();  // pads the pools (the kernels read a few bytes around a field) and resolves sa_tag
    uint32_t size() const { return uint32_t(tid.size()); }
    void view(ptl_read_records* recs, ptl_read_extras* extras) const;
};

enum : int32_t { kFetchAll = -2, kFetchUnmapped = -1 };  // tid values of fetch()
enum : uint32_t {
    kKeepStartInRegion = 1,    // only records whose pos lies in [begin, end)  (:403-406, mod.rs:213-217)
    kSkipSupplementary = 2,    // (:404)
    kSkipUnmappedSecondary = 4,  // (mod.rs:208)
    kOnlyUnmapped = 8,         // (:550-552)
    kKeepRaw = 16,             // keep the undecoded record bytes (unmapped pass-through)
    kDecodeSeqAscii = 32       // (unused flag slot; ASCII decode is done by the caller where needed)
};

class BamFile {
public:
    static BamFile* open(const std::string& path);  // throws BamIoError
    ~BamFile();
    const std::string& header_text() const { return text_; }
    const std::vector<std::string>& ref_names() const { return ref_names_; }
    const std::vector<uint64_t>& ref_len() const { return ref_len_; }
    bool has_index() const { return has_index_; }
    bool has_eof_marker() const { return has_eof_; }
    // Records overlapping [begin, end) of reference `tid` in file order = IndexedReader::fetch(Region) + read loop;
    // tid = kFetchUnmapped: the unplaced reads at the end of the file (FetchDefinition::Unmapped); kFetchAll: every record.
    void fetch(int32_t tid, int64_t begin, int64_t end, uint32_t filter, DecodedBatch& out) const;

private:
    int fd_ = -1;
    std::string path_, text_;
    std::vector<std::string> ref_names_;
    std::vector<uint64_t> ref_len_;
    uint64_t first_record_voff_ = 0;
    bool has_index_ = false, has_eof_ = false;
    BaiIndex index_;
};

// samtools-index equivalent: scan `bam_path` and write a .bai next to it (or to bai_path).
void build_bai(const std::string& bam_path, const std::string& bai_path);
// the same as a .csi (CSIv1: BGZF-compressed, bins of 2^min_shift bases at the finest of depth + 1 levels, a lower-bound
// offset per bin instead of the linear index); depth <= 0: the smallest depth that covers the longest reference
void build_csi(const std::string& bam_path, const std::string& csi_path, int min_shift, int depth);

// FASTA -> (name, upper-cased sequence) per record = get_genome_ref_from_fasta (lib/rust-vc-utils/src/genome_ref.rs:43-79):
// the id is the header up to the first whitespace, every base upper-cased, nothing else changed.
struct FastaRecord {
    std::string name;
    std::vector<uint8_t> seq;
};
std::vector<FastaRecord> read_fasta(const std::string& path, int n_threads);

}  // namespace ptl
