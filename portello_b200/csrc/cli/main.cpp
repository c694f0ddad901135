// portello-b200: the command-line tool, portello's interface on top of libportello_b200.so.
//
// Same flags, inputs, outputs, messages and exit codes as the reference's binary (src/cli.rs:8-75,86-170; src/main.rs:24-122;
// src/logger.rs:5-26):
//   portello-b200 --assembly-to-ref contigs_to_ref.bam --read-to-assembly reads_to_contigs.bam --ref ref.fa
//                 --remapped-read-output out.bam|- --unassembled-read-output unassembled.bam [--threads N]
// Everything goes through the public C-ABI (include/portello_b200.h), the way a Rust host would drive it:
//   phase A  ptl_fasta_load + length checks (get_chrom_array, main.rs:24-62), ptl_scan_contig_bam (scan_contig_bam's record
//            loop), ptl_set_contig_records (primary / supplementary assembly, trim, join, device segment tables)
//   phase B  one task per (contig x <= 20 Mb window) (read_alignment_scanner.rs:495-535,606-660) on --threads worker threads,
//            each owning a GPU batch slot: ptl_bam_fetch (records that start in the window, supplementary skipped) ->
//            ptl_pack_batch -> ptl_lift_submit / ptl_lift_wait -> ptl_assemble_records (ready BAM records) ->
//            BGZF (level 0 on the device for `-`, the reference's pipe mode :66-71; host zlib otherwise) -> the shared writer
//   unmapped reads pass through unchanged into the unassembled output (scan_unmapped_reads, :537-559)
// Extra flags of this build: --gpu N (device index), --batch-reads N (reads per GPU batch).
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../../include/portello_b200.h"

namespace {

constexpr const char* kName = "portello-b200";
constexpr const char* kVersion = "0.2.0";
constexpr int EX_USAGE_ = 64, EX_DATAERR_ = 65, EX_PANIC_ = 101;

std::mutex g_log_mu;
void log_line(const char* level, const std::string& msg) {
    // fern format of src/logger.rs:11-19: [%Y-%m-%d][%H:%M:%S][name][LEVEL] message
    char ts[64];
    const std::time_t t = std::time(nullptr);
    std::tm tmv{};
    localtime_r(&t, &tmv);
    std::strftime(ts, sizeof ts, "[%Y-%m-%d][%H:%M:%S]", &tmv);
    std::lock_guard<std::mutex> lk(g_log_mu);
    std::fprintf(stderr, "%s[%s][%s] %s\n", ts, kName, level, msg.c_str());
}
void info(const std::string& m) { log_line("INFO", m); }
// PORTELLO_B200_TIMING=1: phase boundaries with the elapsed time, at DEBUG level (not part of the reference's log)
const std::chrono::steady_clock::time_point g_t0 = std::chrono::steady_clock::now();
void phase(const char* what) {
    static const bool on = std::getenv("PORTELLO_B200_TIMING") != nullptr;
    if (!on) return;
    char b[160];
    std::snprintf(b, sizeof b, "%s at +%.3f s", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - g_t0).count());
    log_line("DEBUG", b);
}
void error(const std::string& m) { log_line("ERROR", m); }
[[noreturn]] void panic(const std::string& m) {
    std::fprintf(stderr, "%s: fatal: %s\n", kName, m.c_str());
    std::exit(EX_PANIC_);
}
[[noreturn]] void usage_error(const std::string& m) {
    std::fprintf(stderr, "Invalid command-line setting: %s\n", m.c_str());
    std::exit(EX_USAGE_);
}

struct Settings {
    std::string assembly_to_ref, read_to_assembly, remapped_out, unassembled_out, ref, target_region;
    int threads = 0, gpu = 0;
    uint32_t batch_reads = 16384;
};

const char* kHelp =
    "portello-b200 %s\n"
    "Transfer HiFi read alignments from assembly contigs onto the reference genome (B200-native liftover path)\n\n"
    "Usage: portello-b200 [OPTIONS] --assembly-to-ref <FILE> --read-to-assembly <FILE> --remapped-read-output <FILE> "
    "--unassembled-read-output <FILE> --ref <FILE>\n\n"
    "Options:\n"
    "      --assembly-to-ref <FILE>          Assembly contig to reference genome alignment file in BAM format (sorted, indexed)\n"
    "      --read-to-assembly <FILE>         Read to assembly alignment file in BAM format (sorted, indexed)\n"
    "      --remapped-read-output <FILE>     Filename to use for remapped read output, or '-' for stdout (uncompressed BAM)\n"
    "      --unassembled-read-output <FILE>  Filename to use for unmapped reads which are not (well) mapped to any assembly contig\n"
    "      --ref <FILE>                      Genome reference in FASTA format\n"
    "      --target-region <REGION>          (debugging option of the reference; not supported by this build)\n"
    "      --threads <THREAD_COUNT>          Number of threads to use. Defaults to all logical cpus detected\n"
    "      --gpu <INDEX>                     CUDA device to use [default: 0]\n"
    "      --batch-reads <N>                 Reads per GPU batch [default: 16384]\n"
    "  -h, --help                            Print help\n"
    "  -V, --version                         Print version\n";

Settings parse(int argc, char** argv) {
    Settings s;
    std::string threads_opt;
    auto value = [&](int& i, const std::string& flag) -> std::string {
        const std::string a = argv[i];
        const size_t eq = a.find('=');
        if (eq != std::string::npos) return a.substr(eq + 1);
        if (i + 1 >= argc) {
            std::fprintf(stderr, "error: a value is required for '%s <FILE>' but none was supplied\n", flag.c_str());
            std::exit(2);
        }
        return argv[++i];
    };
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        const std::string key = a.substr(0, a.find('='));
        if (key == "--assembly-to-ref") s.assembly_to_ref = value(i, key);
        else if (key == "--read-to-assembly") s.read_to_assembly = value(i, key);
        else if (key == "--remapped-read-output") s.remapped_out = value(i, key);
        else if (key == "--unassembled-read-output") s.unassembled_out = value(i, key);
        else if (key == "--ref") s.ref = value(i, key);
        else if (key == "--target-region") s.target_region = value(i, key);
        else if (key == "--threads") threads_opt = value(i, key);
        else if (key == "--gpu") s.gpu = std::atoi(value(i, key).c_str());
        else if (key == "--batch-reads") s.batch_reads = uint32_t(std::max(1, std::atoi(value(i, key).c_str())));
        else if (key == "-h" || key == "--help") { std::printf(kHelp, kVersion); std::exit(0); }
        else if (key == "-V" || key == "--version") { std::printf("%s %s\n", kName, kVersion); std::exit(0); }
        else { std::fprintf(stderr, "error: unexpected argument '%s' found\n", a.c_str()); std::exit(2); }
    }
    for (const auto& req : {std::make_pair(&s.assembly_to_ref, "--assembly-to-ref <FILE>"), std::make_pair(&s.read_to_assembly, "--read-to-assembly <FILE>"),
                            std::make_pair(&s.remapped_out, "--remapped-read-output <FILE>"),
                            std::make_pair(&s.unassembled_out, "--unassembled-read-output <FILE>"), std::make_pair(&s.ref, "--ref <FILE>")})
        if (req.first->empty()) {
            std::fprintf(stderr, "error: the following required arguments were not provided:\n  %s\n", req.second);
            std::exit(2);
        }
    // validate_and_fix_settings (cli.rs:86-141)
    auto exists = [](const std::string& p) { return ::access(p.c_str(), F_OK) == 0; };
    auto check_in = [&](const std::string& f, const char* label) {
        if (!exists(f)) usage_error(std::string("Can't find specified ") + label + " file: '" + f + "'");
    };
    check_in(s.assembly_to_ref, "contig-to-ref bam");
    check_in(s.read_to_assembly, "read-to-contig bam");
    check_in(s.ref, "reference fasta");
    auto check_out = [&](const std::string& f, const char* label) {
        const size_t slash = f.rfind('/');
        const std::string parent = slash == std::string::npos ? "" : f.substr(0, slash);
        if (!parent.empty() && !exists(parent)) usage_error(std::string("Can't find existing directory for ") + label + " file: '" + f + "'");
    };
    if (s.remapped_out != "-") check_out(s.remapped_out, "remapped read output");
    check_out(s.unassembled_out, "unassembled read output");
    if (!threads_opt.empty()) {
        s.threads = std::atoi(threads_opt.c_str());
        if (s.threads <= 0) usage_error("--threads argument must be greater than 0");
    } else {
        s.threads = int(std::max(1u, std::thread::hardware_concurrency()));
    }
    if (!s.target_region.empty()) usage_error("--target-region is a debugging option of the reference and is not supported by this build");
    return s;
}

// assert_mapped_and_indexed_bam (cli.rs:143-164)
ptl_bam_file* open_checked(const std::string& path) {
    ptl_bam_file* f = nullptr;
    if (ptl_bam_open(path.c_str(), &f) != PTL_OK) panic(std::string("Failed to open input alignment file: ") + ptl_bam_last_error());
    if (!ptl_bam_has_index(f)) panic("Failed to open input alignment file: no index found for '" + path + "'");
    if (!ptl_bam_has_eof_marker(f)) panic("alignment file is missing its EOF marker (truncated?): '" + path + "'");
    if (ptl_bam_n_ref(f) == 0) panic("Input alignment file is not mapped: '" + path + "'");
    return f;
}

struct Writer {
    FILE* fp = nullptr;
    std::mutex mu;
    bool to_stdout = false;
    uint64_t n_records = 0;
    void write(const uint8_t* p, uint64_t n) {
        if (n && std::fwrite(p, 1, n, fp) != n) panic("write error on an output file");
    }
};

std::vector<uint8_t> bgzf(const uint8_t* p, uint64_t n, int level, int threads, bool eof) {
    std::vector<uint8_t> out(ptl_bgzf_bound(n));
    const int64_t k = ptl_bgzf_compress(p, n, level, threads, eof ? 1 : 0, out.data(), out.size());
    if (k < 0) panic("BGZF compression failed");
    out.resize(size_t(k));
    return out;
}

// get_alignment_file_header (read_alignment_scanner.rs:35-59) as a BGZF-framed header block
std::vector<uint8_t> output_header(const std::vector<std::string>& names, const std::vector<uint64_t>& lens, const std::string& cmdline, int level) {
    std::string text = "@HD\tVN:1.6\tSO:unsorted\n";
    for (size_t i = 0; i < names.size(); ++i) text += "@SQ\tSN:" + names[i] + "\tLN:" + std::to_string(lens[i]) + "\n";
    text += std::string("@PG\tPN:") + kName + "\tID:" + kName + "-" + kVersion + "\tVN:" + kVersion + "\tCL:" + cmdline + "\n";
    std::vector<const char*> np;
    for (const auto& n : names) np.push_back(n.c_str());
    std::vector<uint8_t> raw(text.size() + 64 + names.size() * 300);
    const int64_t k = ptl_bam_header(text.c_str(), uint32_t(names.size()), np.data(), lens.data(), raw.data(), raw.size());
    if (k < 0) panic("BAM header does not fit");
    return bgzf(raw.data(), uint64_t(k), level, 1, false);
}

std::string hhmmssxxx(double sec) {
    const long ms = long(sec * 1000.0 + 0.5);
    char b[64];
    std::snprintf(b, sizeof b, "%02ld:%02ld:%02ld.%03ld", ms / 3600000, (ms / 60000) % 60, (ms / 1000) % 60, ms % 1000);
    return b;
}

}  // namespace

int main(int argc, char** argv) {
    const Settings st = parse(argc, argv);
    // validate_settings_data (cli.rs:167-170)
    ptl_bam_file* contig_bam = open_checked(st.assembly_to_ref);
    ptl_bam_file* read_bam = open_checked(st.read_to_assembly);

    std::string cmdline;
    for (int i = 0; i < argc; ++i) cmdline += (i ? " " : "") + std::string(argv[i]);
    info(std::string("Starting ") + kName + " " + kVersion);
    info("cmdline: " + cmdline);
    info("Running on " + std::to_string(st.threads) + " threads");
    const auto t_start = std::chrono::steady_clock::now();

    // ChromList::from_bam_filename x 2 (main.rs:77-78)
    std::vector<std::string> ref_names, contig_names;
    std::vector<uint64_t> ref_len, contig_len;
    for (uint32_t i = 0; i < ptl_bam_n_ref(contig_bam); ++i) { ref_names.push_back(ptl_bam_ref_name(contig_bam, i)); ref_len.push_back(ptl_bam_ref_len(contig_bam, i)); }
    for (uint32_t i = 0; i < ptl_bam_n_ref(read_bam); ++i) { contig_names.push_back(ptl_bam_ref_name(read_bam, i)); contig_len.push_back(ptl_bam_ref_len(read_bam, i)); }
    std::vector<const char*> ref_name_p, contig_name_p;
    for (const auto& n : ref_names) ref_name_p.push_back(n.c_str());
    for (const auto& n : contig_names) contig_name_p.push_back(n.c_str());

    // Worker threads decode, pack, compress and write (the host-bound stages: one per --threads, as the reference's rayon
    // pool); the GPU part of a batch borrows one of a few batch slots, which are busy for milliseconds at a time.
    const int n_workers = std::max(1, st.threads);
    const int n_slots = std::min(n_workers, 4);
    // (the CUDA context takes seconds to come up on a cold driver: it does so behind the reference and contig file reads)
    ptl_ctx* ctx = nullptr;
    int rc_create = PTL_OK;
    std::thread ctx_thread([&]() { rc_create = ptl_create(st.gpu, n_slots, &ctx); });

    // get_chrom_array (main.rs:24-62)
    info("Reading reference genome from file '" + st.ref + "'");
    ptl_fasta* fasta = nullptr;
    if (ptl_fasta_load(st.ref.c_str(), st.threads, &fasta) != PTL_OK) panic(ptl_bam_last_error());
    phase("reference loaded");
    std::unordered_map<std::string, uint32_t> fasta_index;
    for (uint32_t i = 0; i < ptl_fasta_n(fasta); ++i) fasta_index.emplace(ptl_fasta_name(fasta, i), i);
    std::vector<const uint8_t*> chrom_seq;
    bool consistency_error = false;
    for (size_t c = 0; c < ref_names.size(); ++c) {
        auto it = fasta_index.find(ref_names[c]);
        if (it == fasta_index.end()) {
            error("Chromosome \"" + ref_names[c] + "\" specified in the assembly-to-ref alignment file, but not in the reference fasta");
            consistency_error = true;
            continue;
        }
        uint64_t n = 0;
        const uint8_t* p = ptl_fasta_seq(fasta, it->second, &n);
        if (n != ref_len[c]) {
            error("Chromosome \"" + ref_names[c] + "\" specified with inconsistent length: " + std::to_string(ref_len[c]) +
                  " in the assembly-to-ref alignment file, and " + std::to_string(n) + " in the reference fasta");
            consistency_error = true;
            continue;
        }
        chrom_seq.push_back(p);
    }
    if (consistency_error) {
        error("Exiting due to one or more reference consistency issues");
        ctx_thread.join();
        return EX_DATAERR_;
    }

    // ---- phase A: scan_contig_bam (contig_alignment_scanner/mod.rs:290-459)
    info("Processing contig-to-ref alignment file '" + st.assembly_to_ref + "'");
    ptl_contig_scan* scan = nullptr;
    if (ptl_scan_contig_bam(contig_bam, uint32_t(contig_names.size()), contig_name_p.data(), contig_len.data(), st.threads, &scan) != PTL_OK)
        panic(ptl_bam_last_error());
    phase("contig alignment file scanned");
    ptl_contig_records crecs{};
    ptl_contig_scan_view(scan, &crecs);
    ctx_thread.join();
    if (rc_create == PTL_ERR_NO_DEVICE) panic("no usable sm_100 CUDA device: the liftover path has no CPU fallback");
    if (rc_create != PTL_OK) panic("ptl_create failed");
    phase("device context created");
    if (ptl_set_reference(ctx, uint32_t(ref_names.size()), ref_len.data(), chrom_seq.data()) != PTL_OK) panic(ptl_last_error(ctx));
    phase("reference on the device");
    info("Clipping repeated contig matches at split alignment segment boundaries");
    info("Joining colinear split alignment segments in each assembly contig");
    if (ptl_set_contig_records(ctx, &crecs) != PTL_OK) panic(ptl_last_error(ctx));
    if (ptl_set_names(ctx, uint32_t(contig_names.size()), contig_name_p.data(), uint32_t(ref_names.size()), ref_name_p.data()) != PTL_OK) panic(ptl_last_error(ctx));
    phase("contig segments and tables built");
    ptl_contig_scan_free(scan);
    ptl_fasta_free(fasta);  // (the reference bases live on the device now)

    // ---- phase B: scan_and_remap_reads (read_alignment_scanner.rs:566-661)
    info("Processing read-to-contig alignment file '" + st.read_to_assembly + "'");
    const int writer_threads = std::max(1, st.threads / 2);  // (the unmapped pass-through: one big block of records)
    Writer remapped, unassembled;
    remapped.to_stdout = st.remapped_out == "-";
    remapped.fp = remapped.to_stdout ? stdout : std::fopen(st.remapped_out.c_str(), "wb");
    unassembled.fp = std::fopen(st.unassembled_out.c_str(), "wb");
    if (!remapped.fp || !unassembled.fp) panic("cannot open an output file for writing");
    {
        const std::vector<uint8_t> h0 = output_header(ref_names, ref_len, cmdline, remapped.to_stdout ? 0 : 6);
        remapped.write(h0.data(), h0.size());
        const std::vector<uint8_t> h1 = output_header(ref_names, ref_len, cmdline, 6);
        unassembled.write(h1.data(), h1.size());
    }
    // Work units: (contig x window), records assigned by their start (:403-406, :495-535).  The reference's window is 20 Mb;
    // an assembly too small to give every worker a few of those is cut finer (any partition by start position yields the
    // same records; the output is unsorted either way) ...
    uint64_t total_len = 0;
    for (uint64_t l : contig_len) total_len += l;
    // ... and capped at 4 Mb: a worker holds the decoded records of its window (~45 KB per 15 kb read, 30x coverage: 0.9 GB
    // per 20 Mb), and there is one worker per thread.
    const uint64_t window = std::min<uint64_t>(4000000ull, std::max<uint64_t>(1000000ull, total_len / (4ull * uint64_t(n_workers)) + 1));
    struct Unit { uint32_t contig; uint64_t b, e; };
    std::vector<Unit> units;
    for (uint32_t c = 0; c < contig_len.size(); ++c) {
        const uint32_t n = ptl_region_segment_count(contig_len[c], window);
        std::vector<uint64_t> b(n), e(n);
        if (n) ptl_region_segments(contig_len[c], window, b.data(), e.data());
        for (uint32_t k = 0; k < n; ++k) units.push_back(Unit{c, b[k], e[k]});
    }
    std::atomic<size_t> next_unit{0};
    std::atomic<uint64_t> n_reads_done{0}, n_pairs{0}, n_lifted{0};
    struct SlotPool {
        std::mutex mu;
        std::condition_variable cv;
        std::vector<int> free_slots;
        int acquire() {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return !free_slots.empty(); });
            const int s = free_slots.back();
            free_slots.pop_back();
            return s;
        }
        void release(int s) {
            { std::lock_guard<std::mutex> lk(mu); free_slots.push_back(s); }
            cv.notify_one();
        }
    } slots;
    for (int k = 0; k < n_slots; ++k) slots.free_slots.push_back(k);
    auto worker = [&]() {
        std::vector<uint8_t> rec_copy;
        for (;;) {
            const size_t u = next_unit.fetch_add(1);
            if (u >= units.size()) return;
            ptl_decoded_batch* dec = nullptr;
            // records that start in the window, supplementary skipped (:403-406)
            if (ptl_bam_fetch(read_bam, int32_t(units[u].contig), int64_t(units[u].b), int64_t(units[u].e), PTL_BAM_START_IN_REGION | PTL_BAM_SKIP_SUPPLEMENTARY, &dec) != PTL_OK)
                panic(std::string("Failed to parse alignment record: ") + ptl_bam_last_error());
            ptl_read_records recs{};
            ptl_read_extras extras{};
            ptl_decoded_view(dec, &recs, &extras);
            for (uint32_t r = 0; r < recs.n_reads; ++r)
                if (recs.flag[r] & 0x4) panic("assertion failed: !record.is_unmapped()");  // (:396)
            for (uint32_t first = 0; first < recs.n_reads; first += st.batch_reads) {
                const uint32_t count = std::min(st.batch_reads, recs.n_reads - first);
                ptl_packed_batch* pk = nullptr;
                if (ptl_pack_batch(&recs, first, count, uint32_t(contig_names.size()), contig_name_p.data(), 0, &pk) != PTL_OK) panic(ptl_pack_last_error());
                ptl_batch batch{};
                ptl_packed_batch_view(pk, &batch);
                ptl_result res{};
                const int slot = slots.acquire();
                if (ptl_lift_submit(ctx, slot, &batch) != PTL_OK) panic(ptl_last_error(ctx));
                const int rc = ptl_lift_wait(ctx, slot, &res);
                if (rc == PTL_ERR_LIFT_PANIC)
                    panic("read " + std::to_string(res.first_error_read) + " of a batch: the lifted alignment is inconsistent with the read (status " +
                          std::to_string(res.first_error_status) + "; the reference panics at src/read_alignment_scanner.rs:204-229)");
                if (rc != PTL_OK) panic(ptl_last_error(ctx));
                // the extras of this sub-batch: offsets are absolute into the window's pools, so a shifted view is enough
                ptl_read_extras x = extras;
                x.name_off += first; x.aux_off += first; x.mate_tid += first; x.mate_pos += first; x.tlen += first; x.quals.read_qual_off += first;
                ptl_bam_records out{};
                if (ptl_assemble_records(ctx, slot, &x, remapped.to_stdout ? PTL_ASM_NO_DOWNLOAD : 0u, &out) != PTL_OK) panic(ptl_last_error(ctx));
                if (remapped.to_stdout) {  // the reference's pipe mode: uncompressed BAM, framed (and CRC'd) on the device
                    ptl_bgzf_stream z{};
                    if (ptl_bgzf_store_records(ctx, slot, nullptr, 0, 0, &z) != PTL_OK) panic(ptl_last_error(ctx));
                    {
                        std::lock_guard<std::mutex> lk(remapped.mu);
                        remapped.write(z.bytes, z.n_bytes);
                        remapped.n_records += out.n_records;
                    }
                    n_pairs += res.n_pairs;
                    n_lifted += res.n_lifted;
                    slots.release(slot);
                } else {
                    // the records leave the slot's pinned buffer before the slow part (deflate, one thread per worker)
                    const uint64_t total = out.n_records ? out.rec_begin[out.n_records] : 0;
                    const uint64_t n_out = out.n_records;
                    rec_copy.assign(out.bytes, out.bytes + total);
                    n_pairs += res.n_pairs;
                    n_lifted += res.n_lifted;
                    slots.release(slot);
                    const std::vector<uint8_t> z = bgzf(rec_copy.data(), total, 6, 1, false);
                    std::lock_guard<std::mutex> lk(remapped.mu);  // all records of a read are written together (:482-487)
                    remapped.write(z.data(), z.size());
                    remapped.n_records += n_out;
                }
                n_reads_done += count;
                ptl_packed_batch_free(pk);
            }
            ptl_decoded_free(dec);
        }
    };
    std::thread unmapped_thread([&]() {  // scan_unmapped_reads (:537-559): the records pass through unchanged
        ptl_decoded_batch* dec = nullptr;
        if (ptl_bam_fetch(read_bam, PTL_FETCH_UNMAPPED, 0, 0, PTL_BAM_ONLY_UNMAPPED | PTL_BAM_KEEP_RAW, &dec) != PTL_OK)
            panic(std::string("Failed to parse alignment record: ") + ptl_bam_last_error());
        const uint64_t* off = nullptr;
        uint64_t nb = 0;
        const uint8_t* raw = ptl_decoded_raw(dec, &off, &nb);
        ptl_read_records r{};
        ptl_decoded_view(dec, &r, nullptr);
        const std::vector<uint8_t> z = bgzf(raw, nb, 6, writer_threads, false);
        std::lock_guard<std::mutex> lk(unassembled.mu);
        unassembled.write(z.data(), z.size());
        unassembled.n_records += r.n_reads;
        ptl_decoded_free(dec);
    });
    std::vector<std::thread> pool;
    for (int w = 0; w < n_workers; ++w) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
    phase("mapped reads done");
    unmapped_thread.join();
    phase("unmapped reads done");
    {
        const std::vector<uint8_t> eof = bgzf(nullptr, 0, 6, 1, true);
        remapped.write(eof.data(), eof.size());
        unassembled.write(eof.data(), eof.size());
    }
    if (remapped.fp != stdout) std::fclose(remapped.fp); else std::fflush(stdout);
    std::fclose(unassembled.fp);
    info("Lifted " + std::to_string(n_lifted.load()) + " of " + std::to_string(n_pairs.load()) + " read-segment x contig-segment alignments of " +
         std::to_string(n_reads_done.load()) + " reads into " + std::to_string(remapped.n_records) + " records; " + std::to_string(unassembled.n_records) +
         " unmapped reads passed through");
    ptl_destroy(ctx);
    ptl_bam_close(contig_bam);
    ptl_bam_close(read_bam);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    info(std::string(kName) + " completed. Total Runtime: " + hhmmssxxx(sec));
    return 0;
}
