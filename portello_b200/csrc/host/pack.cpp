// Host packer + small host helpers of the product library (no CUDA in this file).
//
//   ptl_pack_split_segments / ptl_pack_batch : a2 — parse_sa_aux_val (lib/rust-vc-utils/src/bam_utils/aux/sa_tag_parser.rs:25-59)
//        + get_seq_order_read_split_segments (lib/rust-vc-utils/src/bam_utils/split_read.rs:56-155), and the record
//        filter of scan_chromosome_segment (src/read_alignment_scanner.rs:396-406).
//   ptl_format_sa_tags : a10 text part (src/read_alignment_scanner.rs:292-301,349-363)
//   ptl_region_segments / ptl_shard_units : a13 work units (lib/rust-vc-utils/src/util.rs:50-67) + LPT sharding
//   ptl_reg2bin : lib/rust-vc-utils/src/bam_utils/util.rs:10-35
#include <algorithm>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <unordered_map>

#include "../../../include/portello_b200.h"
#include "host_util.hpp"

using namespace ptl;

#ifndef PTL_PF_HINT
#define PTL_PF_HINT 3
#endif
#ifndef PTL_WIN_AHEAD
#define PTL_WIN_AHEAD 32
#endif

namespace {

thread_local std::string g_pack_err;

struct Seg {
    uint32_t so_start, so_end, contig;
    int64_t pos;
    uint8_t is_fwd, mapq, from_primary;
    uint32_t cig_off, cig_len;  // into a scratch pool
};

// strict decimal parse of [b,e) as Rust's str::parse would: optional sign, >=1 digit, nothing else
bool parse_i64(const char* b, const char* e, int64_t& out) {
    if (b == e) return false;
    bool neg = false;
    if (*b == '+' || *b == '-') { neg = (*b == '-'); ++b; }
    if (b == e) return false;
    int64_t v = 0;
    for (; b != e; ++b) {
        if (*b < '0' || *b > '9') return false;
        if (v > (INT64_MAX - 9) / 10) return false;
        v = v * 10 + (*b - '0');
    }
    out = neg ? -v : v;
    return true;
}

// CIGAR text -> ops appended to pool; false on syntax error (rust_htslib CigarString::try_from fails -> unwrap panic)
bool parse_cigar_text(const char* b, const char* e, Ops& pool) {
    static const char ops[] = "MIDNSHP=X";
    uint64_t n = 0;
    bool have = false;
    for (; b != e; ++b) {
        if (*b >= '0' && *b <= '9') {
            n = n * 10 + uint64_t(*b - '0');
            if (n > 0x0fffffffull) return false;
            have = true;
        } else {
            const char* p = static_cast<const char*>(std::memchr(ops, *b, 9));
            if (!p || !have) return false;
            pool.push_back(mk(uint32_t(p - ops), n));
            n = 0;
            have = false;
        }
    }
    return !have;
}

// Resolve contig names lazily: the map is built once per pack call.
struct NameMap {
    std::unordered_map<std::string, uint32_t> m;
    NameMap(uint32_t n, const char* const* names) {
        m.reserve(n * 2);
        for (uint32_t i = 0; i < n; ++i) m.emplace(names[i], i);
    }
};

// The segments of one primary record, sequencing order. `pool` receives SA CIGARs; the primary's CIGAR is referenced
// as (prim_off, n_cigar) in the caller's own pool and flagged by from_primary.
void split_segments(const NameMap& names, int32_t tid, int64_t pos, uint16_t flag, uint8_t mapq, const uint32_t* cigar,
                    uint32_t n_cigar, const char* sa, std::vector<Seg>& segs, Ops& pool) {
    segs.clear();
    const bool rev = flag & 0x10;
    const ClipPos pc = clip_positions(cigar, n_cigar);
    auto seq_order = [](const ClipPos& c, bool fwd, uint32_t& s, uint32_t& e) {
        if (fwd) { s = uint32_t(c.left); e = uint32_t(c.right); }
        else { s = uint32_t(c.size - c.right); e = uint32_t(c.size - c.left); }
    };
    Seg p{};
    seq_order(pc, !rev, p.so_start, p.so_end);
    p.contig = uint32_t(tid);
    p.pos = pos;
    p.is_fwd = !rev;
    p.mapq = mapq;
    p.from_primary = 1;
    p.cig_off = 0;
    p.cig_len = n_cigar;
    segs.push_back(p);
    if (sa) {
        const char* s = sa;
        const char* end = sa + std::strlen(sa);
        while (s < end) {
            const char* semi = static_cast<const char*>(std::memchr(s, ';', size_t(end - s)));
            const char* seg_end = semi ? semi : end;
            // split_terminator(','): a trailing empty field is dropped
            const char* f[8];
            const char* fe[8];
            int nf = 0;
            const char* q = s;
            while (true) {
                const char* comma = static_cast<const char*>(std::memchr(q, ',', size_t(seg_end - q)));
                if (!comma) {
                    if (q < seg_end) { if (nf < 8) { f[nf] = q; fe[nf] = seg_end; } ++nf; }
                    break;
                }
                if (nf < 8) { f[nf] = q; fe[nf] = comma; }
                ++nf;
                q = comma + 1;
            }
            if (nf != 6) throw InputError("Unexpected segment in bam SA tag: " + std::string(s, seg_end));
            int64_t sa_pos, sa_mapq, sa_nm;
            if (!parse_i64(f[1], fe[1], sa_pos)) throw InputError("SA tag: bad position");
            const bool fwd = (fe[2] - f[2] == 1 && *f[2] == '+');
            const uint32_t off = uint32_t(pool.size());
            if (!parse_cigar_text(f[3], fe[3], pool)) throw InputError("SA tag: bad CIGAR");
            if (!parse_i64(f[4], fe[4], sa_mapq) || sa_mapq < 0 || sa_mapq > 255) throw InputError("SA tag: bad MAPQ");
            if (!parse_i64(f[5], fe[5], sa_nm) || sa_nm < INT32_MIN || sa_nm > INT32_MAX) throw InputError("SA tag: bad NM");
            const uint32_t len = uint32_t(pool.size()) - off;
            bool aligned = false;
            for (uint32_t i = 0; i < len; ++i) aligned |= is_match(pool[off + i]);
            if (!aligned) throw InputError("Bam record split segment is unaligned");  // split_read.rs:107-110
            const ClipPos c = clip_positions(pool.data() + off, len);
            if (c.size != pc.size) throw InputError("SA segment read size differs from the primary record");  // :113
            auto it = names.m.find(std::string(f[0], fe[0]));
            if (it == names.m.end()) throw InputError("SA tag describes a split read mapped to an unknown contig: " + std::string(f[0], fe[0]));
            Seg g{};
            seq_order(c, fwd, g.so_start, g.so_end);
            g.contig = it->second;
            g.pos = sa_pos - 1;
            g.is_fwd = fwd;
            g.mapq = uint8_t(sa_mapq);
            g.from_primary = 0;
            g.cig_off = off;
            g.cig_len = len;
            segs.push_back(g);
            s = semi ? semi + 1 : end;
        }
        std::stable_sort(segs.begin(), segs.end(), [](const Seg& a, const Seg& b) { return a.so_start < b.so_start; });  // :140
    }
    for (const Seg& g : segs)
        if (g.so_start >= g.so_end) throw InputError("Can't parse consistent split read information from SA tag");  // :145-152
}

size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

// One forward walk of a CIGAR: clip positions as get_read_clip_positions (cigar/mod.rs:85-118), reference span, and --
// with kClusters -- the read offset at which every I/D cluster (a maximal run of I/D ops with at least one non-empty op,
// as CigarShiftBuilder forms them, cigar_indel_shifter.rs:63-85) begins, appended to `cluster_read_begin`.
struct CigarScan { uint64_t left, right, size; int64_t ref; uint32_t n_clusters; };
// Op kinds change unpredictably along a HiFi CIGAR, so the walk is branch-free; `cluster_read_begin` needs room for n entries.
template <bool kClusters>
inline CigarScan scan_cigar(const uint32_t* c, uint32_t n, uint32_t* cluster_read_begin) {
    constexpr uint32_t kIndelMask = (1u << OP_I) | (1u << OP_D);
    uint64_t left = 0, right = 0, size = 0, ref = 0;
    uint32_t in_left = 1, open = 0, n_cl = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t x = c[i], op = op_of(x), l = len_of(x);
        const uint32_t indel = (kIndelMask >> op) & 1u, clip = (kClipMask >> op) & 1u, nonempty = l != 0;
        if (kClusters) {
            cluster_read_begin[n_cl] = uint32_t(size);
            n_cl += indel & nonempty & (open ^ 1u);
            open = indel & (open | nonempty);
        }
        left += uint64_t(l) & (0 - uint64_t(clip & in_left));
        right += uint64_t(l) & (0 - uint64_t(clip & (in_left ^ 1u)));
        in_left &= clip;
        size += uint64_t(l) & (0 - uint64_t((kReadMask >> op) & 1u));
        ref += uint64_t(l) & (0 - uint64_t((kRefMask >> op) & 1u));
    }
    return {left, size - right, size, int64_t(ref), n_cl};
}

#if defined(__x86_64__) && defined(__GNUC__)
#define PTL_HAVE_AVX2_SCAN 1
// The same walk, eight ops per step (AVX2, chosen at run time): the three masked sums are plain vector sums; the cluster
// starts come from two bit masks per 64 ops (I/D ops, non-empty I/D ops) through a carry-lookahead of "an open cluster
// stays open across I/D ops", and their read offsets from an in-register prefix sum kept in `pre` (room for n + 8 words).
// 1.5x (22 ops) to 2.3x (70 ops) the scalar walk; identical results (tests/test_host_logic.py: random CIGARs of every op).
template <bool kClusters>
__attribute__((target("avx2")))
static CigarScan scan_cigar_avx2(const uint32_t* c, uint32_t n, uint32_t* cluster_read_begin, uint32_t* pre) {
    uint64_t left = 0, right = 0, size = 0, ref = 0;
    uint32_t i = 0, n_cl = 0;
    while (i < n && is_clip(c[i])) { left += len_of(c[i]); size += len_of(c[i]); ++i; }
    const __m256i zero = _mm256_setzero_si256(), one = _mm256_set1_epi32(1), two = _mm256_set1_epi32(2), m15 = _mm256_set1_epi32(15);
    const __m256i kread = _mm256_set1_epi32(int(kReadMask)), kref = _mm256_set1_epi32(int(kRefMask)), kclip = _mm256_set1_epi32(int(kClipMask));
    const __m256i lo32 = _mm256_set1_epi64x(0xffffffffll);
    const __m256i lane_idx = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7);
    const __m256i hi_half = _mm256_setr_epi32(0, 0, 0, 0, -1, -1, -1, -1), bc3 = _mm256_set1_epi32(3);
    __m256i a_size = zero, a_ref = zero, a_right = zero;
    uint32_t base32 = uint32_t(size);
    uint64_t Iw = 0, Nw = 0, carry = 0;
    uint32_t word_base = i, in_word = 0;
    auto flush_word = [&]() {
        if (!kClusters) return;
        uint64_t G = Nw | (Iw & carry), P = Iw;  // carry-in opens position 0 when it is an indel op
        // F_j = I_j & (N_j | F_{j-1}):  Kogge-Stone over "propagate through I"
        uint64_t F = G;
        uint64_t Pm = P;
        F |= Pm & (F << 1); Pm &= Pm << 1;
        F |= Pm & (F << 2); Pm &= Pm << 2;
        F |= Pm & (F << 4); Pm &= Pm << 4;
        F |= Pm & (F << 8); Pm &= Pm << 8;
        F |= Pm & (F << 16); Pm &= Pm << 16;
        F |= Pm & (F << 32);
        uint64_t starts = Nw & ~((F << 1) | carry);
        while (starts) {
            const int j = __builtin_ctzll(starts);
            cluster_read_begin[n_cl++] = pre[word_base + j];
            starts &= starts - 1;
        }
        carry = F >> 63;
        Iw = Nw = 0;
        word_base += 64;
        in_word = 0;
    };
    for (; i < n; i += 8) {
        const uint32_t rem = n - i;
        __m256i x;
        if (rem >= 8) x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(c + i));
        else x = _mm256_maskload_epi32(reinterpret_cast<const int*>(c + i), _mm256_cmpgt_epi32(_mm256_set1_epi32(int(rem)), lane_idx));
        const __m256i op = _mm256_and_si256(x, m15), l = _mm256_srli_epi32(x, 4);
        const __m256i rd = _mm256_and_si256(l, _mm256_sub_epi32(zero, _mm256_and_si256(_mm256_srlv_epi32(kread, op), one)));
        const __m256i rf = _mm256_and_si256(l, _mm256_sub_epi32(zero, _mm256_and_si256(_mm256_srlv_epi32(kref, op), one)));
        const __m256i cl = _mm256_and_si256(l, _mm256_sub_epi32(zero, _mm256_and_si256(_mm256_srlv_epi32(kclip, op), one)));
        a_size = _mm256_add_epi64(a_size, _mm256_add_epi64(_mm256_and_si256(rd, lo32), _mm256_srli_epi64(rd, 32)));
        a_ref = _mm256_add_epi64(a_ref, _mm256_add_epi64(_mm256_and_si256(rf, lo32), _mm256_srli_epi64(rf, 32)));
        a_right = _mm256_add_epi64(a_right, _mm256_add_epi64(_mm256_and_si256(cl, lo32), _mm256_srli_epi64(cl, 32)));
        if (kClusters) {
            const __m256i indel = _mm256_or_si256(_mm256_cmpeq_epi32(op, one), _mm256_cmpeq_epi32(op, two));
            const uint32_t Ib = uint32_t(_mm256_movemask_ps(_mm256_castsi256_ps(indel)));
            const uint32_t Zb = uint32_t(_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(l, zero))));
            Iw |= uint64_t(Ib) << in_word;
            Nw |= uint64_t(Ib & ~Zb) << in_word;
            __m256i t = _mm256_add_epi32(rd, _mm256_slli_si256(rd, 4));
            t = _mm256_add_epi32(t, _mm256_slli_si256(t, 8));
            const __m256i lowtot = _mm256_permutevar8x32_epi32(t, bc3);
            t = _mm256_add_epi32(t, _mm256_and_si256(lowtot, hi_half));
            const __m256i excl = _mm256_add_epi32(_mm256_sub_epi32(t, rd), _mm256_set1_epi32(int(base32)));
            _mm256_storeu_si256(reinterpret_cast<__m256i*>(pre + i), excl);
            base32 += uint32_t(_mm256_extract_epi32(t, 7));
            in_word += 8;
            if (in_word == 64) flush_word();
        }
    }
    if (kClusters && in_word) flush_word();
    alignas(32) uint64_t w[12];
    _mm256_store_si256(reinterpret_cast<__m256i*>(w), a_size);
    _mm256_store_si256(reinterpret_cast<__m256i*>(w + 4), a_ref);
    _mm256_store_si256(reinterpret_cast<__m256i*>(w + 8), a_right);
    size += w[0] + w[1] + w[2] + w[3];
    ref += w[4] + w[5] + w[6] + w[7];
    right += w[8] + w[9] + w[10] + w[11];
    return {left, size - right, size, int64_t(ref), n_cl};
}
#endif

// dispatch: the vector walk for CIGARs long enough to fill a few vectors, on CPUs that have it
template <bool kClusters>
inline CigarScan scan_cigar_any(const uint32_t* c, uint32_t n, uint32_t* cluster_read_begin, std::vector<uint32_t>& pre) {
#ifdef PTL_HAVE_AVX2_SCAN
    static const bool avx2 = __builtin_cpu_supports("avx2") && !std::getenv("PTL_PACK_SCALAR");
    if (avx2 && n >= 12u) {
        if (kClusters && pre.size() < size_t(n) + 8u) pre.resize(std::max<size_t>(size_t(n) + 8u, pre.size() * 2));
        return scan_cigar_avx2<kClusters>(c, n, cluster_read_begin, pre.data());
    }
#endif
    (void)pre;
    return scan_cigar<kClusters>(c, n, cluster_read_begin);
}

// The 16 nibbles the shifter may look at for a cluster whose walk starts at read_view[read_end - 1]: nibble q (bits 4q..)
// is read_view[read_end - 1 - q]; steps outside the read stay 0 (the kernel reports the reference's panic before using them).
inline uint64_t window_bits(const uint8_t* sq, uint32_t len, uint32_t read_end, bool flip, const uint8_t* pool_end) {
    if (read_end >= 16u && read_end <= len) {  // 16 steps, all inside the read: two loads instead of 16
        const uint32_t a = flip ? len - read_end : read_end - 16u;  // lowest of the 16 nibble indices
        const uint8_t* p = sq + (a >> 1);
        if (p + 9 <= pool_end) {
            uint64_t x;
            std::memcpy(&x, p, 8);
            if (!flip) {  // descending indices: the byte-swapped word has nibble a (or a - 1) on top and a + 15 (a + 14) at the bottom
                x = __builtin_bswap64(x);
                return (a & 1u) ? (x << 4) | uint64_t(p[8] >> 4) : x;
            }
            x = ((x & 0x0f0f0f0f0f0f0f0full) << 4) | ((x >> 4) & 0x0f0f0f0f0f0f0f0full);  // ascending: nibbles in index order
            return (a & 1u) ? (x >> 4) | (uint64_t(p[8] >> 4) << 60) : x;
        }
    }
    uint64_t bits = 0;
    for (uint32_t q = 0; q < 16 && q < read_end; ++q) {
        const uint32_t idx = read_end - 1u - q;
        if (idx >= len) continue;
        const uint32_t j = flip ? len - 1u - idx : idx;
        const uint32_t nib = (j & 1u) ? (sq[j >> 1] & 0xfu) : (sq[j >> 1] >> 4);
        bits |= uint64_t(nib) << (4u * q);
    }
    return bits;
}

}  // namespace

// Scratch of one pack call, kept with the batch so that a ring of reused batches stops allocating.
struct PackScratch {
    std::vector<Seg> segs, tmp;
    std::vector<uint32_t> seg_begin, kept, win_begin, seg_size;
    std::vector<uint32_t> cluster_read_begin;  // per I/D cluster, CIGAR order: read bases consumed before its first op
    std::vector<uint32_t> prefix;              // scratch of the vector CIGAR walk
    Ops sa_pool;
};

struct ptl_packed_batch {
    PackScratch scratch;
    void* arena = nullptr;
    size_t arena_cap = 0;
    bool pinned = false;
    ptl_batch view{};
    std::vector<uint32_t> record_index;  // batch read -> index into the source records
    uint32_t n_skipped_supplementary = 0;
};

extern "C" {

const char* ptl_pack_last_error(void) { return g_pack_err.c_str(); }

int ptl_pack_split_segments(uint32_t n_names, const char* const* contig_names, int32_t tid, int64_t pos, uint16_t flag,
                            uint8_t mapq, const uint32_t* cigar, uint32_t n_cigar, const char* sa_tag, uint32_t cap_segments,
                            uint32_t cap_cigar, ptl_split_segments* out, uint32_t* n_segments, uint32_t* n_cigar_out) {
    if (!out || !n_segments || !n_cigar_out) return PTL_ERR_INVALID_ARG;
    try {
        NameMap names(n_names, contig_names);
        std::vector<Seg> segs;
        Ops pool;
        split_segments(names, tid, pos, flag, mapq, cigar, n_cigar, sa_tag, segs, pool);
        size_t total = 0;
        for (const Seg& g : segs) total += g.cig_len;
        *n_segments = uint32_t(segs.size());
        *n_cigar_out = uint32_t(total);
        if (segs.size() > cap_segments || total > cap_cigar) return PTL_ERR_INVALID_ARG;
        uint32_t w = 0;
        for (size_t i = 0; i < segs.size(); ++i) {
            const Seg& g = segs[i];
            out->seq_order_start[i] = g.so_start;
            out->seq_order_end[i] = g.so_end;
            out->contig[i] = g.contig;
            out->pos[i] = g.pos;
            out->is_fwd[i] = g.is_fwd;
            out->mapq[i] = g.mapq;
            out->from_primary[i] = g.from_primary;
            out->cigar_begin[i] = w;
            const uint32_t* src = g.from_primary ? cigar : pool.data() + g.cig_off;
            std::memcpy(out->cigar + w, src, size_t(g.cig_len) * 4);
            w += g.cig_len;
        }
        out->cigar_begin[segs.size()] = w;
        return PTL_OK;
    } catch (const InputError& e) {
        g_pack_err = e.what();
        return PTL_ERR_INPUT;
    }
}

int ptl_pack_batch(const ptl_read_records* recs, uint32_t first, uint32_t count, uint32_t n_contigs,
                   const char* const* contig_names, int pinned, ptl_packed_batch** out) {
    return ptl_pack_batch_ex(recs, first, count, n_contigs, contig_names, pinned, PTL_WIN_NONE, nullptr, out);
}

// Packs into `pb`, reusing its arena and its scratch vectors when they are large enough (a streaming host repacks the
// same few buffers).  One forward walk per CIGAR yields everything the batch needs from it: clip positions, read and
// reference spans, and the read offset of every I/D cluster; the window fill then never looks at a CIGAR again.
static int pack_impl(const ptl_read_records* recs, uint32_t first, uint32_t count, uint32_t n_contigs, const char* const* contig_names, int window_mode,
                     const ptl_contig_segments* wsegs, ptl_packed_batch* pb) {
    if (!recs || !pb || uint64_t(first) + count > recs->n_reads) return PTL_ERR_INVALID_ARG;
    if (window_mode < PTL_WIN_NONE || window_mode > PTL_WIN_REVERSE_PAIRS || (window_mode == PTL_WIN_REVERSE_PAIRS && !wsegs)) return PTL_ERR_INVALID_ARG;
    const bool contig_wants_windows = window_mode != PTL_WIN_NONE;
    const bool pinned = pb->pinned;
    try {
        PackScratch& S = pb->scratch;
        std::vector<Seg>& segs = S.segs;
        std::vector<Seg>& tmp = S.tmp;
        std::vector<uint32_t>& seg_begin = S.seg_begin;
        std::vector<uint32_t>& kept = S.kept;
        std::vector<uint32_t>& win_begin = S.win_begin;
        std::vector<uint32_t>& seg_size = S.seg_size;
        std::vector<uint32_t>& rb = S.cluster_read_begin;
        Ops& sa_pool = S.sa_pool;
        segs.clear(); seg_begin.clear(); kept.clear(); win_begin.clear(); seg_size.clear(); sa_pool.clear();
        seg_begin.push_back(0);
        size_t n_rb = 0;  // entries of rb in use (a cluster per op at most: room is made ahead of every walk)
        auto rb_room = [&](size_t ops) { if (rb.size() < n_rb + ops) rb.resize(std::max(n_rb + ops, rb.size() * 2)); };
        std::unique_ptr<NameMap> names;  // built on the first SA tag
        // (the records are caller memory: the CIGAR CSR is checked before it sizes or indexes anything;
        //  cigar_begin[n_reads] is the size of the caller's pool -- tools/fuzz/fuzz_read_records.py)
        if (recs->n_reads && (!recs->cigar_begin || !recs->flag || !recs->tid || !recs->pos || !recs->mapq || !recs->bin || !recs->seq_len || !recs->seq_off))
            throw InputError("read records: missing arrays");
        {
            const uint64_t pool = recs->n_reads ? recs->cigar_begin[recs->n_reads] : 0;
            for (uint32_t r = first; r < first + count; ++r) {
                const uint64_t b0 = recs->cigar_begin[r], b1 = recs->cigar_begin[r + 1];
                if (b0 > b1 || b1 > pool || b1 - b0 > 0xffffffffull) throw InputError("read records: cigar_begin is not a CSR of the CIGAR pool");
            }
            if (pool && !recs->cigar) throw InputError("read records: missing CIGAR pool");
        }
        const uint64_t cig0 = count ? recs->cigar_begin[first] : 0, cig1 = count ? recs->cigar_begin[first + count] : 0;
        const uint32_t prim_ops = uint32_t(cig1 - cig0);
        uint32_t skipped = 0;
        // The pair test of read_alignment_scanner.rs:80-103 against the reverse-strand segments of the read segment's contig:
        // only those read segments go through left_shift_indels and want windows.
        auto wants_windows = [&](uint32_t contig, int64_t start, int64_t end) {
            if (window_mode != PTL_WIN_REVERSE_PAIRS) return true;
            if (contig >= wsegs->n_contigs) return false;
            for (uint32_t q = wsegs->contig_seg_begin[contig]; q < wsegs->contig_seg_begin[contig + 1]; ++q)
                if (!wsegs->seg_is_fwd[q] && end >= int64_t(wsegs->seg_seq_order_start[q]) && start < int64_t(wsegs->seg_seq_order_end[q])) return true;
            return false;
        };
        // pass 1: segment lists (SA tags are rare), and per segment the read offsets of its I/D clusters
        for (uint32_t r = first; r < first + count; ++r) {
            const uint16_t flag = recs->flag[r];
            if (flag & 0x4) throw InputError("unmapped record in the mapped read scan (assert, read_alignment_scanner.rs:396)");
            if (flag & 0x800) { ++skipped; continue; }  // supplementary records are skipped (:404)
            const uint32_t* cg = recs->cigar + recs->cigar_begin[r];
            const uint32_t ncg = uint32_t(recs->cigar_begin[r + 1] - recs->cigar_begin[r]);
            const char* sa = recs->sa_tag ? recs->sa_tag[r] : nullptr;
            if (!sa) {
                // one segment: the primary record itself
                if (contig_wants_windows) rb_room(ncg);
                const CigarScan sc = contig_wants_windows ? scan_cigar_any<true>(cg, ncg, rb.data() + n_rb, S.prefix) : scan_cigar_any<false>(cg, ncg, nullptr, S.prefix);
                const bool rev = flag & 0x10;
                Seg p{};
                if (!rev) { p.so_start = uint32_t(sc.left); p.so_end = uint32_t(sc.right); }
                else { p.so_start = uint32_t(sc.size - sc.right); p.so_end = uint32_t(sc.size - sc.left); }
                if (p.so_start >= p.so_end) throw InputError("Can't parse consistent split read information from SA tag");  // split_read.rs:145-152
                p.contig = uint32_t(recs->tid[r]);
                p.pos = recs->pos[r];
                p.is_fwd = !rev;
                p.mapq = recs->mapq[r];
                p.from_primary = 1;
                p.cig_off = uint32_t(recs->cigar_begin[r] - cig0);
                p.cig_len = ncg;
                if (contig_wants_windows) {
                    win_begin.push_back(uint32_t(n_rb));
                    if (sc.n_clusters && wants_windows(p.contig, p.pos, p.pos + sc.ref)) n_rb += sc.n_clusters;
                    seg_size.push_back(uint32_t(sc.size));
                }
                segs.push_back(p);
            } else {
                if (!names) names.reset(new NameMap(n_contigs, contig_names));
                split_segments(*names, recs->tid[r], recs->pos[r], flag, recs->mapq[r], cg, ncg, sa, tmp, sa_pool);
                for (Seg g : tmp) {
                    if (contig_wants_windows) {
                        const uint32_t* c = g.from_primary ? cg : sa_pool.data() + g.cig_off;
                        rb_room(g.cig_len);
                        const CigarScan sc = scan_cigar_any<true>(c, g.cig_len, rb.data() + n_rb, S.prefix);
                        win_begin.push_back(uint32_t(n_rb));
                        if (sc.n_clusters && wants_windows(g.contig, g.pos, g.pos + sc.ref)) n_rb += sc.n_clusters;
                        seg_size.push_back(uint32_t(sc.size));
                    }
                    // rebase CIGAR offsets into the batch pool: [record cigars of the slice][SA cigars]
                    if (g.from_primary) g.cig_off = uint32_t(recs->cigar_begin[r] - cig0);
                    else g.cig_off = prim_ops + g.cig_off;
                    segs.push_back(g);
                }
            }
            if (n_rb > 0xffffffffull) throw InputError("too many indel clusters in one batch");
            seg_begin.push_back(uint32_t(segs.size()));
            kept.push_back(r);
        }
        const uint32_t n = uint32_t(kept.size()), ns = uint32_t(segs.size());
        const uint64_t n_cig = (cig1 - cig0) + sa_pool.size();
        const uint64_t n_win = n_rb;
        if (contig_wants_windows) win_begin.push_back(uint32_t(n_win));
        // one arena for all small arrays
        size_t off = 0;
        auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes); return o; };
        const size_t o_flag = take(n * 2), o_mapq = take(n), o_bin = take(n * 2), o_len = take(n * 4), o_soff = take(n * 8),
                     o_sb = take((n + 1) * 4), o_ctg = take(ns * 4), o_pos = take(ns * 8), o_fwd = take(ns), o_cb = take(ns * 8),
                     o_cl = take(ns * 4), o_cig = take(n_cig * 4);
        const size_t o_wb = contig_wants_windows ? take((size_t(ns) + 1) * 4) : 0, o_win = contig_wants_windows ? take(n_win * 8) : 0;
        if (off > pb->arena_cap) {
            if (pb->arena) { if (pinned) ptl_host_free(pb->arena); else std::free(pb->arena); }
            const size_t want = std::max<size_t>(off + off / 8, 256);
            pb->arena = pinned ? ptl_host_alloc(want) : std::malloc(want);
            pb->arena_cap = pb->arena ? want : 0;
            if (!pb->arena) {
                g_pack_err = pinned ? "pinned allocation failed (no CUDA device?)" : "out of memory";
                return PTL_ERR_CUDA;
            }
        }
        pb->view = ptl_batch{};
        char* a = static_cast<char*>(pb->arena);
        auto* flag = reinterpret_cast<uint16_t*>(a + o_flag);
        auto* mapq = reinterpret_cast<uint8_t*>(a + o_mapq);
        auto* bin = reinterpret_cast<uint16_t*>(a + o_bin);
        auto* slen = reinterpret_cast<uint32_t*>(a + o_len);
        auto* soff = reinterpret_cast<uint64_t*>(a + o_soff);
        auto* sb = reinterpret_cast<uint32_t*>(a + o_sb);
        auto* ctg = reinterpret_cast<uint32_t*>(a + o_ctg);
        auto* pos = reinterpret_cast<int64_t*>(a + o_pos);
        auto* fwd = reinterpret_cast<uint8_t*>(a + o_fwd);
        auto* cb = reinterpret_cast<uint64_t*>(a + o_cb);
        auto* cl = reinterpret_cast<uint32_t*>(a + o_cl);
        auto* cig = reinterpret_cast<uint32_t*>(a + o_cig);
        // the batch lends out only the byte range of the packed-bases pool that its reads occupy
        uint64_t seq_lo = UINT64_MAX, seq_hi = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t r = kept[i];
            seq_lo = std::min(seq_lo, recs->seq_off[r]);
            seq_hi = std::max(seq_hi, recs->seq_off[r] + (uint64_t(recs->seq_len[r]) + 1) / 2);
        }
        if (n == 0) seq_lo = seq_hi = 0;
        if (seq_hi > recs->seq4_bytes) throw InputError("read bases outside the seq4 pool");
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t r = kept[i];
            flag[i] = recs->flag[r];
            mapq[i] = recs->mapq[r];
            bin[i] = recs->bin[r];
            slen[i] = recs->seq_len[r];
            soff[i] = recs->seq_off[r] - seq_lo;
        }
        std::memcpy(sb, seg_begin.data(), size_t(n + 1) * 4);
        for (uint32_t k = 0; k < ns; ++k) {
            ctg[k] = segs[k].contig;
            pos[k] = segs[k].pos;
            fwd[k] = segs[k].is_fwd;
            cb[k] = segs[k].cig_off;
            cl[k] = segs[k].cig_len;
        }
        if (cig1 > cig0) std::memcpy(cig, recs->cigar + cig0, size_t(cig1 - cig0) * 4);
        if (!sa_pool.empty()) std::memcpy(cig + (cig1 - cig0), sa_pool.data(), sa_pool.size() * 4);
        ptl_batch& v = pb->view;
        if (contig_wants_windows) {
            // The walk of cluster k (in the order of the REVERSED CIGAR, read_alignment_scanner.rs:165-167) compares
            // read_view[read_end - 1 - q], q = 0,1,..; read_view is the record's bases, reverse-complemented when
            // need_flipped_read_alignment (:153-157), which for a reverse-strand contig segment is
            // !(record.is_reverse() == segment.is_fwd_strand).  The window stores the BAM nibble of each step.
            // read_end = (read bases after the cluster) + (its insertions) = size - (read bases before it), so the forward
            // offsets of pass 1 give every window; a segment's windows are stored last cluster first.
            auto* wb = reinterpret_cast<uint32_t*>(a + o_wb);
            auto* win = reinterpret_cast<uint64_t*>(a + o_win);
            std::memcpy(wb, win_begin.data(), (size_t(ns) + 1) * 4);
            // The bases of a cluster sit in a cache line nobody has touched yet (7.5 KB of packed bases per read, one line
            // per cluster), so the fill runs kWinAhead windows behind a cursor that only prefetches: the misses of
            // neighbouring clusters, segments and reads overlap instead of queueing up one by one.
            constexpr uint32_t kWinAhead = PTL_WIN_AHEAD;
            struct Pending { const uint8_t* sq; uint32_t len, read_end, flip; };
            Pending ring[kWinAhead];
            uint64_t issued = 0, done = 0;
            const uint8_t* pool_end = recs->seq4 + recs->seq4_bytes;
            auto retire = [&]() {
                const Pending& t = ring[done % kWinAhead];
                win[done++] = window_bits(t.sq, t.len, t.read_end, t.flip != 0, pool_end);
            };
            for (uint32_t i = 0; i < n; ++i) {
                const uint32_t r = kept[i];
                const uint8_t* sq = recs->seq4 + recs->seq_off[r];
                const uint32_t len = recs->seq_len[r];
                const bool rec_rev = (recs->flag[r] & 0x10) != 0;
                for (uint32_t k = seg_begin[i]; k < seg_begin[i + 1]; ++k) {
                    const uint32_t w0 = win_begin[k], w1 = win_begin[k + 1];
                    if (w0 == w1) continue;
                    const uint32_t flip = !(rec_rev == (segs[k].is_fwd != 0));
                    const uint32_t size32 = seg_size[k];
                    for (uint32_t c = w1; c-- > w0;) {
                        const uint32_t read_end = size32 - rb[c];
                        if (issued - done == kWinAhead) retire();
                        ring[issued++ % kWinAhead] = Pending{sq, len, read_end, flip};
                        if (read_end > 0 && read_end <= len) {
                            const uint32_t hi = read_end - 1u, lo = read_end >= 16u ? read_end - 16u : 0u;
                            const uint32_t j0 = flip ? len - 1u - hi : lo, j1 = flip ? len - 1u - lo : hi;
                            __builtin_prefetch(sq + (j0 >> 1), 0, PTL_PF_HINT);
                            __builtin_prefetch(sq + (j1 >> 1), 0, PTL_PF_HINT);
                        }
                    }
                }
            }
            while (done < issued) retire();
            v.indel_win = win;
            v.rseg_win_begin = wb;
            v.n_indel_win = n_win;
        }
        v.n_reads = n;
        v.read_flag = flag; v.read_mapq = mapq; v.read_bin = bin; v.read_seq_len = slen; v.read_seq_off = soff;
        v.read_seg_begin = sb;
        v.n_read_segments = ns;
        v.rseg_contig = ctg; v.rseg_pos = pos; v.rseg_is_fwd = fwd; v.rseg_cigar_begin = cb; v.rseg_cigar_len = cl;
        v.cigar = cig;
        v.n_cigar = n_cig;
        v.seq4 = recs->seq4 + seq_lo;  // borrowed
        v.seq4_bytes = seq_hi - seq_lo;
        pb->record_index.assign(kept.begin(), kept.end());
        pb->n_skipped_supplementary = skipped;
        return PTL_OK;
    } catch (const InputError& e) {
        g_pack_err = e.what();
        return PTL_ERR_INPUT;
    }
}

int ptl_pack_batch_ex(const ptl_read_records* recs, uint32_t first, uint32_t count, uint32_t n_contigs,
                      const char* const* contig_names, int pinned, int window_mode, const ptl_contig_segments* wsegs, ptl_packed_batch** out) {
    if (!out) return PTL_ERR_INVALID_ARG;
    auto* pb = new ptl_packed_batch();
    pb->pinned = pinned != 0;
    const int rc = pack_impl(recs, first, count, n_contigs, contig_names, window_mode, wsegs, pb);
    if (rc != PTL_OK) {
        ptl_packed_batch_free(pb);
        return rc;
    }
    *out = pb;
    return PTL_OK;
}
int ptl_pack_batch_into(ptl_packed_batch* reuse, const ptl_read_records* recs, uint32_t first, uint32_t count, uint32_t n_contigs,
                        const char* const* contig_names, int window_mode, const ptl_contig_segments* wsegs) {
    return pack_impl(recs, first, count, n_contigs, contig_names, window_mode, wsegs, reuse);
}
void ptl_packed_batch_view(const ptl_packed_batch* p, ptl_batch* out) { *out = p->view; }
const uint32_t* ptl_packed_batch_record_index(const ptl_packed_batch* p) { return p->record_index.data(); }
void ptl_packed_batch_free(ptl_packed_batch* p) {
    if (!p) return;
    if (p->arena) { if (p->pinned) ptl_host_free(p->arena); else std::free(p->arena); }
    delete p;
}

int ptl_format_sa_tags(const ptl_result* res, uint32_t n_chrom, const char* const* chrom_names, char* buf, uint64_t cap,
                       uint64_t* sa_begin, uint64_t* need) {
    if (!res || !need) return PTL_ERR_INVALID_ARG;
    static const char opc[] = "MIDNSHP=X";
    std::string all;
    std::vector<std::string> seg;
    std::vector<uint64_t> begin;
    begin.reserve(size_t(res->n_records) + 1);
    for (uint32_t r = 0; r < res->n_reads; ++r) {
        const uint32_t k0 = res->read_rec_begin[r], k1 = res->read_rec_begin[r + 1];
        seg.clear();
        const bool lifted = !(k1 - k0 == 1 && res->rec_status[k0] != PTL_REC_LIFTED);
        if (lifted) {
            for (uint32_t k = k0; k < k1; ++k) {  // get_sa_tag_segment, :292-301
                const int32_t tid = res->rec_tid[k];
                if (tid < 0 || uint32_t(tid) >= n_chrom) return PTL_ERR_INVALID_ARG;
                std::string s = chrom_names[tid];
                s += ',';
                s += std::to_string(res->rec_pos[k] + 1);
                s += (res->rec_flag[k] & 0x10) ? ",-," : ",+,";
                for (uint64_t c = res->rec_cigar_begin[k]; c < res->rec_cigar_begin[k + 1]; ++c) {
                    s += std::to_string(len_of(res->cigar[c]));
                    s += opc[op_of(res->cigar[c])];
                }
                s += ',';
                s += std::to_string(unsigned(res->rec_mapq[k]));
                s += ",0;";
                seg.push_back(std::move(s));
            }
        }
        for (uint32_t k = k0; k < k1; ++k) {  // :349-363
            begin.push_back(all.size());
            if (lifted)
                for (uint32_t j = k0; j < k1; ++j)
                    if (j != k) all += seg[j - k0];
            all.push_back('\0');
        }
    }
    begin.push_back(all.size());
    *need = all.size();
    if (all.size() > cap || !buf || !sa_begin) return PTL_ERR_INVALID_ARG;
    std::memcpy(buf, all.data(), all.size());
    std::memcpy(sa_begin, begin.data(), begin.size() * sizeof(uint64_t));
    return PTL_OK;
}

uint32_t ptl_region_segment_count(uint64_t size, uint64_t segment_size) {
    if (!size || !segment_size) return 0;
    return uint32_t(1 + (size - 1) / segment_size);
}
void ptl_region_segments(uint64_t size, uint64_t segment_size, uint64_t* begin, uint64_t* end) {
    const uint64_t count = ptl_region_segment_count(size, segment_size);
    if (!count) return;
    const uint64_t base = size / count, extra = size % count;
    uint64_t start = 0;
    for (uint64_t i = 0; i < count; ++i) {
        const uint64_t e = std::min(start + base + (i < extra ? 1 : 0), size);
        begin[i] = start;
        end[i] = e;
        start = e;
    }
}
void ptl_shard_units(uint32_t n_units, const uint64_t* weight, uint32_t n_ranks, uint32_t* owner) {
    if (!n_ranks) return;
    std::vector<uint32_t> order(n_units);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return weight[a] > weight[b]; });
    std::vector<uint64_t> load(n_ranks, 0);
    for (uint32_t u : order) {
        const uint32_t r = uint32_t(std::min_element(load.begin(), load.end()) - load.begin());  // lowest rank wins ties
        owner[u] = r;
        load[r] += weight[u];
    }
}

uint16_t ptl_reg2bin(int64_t begin, int64_t end) {
    const uint64_t b = uint64_t(begin), e = uint64_t(end) - 1;
    uint32_t shift = 14;
    uint32_t offset = ((1u << 15) - 1) / 7;
    for (int level = 5; level > 0; --level) {
        if ((b >> shift) == (e >> shift)) return uint16_t(offset + (b >> shift));
        shift += 3;
        offset -= 1u << ((level - 1) * 3);
    }
    return 0;
}

}  // extern "C"
