// BGZF framing on the device, compression level 0 (the reference's stdout mode writes exactly this:
// `--remapped-read-output -` sets CompressionLevel::Uncompressed, src/read_alignment_scanner.rs:66-71, "to optimize piping
// into samtools sort").  One BGZF block (SAM spec 4.1) per thread block:
//     18-byte gzip header with the BC size subfield | 01 LEN NLEN (one stored deflate block) | <= 0xff00 data bytes | CRC32 | ISIZE
// Input offsets and output offsets are closed forms of the block index (every block but the last is full), so there is
// no size pass.  HBM-bound: 64 KB in, 64 KB out per block, the CRC computed from shared memory in between.
//   1. CRC-32 (IEEE, reflected): every thread reads its own 256-byte slice of the payload straight from global memory and
//      runs two zero-state CRC chains over its halves, one 32-bit word per step through eight per-lane 16-entry nibble
//      tables (conflict-free; see below); the slices are laid out from the END of the data (leading zeros do not change
//      a zero-state CRC, so a short last block is no special case) and combined pairwise up a tree:
//      crc(A || B) = Z^|B|(crc(A)) ^ crc(B), where Z^m = "advance the state through m zero bytes" is linear over GF(2)
//      and held as nibble tables for m = 128 B and 256 * 2^k B;
//   2. the real CRC is Z^n(0xffffffff) ^ crc_zero_state(data) ^ 0xffffffff; Z^n(0xffffffff) is a per-launch constant for
//      full blocks and one more for the last block (host-computed);
//   3. the payload is copied by the record copy of assemble_bam.cuh (the second read of the block hits L1 / L2).
// (First versions staged the block in 66 KB of shared memory: 3 blocks per SM, and the 256-entry byte tables shared by
//  the lanes made the shared-memory pipe the bound - 68 % of its peak, 57 % bank conflicts: 2.45 ms per 47 k blocks.)
#pragma once
#include <cstdint>

#include "assemble_bam.cuh"

namespace ptl {

constexpr uint32_t kBgzfIn = 0xff00u;             // payload bytes of a full block (htslib BGZF_BLOCK_SIZE)
constexpr uint32_t kBgzfOverhead = 18u + 5u + 8u;  // gzip header + stored-block header + CRC32 + ISIZE
// Tables (words), built by the host (context.cu: bgzf_tables):
//   [0, 256)        T0, the byte-wise CRC table (only the < 16 odd bytes in front of a slice half use it)
//   [256, 1280)     nibble tables of the zero-byte shifts by 128 B (the two chains of a thread) and by 256 * 2^k B,
//                   k = 0..7 (the combine tree): [9 levels][8 nibbles][16]
//   [1280, 1408)    W[8][16]: W[j][x] = (x << 4 j) advanced through 4 zero bytes -- one CRC step over a 32-bit word is
//                   state' = XOR_j W[j][nibble j of (state ^ word)].  The kernel copies W into shared memory once PER LANE
//                   (bank = lane), so the eight lookups of a step never conflict; 256-entry byte tables would need 32 KB per
//                   table for that, and shared among the lanes they cost 3.5 wavefronts per lookup (57 % of all shared-memory
//                   wavefronts of the first version of this kernel, ncu r04c).
constexpr uint32_t kBgzfShift = 256u, kBgzfWord = kBgzfShift + 9u * 128u, kBgzfTableWords = kBgzfWord + 128u;
constexpr uint32_t kBgzfSmemWords = 256u + 9u * 128u + 128u * 32u + 8u;  // T0 | shift tables | per-lane W | 8 warp CRCs

struct BgzfArgs {
    const uint8_t* in;       // the stream (device); readable 64 KB in front (never used as data) and 32 bytes behind
    uint64_t n;              // stream bytes
    uint8_t* out;            // n + 31 * n_blocks bytes
    const uint32_t* tables;  // kBgzfTableWords words (layout above)
    uint32_t init_full;      // Z^0xff00(0xffffffff)
    uint32_t init_last;      // Z^(bytes of the last block)(0xffffffff)
    uint64_t n_blocks;
};

// Z^m(v) from eight 16-entry tables, one per nibble of v (the map is linear over GF(2))
__device__ __forceinline__ uint32_t gf2_apply_nibbles(const uint32_t* __restrict__ t, uint32_t v) {
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) r ^= t[16 * j + ((v >> (4 * j)) & 15u)];
    return r;
}

// blockDim.x == 256; kBgzfSmemWords words of dynamic shared memory
// the tables into shared memory (once per resident thread block: the kernel is persistent, 22 KB of table stores per
// 64 KB block of payload were 35 % of its stall samples)
__device__ __forceinline__ void bgzf_store_init(const BgzfArgs& A, uint32_t tid, uint32_t* sm) {
    uint32_t* wl = sm + 256 + 9 * 128;
    for (uint32_t i = tid; i < kBgzfWord; i += 256u) sm[i] = A.tables[i];
    for (uint32_t i = tid; i < 128u * 32u; i += 256u) wl[i] = A.tables[kBgzfWord + (i >> 5)];
    __syncthreads();
}
// gzip header + the stored-block header of a block with n payload bytes (unaligned destination: byte stores by one thread)
__device__ __forceinline__ void bgzf_write_header(uint8_t* dst, uint32_t n, uint32_t tid) {
    if (tid == 0) {
        const uint32_t bsize1 = n + kBgzfOverhead - 1u;
        const uint8_t h[23] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, uint8_t(bsize1 & 0xffu), uint8_t(bsize1 >> 8),
                               0x01, uint8_t(n & 0xffu), uint8_t(n >> 8), uint8_t(~n & 0xffu), uint8_t((~n >> 8) & 0xffu)};
        for (int i = 0; i < 23; ++i) dst[i] = h[i];
    }
}
// The zero-state CRCs of the 256 slices (slice of thread t ends 256 (255 - t) bytes before the end of the sliced data) ->
// the block's CRC-32 and the footer.  `tail_len` (< 16) bytes with zero-state CRC `tail_crc` may follow the sliced data.
__device__ __forceinline__ void bgzf_crc_combine_footer(const BgzfArgs& A, uint32_t crc, uint32_t n, uint8_t* data_dst, uint32_t tid, uint32_t* sm,
                                                        uint32_t tail_len, uint32_t tail_crc) {
    uint32_t* t0 = sm;
    uint32_t* shift = sm + 256;
    uint32_t* warp_crc = shift + 9 * 128 + 128 * 32;
    const uint32_t lane = tid & 31u;
    // tree combine: after level k a thread with tid % 2^(k+1) == 0 holds the CRC of 2^(k+1) slices
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const uint32_t right = __shfl_down_sync(0xffffffffu, crc, 1u << k);
        if ((lane & ((2u << k) - 1u)) == 0u) crc = gf2_apply_nibbles(shift + 128 * (k + 1), crc) ^ right;
    }
    if (lane == 0) warp_crc[tid >> 5] = crc;
    __syncthreads();
    if (tid < 32u) {
        crc = tid < 8u ? warp_crc[tid] : 0u;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t right = __shfl_down_sync(0xffffffffu, crc, 1u << k);
            if ((tid & ((2u << k) - 1u)) == 0u) crc = gf2_apply_nibbles(shift + 128 * (6 + k), crc) ^ right;
        }
        if (tid == 0) {  // footer: CRC32, ISIZE
            for (uint32_t i = 0; i < tail_len; ++i) crc = t0[crc & 0xffu] ^ (crc >> 8);  // through tail_len zero bytes
            crc ^= tail_crc;
            const uint32_t full = ((n == kBgzfIn) ? A.init_full : A.init_last) ^ crc ^ 0xffffffffu;
            uint8_t* f = data_dst + n;
            for (int i = 0; i < 4; ++i) { f[i] = uint8_t(full >> (8 * i)); f[4 + i] = uint8_t(n >> (8 * i)); }
        }
    }
}
// CRC-32 of the n payload bytes at `src` (block-cooperative, 256 threads) and the footer (CRC32, ISIZE) behind data_dst.
__device__ __forceinline__ void bgzf_crc_and_footer(const BgzfArgs& A, const uint8_t* src, uint32_t n, uint8_t* data_dst, uint32_t tid, uint32_t* sm) {
    uint32_t* t0 = sm;
    uint32_t* shift = sm + 256;
    uint32_t* wl = shift + 9 * 128;  // per-lane word-step tables: entry e of lane l at wl[32 e + l]
    const uint32_t lane = tid & 31u;
    // ---- zero-state CRC of this thread's 256-byte slice, slices counted back from the end of the data; its two halves run
    //      as two independent chains; every thread reads its slice straight from global memory (16 bytes per step: five
    //      aligned words + funnel shifts; a sector is used by two consecutive steps of the same thread)
    const int32_t s_end = int32_t(n) - 256 * int32_t(255u - tid);
    const int32_t s_begin = max(s_end - 256, 0);
    uint32_t crc_a = 0, crc_b = 0;
    auto word_step = [&](uint32_t crc, uint32_t w) {
        crc ^= w;
        uint32_t r = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) r ^= wl[32u * (16u * j + ((crc >> (4 * j)) & 15u)) + lane];
        return r;
    };
    auto chain = [&](int32_t from, int32_t to) {  // CRC of data[from, to), zero initial state
        uint32_t crc = 0;
        int32_t p = from;
        for (const int32_t odd = (to - from) & 15; p < from + odd; ++p) crc = t0[(crc ^ src[p]) & 0xffu] ^ (crc >> 8);
        for (; p < to; p += 16) {
            const uint4 v = window128_body(src + p);
            crc = word_step(word_step(word_step(word_step(crc, v.x), v.y), v.z), v.w);
        }
        return crc;
    };
    if (s_end > 0) {
        if (s_end - s_begin == 256) {  // (every slice of a full block)
            int32_t pa = s_begin, pb = s_begin + 128;
            const bool aligned = ((reinterpret_cast<uint64_t>(src) + uint64_t(s_begin)) & 15ull) == 0ull;  // (block-uniform)
#pragma unroll 2
            for (int i = 0; i < 8; ++i, pa += 16, pb += 16) {
                uint4 va, vb;
                if (aligned) {
                    va = *reinterpret_cast<const uint4*>(src + pa);
                    vb = *reinterpret_cast<const uint4*>(src + pb);
                } else {
                    va = window128_body(src + pa);
                    vb = window128_body(src + pb);
                }
                crc_a = word_step(crc_a, va.x); crc_b = word_step(crc_b, vb.x);
                crc_a = word_step(crc_a, va.y); crc_b = word_step(crc_b, vb.y);
                crc_a = word_step(crc_a, va.z); crc_b = word_step(crc_b, vb.z);
                crc_a = word_step(crc_a, va.w); crc_b = word_step(crc_b, vb.w);
            }
        } else {  // the first slice of a short last block
            const int32_t mid = max(s_end - 128, s_begin);
            crc_a = chain(s_begin, mid);
            crc_b = chain(mid, s_end);
        }
    }
    bgzf_crc_combine_footer(A, gf2_apply_nibbles(shift, crc_a) ^ crc_b, n, data_dst, tid, sm, 0u, 0u);  // (|B| = 128 whenever A is not empty)
}
__device__ __forceinline__ void bgzf_store_block_body(const BgzfArgs& A, uint64_t b, uint32_t tid, uint32_t* sm) {
    const uint64_t in_off = b * kBgzfIn;
    const uint32_t n = uint32_t(min(uint64_t(kBgzfIn), A.n - in_off));
    uint8_t* dst = A.out + b * uint64_t(kBgzfIn + kBgzfOverhead);
    uint8_t* data_dst = dst + 23;
    const uint8_t* src = A.in + in_off;
    bgzf_write_header(dst, n, tid);
    bgzf_crc_and_footer(A, src, n, data_dst, tid, sm);
    // ---- the payload itself: the record copy of assemble_bam.cuh (aligned 16-byte stores, funnel-shifted source)
    copy_field(data_dst, src, n, tid, 256u);
    __syncthreads();  // (warp_crc is reused by the next block of this persistent thread block)
}

// ------------------------------------------------------------------------------------------------------------------
// Fused record assembly + framing (ptl_frame_records): the payload of BGZF block b is produced STRAIGHT FROM THE SOURCES
// (packed bases and qualities of the reads, re-oriented for flipped records; the few hundred bytes of core / name / CIGAR /
// aux / tag text of every record come from the sparse scratch bam_write_meta_kernel fills), so the assembled record stream
// is never written and read back: 3 GB in, 3 GB out per 131 k reads instead of 12 GB over two kernels.  The block then
// reads its own 64 KB back (from L2: it has just written them) for the CRC.
struct FrameArgs {
    BgzfArgs Z;            // Z.in is unused; Z.n = prefix_bytes + the record bytes
    BamAsmArgs A;          // A.out = the sparse scratch with the small fields of every record at its stream offset
    const uint8_t* prefix; // device copy of the stream prefix (the BAM header), may be nullptr
    uint64_t prefix_bytes;
};
// A run = a maximal stretch of the block's payload that comes from ONE source field: the prefix, the small fields of a
// record (from the sparse scratch), its packed bases or its qualities (plain, or re-oriented for a flipped record).
// Warp 0 builds the table of the block's runs in shared memory (lanes over records: the descriptor loads of up to 32
// records are in flight together), then ALL threads walk the block's destination-aligned 16-byte chunks as one flat index
// space, so that every load of the block is independent of every other (one copy_field call per field, one after the
// other, left each block waiting through a dozen dependent DRAM round trips: 6.2 ms per 47 k blocks instead of 3.8 ms for
// the two separate kernels).
struct FrameRun {
    int32_t begin;       // payload offset of the run's first byte (may be negative: the field starts in the previous block)
    uint32_t len;        // bytes of the run inside [0, n)... counted from `begin` (clipped only at the far end)
    uint32_t kind;       // 0 plain copy, 1 reverse-complemented packed bases, 2 reversed qualities
    uint32_t l_seq;      // kinds 1, 2: bases of the read
    const uint8_t* src;  // kind 0: source byte of payload offset `begin`; kinds 1, 2: the read's packed bases / qualities
    int64_t off;         // kinds 1, 2: field offset of the run's first byte
};
constexpr uint32_t kFrameMaxRuns = 132;  // 32 records x 4 fields + the prefix (+ slack)
constexpr uint32_t kFrameSmemWords = kBgzfSmemWords + kFrameMaxRuns * (sizeof(FrameRun) / 4u) + 8u;

// 16 bytes of run r at run-relative offset o (bytes outside the run: unspecified), bounds-guarded
// (the rare paths are kept out of line: fully inlined the kernel was 12 k instructions, 200 KB of code, and ran 3x slower
//  than the two-pass path on instruction fetch alone)
static __device__ __noinline__ uint4 frame_produce_guarded(const FrameRun& r, int64_t o) {
    if (r.kind == 0u) return window128_guarded(r.src, o, int64_t(r.len));
    const int64_t len = r.l_seq;
    if (r.kind == 1u) return revcomp_chunk(r.src, len, o + r.off);
    const uint4 q = window128_lean(r.src, len - 16 - (o + r.off), len);
    uint4 v;
    v.x = __byte_perm(q.w, 0u, 0x0123u);
    v.y = __byte_perm(q.z, 0u, 0x0123u);
    v.z = __byte_perm(q.y, 0u, 0x0123u);
    v.w = __byte_perm(q.x, 0u, 0x0123u);
    return v;
}

// ---- the run table of (a pass over) one block: warp 0, lanes over records -------------------------------------------
// Fills runs[] with the prefix run (first pass) and the runs of up to 32 records from k_next on; ctl[0] = runs,
// ctl[1] = next record, ctl[2] = 1 when no record of the block is left.
static __device__ __noinline__ void frame_build_runs(const FrameArgs& F, uint64_t x0, uint64_t x1, bool first_pass, uint32_t k_next, FrameRun* runs,
                                                     uint32_t* ctl, uint32_t lane) {
    const BamAsmArgs& A = F.A;
    const bool has_recs = x1 > F.prefix_bytes && A.n_records;
    const uint64_t xs = x0 > F.prefix_bytes ? x0 : F.prefix_bytes;
    const uint64_t r0 = xs - F.prefix_bytes, r1 = has_recs ? x1 - F.prefix_bytes : 0ull;
    const int64_t rec_shift = int64_t(F.prefix_bytes) - int64_t(x0);  // payload offset = record-space offset + rec_shift
    uint32_t base = 0;
    if (first_pass) {
        if (x0 < F.prefix_bytes) {
            if (lane == 0) runs[0] = FrameRun{0, uint32_t((x1 < F.prefix_bytes ? x1 : F.prefix_bytes) - x0), 0u, 0u, F.prefix + x0, 0};
            base = 1;
        }
        if (has_recs) {  // first record = last k with rec_begin[k] <= r0: a 32-ary search, one probe per lane and round
            uint32_t lo = 0, hi = A.n_records;  // invariant: rec_begin[lo] <= r0 < rec_begin[hi] (rec_begin[n_records] = total > r0)
            while (hi - lo > 1u) {
                const uint32_t step = (hi - lo + 31u) / 32u;
                const uint32_t probe = lo + lane * step;
                const bool le = probe < hi && A.rec_begin[probe] <= r0;
                const uint32_t m = __ballot_sync(0xffffffffu, le);  // (lane 0 probes lo: always set)
                const uint32_t top = 31u - uint32_t(__clz(int(m)));
                const uint32_t nlo = lo + top * step;
                hi = min(hi, nlo + step);
                lo = nlo;
            }
            k_next = lo;
        }
    }
    uint32_t cnt = 0;
    FrameRun mine[4];
    bool started = false;
    if (has_recs) {
        const uint32_t k = k_next + lane;
        if (k < A.n_records) {
            const uint64_t rb = A.rec_begin[k];
            if (rb < r1) {
                started = true;
                const BamRecLayout L = BamRecLayout::unpack(A.rec_desc + 2 * size_t(k));
                const uint64_t o_seq = L.o_seq(), o_qual = o_seq + L.seq_bytes, o_aux = o_qual + L.l_seq;
                const bool flip = A.rec_need_flip[k] != 0;
                const uint8_t* src_s = A.seq4 + A.read_seq_off[L.r];
                const uint8_t* src_q = A.qual + A.qual_off[L.r];
                const uint64_t fa[5] = {0, o_seq, o_qual, o_aux, L.total};
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    const uint64_t lo_c = r0 > rb ? r0 - rb : 0ull, hi_c = r1 - rb;
                    const uint64_t ia = fa[f] > lo_c ? fa[f] : lo_c, ib = fa[f + 1] < hi_c ? fa[f + 1] : hi_c;
                    if (ia >= ib) continue;
                    FrameRun r;
                    r.begin = int32_t(int64_t(rb + ia) + rec_shift);
                    r.len = uint32_t(ib - ia);
                    r.l_seq = L.l_seq;
                    if (f == 0 || f == 3) { r.kind = 0u; r.src = A.out + rb + ia; r.off = 0; }
                    else if (f == 1) { r.kind = flip ? 1u : 0u; r.src = flip ? src_s : src_s + (ia - o_seq); r.off = int64_t(ia - o_seq); }
                    else { r.kind = flip ? 2u : 0u; r.src = flip ? src_q : src_q + (ia - o_qual); r.off = int64_t(ia - o_qual); }
                    mine[cnt++] = r;
                }
            }
        }
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (int(lane) >= d) incl += o;
    }
    const uint32_t at = base + incl - cnt;
    for (uint32_t i = 0; i < cnt; ++i) runs[at + i] = mine[i];
    const uint32_t n_started = uint32_t(__popc(__ballot_sync(0xffffffffu, started)));
    if (lane == 31u) {
        const uint32_t kn = k_next + n_started;
        ctl[0] = base + incl;
        ctl[1] = kn;
        ctl[2] = (!has_recs || n_started < 32u || kn >= A.n_records || A.rec_begin[kn] >= r1) ? 1u : 0u;
    }
}

// ---- single pass: produce -> CRC -> store, every byte handled once (blocks whose runs all fit the table) ----------------
__device__ __forceinline__ uint32_t crc_word_step(const uint32_t* __restrict__ wl, uint32_t lane, uint32_t crc, uint32_t w) {
    crc ^= w;
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) r ^= wl[32u * (16u * j + ((crc >> (4 * j)) & 15u)) + lane];
    return r;
}
// 16 payload bytes at payload offset o from the run table (bytes outside the payload come out as zero: the chunks in
// front of the first and behind the last 16-byte boundary of the destination); `ri` = a cursor that only moves forward
static __device__ __noinline__ uint4 frame_chunk(const FrameRun* __restrict__ runs, uint32_t n_runs, uint32_t& ri, int32_t o) {
    while (ri + 1u < n_runs && runs[ri + 1u].begin <= o) ++ri;
    {
        const FrameRun& r = runs[ri];
        const int32_t rel = o - r.begin;
        if (rel >= 0 && rel + 16 <= int32_t(r.len)) {
            if (r.kind == 0u && rel + 20 <= int32_t(r.len)) return window128_body(r.src + rel);
            return frame_produce_guarded(r, rel);
        }
    }
    uint32_t w[4] = {0u, 0u, 0u, 0u};  // across a run boundary: every run that touches the chunk contributes its own bytes
    for (uint32_t rj = ri; rj < n_runs && runs[rj].begin < o + 16; ++rj) {
        const FrameRun& q = runs[rj];
        const int32_t rl = o - q.begin;
        const int32_t lo = max(0, -rl), hi = min(16, int32_t(q.len) - rl);  // bytes [lo, hi) of the chunk belong to this run
        if (lo >= hi) continue;
        const uint4 v = frame_produce_guarded(q, rl);
        const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int32_t blo = min(max(lo - 4 * k, 0), 4), bhi = min(max(hi - 4 * k, 0), 4);
            const uint32_t m = uint32_t(((1ull << (8 * bhi)) - 1ull) & ~((1ull << (8 * blo)) - 1ull));
            w[k] |= x[k] & m;
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
// two aligned 16-byte loads -> the 16 bytes at byte offset 4 ws + bs / 8 of their 32
__device__ __forceinline__ uint4 shift128(const uint4& a, const uint4& b, uint32_t ws, uint32_t bs) {
    const uint32_t x0 = ws == 0u ? a.x : ws == 1u ? a.y : ws == 2u ? a.z : a.w;
    const uint32_t x1 = ws == 0u ? a.y : ws == 1u ? a.z : ws == 2u ? a.w : b.x;
    const uint32_t x2 = ws == 0u ? a.z : ws == 1u ? a.w : ws == 2u ? b.x : b.y;
    const uint32_t x3 = ws == 0u ? a.w : ws == 1u ? b.x : ws == 2u ? b.y : b.z;
    const uint32_t x4 = ws == 0u ? b.x : ws == 1u ? b.y : ws == 2u ? b.z : b.w;
    return make_uint4(__funnelshift_r(x0, x1, bs), __funnelshift_r(x1, x2, bs), __funnelshift_r(x2, x3, bs), __funnelshift_r(x3, x4, bs));
}
// The block's payload, produced, CRC'd and stored in ONE pass over destination-aligned 256-byte slices (thread t: the slice
// ending 256 (255 - t) bytes before the last 16-byte boundary of the destination).  A slice that is a plain copy of one
// source runs its two halves as two interleaved chains with ONE new aligned 128-bit load per chunk and chain
// (thread-contiguous addresses: a window of five 32-bit words costs five L1 wavefronts per lane and chunk, which made a
// re-reading CRC phase L1-bound); anything else goes chunk by chunk through frame_chunk.
__device__ __forceinline__ void bgzf_frame_single_pass(const FrameArgs& F, const FrameRun* __restrict__ runs, uint32_t n_runs, uint32_t n, uint8_t* data_dst,
                                                       uint32_t tid, uint32_t* sm) {
    const uint32_t* t0 = sm;
    const uint32_t* shift = sm + 256;
    const uint32_t* wl = shift + 9 * 128;
    const uint32_t lane = tid & 31u;
    const int32_t a0 = int32_t(reinterpret_cast<uint64_t>(data_dst) & 15ull);
    int32_t tail = (int32_t(n) + a0) & 15;  // the < 16 bytes behind the last boundary: folded in by thread 0 (bgzf_crc_combine_footer)
    if (tail > int32_t(n)) tail = int32_t(n);
    const int32_t e_al = int32_t(n) - tail;
    const int32_t s_end = e_al - 256 * int32_t(255u - tid);
    const int32_t s_begin = max(s_end - 256, 0);
    uint32_t crc_a = 0, crc_b = 0;
    if (s_end > 0) {
        bool fast = false;
        if (s_end - s_begin == 256) {
            uint32_t ri = 0;
            while (ri + 1u < n_runs && runs[ri + 1u].begin <= s_begin) ++ri;
            const int32_t rel = s_begin - runs[ri].begin;
            if (runs[ri].kind == 0u && rel >= 0 && rel + 256 + 32 <= int32_t(runs[ri].len)) {  // (+ 32: the look-ahead load stays inside the run)
                fast = true;
                const uint64_t sa = reinterpret_cast<uint64_t>(runs[ri].src + rel);
                const uint4* __restrict__ base = reinterpret_cast<const uint4*>(sa & ~15ull);
                const uint32_t ws = uint32_t(sa & 15ull) >> 2, bs = uint32_t(sa & 3ull) * 8u;
                uint4* __restrict__ d = reinterpret_cast<uint4*>(data_dst + s_begin);
                uint4 pa = base[0], pb = base[8];
#pragma unroll 1
                for (int i = 0; i < 8; ++i) {
                    const uint4 na = base[i + 1], nb = base[i + 9];
                    const uint4 va = shift128(pa, na, ws, bs), vb = shift128(pb, nb, ws, bs);
                    d[i] = va;
                    d[i + 8] = vb;
                    crc_a = crc_word_step(wl, lane, crc_a, va.x); crc_b = crc_word_step(wl, lane, crc_b, vb.x);
                    crc_a = crc_word_step(wl, lane, crc_a, va.y); crc_b = crc_word_step(wl, lane, crc_b, vb.y);
                    crc_a = crc_word_step(wl, lane, crc_a, va.z); crc_b = crc_word_step(wl, lane, crc_b, vb.z);
                    crc_a = crc_word_step(wl, lane, crc_a, va.w); crc_b = crc_word_step(wl, lane, crc_b, vb.w);
                    pa = na;
                    pb = nb;
                }
            }
        }
        if (!fast) {  // chain A over [s_begin, mid), chain B over [mid, s_end), one after the other
            const int32_t mid = (s_end - s_begin == 256) ? s_begin + 128 : max(s_end - 128, s_begin);
            uint32_t ri = 0;
            auto range = [&](int32_t p0, int32_t p1, uint32_t crc) {  // produce, store and CRC payload [p0, p1); p1 is chunk-aligned
                int32_t p = p0;
                if (p < p1 && ((p + a0) & 15)) {  // (only in front of the very first chunk of the block: p0 == 0)
                    const int32_t h = min(p1, 16 - a0);  // the bytes [0, h) sit at the end of the virtual chunk [h - 16, h)
                    const uint4 v = frame_chunk(runs, n_runs, ri, h - 16);
                    const uint32_t x[4] = {v.x, v.y, v.z, v.w};
                    for (; p < h; ++p) {
                        const int32_t t = p - (h - 16);
                        const uint32_t bv = (x[t >> 2] >> (8 * (t & 3))) & 0xffu;
                        data_dst[p] = uint8_t(bv);
                        crc = t0[(crc ^ bv) & 0xffu] ^ (crc >> 8);
                    }
                }
#pragma unroll 1
                for (; p + 16 <= p1; p += 16) {
                    const uint4 v = frame_chunk(runs, n_runs, ri, p);
                    *reinterpret_cast<uint4*>(data_dst + p) = v;
                    crc = crc_word_step(wl, lane, crc_word_step(wl, lane, crc_word_step(wl, lane, crc_word_step(wl, lane, crc, v.x), v.y), v.z), v.w);
                }
                return crc;
            };
            crc_a = range(s_begin, mid, 0u);
            crc_b = range(mid, s_end, 0u);
        }
    }
    uint32_t tail_crc = 0;
    if (tid == 0 && tail > 0) {
        uint32_t ri = 0;
        const uint4 v = frame_chunk(runs, n_runs, ri, e_al);  // the virtual chunk [e_al, e_al + 16): its first `tail` bytes exist
        const uint32_t x[4] = {v.x, v.y, v.z, v.w};
        for (int32_t t = 0; t < tail; ++t) {
            const uint32_t bv = (x[t >> 2] >> (8 * (t & 3))) & 0xffu;
            data_dst[e_al + t] = uint8_t(bv);
            tail_crc = t0[(tail_crc ^ bv) & 0xffu] ^ (tail_crc >> 8);
        }
    }
    bgzf_crc_combine_footer(F.Z, gf2_apply_nibbles(shift, crc_a) ^ crc_b, n, data_dst, tid, sm, uint32_t(tail), tail_crc);
}

// ---- the general path (a block with more than 32 records, i.e. records of a few hundred bytes): pass after pass, every
//      run stores its own bytes of every chunk it touches, then the block reads its payload back for the CRC
static __device__ __noinline__ void bgzf_frame_general(const FrameArgs& F, uint64_t x0, uint64_t x1, FrameRun* runs, uint32_t* ctl, uint32_t n,
                                                       uint8_t* data_dst, uint32_t tid, uint32_t* sm) {
    for (;;) {  // (entered with the table of the first pass built)
        const uint32_t n_runs = ctl[0];
        const bool done = ctl[2] != 0u;
        const uint32_t k_next = ctl[1];
        if (n_runs) {
            const int32_t span0 = max(runs[0].begin, 0);
            const int32_t span1 = min(int32_t(n), runs[n_runs - 1].begin + int32_t(runs[n_runs - 1].len));
            const int32_t a0 = int32_t(reinterpret_cast<uint64_t>(data_dst) & 15ull);  // chunk c covers payload [16 c - a0, + 16)
            const int32_t c0 = (span0 + a0) >> 4, c1 = (span1 + a0 + 15) >> 4;
            uint32_t ri = 0;
            for (int32_t c = c0 + int32_t(tid); c < c1; c += 256) {
                const int32_t o = 16 * c - a0;
                while (ri + 1u < n_runs && runs[ri + 1u].begin <= o) ++ri;
                for (uint32_t rj = ri; rj < n_runs && runs[rj].begin < o + 16; ++rj) {
                    const FrameRun q = runs[rj];
                    const int64_t rl = int64_t(o) - q.begin;
                    if (rl + 16 <= 0) continue;
                    const uint4 v = frame_produce_guarded(q, rl);
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const int64_t pr = rl + t;  // run-relative
                        const int32_t pp = o + t;   // payload
                        if (pr >= 0 && pr < int64_t(q.len) && pp >= 0 && pp < int32_t(n)) data_dst[pp] = uint8_t(w[t >> 2] >> (8 * (t & 3)));
                    }
                }
            }
        }
        __syncthreads();  // (the table is rebuilt by the next pass; after the last pass: the payload is complete and visible)
        if (done) break;
        if (tid < 32u) frame_build_runs(F, x0, x1, false, k_next, runs, ctl, tid);
        __syncthreads();
    }
    bgzf_crc_and_footer(F.Z, data_dst, n, data_dst, tid, sm);
}

__device__ __forceinline__ void bgzf_frame_block_body(const FrameArgs& F, uint64_t b, uint32_t tid, uint32_t* sm) {
    FrameRun* runs = reinterpret_cast<FrameRun*>(sm + kBgzfSmemWords);
    uint32_t* ctl = sm + kBgzfSmemWords + kFrameMaxRuns * (sizeof(FrameRun) / 4u);  // [0] runs of this pass, [1] next record, [2] done
    const uint64_t x0 = b * kBgzfIn;
    const uint32_t n = uint32_t(min(uint64_t(kBgzfIn), F.Z.n - x0));
    const uint64_t x1 = x0 + n;
    uint8_t* dst = F.Z.out + b * uint64_t(kBgzfIn + kBgzfOverhead);
    uint8_t* data_dst = dst + 23;
    bgzf_write_header(dst, n, tid);
    if (tid < 32u) frame_build_runs(F, x0, x1, true, 0u, runs, ctl, tid);
    __syncthreads();
#ifndef FRAME_TWO_PHASE
    if (ctl[2] != 0u) bgzf_frame_single_pass(F, runs, ctl[0], n, data_dst, tid, sm);  // every run of the block is in the table (<= 32 records: any real data)
    else
#endif
        bgzf_frame_general(F, x0, x1, runs, ctl, n, data_dst, tid, sm);
    __syncthreads();  // (the table and warp_crc are reused by the next block of this persistent thread block)
}

}  // namespace ptl
