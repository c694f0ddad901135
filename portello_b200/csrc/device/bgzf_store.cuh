// BGZF framing on the device, compression level 0 (the reference's stdout mode writes exactly this:
// `--remapped-read-output -` sets CompressionLevel::Uncompressed, src/read_alignment_scanner.rs:66-71, "to optimize piping
// into samtools sort").  One BGZF block (SAM spec 4.1) per thread block:
//     18-byte gzip header with the BC size subfield | 01 LEN NLEN (one stored deflate block) | <= 0xff00 data bytes | CRC32 | ISIZE
// Input offsets and output offsets are closed forms of the block index (every block but the last is full), so there is
// no size pass.  HBM-bound: 64 KB in, 64 KB out per block, the CRC computed from shared memory in between.
//   1. the block's bytes are staged in shared memory so that they END at a fixed word-aligned offset (five aligned 32-bit
//      global loads + funnel shifts per 16 bytes, as in assemble_bam.cuh); the copy-out realigns them to the destination
//      the same way from shared memory;
//   2. CRC-32 (IEEE, reflected): the 256 threads each run the slice-by-4 table update over 64 consecutive words with a
//      ZERO initial state; the slices are laid out from the END of the data (leading zeros do not change a zero-state CRC,
//      so a short last block needs no special case), and partial CRCs are combined pairwise up a tree:
//      crc(A || B) = Z^|B|(crc(A)) ^ crc(B), where Z^m = "advance the state through m zero bytes" is a linear map over
//      GF(2) held as a 32 x 32 bit matrix for m = 256 * 2^k bytes, k = 0..7 (host-precomputed).  (First version: byte-wise
//      update over 255-byte slices, 4-way bank conflicts on the data: 0.82 TB/s of input; the staging buffer now carries a
//      padding word per 64 words so that the slices of a warp fall on different banks);
//   3. the real CRC is Z^n(0xffffffff) ^ crc_zero_state(data) ^ 0xffffffff; Z^n(0xffffffff) is a per-launch constant for
//      full blocks and one more for the last block (host-computed).
#pragma once
#include <cstdint>

#include "assemble_bam.cuh"

namespace ptl {

constexpr uint32_t kBgzfIn = 0xff00u;             // payload bytes of a full block (htslib BGZF_BLOCK_SIZE)
constexpr uint32_t kBgzfOverhead = 18u + 5u + 8u;  // gzip header + stored-block header + CRC32 + ISIZE
// slice-by-4 CRC tables [4][256]; nibble tables [6][8][16] of the zero-byte shifts by 128 B (the two chains of a thread)
// and by 256 * 2^k B, k = 0..4 (the warp levels of the tree); bit matrices [3][32] for k = 5..7 (the block levels)
constexpr uint32_t kBgzfNibble = 1024u, kBgzfMatrix = kBgzfNibble + 6u * 128u, kBgzfTableWords = kBgzfMatrix + 3u * 32u;
constexpr uint32_t kBgzfEnd = kBgzfIn + 16u;       // the data ENDS at this (16-byte aligned) logical byte of the staging buffer
// One padding word per 64 data words: the 256 CRC slices start 64 words apart, which would put all lanes of a warp on one
// shared-memory bank; with the padding they are 65 words apart.
__device__ __forceinline__ uint32_t bgzf_pw(uint32_t logical_word) { return logical_word + (logical_word >> 6); }
constexpr uint32_t kBgzfDataWords = (kBgzfEnd / 4u + 8u) + ((kBgzfEnd / 4u + 8u) >> 6) + 1u;
constexpr uint32_t kBgzfSmemBytes = (kBgzfDataWords + kBgzfTableWords + 8u) * 4u;

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
__device__ __forceinline__ uint32_t gf2_apply(const uint32_t* __restrict__ m, uint32_t v) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) r ^= ((v >> i) & 1u) ? m[i] : 0u;
    return r;
}

// blockDim.x == 256; kBgzfSmemBytes of dynamic shared memory (words): [staging buffer | tables | 8 warp CRCs]
__device__ __forceinline__ void bgzf_store_block_body(const BgzfArgs& A, uint64_t b, uint32_t tid, uint32_t* sm) {
    uint32_t* tab = sm + kBgzfDataWords;
    uint32_t* warp_crc = tab + kBgzfTableWords;
    const uint64_t in_off = b * kBgzfIn;
    const uint32_t n = uint32_t(min(uint64_t(kBgzfIn), A.n - in_off));
    uint8_t* dst = A.out + b * uint64_t(kBgzfIn + kBgzfOverhead);
    uint8_t* data_dst = dst + 23;
    const uint8_t* src = A.in + in_off;
    const uint32_t start = kBgzfEnd - n;  // the data sits at logical bytes [start, kBgzfEnd)
    for (uint32_t i = tid; i < kBgzfTableWords; i += 256u) tab[i] = A.tables[i];
    // ---- 1. stage: logical chunk c = bytes [16 c, 16 c + 16) = stream bytes from 16 c - start; bytes in front of the data are zero
    const uint32_t c0 = start >> 4, c1 = kBgzfEnd >> 4;
    // (four chunks = twenty global loads in flight per thread: one chunk at a time, the 16 dependent round trips of this loop
    //  were most of the block's 29 us)
    for (uint32_t cb = c0 + tid; cb < c1; cb += 1024u) {
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = cb + 256u * j;
            if (c < c1) v[j] = window128_body(src + int64_t(16u * c) - int64_t(start));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = cb + 256u * j;
            if (c >= c1) continue;
            uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            if (c == c0) {
                const uint32_t k = start & 15u;  // the first k bytes of this chunk precede the data
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int cut = int(k) - 4 * i;  // bytes of word i to clear
                    if (cut >= 4) w[i] = 0u;
                    else if (cut > 0) w[i] &= 0xffffffffu << (8 * cut);
                }
            }
            const uint32_t p = bgzf_pw(4u * c);  // (a chunk never straddles a padding word: 64 is a multiple of 4)
            sm[p] = w[0]; sm[p + 1] = w[1]; sm[p + 2] = w[2]; sm[p + 3] = w[3];
        }
    }
    if (tid == 0) {  // gzip header + the stored-block header (unaligned destination: byte stores)
        const uint32_t bsize1 = n + kBgzfOverhead - 1u;
        const uint8_t h[23] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, uint8_t(bsize1 & 0xffu), uint8_t(bsize1 >> 8),
                               0x01, uint8_t(n & 0xffu), uint8_t(n >> 8), uint8_t(~n & 0xffu), uint8_t((~n >> 8) & 0xffu)};
        for (int i = 0; i < 23; ++i) dst[i] = h[i];
    }
    __syncthreads();
    // ---- 2. zero-state CRC of this thread's 64 words (slice-by-4), slices counted back from the end of the data; the two
    //         halves of a slice run as two independent chains (the table lookups of one chain are a dependent sequence)
    const int32_t lw_end = int32_t(kBgzfEnd / 4u) - 64 * int32_t(255u - tid);
    const int32_t lw_min = int32_t(4u * c0);
    uint32_t crc_a = 0, crc_b = 0;
    auto step = [&](uint32_t crc, int32_t lw) {
        if (lw < lw_min) return crc;  // (in front of the data: zero bytes on a zero state)
        crc ^= sm[bgzf_pw(uint32_t(lw))];
        return tab[768u + (crc & 0xffu)] ^ tab[512u + ((crc >> 8) & 0xffu)] ^ tab[256u + ((crc >> 16) & 0xffu)] ^ tab[crc >> 24];
    };
    if (lw_end > lw_min) {
#pragma unroll 4
        for (int32_t i = 0; i < 32; ++i) {
            crc_a = step(crc_a, lw_end - 64 + i);
            crc_b = step(crc_b, lw_end - 32 + i);
        }
    }
    uint32_t crc = gf2_apply_nibbles(tab + kBgzfNibble, crc_a) ^ crc_b;
    // tree combine: after level k a thread with tid % 2^(k+1) == 0 holds the CRC of 2^(k+1) slices
    const uint32_t lane = tid & 31u;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const uint32_t right = __shfl_down_sync(0xffffffffu, crc, 1u << k);
        if ((lane & ((2u << k) - 1u)) == 0u) crc = gf2_apply_nibbles(tab + kBgzfNibble + 128 * (k + 1), crc) ^ right;
    }
    if (lane == 0) warp_crc[tid >> 5] = crc;
    __syncthreads();
    if (tid < 32u) {
        crc = tid < 8u ? warp_crc[tid] : 0u;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t right = __shfl_down_sync(0xffffffffu, crc, 1u << k);
            if ((tid & ((2u << k) - 1u)) == 0u) crc = gf2_apply(tab + kBgzfMatrix + 32 * k, crc) ^ right;
        }
        if (tid == 0) {  // ---- 3. footer: CRC32, ISIZE
            const uint32_t full = ((n == kBgzfIn) ? A.init_full : A.init_last) ^ crc ^ 0xffffffffu;
            uint8_t* f = data_dst + n;
            for (int i = 0; i < 4; ++i) { f[i] = uint8_t(full >> (8 * i)); f[4 + i] = uint8_t(n >> (8 * i)); }
        }
    }
    // ---- copy out: 16-byte chunks aligned to the destination (five staged words + funnel shifts), bytes at the two ends
    const uint32_t a0 = uint32_t(reinterpret_cast<uint64_t>(data_dst) & 15ull);
    uint8_t* dst16 = data_dst - a0;  // 16-byte aligned; destination chunk d holds data bytes [16 d - a0, 16 d - a0 + 16)
    auto byte_at = [&](uint32_t i) {  // data byte i
        const uint32_t q = start + i;
        return uint8_t(sm[bgzf_pw(q >> 2)] >> (8u * (q & 3u)));
    };
    const uint32_t d_lo = (a0 + 15u) >> 4, d_hi = (a0 + n) >> 4;  // full chunks [d_lo, d_hi)
    for (uint32_t d = d_lo + tid; d < d_hi; d += 256u) {
        const uint32_t q = start + 16u * d - a0, w0 = q >> 2, sh = (q & 3u) * 8u;
        const uint32_t x0 = sm[bgzf_pw(w0)], x1 = sm[bgzf_pw(w0 + 1)], x2 = sm[bgzf_pw(w0 + 2)], x3 = sm[bgzf_pw(w0 + 3)], x4 = sm[bgzf_pw(w0 + 4)];
        uint4 v;
        v.x = __funnelshift_r(x0, x1, sh);
        v.y = __funnelshift_r(x1, x2, sh);
        v.z = __funnelshift_r(x2, x3, sh);
        v.w = __funnelshift_r(x3, x4, sh);
        *reinterpret_cast<uint4*>(dst16 + 16u * d) = v;
    }
    if (d_lo <= d_hi) {
        for (uint32_t i = tid; i < min(16u * d_lo - a0, n); i += 256u) data_dst[i] = byte_at(i);
        for (uint32_t i = max(16u * d_hi, a0) - a0 + tid; i < n; i += 256u) data_dst[i] = byte_at(i);
    } else {  // fewer than 16 bytes inside one chunk
        for (uint32_t i = tid; i < n; i += 256u) data_dst[i] = byte_at(i);
    }
}

}  // namespace ptl
