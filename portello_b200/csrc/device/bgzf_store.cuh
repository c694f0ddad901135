// BGZF framing on the device, compression level 0 (the reference's stdout mode writes exactly this:
// `--remapped-read-output -` sets CompressionLevel::Uncompressed, src/read_alignment_scanner.rs:66-71, "to optimize piping
// into samtools sort").  One BGZF block (SAM spec 4.1) per thread block:
//     18-byte gzip header with the BC size subfield | 01 LEN NLEN (one stored deflate block) | <= 0xff00 data bytes | CRC32 | ISIZE
// Input offsets and output offsets are closed forms of the block index (every block but the last is full), so there is
// no size pass.  HBM-bound: 64 KB in, 64 KB out per block, the CRC computed from shared memory in between.
//   1. the block's bytes are staged in shared memory at the alignment of their DESTINATION (five aligned 32-bit loads +
//      funnel shifts per 16 bytes, as in assemble_bam.cuh), so that the copy-out is aligned 128-bit loads and stores;
//   2. CRC-32 (IEEE, reflected): the 256 threads each run the byte-wise table update over 255 consecutive bytes with a
//      ZERO initial state; the slices are laid out from the END of the data (leading zeros do not change a zero-state CRC,
//      so a short last block needs no special case), and partial CRCs are combined pairwise up a tree:
//      crc(A || B) = Z^|B|(crc(A)) ^ crc(B), where Z^m = "advance the state through m zero bytes" is a linear map over
//      GF(2) held as a 32 x 32 bit matrix for m = 255 * 2^k bytes, k = 0..7 (host-precomputed);
//   3. the real CRC is Z^n(0xffffffff) ^ crc_zero_state(data) ^ 0xffffffff; Z^n(0xffffffff) is a per-launch constant for
//      full blocks and one more for the last block (host-computed).
#pragma once
#include <cstdint>

#include "assemble_bam.cuh"

namespace ptl {

constexpr uint32_t kBgzfIn = 0xff00u;             // payload bytes of a full block (htslib BGZF_BLOCK_SIZE)
constexpr uint32_t kBgzfOverhead = 18u + 5u + 8u;  // gzip header + stored-block header + CRC32 + ISIZE
constexpr uint32_t kBgzfTableWords = 256u + 8u * 32u;

struct BgzfArgs {
    const uint8_t* in;       // the stream (device); readable 16 bytes in front and 32 bytes behind
    uint64_t n;              // stream bytes
    uint8_t* out;            // n + 31 * n_blocks bytes
    const uint32_t* tables;  // [256] CRC table, [8][32] shift matrices
    uint32_t init_full;      // Z^0xff00(0xffffffff)
    uint32_t init_last;      // Z^(bytes of the last block)(0xffffffff)
    uint64_t n_blocks;
};

__device__ __forceinline__ uint32_t gf2_apply(const uint32_t* __restrict__ m, uint32_t v) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) r ^= ((v >> i) & 1u) ? m[i] : 0u;
    return r;
}

// blockDim.x == 256; dynamic shared memory: kBgzfIn + 48 bytes of data, then kBgzfTableWords words, then 8 words
__device__ __forceinline__ void bgzf_store_block_body(const BgzfArgs& A, uint64_t b, uint32_t tid, uint8_t* smem) {
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem + kBgzfIn + 48u);
    uint32_t* warp_crc = tab + kBgzfTableWords;
    const uint64_t in_off = b * kBgzfIn;
    const uint32_t n = uint32_t(min(uint64_t(kBgzfIn), A.n - in_off));
    uint8_t* dst = A.out + b * uint64_t(kBgzfIn + kBgzfOverhead);
    uint8_t* data_dst = dst + 23;
    const uint32_t a0 = uint32_t(reinterpret_cast<uint64_t>(data_dst) & 15ull);  // data sits at smem[a0, a0 + n)
    const uint8_t* src = A.in + in_off;
    for (uint32_t i = tid; i < kBgzfTableWords; i += 256u) tab[i] = A.tables[i];
    // ---- 1. stage (chunk c = smem bytes [16 c, 16 c + 16) = stream bytes starting at 16 c - a0)
    const uint32_t n_chunks = (a0 + n + 15u) >> 4;
    for (uint32_t c = tid; c < n_chunks; c += 256u)
        *reinterpret_cast<uint4*>(smem + 16u * c) = window128_body(src + int64_t(16u * c) - int64_t(a0));
    if (tid == 0) {  // gzip header + the stored-block header (unaligned destination: byte stores)
        const uint32_t bsize1 = n + kBgzfOverhead - 1u;
        const uint8_t h[23] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, uint8_t(bsize1 & 0xffu), uint8_t(bsize1 >> 8),
                               0x01, uint8_t(n & 0xffu), uint8_t(n >> 8), uint8_t(~n & 0xffu), uint8_t((~n >> 8) & 0xffu)};
        for (int i = 0; i < 23; ++i) dst[i] = h[i];
    }
    __syncthreads();
    // ---- 2. zero-state CRC of this thread's 255 bytes, counted from the end of the data
    const int32_t e = int32_t(a0 + n);
    const int32_t s_end = e - 255 * int32_t(255u - tid);
    int32_t p = max(s_end - 255, int32_t(a0));
    uint32_t crc = 0;
    for (; p < s_end; ++p) crc = tab[(crc ^ smem[p]) & 0xffu] ^ (crc >> 8);
    // tree combine: after level k a thread with tid % 2^(k+1) == 0 holds the CRC of 2^(k+1) slices
    const uint32_t lane = tid & 31u;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const uint32_t right = __shfl_down_sync(0xffffffffu, crc, 1u << k);
        if ((lane & ((2u << k) - 1u)) == 0u) crc = gf2_apply(tab + 256 + 32 * k, crc) ^ right;
    }
    if (lane == 0) warp_crc[tid >> 5] = crc;
    __syncthreads();
    if (tid < 32u) {
        crc = tid < 8u ? warp_crc[tid] : 0u;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t right = __shfl_down_sync(0xffffffffu, crc, 1u << k);
            if ((tid & ((2u << k) - 1u)) == 0u) crc = gf2_apply(tab + 256 + 32 * (5 + k), crc) ^ right;
        }
        if (tid == 0) {  // ---- 3. footer: CRC32, ISIZE
            const uint32_t full = ((n == kBgzfIn) ? A.init_full : A.init_last) ^ crc ^ 0xffffffffu;
            uint8_t* f = data_dst + n;
            for (int i = 0; i < 4; ++i) { f[i] = uint8_t(full >> (8 * i)); f[4 + i] = uint8_t(n >> (8 * i)); }
        }
    }
    // ---- copy out: aligned 16-byte chunks strictly inside [a0, a0 + n), bytes at the two ends
    uint8_t* dst16 = data_dst - a0;  // 16-byte aligned
    const uint32_t c_lo = (a0 + 15u) >> 4, c_hi = (a0 + n) >> 4;  // full chunks [c_lo, c_hi)
    for (uint32_t c = c_lo + tid; c < c_hi; c += 256u) *reinterpret_cast<uint4*>(dst16 + 16u * c) = *reinterpret_cast<const uint4*>(smem + 16u * c);
    if (c_lo <= c_hi) {
        for (uint32_t i = a0 + tid; i < min(16u * c_lo, a0 + n); i += 256u) dst16[i] = smem[i];
        for (uint32_t i = max(16u * c_hi, a0) + tid; i < a0 + n; i += 256u) dst16[i] = smem[i];
    } else {  // fewer than 16 bytes inside one chunk
        for (uint32_t i = a0 + tid; i < a0 + n; i += 256u) dst16[i] = smem[i];
    }
}

}  // namespace ptl
