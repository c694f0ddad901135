// Record assembly, bases: seq + qual of every output record, re-oriented like
// reverse_alignment_seq_and_qual (src/read_alignment_scanner.rs:125-133) when the record was flipped.
//
// HBM-bound streaming work (22.5 KB in and out per 15 kb record): one block per record, every thread produces aligned
// 32-bit output words from two aligned 32-bit loads of the (arbitrarily aligned, possibly mirrored) source range.
//   * qualities reversed: the 4 source bytes in front of the mirrored position, byte-swapped
//   * bases reverse-complemented WITHOUT decoding: with the two nibbles of every byte swapped, nibble t of the read sits at
//     bits [4t, 4t+4) of the byte stream; __brev of a 32-bit window then reverses the ORDER of its 8 nibbles and the BITS
//     of each one, and reversing the 4 bits of a one-hot BAM code is its complement (A=1 <-> T=8, C=2 <-> G=4); every
//     other code ('=', IUPAC, N) must become N=15, which is "popcount of the nibble != 1" in SWAR form
//     (rev_comp_in_place + comp_base, lib/rust-vc-utils/src/seq_util.rs:1-40; Record::set re-encodes A,C,G,T,N as 1,2,4,8,15)
#pragma once
#include <cstdint>

namespace ptl {

struct AsmArgs {
    uint32_t n_records;
    const uint32_t* rec_read;      // record -> batch read index
    const uint8_t* rec_flip;
    const uint32_t* read_seq_len;
    const uint64_t* read_seq_off;  // into seq4
    const uint8_t* seq4;
    const uint64_t* read_qual_off;
    const uint8_t* qual;
    const uint64_t* rec_seq_begin;   // [n_records+1], multiples of 4
    const uint64_t* rec_qual_begin;  // [n_records+1], multiples of 4
    uint8_t* out_seq4;
    uint8_t* out_qual;
};

// 32 bits of a byte stream starting at byte offset `off` of `base` (off may be negative or run past `len`: bytes outside
// [0, len) read as 0).  Two aligned loads inside [0, len) rounded out to words (the pools are padded).
__device__ __forceinline__ uint32_t window32(const uint8_t* __restrict__ base, int64_t off, int64_t len) {
    const uint64_t addr = reinterpret_cast<uint64_t>(base) + uint64_t(off);  // (two's complement: fine for negative off)
    const uint32_t sh = uint32_t(addr & 3ull) * 8u;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(addr & ~3ull);
    // which of the two words touch [0, len)?
    const int64_t w0_off = off - int64_t(addr & 3ull);
    const uint32_t lo = (w0_off + 4 > 0 && w0_off < len) ? w[0] : 0u;
    const uint32_t hi = (sh && w0_off + 8 > 0 && w0_off + 4 < len) ? w[1] : 0u;
    uint32_t v = __funnelshift_r(lo, hi, sh);
    // zero the bytes outside [0, len)
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (off + k < 0 || off + k >= len) v &= ~(0xffu << (8 * k));
    return v;
}

__device__ __forceinline__ uint32_t swap_nibbles(uint32_t w) { return ((w & 0x0f0f0f0fu) << 4) | ((w >> 4) & 0x0f0f0f0fu); }

// nibbles that are not one-hot (popcount != 1) become 0xf
__device__ __forceinline__ uint32_t non_acgt_to_n(uint32_t w) {
    const uint32_t pc = w - ((w >> 1) & 0x77777777u) - ((w >> 2) & 0x33333333u) - ((w >> 3) & 0x11111111u);  // popcount per nibble
    const uint32_t x = pc ^ 0x11111111u;                                                                   // 0 where popcount == 1
    const uint32_t bad = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;                               // bit 0 of the nibble set where != 1
    return w | (bad * 0xfu);
}

// One block per record.
__device__ __forceinline__ void assemble_record_body(const AsmArgs& A, uint32_t k, uint32_t tid, uint32_t n_threads) {
    const uint32_t r = A.rec_read[k];
    const int64_t len = A.read_seq_len[r];
    const bool flip = A.rec_flip[k] != 0;
    const uint8_t* src_q = A.qual + A.read_qual_off[r];
    const uint8_t* src_s = A.seq4 + A.read_seq_off[r];
    uint32_t* dst_q = reinterpret_cast<uint32_t*>(A.out_qual + A.rec_qual_begin[k]);
    uint32_t* dst_s = reinterpret_cast<uint32_t*>(A.out_seq4 + A.rec_seq_begin[k]);
    const int64_t seq_bytes = (len + 1) >> 1;
    const uint32_t n_qw = uint32_t((len + 3) >> 2), n_sw = uint32_t((seq_bytes + 3) >> 2);
    // qualities: out[i] = in[len-1-i]
    for (uint32_t j = tid; j < n_qw; j += n_threads) {
        uint32_t v;
        if (!flip) v = window32(src_q, int64_t(j) * 4, len);
        else v = __byte_perm(window32(src_q, len - 4 - int64_t(j) * 4, len), 0u, 0x0123u);
        dst_q[j] = v;
    }
    // bases: out nibble i = comp(in nibble len-1-i)
    for (uint32_t j = tid; j < n_sw; j += n_threads) {
        uint32_t v;
        if (!flip) {
            v = window32(src_s, int64_t(j) * 4, seq_bytes);
        } else {
            // source nibbles [a, a+8), a = len - 8(j+1), in nibble-monotonic form (low nibble of a byte = the earlier base)
            const int64_t a = len - 8 * (int64_t(j) + 1);
            const int64_t byte0 = a >> 1;  // floor: a may be negative
            const uint32_t w0 = swap_nibbles(window32(src_s, byte0, seq_bytes));
            uint32_t win = w0;
            if (a & 1) {  // the window starts at the later base of byte0: 4 more bits from the next byte
                const uint32_t w1 = swap_nibbles(window32(src_s, byte0 + 4, seq_bytes));
                win = __funnelshift_r(w0, w1, 4u);
            }
            uint32_t rc = non_acgt_to_n(__brev(win));
            // nibbles that lie beyond the read (source index < 0, i.e. output index >= len) are padding: zero
            const int64_t first_out = int64_t(j) * 8;
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (first_out + t >= len) rc &= ~(0xfu << (4 * t));
            v = swap_nibbles(rc);
        }
        dst_s[j] = v;
    }
}

}  // namespace ptl
