// Record assembly, bases: seq + qual of every output record, re-oriented like
// reverse_alignment_seq_and_qual (src/read_alignment_scanner.rs:125-133) when the record was flipped.
//
// HBM-bound streaming work (22.5 KB in and out per 15 kb record): one block per record, every thread produces aligned
// 16-byte output vectors from five aligned 32-bit loads of the (arbitrarily aligned, possibly mirrored) source range.
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
    const uint64_t* rec_seq_begin;   // [n_records+1], multiples of 16
    const uint64_t* rec_qual_begin;  // [n_records+1], multiples of 16
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

// 16 bytes of a byte stream starting at byte offset `off` of `base`, as 4 words.  Fast path (the 5 aligned words lie inside
// the read's own bytes [0, len)): 5 aligned loads + 4 funnel shifts; otherwise the guarded, zero-filling window32.
__device__ __forceinline__ uint4 window128(const uint8_t* __restrict__ base, int64_t off, int64_t len) {
    const uint64_t addr = reinterpret_cast<uint64_t>(base) + uint64_t(off);
    const int64_t a0 = off - int64_t(addr & 3ull);  // stream offset of the first aligned word
    uint4 v;
    if (a0 >= 0 && a0 + 20 <= len) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(addr & ~3ull);
        const uint32_t sh = uint32_t(addr & 3ull) * 8u;
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
        v.x = __funnelshift_r(w0, w1, sh);
        v.y = __funnelshift_r(w1, w2, sh);
        v.z = __funnelshift_r(w2, w3, sh);
        v.w = __funnelshift_r(w3, w4, sh);
    } else {
        v.x = window32(base, off, len);
        v.y = window32(base, off + 4, len);
        v.z = window32(base, off + 8, len);
        v.w = window32(base, off + 12, len);
    }
    return v;
}

// reverse-complement of 8 nibbles held in nibble-monotonic form
__device__ __forceinline__ uint32_t revcomp8(uint32_t win) { return non_acgt_to_n(__brev(win)); }

// One block per record; every thread produces 16 output bytes per step (records start 16-byte aligned in the pools).
__device__ __forceinline__ void assemble_record_body(const AsmArgs& A, uint32_t k, uint32_t tid, uint32_t n_threads) {
    const uint32_t r = A.rec_read[k];
    const int64_t len = A.read_seq_len[r];
    const bool flip = A.rec_flip[k] != 0;
    const uint8_t* src_q = A.qual + A.read_qual_off[r];
    const uint8_t* src_s = A.seq4 + A.read_seq_off[r];
    uint4* dst_q = reinterpret_cast<uint4*>(A.out_qual + A.rec_qual_begin[k]);
    uint4* dst_s = reinterpret_cast<uint4*>(A.out_seq4 + A.rec_seq_begin[k]);
    const int64_t seq_bytes = (len + 1) >> 1;
    const uint32_t n_qv = uint32_t((len + 15) >> 4), n_sv = uint32_t((seq_bytes + 15) >> 4);
    // qualities: out[i] = in[len-1-i]
    for (uint32_t j = tid; j < n_qv; j += n_threads) {
        uint4 v;
        if (!flip) {
            v = window128(src_q, int64_t(j) * 16, len);
        } else {
            const uint4 s = window128(src_q, len - 16 - int64_t(j) * 16, len);  // the 16 bytes in front of the mirrored position
            v.x = __byte_perm(s.w, 0u, 0x0123u);
            v.y = __byte_perm(s.z, 0u, 0x0123u);
            v.z = __byte_perm(s.y, 0u, 0x0123u);
            v.w = __byte_perm(s.x, 0u, 0x0123u);
        }
        dst_q[j] = v;
    }
    // bases: out nibble i = comp(in nibble len-1-i)
    for (uint32_t j = tid; j < n_sv; j += n_threads) {
        uint4 v;
        if (!flip) {
            v = window128(src_s, int64_t(j) * 16, seq_bytes);
        } else {
            // source nibbles [a, a+32), a = len - 32(j+1), in nibble-monotonic form (low nibble of a byte = the earlier base)
            const int64_t a = len - 32 * (int64_t(j) + 1);
            const int64_t byte0 = a >> 1;  // floor: a may be negative
            uint4 s = window128(src_s, byte0, seq_bytes);
            s.x = swap_nibbles(s.x); s.y = swap_nibbles(s.y); s.z = swap_nibbles(s.z); s.w = swap_nibbles(s.w);
            if (a & 1) {  // the window starts at the later base of byte0: shift by one nibble, 4 more bits from the 17th byte
                const uint32_t t = swap_nibbles(window32(src_s, byte0 + 16, seq_bytes));
                s.x = __funnelshift_r(s.x, s.y, 4u);
                s.y = __funnelshift_r(s.y, s.z, 4u);
                s.z = __funnelshift_r(s.z, s.w, 4u);
                s.w = __funnelshift_r(s.w, t, 4u);
            }
            uint32_t o[4] = {revcomp8(s.w), revcomp8(s.z), revcomp8(s.y), revcomp8(s.x)};
            // nibbles beyond the read (output index >= len) are padding: zero
            const int64_t first_out = int64_t(j) * 32;
            if (first_out + 32 > len) {
#pragma unroll
                for (int t = 0; t < 32; ++t)
                    if (first_out + t >= len) o[t >> 3] &= ~(0xfu << (4 * (t & 7)));
            }
            v.x = swap_nibbles(o[0]); v.y = swap_nibbles(o[1]); v.z = swap_nibbles(o[2]); v.w = swap_nibbles(o[3]);
        }
        dst_s[j] = v;
    }
}

}  // namespace ptl
