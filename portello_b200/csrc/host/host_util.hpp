// Host-side CIGAR primitives of the product library (flat u32 BAM ops, no per-op objects).
// Behaviour follows lib/rust-vc-utils/src/bam_utils/cigar/mod.rs:16-327 of the reference; independent of oracle/.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace ptl {

enum : uint32_t { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };

inline uint32_t op_of(uint32_t c) { return c & 0xf; }
inline uint32_t len_of(uint32_t c) { return c >> 4; }
inline uint32_t mk(uint32_t op, uint64_t len) { return uint32_t(len << 4) | op; }

// bit masks over op codes
constexpr uint32_t kMatchMask = (1u << OP_M) | (1u << OP_EQ) | (1u << OP_X);
constexpr uint32_t kRefMask = kMatchMask | (1u << OP_D) | (1u << OP_N);                    // get_cigarseg_ref_offset :41-47
constexpr uint32_t kReadMask = kMatchMask | (1u << OP_I) | (1u << OP_S) | (1u << OP_H);    // get_cigarseg_read_offset, ignore_hard_clip=false :26-39
constexpr uint32_t kClipMask = (1u << OP_S) | (1u << OP_H);

inline bool is_match(uint32_t c) { return (kMatchMask >> op_of(c)) & 1; }
inline bool is_clip(uint32_t c) { return (kClipMask >> op_of(c)) & 1; }
inline uint64_t ref_adv(uint32_t c) { return ((kRefMask >> op_of(c)) & 1) ? len_of(c) : 0; }
inline uint64_t read_adv(uint32_t c) { return ((kReadMask >> op_of(c)) & 1) ? len_of(c) : 0; }

using Ops = std::vector<uint32_t>;

inline int64_t ref_span(const uint32_t* c, size_t n) {
    int64_t r = 0;
    for (size_t i = 0; i < n; ++i) r += int64_t(ref_adv(c[i]));
    return r;
}
inline uint64_t read_span(const uint32_t* c, size_t n) {
    uint64_t r = 0;
    for (size_t i = 0; i < n; ++i) r += read_adv(c[i]);
    return r;
}

// get_read_clip_positions(cigar, ignore_hard_clip = false), cigar/mod.rs:85-118 -> (left clip end, right clip start, size)
struct ClipPos { uint64_t left, right, size; };
inline ClipPos clip_positions(const uint32_t* c, size_t n) {
    uint64_t left = 0, right = 0, size = 0;
    bool in_left = true;
    for (size_t i = 0; i < n; ++i) {
        if (is_clip(c[i])) (in_left ? left : right) += len_of(c[i]);
        else in_left = false;
        size += read_adv(c[i]);
    }
    return {left, size - right, size};
}

// compress_cigar, cigar/mod.rs:204-228 (incl. the Pad-after-Pad drop), appending to `out`
inline void compress_into(const uint32_t* c, size_t n, Ops& out) {
    uint32_t last = mk(OP_M, 0);
    for (size_t i = 0; i < n; ++i) {
        if (len_of(c[i]) == 0) continue;
        if (op_of(c[i]) == op_of(last)) {
            if (op_of(last) != OP_P) last += c[i] & ~0xfu;
        } else {
            if (len_of(last)) out.push_back(last);
            last = c[i];
        }
    }
    if (len_of(last)) out.push_back(last);
}

struct InputError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

}  // namespace ptl
