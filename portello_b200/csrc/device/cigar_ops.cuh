// BAM CIGAR op codes, op classes and small helpers shared by the per-thread stages (lift_device.cuh) and the
// warp-cooperative stages (lift_warp.cuh).  No warp intrinsics here.
#pragma once
#include <cstdint>

namespace ptl {

enum : uint32_t { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };
constexpr uint32_t kMatchMask = (1u << OP_M) | (1u << OP_EQ) | (1u << OP_X);
constexpr uint32_t kRefMask = kMatchMask | (1u << OP_D) | (1u << OP_N);
constexpr uint32_t kReadMask = kMatchMask | (1u << OP_I) | (1u << OP_S) | (1u << OP_H);
constexpr uint32_t NO_OP = 0xffffffffu;
constexpr uint32_t FULL = 0xffffffffu;

constexpr int ST_LIFTED = 1, ST_NONE = 0, ST_PENDING_SIMPLIFY = 2, ST_PENDING_LIFT = 3, ST_ERR_LENGTH = -1, ST_ERR_BOUNDS = -2, ST_ERR_CAPACITY = -3;

__device__ __forceinline__ bool op_is_match(uint32_t op) { return (kMatchMask >> op) & 1u; }
__device__ __forceinline__ uint32_t op_ref_adv(uint32_t c) { return ((kRefMask >> (c & 0xf)) & 1u) ? (c >> 4) : 0u; }
__device__ __forceinline__ uint32_t op_read_adv(uint32_t c) { return ((kReadMask >> (c & 0xf)) & 1u) ? (c >> 4) : 0u; }

// bam_reg2bin (lib/rust-vc-utils/src/bam_utils/util.rs:10-35)
__device__ __forceinline__ uint16_t reg2bin(int64_t begin, int64_t end) {
    const uint64_t b = uint64_t(begin), e = uint64_t(end) - 1ull;
    if ((b >> 14) == (e >> 14)) return uint16_t(4681u + (b >> 14));
    if ((b >> 17) == (e >> 17)) return uint16_t(585u + (b >> 17));
    if ((b >> 20) == (e >> 20)) return uint16_t(73u + (b >> 20));
    if ((b >> 23) == (e >> 23)) return uint16_t(9u + (b >> 23));
    if ((b >> 26) == (e >> 26)) return uint16_t(1u + (b >> 26));
    return 0;
}

}  // namespace ptl
