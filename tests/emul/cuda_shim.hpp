// TEST INFRASTRUCTURE ONLY.  A one-lane SIMT shim: lets g++ compile the product's per-index device bodies
// (portello_b200/csrc/device/{lift_device,pair_bodies}.cuh) for the host, every "warp" holding a single active lane
// (warp votes degenerate to the lane's own predicate).  Nothing under portello_b200/ includes this file.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __noinline__

struct uint2 { uint32_t x, y; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

struct uint4 { uint32_t x, y, z, w; };  // (no alignment attribute: the host may store it to any address)
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return uint32_t(((uint64_t(hi) << 32) | lo) >> (sh & 31u)); }
inline uint32_t __brev(uint32_t v) { uint32_t r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t v = (uint64_t(y) << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= uint32_t((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}

using std::max;
using std::min;

inline bool __any_sync(unsigned, bool pred) { return pred; }
inline uint32_t __shfl_up_sync(unsigned, uint32_t v, int) { return v; }  // (one lane: callers guard the use with lane >= delta)
struct EmulDim3 { uint32_t x = 0, y = 0, z = 0; };
static const EmulDim3 threadIdx{};  // the single lane of the shim is lane 0

inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { const unsigned int o = *p; *p += v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }
inline unsigned int atomicOr(unsigned int* p, unsigned int v) { const unsigned int o = *p; *p |= v; return o; }
inline long long atomicMin(long long* p, long long v) { const long long o = *p; *p = std::min(o, v); return o; }
