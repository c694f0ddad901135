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

struct uint2 { uint32_t x, y; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

using std::max;
using std::min;

inline bool __any_sync(unsigned, bool pred) { return pred; }

inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { const unsigned int o = *p; *p += v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }
inline unsigned int atomicOr(unsigned int* p, unsigned int v) { const unsigned int o = *p; *p |= v; return o; }
inline long long atomicMin(long long* p, long long v) { const long long o = *p; *p = std::min(o, v); return o; }
