// TEST INFRASTRUCTURE ONLY.  A 32-lane SIMT shim for the warp-cooperative device code (portello_b200/csrc/device/
// lift_warp.cuh): every lane of a "warp" is a host thread, every warp collective (shuffle, ballot, any, syncwarp) is an
// exchange through a shared slot array bracketed by ONE barrier (slots are double-buffered by collective parity, so a
// lane can never overwrite a value another lane has not read yet).  Slow (~13 us per collective) but bit-faithful to
// the lock-step semantics the kernels rely on.  Nothing under portello_b200/ includes this file.
#pragma once
#include <algorithm>
#include <barrier>
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

namespace warp_emul {
struct Warp {
    std::barrier<> bar{32};
    uint64_t slot[2][32];
};
inline thread_local Warp* tl_warp = nullptr;
inline thread_local uint32_t tl_lane = 0;
inline thread_local uint32_t tl_parity = 0;
inline thread_local std::barrier<>* tl_block = nullptr;  // __syncthreads of a multi-warp block (emul_bgzf.cpp)

// deposit `v`, wait for all lanes, return the whole slot array of this collective
inline const uint64_t* exchange(uint64_t v) {
    Warp* w = tl_warp;
    const uint32_t b = tl_parity;
    tl_parity ^= 1u;
    w->slot[b][tl_lane] = v;
    w->bar.arrive_and_wait();
    return w->slot[b];
}
template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
}  // namespace warp_emul

template <class T> inline T __shfl_sync(unsigned, T v, int src) { return warp_emul::from_bits<T>(warp_emul::exchange(warp_emul::to_bits(v))[src & 31]); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int d) {
    const uint64_t* s = warp_emul::exchange(warp_emul::to_bits(v));
    const int src = int(warp_emul::tl_lane) - d;
    return src >= 0 ? warp_emul::from_bits<T>(s[src]) : v;
}
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) {
    const uint64_t* s = warp_emul::exchange(warp_emul::to_bits(v));
    const int src = int(warp_emul::tl_lane) + d;
    return src < 32 ? warp_emul::from_bits<T>(s[src]) : v;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return warp_emul::from_bits<T>(warp_emul::exchange(warp_emul::to_bits(v))[(warp_emul::tl_lane ^ unsigned(m)) & 31]); }
inline unsigned __ballot_sync(unsigned, bool pred) {
    const uint64_t* s = warp_emul::exchange(pred ? 1u : 0u);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= unsigned(s[i] & 1u) << i;
    return m;
}
inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0u; }
inline void __syncwarp() { warp_emul::exchange(0); }
inline void __syncthreads() { warp_emul::tl_block->arrive_and_wait(); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline int __ffs(unsigned v) { return __builtin_ffs(int(v)); }

// debug hook of lift_warp.cuh: values the kernels treat as warp-uniform really are identical in all 32 lanes
#include <cstdio>
#include <cstdlib>
template <class T> inline void ptl_assert_uniform(T v, const char* what, int line) {
    const uint64_t* s = warp_emul::exchange(warp_emul::to_bits(v));
    for (int i = 1; i < 32; ++i)
        if (s[i] != s[0]) { std::fprintf(stderr, "lift_warp.cuh:%d: %s differs between lanes (lane 0: %llu, lane %d: %llu)\n", line, what, (unsigned long long)s[0], i, (unsigned long long)s[i]); std::abort(); }
}
#define PTL_ASSERT_UNIFORM(v) ptl_assert_uniform((v), #v, __LINE__)

// only lane 0 of a warp calls these, and the emulation runs one warp at a time
inline unsigned int atomicOr(unsigned int* p, unsigned int v) { const unsigned int o = *p; *p |= v; return o; }
inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { const unsigned int o = *p; *p += v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }
