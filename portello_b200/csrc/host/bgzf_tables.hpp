// Host-side tables of the device BGZF framing (bgzf_store.cuh): the CRC-32 byte table and the nibble tables of the zero-byte
// shift maps.  Shared by the product (context.cu) and the host emulation of the kernel (tests/emul).
#pragma once
#include <cstdint>
#include <vector>

#include "../device/bgzf_store.cuh"

namespace ptl {

// CRC-32 (IEEE 802.3, reflected): the state advanced through n zero bytes (a linear map over GF(2))
inline uint32_t crc_zero_bytes(const uint32_t* t, uint32_t v, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) v = t[v & 0xffu] ^ (v >> 8);
    return v;
}
inline std::vector<uint32_t> bgzf_tables() {
    std::vector<uint32_t> t(kBgzfTableWords);
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
        t[i] = c;
    }
    // nibble tables of the zero-byte shifts by 128 B and by 256 * 2^k B (k = 0..7), then of the 4-byte word step
    for (int lvl = 0; lvl < 9; ++lvl) {
        const uint64_t m = lvl == 0 ? 128ull : (256ull << (lvl - 1));
        for (int j = 0; j < 8; ++j)
            for (uint32_t x = 0; x < 16; ++x) t[kBgzfShift + 128 * lvl + 16 * j + x] = crc_zero_bytes(t.data(), x << (4 * j), m);
    }
    for (int j = 0; j < 8; ++j)
        for (uint32_t x = 0; x < 16; ++x) t[kBgzfWord + 16 * j + x] = crc_zero_bytes(t.data(), x << (4 * j), 4);
    return t;
}

}  // namespace ptl
