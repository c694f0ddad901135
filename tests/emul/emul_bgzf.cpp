// TEST INFRASTRUCTURE ONLY.  Host emulation of bgzf_store_kernel (portello_b200/csrc/device/bgzf_store.cuh): one thread block =
// 256 host threads, 8 warps in lock step (cuda_shim_warp.hpp) plus a block barrier for __syncthreads, persistent over the
// BGZF blocks of the stream like the kernel.  The CPU suite checks its output against a Python model of htslib's level-0
// framing and against python's gzip reader.
#include "cuda_shim_warp.hpp"

#include <thread>
#include <vector>

#include "../../include/portello_b200.h"
#include "../../portello_b200/csrc/host/bgzf_tables.hpp"

extern "C" int64_t ptl_emul_bgzf_store(const uint8_t* in, uint64_t n, int append_eof, uint8_t* out, uint64_t cap) {
    using namespace ptl;
    static const uint8_t kEof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const uint64_t n_blocks = (n + kBgzfIn - 1) / kBgzfIn;
    const uint64_t framed = n + uint64_t(kBgzfOverhead) * n_blocks, total = framed + (append_eof ? sizeof(kEof) : 0);
    if ((n && !in) || !out || total > cap) return PTL_ERR_INVALID_ARG;
    // the stream with the readable margins the kernel assumes (the caller's alignment of `in` is kept modulo 16)
    std::vector<uint8_t> buf(n + 65536 + 64 + 16, 0);
    uint8_t* src = buf.data() + 65536;
    src += (16 - (reinterpret_cast<uintptr_t>(src) & 15u) + (reinterpret_cast<uintptr_t>(in) & 15u)) & 15u;
    if (n) std::memcpy(src, in, n);
    const std::vector<uint32_t> tables = bgzf_tables();
    BgzfArgs A{};
    A.in = src; A.n = n; A.out = out; A.tables = tables.data(); A.n_blocks = n_blocks;
    A.init_full = crc_zero_bytes(tables.data(), 0xffffffffu, kBgzfIn);
    A.init_last = crc_zero_bytes(tables.data(), 0xffffffffu, n_blocks ? n - (n_blocks - 1) * kBgzfIn : 0);
    if (n_blocks) {
        std::vector<uint32_t> smem(kBgzfSmemWords);
        std::vector<warp_emul::Warp> warps(8);
        std::barrier<> block(256);
        std::vector<std::thread> threads;
        for (uint32_t tid = 0; tid < 256; ++tid)
            threads.emplace_back([&, tid] {
                warp_emul::tl_warp = &warps[tid >> 5];
                warp_emul::tl_lane = tid & 31u;
                warp_emul::tl_parity = 0;
                warp_emul::tl_block = &block;
                bgzf_store_init(A, tid, smem.data());
                for (uint64_t b = 0; b < n_blocks; ++b) bgzf_store_block_body(A, b, tid, smem.data());
            });
        for (auto& t : threads) t.join();
    }
    if (append_eof) std::memcpy(out + framed, kEof, sizeof(kEof));
    return int64_t(total);
}
