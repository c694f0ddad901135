// TEST INFRASTRUCTURE ONLY.  Host emulation of warp_pairs_kernel (portello_b200/csrc/device/kernels.cu): the
// warp-cooperative liftover of lift_warp.cuh run by 32 host threads in lock step (cuda_shim_warp.hpp), one warp at a
// time over DevWork::long_list.  Compiled as its own translation unit because the rest of the emulation uses the
// one-lane shim.
#include "cuda_shim_warp.hpp"

#include <thread>
#include <vector>

#include "../../portello_b200/csrc/device/lift_warp.cuh"
#include "../../portello_b200/csrc/device/table_warp.cuh"

void emul_lift_long_pairs(const ptl::DevStatic& S, const ptl::DevBatch& B, const ptl::DevWork& W, ptl::DevTotals* T, uint32_t stage_mask) {
    const uint32_t n_long = std::min<uint32_t>(T->n_long, W.pair_cap);
    const uint32_t n_simp = ((stage_mask & 6u) == 6u) ? std::min<uint32_t>(T->n_simplify, W.pair_cap) : 0u;
    if (!(n_long + n_simp)) return;
    warp_emul::Warp warp;
    std::vector<std::thread> lanes;
    for (uint32_t lane = 0; lane < 32; ++lane)
        lanes.emplace_back([&, lane] {
            warp_emul::tl_warp = &warp;
            warp_emul::tl_lane = lane;
            warp_emul::tl_parity = 0;
            for (uint32_t t = 0; t < n_long + n_simp; ++t) {
                if (t < n_long) ptl::lift_long_pair_body(S, B, W, T, W.long_list[t], lane, stage_mask);
                // (odd worklist entries: the warp-per-pair formulation; emul.cpp runs the even ones through the thread-per-pair
                //  body the product launches, so both formulations of a9 stay covered against the oracle)
                else if ((t - n_long) & 1u) ptl::simplify_warp_pair_body(S, B, W, T, W.simplify_list[t - n_long], lane);
            }
        });
    for (auto& th : lanes) th.join();
}

// table_build_kernel (one warp per segment, lanes over CIGAR ops): count pass into `counts` (out == nullptr) or fill pass.
void emul_table_build_warp(const ptl::DevStatic& S, uint32_t* counts, ptl::TabEntry* out) {
    warp_emul::Warp warp;
    std::vector<std::thread> lanes;
    for (uint32_t lane = 0; lane < 32; ++lane)
        lanes.emplace_back([&, lane] {
            warp_emul::tl_warp = &warp;
            warp_emul::tl_lane = lane;
            warp_emul::tl_parity = 0;
            for (uint32_t g = 0; g < S.n_segments; ++g) ptl::table_build_warp_body(S, g, lane, counts, out);
        });
    for (auto& th : lanes) th.join();
}

// pair_count_warp_kernel: every read counted by one warp (32 lanes in lock step)
void emul_pair_count_warp(const ptl::DevStatic& S, const ptl::DevBatch& B, const ptl::DevWork& W, ptl::DevTotals* T) {
    warp_emul::Warp warp;
    std::vector<std::thread> lanes;
    for (uint32_t lane = 0; lane < 32; ++lane)
        lanes.emplace_back([&, lane] {
            warp_emul::tl_warp = &warp;
            warp_emul::tl_lane = lane;
            warp_emul::tl_parity = 0;
            for (uint32_t r = 0; r < B.n_reads; ++r) ptl::pair_count_warp_body(S, B, W, T, r, lane);
        });
    for (auto& th : lanes) th.join();
}
