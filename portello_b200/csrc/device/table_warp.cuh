// Warp-cooperative bodies of the setup / enumeration kernels.
// a7, warp-cooperative: the flat ReadToRefTreeMap of one contig->reference segment built by ONE WARP, lanes over CIGAR ops.
// Kept in a header so that tests/emul can run it with 32 host threads in lock step against the scalar statement of the
// algorithm (pair_bodies.cuh: table_build_body).  See kernels.cu for the description.
#pragma once
#include <cstdint>

#include "cigar_ops.cuh"
#include "device_types.hpp"

namespace ptl {

template <class T>
__device__ __forceinline__ T warp_incl_add(T v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const T o = __shfl_up_sync(FULL, v, d);
        if (int(lane) >= d) v += o;
    }
    return v;
}

__device__ __forceinline__ void table_build_warp_body(const DevStatic& S, uint32_t g, uint32_t lane, uint32_t* counts, TabEntry* out) {
    const uint32_t* c = S.seg_cigar + S.seg_cigar_begin[g];
    const uint64_t n = S.seg_cigar_begin[g + 1] - S.seg_cigar_begin[g];
    TabEntry* o = out ? out + S.seg_tab_begin[g] : nullptr;
    const uint32_t lt = (1u << lane) - 1u;
    // warp-uniform carries
    long long ref_pos = S.seg_pos[g];
    unsigned long long read_pos = 0, ml = 0;  // ml: length of the open match stretch
    uint32_t n_close = 0, n_over = 0, last_key = 0;
    long long last_ref_end = 0;
    bool have_last = false;
    auto emit = [&](bool closer, uint32_t idx, uint32_t k0, long long v0, uint32_t end_key, bool have_prev, long long prev_ref_end) {
        if (!o) return;
        if (closer) o[idx + 1] = TabEntry{end_key, -1, 0u, 0u};
        __syncwarp();
        if (closer) {
            const long long d = have_prev ? v0 - prev_ref_end : 0;
            o[idx] = TabEntry{k0, int32_t(v0), d > 0 ? uint32_t(d) : 0u, 0u};
        }
        __syncwarp();
    };
    for (uint64_t base = 0; base < n; base += 32) {
        const bool active = base + lane < n;
        const uint32_t x = active ? c[base + lane] : 0u;  // (padding = a match of length 0: neutral)
        const bool m = op_is_match(x & 0xfu);
        const unsigned long long mlen = m ? (x >> 4) : 0u, qa = op_read_adv(x), ra = op_ref_adv(x);
        const unsigned long long rp_before = read_pos + warp_incl_add(qa, lane) - qa;
        const long long fp_before = ref_pos + (long long)(warp_incl_add(ra, lane) - ra);
        const unsigned long long pm = warp_incl_add(mlen, lane);
        // match stretch in front of this op: back to the last non-match lane, or into the previous rounds
        const uint32_t nm = __ballot_sync(FULL, !m);
        const uint32_t nm_before = nm & lt;
        const int j = nm_before ? 31 - __clz(nm_before) : -1;
        const unsigned long long pm_j = __shfl_sync(FULL, pm, j < 0 ? 0 : j);
        const unsigned long long ml_before = (j < 0) ? ml + (pm - mlen) : (pm - mlen) - pm_j;
        const bool closer = !m && ml_before > 0;
        const uint32_t cb = __ballot_sync(FULL, closer);
        const uint32_t end_key = uint32_t(rp_before), k0 = uint32_t(rp_before - ml_before);
        const long long v0 = fp_before - (long long)ml_before;
        // the closer in front of this one (key and reference end of its run)
        const uint32_t cb_before = cb & lt;
        const int pl = cb_before ? 31 - __clz(cb_before) : -1;
        const uint32_t pk = __shfl_sync(FULL, end_key, pl < 0 ? 0 : pl);
        const long long pr = __shfl_sync(FULL, fp_before, pl < 0 ? 0 : pl);
        const bool have_prev = pl >= 0 || have_last;
        const uint32_t prev_key = pl >= 0 ? pk : last_key;
        const long long prev_ref_end = pl >= 0 ? pr : last_ref_end;
        const bool over = closer && have_prev && prev_key == k0;
        const uint32_t ob = __ballot_sync(FULL, over);
        const uint32_t idx = 2u * (n_close + __popc(cb_before)) - (n_over + __popc(ob & (lt | (1u << lane))));
        emit(closer, idx, k0, v0, end_key, have_prev, prev_ref_end);
        // carries
        const unsigned long long pm_all = __shfl_sync(FULL, pm, 31);
        if (nm) {
            const int jl = 31 - __clz(nm);
            ml = pm_all - __shfl_sync(FULL, pm, jl);
        } else {
            ml += pm_all;
        }
        if (cb) {
            const int cl = 31 - __clz(cb);
            last_key = __shfl_sync(FULL, end_key, cl);
            last_ref_end = __shfl_sync(FULL, fp_before, cl);
            have_last = true;
        }
        n_close += __popc(cb);
        n_over += __popc(ob);
        read_pos = __shfl_sync(FULL, rp_before + qa, 31);
        ref_pos = __shfl_sync(FULL, fp_before + (long long)ra, 31);
    }
    if (ml > 0) {  // the CIGAR ends inside a run
        const uint32_t k0 = uint32_t(read_pos - ml);
        const bool over = have_last && last_key == k0;
        n_over += over ? 1u : 0u;
        emit(lane == 0, 2u * n_close - n_over, k0, ref_pos - (long long)ml, uint32_t(read_pos), have_last, last_ref_end);
        ++n_close;
    }
    if (!out && lane == 0) counts[g] = 2u * n_close - n_over;
}


// a3 (count) of read r by one WARP (lanes over the CIGAR ops and over the contig's segments): same results as
// pair_count_body in pair_bodies.cuh.  Here, in a header, so that tests/emul can run it in lock step against the scalar body.
__device__ __forceinline__ void pair_count_warp_body(const DevStatic& S, const DevBatch& B, const DevWork& W, DevTotals* T, uint32_t r, uint32_t lane) {

    const uint32_t s0 = B.read_seg_begin[r], s1 = B.read_seg_begin[r + 1];
    bool bad = s0 > s1 || s1 > B.n_rsegs || !range_in_pool(B.read_seq_off[r], (uint64_t(B.read_seq_len[r]) + 1u) / 2u, B.seq4_bytes);
    for (uint32_t s = s0; s < s1 && s < B.n_rsegs; ++s) {
        if (lane == 0) W.rseg_read[s] = r;
        const uint64_t c0 = B.rseg_cigar_begin[s];
        const uint32_t n = B.rseg_cigar_len[s];
        const int64_t pos = B.rseg_pos[s];
        if (bad || B.rseg_contig[s] >= S.n_contigs || !range_in_pool(c0, n, B.n_cigar) || pos < 0 || pos > 0x7fffffffLL) {
            bad = true;
            if (lane == 0) {
                W.rseg_ref_len[s] = 0;
                W.rseg_n_id[s] = 0;
                W.rseg_read_len[s] = 0;
                W.rseg_pair_begin[s] = 0;
            }
            continue;
        }
        const uint32_t* c = B.cigar + c0;
        unsigned long long ref_len = 0;
        uint32_t n_id = 0, read_len = 0;
        for (uint32_t i = lane; i < n; i += 32u) {
            const uint32_t x = c[i];
            ref_len += op_ref_adv(x);
            read_len += op_read_adv(x);
            n_id += op_is_match(x & 0xfu) ? 0u : 1u;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            ref_len += __shfl_xor_sync(FULL, ref_len, d);
            read_len += __shfl_xor_sync(FULL, read_len, d);
            n_id += __shfl_xor_sync(FULL, n_id, d);
        }
        const int64_t start = pos, end = start + int64_t(ref_len);
        const uint32_t ctg = B.rseg_contig[s];
        uint32_t cnt = 0;
        const uint32_t g1 = S.contig_seg_begin[ctg + 1];
        for (uint32_t g0 = S.contig_seg_begin[ctg]; g0 < g1; g0 += 32u) {
            const uint32_t g = g0 + lane;
            const bool hit = g < g1 && end >= int64_t(S.seg_so_start[g]) && start < int64_t(S.seg_so_end[g]);
            cnt += __popc(__ballot_sync(FULL, hit));
        }
        if (lane == 0) {
            W.rseg_ref_len[s] = int64_t(ref_len);
            W.rseg_n_id[s] = n_id;
            W.rseg_read_len[s] = read_len;
            W.rseg_pair_begin[s] = cnt;
        }
    }
    if (bad && lane == 0) atomicOr(&T->overflow, OVF_INVALID);
}

}  // namespace ptl
