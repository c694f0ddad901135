// Launch wrappers of the CUDA kernels (kernels.cu) used by the host orchestration (context.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "device_types.hpp"

namespace ptl {

// CUDA events bracketing the stages of one launch_lift: [0] start, [1] pairs enumerated, [2] pairs lifted, [3] records emitted
struct StageEvents {
    static constexpr int N = 4;
    cudaEvent_t e[N];
};

template <class T>
void exclusive_scan_inplace(T* a, uint64_t n, void* tmp, size_t tmp_bytes, cudaStream_t st, uint64_t* launches,
                            const unsigned long long* n_dev = nullptr);
size_t scan_tmp_bytes(uint64_t n);

// counts != nullptr, out == nullptr : count table entries per segment; then (after a scan into S.seg_tab_begin) fill `out`.
void launch_table_build(const DevStatic& S, uint32_t* counts, TabEntry* out, cudaStream_t st);

// `arena`: the compact result arena (device_types.hpp: result_layout) of `arena_cap` bytes, header = the batch totals.
void launch_lift(const DevStatic& S, const DevBatch& B, const DevWork& W, char* arena, uint64_t arena_cap, DevTotals* T, uint32_t stage_mask,
                 void* scan_tmp, size_t scan_tmp_bytes, cudaStream_t st, uint64_t* launches, StageEvents* ev);

// Record assembly, bases (assemble.cuh).  Sizes: per record the 4-byte-rounded byte counts of its bases and qualities into
// seq_begin / qual_begin ([n_records+1], then exclusive scans), and the record's read index into rec_read.
struct AsmArgs;
void launch_assemble_sizes(uint32_t n_records, uint32_t n_reads, const uint32_t* read_rec_begin, const uint32_t* read_seq_len,
                           const uint64_t* read_qual_off, uint64_t qual_bytes, uint32_t* rec_read, uint64_t* seq_begin, uint64_t* qual_begin,
                           unsigned int* error, void* scan_tmp, size_t scan_tmp_bytes, cudaStream_t st, uint64_t* launches);
void launch_assemble_records(const AsmArgs& A, cudaStream_t st, uint64_t* launches);

// Record assembly, whole BAM records (assemble_bam.cuh): sizes (aux walk per read, SA entry length per record, record
// size -> exclusive scan into A.rec_begin), then the writer (one block per record).
struct BamAsmArgs;
void launch_bam_sizes(const BamAsmArgs& A, void* scan_tmp, size_t scan_tmp_bytes, cudaStream_t st, uint64_t* launches);
void launch_bam_write(const BamAsmArgs& A, cudaStream_t st, cudaStream_t st_meta, uint64_t* launches);

// BGZF framing at compression level 0 (bgzf_store.cuh): one thread block per BGZF block.
struct BgzfArgs;
void launch_bgzf_store(const BgzfArgs& A, cudaStream_t st, uint64_t* launches);
// Fused record assembly + framing (bgzf_store.cuh: bgzf_frame_block_body).
struct FrameArgs;
void launch_bam_frame(const FrameArgs& F, cudaStream_t st, uint64_t* launches);

}  // namespace ptl
