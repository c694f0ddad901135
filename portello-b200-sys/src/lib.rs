//! Raw FFI declarations of `libportello_b200.so`: the B200-native read-mapping transfer ("liftover") path of portello.
//!
//! GENERATED from `include/portello_b200.h` by `tools/gen_rust_sys.py`; do not edit by hand.  The header is the contract
//! and documents every item (with the portello source lines each entry point replaces); `tests/test_ffi_binding.py`
//! checks this file against it field for field.
//!
//! Safe wrappers (`Context`, `Batch`, `LiftResult`) belong in a `portello-b200` crate on top of this one; portello's
//! `read_alignment_scanner` would call `ptl_lift_submit` / `ptl_lift_wait` from its per-window workers.
#![allow(non_camel_case_types)]
#![allow(clippy::too_many_arguments)]
use std::os::raw::{c_char, c_int, c_void};

// ---------------------------------------------------------------- constants (status codes, stage masks, flags)
pub const PTL_ASM_RESIDENT_QUAL: u32 = 1;
pub const PTL_ASM_NO_DOWNLOAD: u32 = 2;
pub const PTL_BGZF_EOF: u32 = 4;
pub const PTL_OK: i32 = 0;
pub const PTL_ERR_INVALID_ARG: i32 = 1;
pub const PTL_ERR_CUDA: i32 = 2;
pub const PTL_ERR_NO_DEVICE: i32 = 3;
pub const PTL_ERR_STATE: i32 = 4;
pub const PTL_ERR_LIFT_PANIC: i32 = 5;
pub const PTL_ERR_INPUT: i32 = 6;
pub const PTL_REC_LIFTED: i32 = 1;
pub const PTL_REC_UNMAPPED: i32 = 0;
pub const PTL_PAIR_NONE: i32 = 0;
pub const PTL_PAIR_ERR_LENGTH: i32 = -1;
pub const PTL_PAIR_ERR_BOUNDS: i32 = -2;
pub const PTL_PAIR_ERR_CAPACITY: i32 = -3;
pub const PTL_STAGE_LEFT_SHIFT: u32 = 1;
pub const PTL_STAGE_LIFTOVER: u32 = 2;
pub const PTL_STAGE_SIMPLIFY: u32 = 4;
pub const PTL_STAGE_ALL: u32 = 7;
pub const PTL_WIN_NONE: i32 = 0;
pub const PTL_WIN_ALL: i32 = 1;
pub const PTL_WIN_REVERSE_PAIRS: i32 = 2;
pub const PTL_FETCH_ALL: i32 = -2;
pub const PTL_FETCH_UNMAPPED: i32 = -1;
pub const PTL_BAM_START_IN_REGION: u32 = 1;
pub const PTL_BAM_SKIP_SUPPLEMENTARY: u32 = 2;
pub const PTL_BAM_SKIP_UNMAPPED_SECONDARY: u32 = 4;
pub const PTL_BAM_ONLY_UNMAPPED: u32 = 8;
pub const PTL_BAM_KEEP_RAW: u32 = 16;

// ---------------------------------------------------------------- structs
#[repr(C)]
pub struct ptl_ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct ptl_prepared_contigs {
    _private: [u8; 0],
}

#[repr(C)]
pub struct ptl_packed_batch {
    _private: [u8; 0],
}

#[repr(C)]
pub struct ptl_bam_file {
    _private: [u8; 0],
}

#[repr(C)]
pub struct ptl_decoded_batch {
    _private: [u8; 0],
}

#[repr(C)]
pub struct ptl_fasta {
    _private: [u8; 0],
}

#[repr(C)]
pub struct ptl_contig_scan {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ptl_contig_segments {
    pub n_contigs: u32,
    pub contig_len: *const u64,
    pub contig_seg_begin: *const u32,
    pub rev_contig_seq: *const *const u8,
    pub n_segments: u32,
    pub seg_seq_order_start: *const u32,
    pub seg_seq_order_end: *const u32,
    pub seg_chrom_index: *const i32,
    pub seg_pos: *const i64,
    pub seg_is_fwd: *const u8,
    pub seg_mapq: *const u8,
    pub seg_cigar_begin: *const u64,
    pub cigar: *const u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ptl_batch {
    pub n_reads: u32,
    pub read_flag: *const u16,
    pub read_mapq: *const u8,
    pub read_bin: *const u16,
    pub read_seq_len: *const u32,
    pub read_seq_off: *const u64,
    pub read_seg_begin: *const u32,
    pub n_read_segments: u32,
    pub rseg_contig: *const u32,
    pub rseg_pos: *const i64,
    pub rseg_is_fwd: *const u8,
    pub rseg_cigar_begin: *const u64,
    pub rseg_cigar_len: *const u32,
    pub cigar: *const u32,
    pub n_cigar: u64,
    pub seq4: *const u8,
    pub seq4_bytes: u64,
    pub indel_win: *const u64,
    pub rseg_win_begin: *const u32,
    pub n_indel_win: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ptl_result {
    pub n_reads: u32,
    pub read_rec_begin: *const u32,
    pub n_records: u32,
    pub rec_status: *const i8,
    pub rec_read_segment: *const u32,
    pub rec_contig_segment: *const u32,
    pub rec_tid: *const i32,
    pub rec_pos: *const i64,
    pub rec_mapq: *const u8,
    pub rec_flag: *const u16,
    pub rec_bin: *const u16,
    pub rec_need_flip: *const u8,
    pub rec_cigar_begin: *const u64,
    pub cigar: *const u32,
    pub n_cigar: u64,
    pub n_pairs: u64,
    pub n_lifted: u64,
    pub n_errors: u64,
    pub first_error_read: i64,
    pub first_error_status: i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ptl_contig_records {
    pub n_records: u32,
    pub contig_id: *const u32,
    pub flag: *const u16,
    pub tid: *const i32,
    pub pos: *const i64,
    pub mapq: *const u8,
    pub cigar_begin: *const u64,
    pub cigar: *const u32,
    pub sa_tag: *const *const c_char,
    pub seq: *const *const u8,
    pub n_contigs: u32,
    pub contig_len: *const u64,
    pub contig_names: *const *const c_char,
    pub n_ref_chrom: u32,
    pub ref_chrom_names: *const *const c_char,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ptl_read_quals {
    pub qual: *const u8,
    pub read_qual_off: *const u64,
    pub qual_bytes: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ptl_record_bases {
    pub n_records: u32,
    pub rec_seq_begin: *const u64,
    pub seq4: *const u8,
    pub rec_qual_begin: *const u64,
    pub qual: *const u8,
    pub kernel_ms: f32,
    pub bytes_read: u64,
    pub bytes_written: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ptl_read_extras {
    pub name_off: *const u64,
    pub names: *const u8,
    pub aux_off: *const u64,
    pub aux: *const u8,
    pub mate_tid: *const i32,
    pub mate_pos: *const i32,
    pub tlen: *const i32,
    pub quals: ptl_read_quals,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ptl_bam_records {
    pub n_records: u32,
    pub rec_begin: *const u64,
    pub bytes: *const u8,
    pub kernel_ms: f32,
    pub bytes_read: u64,
    pub bytes_written: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ptl_bgzf_stream {
    pub n_bytes: u64,
    pub bytes: *const u8,
    pub n_blocks: u64,
    pub kernel_ms: f32,
    pub bytes_read: u64,
    pub bytes_written: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ptl_read_records {
    pub n_reads: u32,
    pub tid: *const i32,
    pub pos: *const i64,
    pub flag: *const u16,
    pub mapq: *const u8,
    pub bin: *const u16,
    pub seq_len: *const u32,
    pub seq_off: *const u64,
    pub seq4: *const u8,
    pub seq4_bytes: u64,
    pub cigar_begin: *const u64,
    pub cigar: *const u32,
    pub sa_tag: *const *const c_char,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ptl_split_segments {
    pub seq_order_start: *mut u32,
    pub seq_order_end: *mut u32,
    pub contig: *mut u32,
    pub pos: *mut i64,
    pub is_fwd: *mut u8,
    pub mapq: *mut u8,
    pub from_primary: *mut u8,
    pub cigar_begin: *mut u32,
    pub cigar: *mut u32,
}

// ---------------------------------------------------------------- functions
#[link(name = "portello_b200")]
extern "C" {
    pub fn ptl_create(device: c_int, n_slots: c_int, out: *mut *mut ptl_ctx) -> c_int;
    pub fn ptl_destroy(ctx: *mut ptl_ctx);
    pub fn ptl_last_error(ctx: *const ptl_ctx) -> *const c_char;
    pub fn ptl_version() -> *const c_char;
    pub fn ptl_set_reference(ctx: *mut ptl_ctx, n_chrom: u32, chrom_len: *const u64, chrom_seq: *const *const u8) -> c_int;
    pub fn ptl_set_contig_segments(ctx: *mut ptl_ctx, segs: *const ptl_contig_segments) -> c_int;
    pub fn ptl_set_raw_contig_segments(ctx: *mut ptl_ctx, raw: *const ptl_contig_segments) -> c_int;
    pub fn ptl_set_contig_records(ctx: *mut ptl_ctx, recs: *const ptl_contig_records) -> c_int;
    pub fn ptl_prepare_contig_records(recs: *const ptl_contig_records, out: *mut *mut ptl_prepared_contigs) -> c_int;
    pub fn ptl_prepare_raw_contig_segments(raw: *const ptl_contig_segments, out: *mut *mut ptl_prepared_contigs) -> c_int;
    pub fn ptl_prepared_contigs_view(p: *const ptl_prepared_contigs, out: *mut ptl_contig_segments);
    pub fn ptl_prepared_contigs_free(p: *mut ptl_prepared_contigs);
    pub fn ptl_prepare_last_error() -> *const c_char;
    pub fn ptl_get_contig_segments(ctx: *const ptl_ctx, out: *mut ptl_contig_segments) -> c_int;
    pub fn ptl_get_segment_table(ctx: *mut ptl_ctx, segment: u32, cap: u32, keys: *mut u32, vals: *mut i32, n: *mut u32) -> c_int;
    pub fn ptl_lift_submit(ctx: *mut ptl_ctx, slot: c_int, batch: *const ptl_batch) -> c_int;
    pub fn ptl_lift_submit_ex(ctx: *mut ptl_ctx, slot: c_int, batch: *const ptl_batch, stage_mask: u32) -> c_int;
    pub fn ptl_lift_wait(ctx: *mut ptl_ctx, slot: c_int, out: *mut ptl_result) -> c_int;
    pub fn ptl_lift_upload(ctx: *mut ptl_ctx, slot: c_int, batch: *const ptl_batch) -> c_int;
    pub fn ptl_lift_run(ctx: *mut ptl_ctx, slot: c_int, stage_mask: u32) -> c_int;
    pub fn ptl_lift_download(ctx: *mut ptl_ctx, slot: c_int, out: *mut ptl_result) -> c_int;
    pub fn ptl_assemble_bases(ctx: *mut ptl_ctx, slot: c_int, quals: *const ptl_read_quals, flags: u32, out: *mut ptl_record_bases) -> c_int;
    pub fn ptl_set_names(ctx: *mut ptl_ctx, n_contigs: u32, contig_names: *const *const c_char, n_chrom: u32, chrom_names: *const *const c_char) -> c_int;
    pub fn ptl_assemble_records(ctx: *mut ptl_ctx, slot: c_int, extras: *const ptl_read_extras, flags: u32, out: *mut ptl_bam_records) -> c_int;
    pub fn ptl_bam_header(sam_text: *const c_char, n_ref: u32, ref_names: *const *const c_char, ref_len: *const u64, out: *mut u8, cap: u64) -> i64;
    pub fn ptl_bgzf_bound(n: u64) -> u64;
    pub fn ptl_bgzf_compress(in_: *const u8, n: u64, level: c_int, n_threads: c_int, append_eof: c_int, out: *mut u8, cap: u64) -> i64;
    pub fn ptl_bgzf_store_records(ctx: *mut ptl_ctx, slot: c_int, prefix: *const u8, prefix_bytes: u64, flags: u32, out: *mut ptl_bgzf_stream) -> c_int;
    pub fn ptl_frame_records(ctx: *mut ptl_ctx, slot: c_int, extras: *const ptl_read_extras, prefix: *const u8, prefix_bytes: u64, flags: u32, out: *mut ptl_bgzf_stream) -> c_int;
    pub fn ptl_slot_stream(ctx: *mut ptl_ctx, slot: c_int) -> *mut c_void;
    pub fn ptl_slot_kernel_times(ctx: *mut ptl_ctx, slot: c_int, cap: c_int, names: *mut *const c_char, ms: *mut f32) -> c_int;
    pub fn ptl_launch_count(ctx: *const ptl_ctx) -> u64;
    pub fn ptl_slot_counters(ctx: *mut ptl_ctx, slot: c_int, out: *mut u64) -> c_int;
    pub fn ptl_set_seq_zero_copy(ctx: *mut ptl_ctx, enable: c_int) -> c_int;
    pub fn ptl_set_long_pair_ops(ctx: *mut ptl_ctx, n_ops: u32) -> c_int;
    pub fn ptl_host_alloc(bytes: usize) -> *mut c_void;
    pub fn ptl_host_free(p: *mut c_void);
    pub fn ptl_pack_batch(recs: *const ptl_read_records, first: u32, count: u32, n_contigs: u32, contig_names: *const *const c_char, pinned: c_int, out: *mut *mut ptl_packed_batch) -> c_int;
    pub fn ptl_pack_batch_ex(recs: *const ptl_read_records, first: u32, count: u32, n_contigs: u32, contig_names: *const *const c_char, pinned: c_int, window_mode: c_int, segs: *const ptl_contig_segments, out: *mut *mut ptl_packed_batch) -> c_int;
    pub fn ptl_pack_batch_into(reuse: *mut ptl_packed_batch, recs: *const ptl_read_records, first: u32, count: u32, n_contigs: u32, contig_names: *const *const c_char, window_mode: c_int, segs: *const ptl_contig_segments) -> c_int;
    pub fn ptl_packed_batch_view(p: *const ptl_packed_batch, out: *mut ptl_batch);
    pub fn ptl_packed_batch_record_index(p: *const ptl_packed_batch) -> *const u32;
    pub fn ptl_packed_batch_free(p: *mut ptl_packed_batch);
    pub fn ptl_pack_split_segments(n_contig_names: u32, contig_names: *const *const c_char, tid: i32, pos: i64, flag: u16, mapq: u8, cigar: *const u32, n_cigar: u32, sa_tag: *const c_char, cap_segments: u32, cap_cigar: u32, out: *mut ptl_split_segments, n_segments: *mut u32, n_cigar_out: *mut u32) -> c_int;
    pub fn ptl_pack_last_error() -> *const c_char;
    pub fn ptl_format_sa_tags(res: *const ptl_result, n_chrom: u32, chrom_names: *const *const c_char, buf: *mut c_char, cap: u64, sa_begin: *mut u64, need: *mut u64) -> c_int;
    pub fn ptl_region_segment_count(size: u64, segment_size: u64) -> u32;
    pub fn ptl_region_segments(size: u64, segment_size: u64, begin: *mut u64, end: *mut u64);
    pub fn ptl_shard_units(n_units: u32, weight: *const u64, n_ranks: u32, owner: *mut u32);
    pub fn ptl_bam_open(path: *const c_char, out: *mut *mut ptl_bam_file) -> c_int;
    pub fn ptl_bam_close(f: *mut ptl_bam_file);
    pub fn ptl_bam_last_error() -> *const c_char;
    pub fn ptl_bam_n_ref(f: *const ptl_bam_file) -> u32;
    pub fn ptl_bam_ref_name(f: *const ptl_bam_file, i: u32) -> *const c_char;
    pub fn ptl_bam_ref_len(f: *const ptl_bam_file, i: u32) -> u64;
    pub fn ptl_bam_header_text(f: *const ptl_bam_file) -> *const c_char;
    pub fn ptl_bam_has_index(f: *const ptl_bam_file) -> c_int;
    pub fn ptl_bam_has_eof_marker(f: *const ptl_bam_file) -> c_int;
    pub fn ptl_bam_fetch(f: *const ptl_bam_file, tid: i32, begin: i64, end: i64, filter: u32, out: *mut *mut ptl_decoded_batch) -> c_int;
    pub fn ptl_decoded_view(d: *const ptl_decoded_batch, recs: *mut ptl_read_records, extras: *mut ptl_read_extras);
    pub fn ptl_decoded_raw(d: *const ptl_decoded_batch, rec_off: *mut *const u64, n_bytes: *mut u64) -> *const u8;
    pub fn ptl_decoded_free(d: *mut ptl_decoded_batch);
    pub fn ptl_bam_index_build(bam_path: *const c_char, bai_path: *const c_char) -> c_int;
    pub fn ptl_bam_index_build_csi(bam_path: *const c_char, csi_path: *const c_char, min_shift: c_int, depth: c_int) -> c_int;
    pub fn ptl_fasta_load(path: *const c_char, n_threads: c_int, out: *mut *mut ptl_fasta) -> c_int;
    pub fn ptl_fasta_n(f: *const ptl_fasta) -> u32;
    pub fn ptl_fasta_name(f: *const ptl_fasta, i: u32) -> *const c_char;
    pub fn ptl_fasta_seq(f: *const ptl_fasta, i: u32, len: *mut u64) -> *const u8;
    pub fn ptl_fasta_free(f: *mut ptl_fasta);
    pub fn ptl_scan_contig_bam(contig_bam: *const ptl_bam_file, n_contigs: u32, contig_names: *const *const c_char, contig_len: *const u64, n_threads: c_int, out: *mut *mut ptl_contig_scan) -> c_int;
    pub fn ptl_contig_scan_view(s: *const ptl_contig_scan, out: *mut ptl_contig_records);
    pub fn ptl_contig_scan_free(s: *mut ptl_contig_scan);
    pub fn ptl_reg2bin(begin: i64, end: i64) -> u16;
}
