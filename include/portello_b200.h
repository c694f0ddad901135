/*
 * portello_b200.h — C-ABI of the B200-native read-mapping transfer ("liftover") path.
 *
 * The reference (PacificBiosciences/portello, Rust) has no FFI/plugin seam; the path sits behind ordinary Rust
 * calls.  This header is the seam a maintainer would bind from Rust (`extern "C"` in a `portello-b200-sys` crate,
 * see INTEGRATION.md): plain pointers and sizes, no C++/torch types, integer status codes, no unwinding.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the reference repo root).
 *
 * The same ABI is exported twice:
 *   - libportello_b200.so : `ptl_*`        — the product, hand-written sm_100a CUDA kernels, NO CPU fallback
 *   - oracle/_build/libptl_oracle.so : `ptl_oracle_*` — CPU restatement of the reference (TEST INFRASTRUCTURE ONLY)
 * so one harness drives both through identical structs.
 *
 * Conventions
 *   - CIGAR ops are BAM-encoded u32: (len << 4) | op, op codes M=0 I=1 D=2 N=3 S=4 H=5 P=6 '='=7 X=8
 *     (= rust_htslib::bam::record::Cigar variants Match/Ins/Del/RefSkip/SoftClip/HardClip/Pad/Equal/Diff).
 *   - read bases are BAM-native 4-bit packed (high nibble first), decode table "=ACMGRSVTWYHKDBN"
 *     (what `Record::seq().as_bytes()` yields, src/read_alignment_scanner.rs:127,170,238).
 *   - reference / rev_contig bases are ASCII bytes exactly as the reference holds them
 *     (upper-cased FASTA, lib/rust-vc-utils/src/genome_ref.rs:53; rev_contig_seq,
 *     src/contig_alignment_scanner/mod.rs:113-125).
 *   - positions are 0-based; CSR arrays have n+1 entries.
 *   - all input pointers are borrowed for the duration of the call only, unless stated otherwise.
 */
#ifndef PORTELLO_B200_H
#define PORTELLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- status codes */
enum {
    PTL_OK = 0,
    PTL_ERR_INVALID_ARG = 1,
    PTL_ERR_CUDA = 2,          /* CUDA runtime failure (message in ptl_last_error) */
    PTL_ERR_NO_DEVICE = 3,     /* no usable sm_100 device: the product path has no CPU fallback */
    PTL_ERR_STATE = 4,         /* call sequence error (e.g. wait without submit) */
    PTL_ERR_LIFT_PANIC = 5,    /* >=1 pair hit a condition on which the reference panics (see rec status / n_errors) */
    PTL_ERR_INPUT = 6          /* malformed input on which the reference panics during table preparation */
};

/* per-pair / per-record status (ptl_result.rec_status, stage results) */
enum {
    PTL_REC_LIFTED = 1,        /* Some(record) from get_liftover_alignment_for_read_and_contig_segment */
    PTL_REC_UNMAPPED = 0,      /* unmapped fallback copy, finish_remapped_alignment_set :317-335 */
    PTL_PAIR_NONE = 0,         /* liftover returned None (stage API) */
    PTL_PAIR_ERR_LENGTH = -1,  /* lifted CIGAR read length != seq_len  (panic, src/read_alignment_scanner.rs:204-229) */
    PTL_PAIR_ERR_BOUNDS = -2,  /* sequence index out of bounds (Rust slice-index panic in a5/a9 base compares) */
    PTL_PAIR_ERR_CAPACITY = -3 /* internal op-slot bound exceeded (library bug guard; never expected) */
};

/* stage mask for ptl_lift_submit_ex (testing individual kernels against the reference's unit vectors) */
enum {
    PTL_STAGE_LEFT_SHIFT = 1,  /* a5: left_shift_indels on reverse-strand contig segments */
    PTL_STAGE_LIFTOVER = 2,    /* a6: liftover_read_alignment */
    PTL_STAGE_SIMPLIFY = 4,    /* a9: simplify_alignment_indels */
    PTL_STAGE_ALL = 7
};

typedef struct ptl_ctx ptl_ctx;

/* ---------------------------------------------------------------- inputs */

/* Contig->reference mapping segments AFTER trim/join, i.e. the value returned by `scan_contig_bam`
 * (src/contig_alignment_scanner/mod.rs:452-458): per contig an ordered (sequencing-order) list of
 * `ContigMappingSegmentInfo` (mod.rs:25-32) + optional `rev_contig_seq` (mod.rs:38-47).
 * Contig ids index the read->assembly BAM header (AllContigMappingInfo, mod.rs:72-76). */
typedef struct ptl_contig_segments {
    uint32_t n_contigs;
    const uint64_t* contig_len;            /* [n_contigs] length from the read->asm header (read_alignment_scanner.rs:164) */
    const uint32_t* contig_seg_begin;      /* [n_contigs+1] CSR into the segment arrays */
    const uint8_t* const* rev_contig_seq;  /* [n_contigs] ASCII, contig_len bytes each, NULL where the reference holds None */
    uint32_t n_segments;
    const uint32_t* seg_seq_order_start;   /* [n_segments] SeqOrderSplitReadSegment.seq_order_read_start (split_read.rs:15-32) */
    const uint32_t* seg_seq_order_end;     /* [n_segments] .seq_order_read_end */
    const int32_t* seg_chrom_index;        /* [n_segments] .chrom_index (reference chromosome) */
    const int64_t* seg_pos;                /* [n_segments] .pos */
    const uint8_t* seg_is_fwd;             /* [n_segments] .is_fwd_strand */
    const uint8_t* seg_mapq;               /* [n_segments] .mapq */
    const uint64_t* seg_cigar_begin;       /* [n_segments+1] CSR into cigar */
    const uint32_t* cigar;                 /* contig->ref CIGAR ops */
} ptl_contig_segments;

/* One batch of primary read->contig records with their split segments already parsed
 * (get_seq_order_read_split_segments, lib/rust-vc-utils/src/bam_utils/split_read.rs:56-155; the packer
 * ptl_pack_* below does that parsing).  Mirrors the per-read loop body src/read_alignment_scanner.rs:419-472. */
typedef struct ptl_batch {
    uint32_t n_reads;
    const uint16_t* read_flag;        /* [n_reads] BAM flag of the primary record */
    const uint8_t* read_mapq;         /* [n_reads] MAPQ of the primary record (-> ZM:C, :251,267-269) */
    const uint16_t* read_bin;         /* [n_reads] BAM bin of the primary record (kept by the unmapped fallback, :322-334) */
    const uint32_t* read_seq_len;     /* [n_reads] record.seq_len() */
    const uint64_t* read_seq_off;     /* [n_reads] byte offset of the read's packed bases in seq4 */
    const uint32_t* read_seg_begin;   /* [n_reads+1] CSR into the read-segment arrays (sequencing order, split_read.rs:140) */
    uint32_t n_read_segments;
    const uint32_t* rseg_contig;      /* [n_read_segments] .chrom_index (assembly contig id; may differ from the primary's) */
    const int64_t* rseg_pos;          /* [n_read_segments] .pos on the contig */
    const uint8_t* rseg_is_fwd;       /* [n_read_segments] .is_fwd_strand */
    const uint64_t* rseg_cigar_begin; /* [n_read_segments] first op of the segment CIGAR in `cigar` */
    const uint32_t* rseg_cigar_len;   /* [n_read_segments] number of ops */
    const uint32_t* cigar;            /* CIGAR pool */
    uint64_t n_cigar;                 /* ops in the pool */
    const uint8_t* seq4;              /* packed bases pool; for ptl_lift_submit pinned host memory gives async copies */
    uint64_t seq4_bytes;
    /* Optional "indel windows" (NULL / 0 = absent).  left_shift_indels (a5) compares, for every I/D cluster of a read
     * segment that sits on a reverse-strand contig segment, a few read bases next to the cluster
     * (lib/rust-vc-utils/src/indel_breakend_homology.rs:35-49); which ones is data dependent, but they always START at
     * the cluster.  A window holds the first 16 bases the walk of one cluster can touch, 4 bits each (bits [4q,4q+4) =
     * the stored BAM nibble compared at walk step q), so the kernels need the packed bases themselves (seq4, over PCIe
     * in zero-copy mode) only for walks longer than 16 bases and for simplify_alignment_indels.  Windows of read
     * segment s are indel_win[rseg_win_begin[s] .. rseg_win_begin[s+1]) in the cluster order of the REVERSED segment
     * CIGAR (:165-167); a segment may have none (the kernels then read seq4).  Built by ptl_pack_batch_ex. */
    const uint64_t* indel_win;
    const uint32_t* rseg_win_begin;   /* [n_read_segments+1] */
    uint64_t n_indel_win;
} ptl_batch;

/* ---------------------------------------------------------------- outputs */

/* Result of one batch, in submission order.  Records of a read are contiguous and ordered
 * (read-segment sequencing order) x (contig-segment index) exactly as src/read_alignment_scanner.rs:430-471 pushes
 * them; a read with no lifted pair yields one PTL_REC_UNMAPPED record (:317-335).
 * Buffers are owned by the ctx slot and stay valid until the next submit on that slot. */
typedef struct ptl_result {
    uint32_t n_reads;
    const uint32_t* read_rec_begin;       /* [n_reads+1] CSR into the record arrays */
    uint32_t n_records;
    const int8_t* rec_status;             /* PTL_REC_LIFTED / PTL_REC_UNMAPPED */
    const uint32_t* rec_read_segment;     /* batch-global read-segment index (-> contig name for PS:Z, :255-265) */
    const uint32_t* rec_contig_segment;   /* contig-segment index within its contig (the PS "split" index, :260) */
    const int32_t* rec_tid;               /* reference chromosome index, -1 if unmapped */
    const int64_t* rec_pos;               /* 0-based, -1 if unmapped */
    const uint8_t* rec_mapq;              /* contig segment MAPQ (:250-252); 255 if unmapped (:326) */
    const uint16_t* rec_flag;             /* final BAM flag incl. reverse flip, supplementary, unmapped */
    const uint16_t* rec_bin;              /* bam_reg2bin(pos,end) (:278-279); original bin if unmapped */
    const uint8_t* rec_need_flip;         /* 1: host must revcomp seq + reverse qual (:274-276, :330-332) */
    const uint64_t* rec_cigar_begin;      /* [n_records+1] CSR into cigar */
    const uint32_t* cigar;                /* output CIGAR pool, dense, record order */
    uint64_t n_cigar;
    /* counters (work done; used for the roofline arithmetic) */
    uint64_t n_pairs;                     /* attempted (read segment x contig segment) pairs = a4 calls */
    uint64_t n_lifted;                    /* pairs that produced a record */
    uint64_t n_errors;                    /* pairs with a negative status (reference would have panicked) */
    int64_t first_error_read;             /* batch read index of the first such pair, -1 if none */
    int32_t first_error_status;
} ptl_result;

/* ---------------------------------------------------------------- lifecycle */

/* One ctx per GPU (replaces the per-thread reader set of src/worker_thread_data.rs:8-30 as the unit of concurrency).
 * `n_slots` >= 1 double/triple-buffered batch slots, each with its own CUDA stream. */
int ptl_create(int device, int n_slots, ptl_ctx** out);
void ptl_destroy(ptl_ctx* ctx);
/* Message for the last non-OK status on this ctx (owned by ctx; "" if none). Thread-compatible, not thread-safe. */
const char* ptl_last_error(const ptl_ctx* ctx);
/* Library/ABI version string, e.g. "portello_b200 0.1 (sm_100a)". */
const char* ptl_version(void);

/* Reference genome bytes = `reference: &[Vec<u8>]` built by get_chrom_array (src/main.rs:24-62). Copied to device. */
int ptl_set_reference(ptl_ctx* ctx, uint32_t n_chrom, const uint64_t* chrom_len, const uint8_t* const* chrom_seq);

/* Install `AllContigMappingInfo` (post trim/join) and build the device segment tables, the flat form of
 * ReadToRefTreeMap (lib/rust-vc-utils/src/bam_utils/read_to_ref_map.rs:59-137), with a CUDA kernel. */
int ptl_set_contig_segments(ptl_ctx* ctx, const ptl_contig_segments* segs);

/* Same, but from RAW (pre-trim) segments as assembled by add_primary_read + supplementary fill-in
 * (mod.rs:91-183,360-439): runs clip_repeated_contig_matches (contig_repeated_match_trimmer.rs:214-303) and
 * join_colinear_contig_segments (contig_colinear_segment_joiner.rs:124-186) in-library (host C++), then installs.
 * The processed segments can be read back with ptl_get_contig_segments. */
int ptl_set_raw_contig_segments(ptl_ctx* ctx, const ptl_contig_segments* raw);
/* The contig->reference BAM records as scan_contig_bam's record loop sees them (mod.rs:186-240), BAM decode excluded. */
typedef struct ptl_contig_records {
    uint32_t n_records;
    const uint32_t* contig_id;         /* [n_records] assembly contig index of the record's qname (mod.rs:219-220) */
    const uint16_t* flag;              /* BAM flag (unmapped/secondary skipped :208, supplementary :222, reverse) */
    const int32_t* tid;                /* reference chromosome index */
    const int64_t* pos;
    const uint8_t* mapq;
    const uint64_t* cigar_begin;       /* [n_records+1] */
    const uint32_t* cigar;
    const char* const* sa_tag;         /* [n_records] SA:Z value or NULL */
    const uint8_t* const* seq;         /* [n_records] ASCII decode of the stored bases of a primary record: contig_len[contig_id]
                                        * bytes (the whole contig; ptl_scan_contig_bam checks that), else NULL */
    uint32_t n_contigs;
    const uint64_t* contig_len;        /* [n_contigs] */
    const char* const* contig_names;   /* [n_contigs] (error messages only) */
    uint32_t n_ref_chrom;
    const char* const* ref_chrom_names; /* [n_ref_chrom] resolves SA rname -> chrom index */
} ptl_contig_records;
/* = scan_contig_bam minus BAM I/O (src/contig_alignment_scanner/mod.rs:290-459): add_primary_read (:91-133),
 * supplementary exact-CIGAR fill-in keyed by SplitReadKey (:49-56,135-183,360-439), rev_contig_seq (:113-125),
 * then trim + join, then install (host C++ for the O(#contig records) part, CUDA for the tables). */
int ptl_set_contig_records(ptl_ctx* ctx, const ptl_contig_records* recs);
/* Host-only form of the two calls above (no device needed): prepare, inspect, free.  Used by hosts that want the
 * post trim/join segments (e.g. for PS:Z split indices) before a GPU is involved, and by the CPU test-suite. */
typedef struct ptl_prepared_contigs ptl_prepared_contigs;
int ptl_prepare_contig_records(const ptl_contig_records* recs, ptl_prepared_contigs** out);
int ptl_prepare_raw_contig_segments(const ptl_contig_segments* raw, ptl_prepared_contigs** out);
void ptl_prepared_contigs_view(const ptl_prepared_contigs* p, ptl_contig_segments* out);
void ptl_prepared_contigs_free(ptl_prepared_contigs* p);
const char* ptl_prepare_last_error(void);
/* Borrow the installed (post trim/join) segments; pointers owned by ctx, valid until the next set call. */
int ptl_get_contig_segments(const ptl_ctx* ctx, ptl_contig_segments* out);
/* Copy the device-built table of one global segment index back to the host (testing a7):
 * keys[i] = contig read_pos starting a block, vals[i] = ref pos or -1 for None. Returns count via *n; pass cap. */
int ptl_get_segment_table(ptl_ctx* ctx, uint32_t segment, uint32_t cap, uint32_t* keys, int32_t* vals, uint32_t* n);

/* ---------------------------------------------------------------- the hot path */

/* Lift one batch: = the loop body of scan_chromosome_segment (src/read_alignment_scanner.rs:419-487) for n_reads
 * records: pair enumeration (:80-103), per-pair get_liftover_alignment_for_read_and_contig_segment (:136-288:
 * left_shift_indels, liftover_read_alignment, length check, simplify_alignment_indels, field updates) and
 * finish_remapped_alignment_set (:310-366: primary selection / unmapped fallback; SA text is ptl_format_sa_tags).
 * Asynchronous on the slot's stream (H2D copies, kernels, D2H); returns after enqueue.
 * Distinct slots may be driven from distinct host threads. */
int ptl_lift_submit(ptl_ctx* ctx, int slot, const ptl_batch* batch);
/* As above with a stage mask (PTL_STAGE_*), used to pin single kernels to the reference's unit vectors.
 * Without PTL_STAGE_LIFTOVER the pair's (pos,CIGAR) after the enabled stage is returned in contig coordinates
 * with tid = -2. */
int ptl_lift_submit_ex(ptl_ctx* ctx, int slot, const ptl_batch* batch, uint32_t stage_mask);
/* Block until the slot's batch is done and expose its result. Returns PTL_ERR_LIFT_PANIC if n_errors > 0
 * (the reference aborts the whole run in that case; here the remaining records are still valid). */
int ptl_lift_wait(ptl_ctx* ctx, int slot, ptl_result* out);

/* Split phases of submit/wait for device-resident measurement (bench.py `value`): */
int ptl_lift_upload(ptl_ctx* ctx, int slot, const ptl_batch* batch);   /* H2D only (async) */
int ptl_lift_run(ptl_ctx* ctx, int slot, uint32_t stage_mask);         /* kernels only, on the resident batch (async) */
int ptl_lift_download(ptl_ctx* ctx, int slot, ptl_result* out);        /* D2H + sync */
/* ---------------------------------------------------------------- record assembly, bases
 *
 * The read bases and base qualities of every OUTPUT record of the slot's last batch, oriented the way the reference
 * leaves them: a record with rec_need_flip = 1 went through reverse_alignment_seq_and_qual
 * (src/read_alignment_scanner.rs:125-133, called at :274-276 and :330-332): bases decoded, reverse-complemented with
 * rev_comp_in_place (lib/rust-vc-utils/src/seq_util.rs:29-40: A<->T, C<->G, N->N, anything else incl. '=' and IUPAC
 * codes -> N) and re-encoded by Record::set; qualities reversed.  Other records carry the read's bytes unchanged (every
 * record of a read is a clone_record of the same input, :105-117).  BAM layout: 4-bit bases, high nibble first, the unused
 * low nibble of an odd-length sequence zero.
 * Must follow ptl_lift_wait / ptl_lift_download on the same slot (the batch's packed bases are still resident). */
typedef struct {
    const uint8_t* qual;            /* pooled qualities: read r owns qual[read_qual_off[r] .. + read_seq_len[r]) */
    const uint64_t* read_qual_off;  /* [n_reads] */
    uint64_t qual_bytes;
} ptl_read_quals;
typedef struct {
    uint32_t n_records;
    const uint64_t* rec_seq_begin;   /* [n_records+1] byte offsets into seq4; every record starts 16-byte aligned, zero padded */
    const uint8_t* seq4;
    const uint64_t* rec_qual_begin;  /* [n_records+1] byte offsets into qual; every record starts 16-byte aligned, zero padded */
    const uint8_t* qual;
    float kernel_ms;                 /* device time of the assembly kernel (CUDA events on the slot stream) */
    uint64_t bytes_read;
    uint64_t bytes_written;  /* algorithmic bytes of that kernel: bases + qualities in, bases + qualities out */
} ptl_record_bases;
/* flags: PTL_ASM_RESIDENT_QUAL = the qualities uploaded by the previous call on this slot are reused (no H2D);
 *        PTL_ASM_NO_DOWNLOAD   = results stay on the device (out->seq4 / out->qual are NULL): kernel timing only. */
#define PTL_ASM_RESIDENT_QUAL 1u
#define PTL_ASM_NO_DOWNLOAD 2u
int ptl_assemble_bases(ptl_ctx* ctx, int slot, const ptl_read_quals* quals, uint32_t flags, ptl_record_bases* out);

/* ---------------------------------------------------------------- record assembly, whole BAM records
 *
 * Every OUTPUT record of the slot's last batch as the bytes bam_write1 would emit for it (SAM spec 4.2: block_size,
 * refID, pos, l_read_name, mapq, bin, n_cigar_op, flag, l_seq, next_refID, next_pos, tlen, read_name, cigar, seq, qual,
 * aux), i.e. the record get_liftover_alignment_for_read_and_contig_segment + finish_remapped_alignment_set leave behind:
 *   clone_record (src/read_alignment_scanner.rs:105-117): the input record minus the FIRST NM, SA, PS and ZM tag;
 *   set_tid / set_mapq / set_pos / set_cigar / set_bin / flags (:245-282); mate fields and every other tag untouched;
 *   push_aux PS:Z:{contig}_split{index}{+|-} and ZM:C:{MAPQ of the input record} (:255-269);
 *   reverse_alignment_seq_and_qual when the record was flipped (:125-133, as ptl_assemble_bases);
 *   push_aux SA:Z = "{chrom},{pos+1},{+|-},{CIGAR},{mapq},0;" of every OTHER record of the read, in order (:292-301,349-363);
 *   the unmapped fallback (:317-335): no CIGAR, tid/pos -1, MAPQ 255, original bin, no tag appended.
 * Aux order: surviving input tags, PS, ZM, SA.  A lifted CIGAR with more than 65535 ops is written the way htslib's
 * bam_write1 writes it: n_cigar_op = 2, the CIGAR field holds <l_seq>S<ref_len>N, and the real ops follow every other
 * tag as CG:B,I (SAM spec 4.2.2); an error only if ref_len >= 2^28 (bam_write1 refuses that record too).
 * Needs ptl_set_names once (contig names of the read->assembly header for PS, reference names for SA). */
int ptl_set_names(ptl_ctx* ctx, uint32_t n_contigs, const char* const* contig_names, uint32_t n_chrom, const char* const* chrom_names);
typedef struct {
    const uint64_t* name_off;       /* [n_reads+1] into names: qname bytes of read r WITHOUT the trailing NUL */
    const uint8_t* names;
    const uint64_t* aux_off;        /* [n_reads+1] into aux: the raw BAM aux block of read r */
    const uint8_t* aux;
    const int32_t* mate_tid;        /* [n_reads] next_refID */
    const int32_t* mate_pos;        /* [n_reads] next_pos */
    const int32_t* tlen;            /* [n_reads] */
    ptl_read_quals quals;           /* as ptl_assemble_bases */
} ptl_read_extras;
typedef struct {
    uint32_t n_records;
    const uint64_t* rec_begin;       /* [n_records+1] byte offsets into bytes; record k = block_size (u32) + block_size bytes */
    const uint8_t* bytes;
    float kernel_ms;                 /* device time of the record-writing kernel (CUDA events on the slot stream) */
    uint64_t bytes_read;
    uint64_t bytes_written;  /* algorithmic bytes of that kernel: every input byte of a record once, every output byte once */
} ptl_bam_records;
/* flags: PTL_ASM_RESIDENT_QUAL = everything uploaded by the previous call on this slot is reused (no H2D; extras may be NULL);
 *        PTL_ASM_NO_DOWNLOAD   = results stay on the device (out->bytes is NULL): kernel timing only. */
int ptl_assemble_records(ptl_ctx* ctx, int slot, const ptl_read_extras* extras, uint32_t flags, ptl_bam_records* out);

/* ---------------------------------------------------------------- BAM container (output half; host C++ + zlib)
 *
 * What bam::Writer does around the records (src/read_alignment_scanner.rs:537-559), written from the SAM specification
 * without htslib: the header block (4.2: magic, l_text, text, n_ref, {l_name, name, l_ref}) and BGZF framing (4.1:
 * independent gzip members of <= 0xff00 payload bytes with the BC size subfield, CRC32, ISIZE; the 28-byte EOF marker).
 * Feed it the header bytes and then the bytes of ptl_assemble_records; blocks are compressed on n_threads host threads
 * (the reference's --threads).  Both return the number of bytes written, or a negative PTL_ERR_* code. */
int64_t ptl_bam_header(const char* sam_text, uint32_t n_ref, const char* const* ref_names, const uint64_t* ref_len, uint8_t* out, uint64_t cap);
uint64_t ptl_bgzf_bound(uint64_t n);   /* output capacity that always suffices for n input bytes (+ EOF marker) */
int64_t ptl_bgzf_compress(const uint8_t* in, uint64_t n, int level, int n_threads, int append_eof, uint8_t* out, uint64_t cap);

/* ---------------------------------------------------------------- BGZF framing on the device, compression level 0
 *
 * The reference's documented pipe mode (`--remapped-read-output -`, "to optimize piping into samtools sort") writes
 * UNCOMPRESSED BAM: bam::Writer::from_stdout + CompressionLevel::Uncompressed (src/read_alignment_scanner.rs:66-71), i.e.
 * BGZF blocks holding one stored deflate block each.  ptl_bgzf_store_records frames [prefix | the records of the slot's last
 * ptl_assemble_records] that way on the device (SAM spec 4.1: 18-byte header with the BC subfield, 01 LEN NLEN, <= 0xff00
 * bytes, CRC32, ISIZE; the CRC is computed in the kernel), so the host only moves bytes to its pipe.  `prefix` (may be NULL)
 * is typically ptl_bam_header's output on the first batch; at most 32 KB.
 * flags: PTL_ASM_NO_DOWNLOAD (timing only), PTL_BGZF_EOF (append the 28-byte EOF marker). */
#define PTL_BGZF_EOF 4u
typedef struct {
    uint64_t n_bytes;               /* BGZF bytes */
    const uint8_t* bytes;           /* NULL with PTL_ASM_NO_DOWNLOAD */
    uint64_t n_blocks;
    float kernel_ms;                /* device time of the framing kernel (CUDA events on the slot stream) */
    uint64_t bytes_read;
    uint64_t bytes_written;  /* algorithmic bytes: the stream once in, the framed stream once out */
} ptl_bgzf_stream;
int ptl_bgzf_store_records(ptl_ctx* ctx, int slot, const uint8_t* prefix, uint64_t prefix_bytes, uint32_t flags, ptl_bgzf_stream* out);
/* ptl_assemble_records + ptl_bgzf_store_records in ONE pass: [prefix | every output record of the slot's last batch] as
 * level-0 BGZF, byte-identical to calling the two, but the record stream is never materialised: the payload of every BGZF
 * block is produced straight from the packed bases / qualities / names / aux of the reads (and the liftover result) and
 * its CRC from what the block has just written, so the device moves every byte once instead of twice.
 * `extras`, PTL_ASM_RESIDENT_QUAL and PTL_ASM_NO_DOWNLOAD as in ptl_assemble_records; PTL_BGZF_EOF appends the EOF marker. */
int ptl_frame_records(ptl_ctx* ctx, int slot, const ptl_read_extras* extras, const uint8_t* prefix, uint64_t prefix_bytes, uint32_t flags,
                      ptl_bgzf_stream* out);

/* cudaStream_t of a slot (as void*), so callers can bracket work with their own CUDA events. */
void* ptl_slot_stream(ptl_ctx* ctx, int slot);
/* Per-kernel device time of the LAST ptl_lift_run on the slot, CUDA events on the slot stream.
 * names[i] are static strings; returns the number of stages filled (<= cap). Syncs the slot. */
int ptl_slot_kernel_times(ptl_ctx* ctx, int slot, int cap, const char** names, float* ms);
/* Number of kernel launches issued by this ctx so far (bench.py `gpu_launches`). */
uint64_t ptl_launch_count(const ptl_ctx* ctx);
/* Work counters of the last finished batch on a slot (syncs the slot): out[0..6) = attempted pairs, lifted pairs,
 * input CIGAR ops walked, output CIGAR ops, base bytes compared (both operands), scratch op slots used.
 * These feed the algorithmic-bytes formula of the roofline. */
int ptl_slot_counters(ptl_ctx* ctx, int slot, uint64_t* out);
/* 0: copy seq4 to the device (default). 1: seq4 of subsequent batches must be pinned+mapped host memory
 * (ptl_host_alloc); kernels read the few bases they need over PCIe instead of uploading every base. */
int ptl_set_seq_zero_copy(ptl_ctx* ctx, int enable);
/* Tuning: a pair whose read->contig CIGAR has more than `n_ops` ops is lifted by a whole warp (lanes over ops) instead
 * of one thread; 0 sends every pair down the warp-cooperative path.  Same results either way (default 64). */
int ptl_set_long_pair_ops(ptl_ctx* ctx, uint32_t n_ops);


/* Pinned (page-locked, device-mapped) host memory for batches: replaces nothing in the reference, it is the
 * "pinned structure-of-arrays batches" of the north star.  The size is rounded up to whole 256-byte granules: the
 * kernels fetch aligned 4 / 8 / 16-byte words around the bytes they need, so a pool they read in place (seq4 in zero-copy
 * mode) must be readable up to the next 16-byte boundary behind its last byte -- memory from this call always is. */
void* ptl_host_alloc(size_t bytes);
void ptl_host_free(void* p);

/* ---------------------------------------------------------------- host-side helpers (C++, no GPU needed) */

/* Read->contig primary records as scan_chromosome_segment sees them after BAM decode
 * (src/read_alignment_scanner.rs:382-406: mapped, non-supplementary), in BAM order. */
typedef struct ptl_read_records {
    uint32_t n_reads;
    const int32_t* tid;            /* contig index */
    const int64_t* pos;
    const uint16_t* flag;
    const uint8_t* mapq;
    const uint16_t* bin;
    const uint32_t* seq_len;
    const uint64_t* seq_off;       /* byte offset of the record's packed bases in seq4 */
    const uint8_t* seq4;
    uint64_t seq4_bytes;
    const uint64_t* cigar_begin;   /* [n_reads+1] */
    const uint32_t* cigar;
    const char* const* sa_tag;     /* [n_reads] SA:Z value, NULL if absent */
} ptl_read_records;

/* The host packer (replaces the record loop head of scan_chromosome_segment, :393-421): turns records
 * [first, first+count) into one SoA ptl_batch, running get_seq_order_read_split_segments (split_read.rs:56-155) for
 * every record that carries an SA tag.  Small arrays are (optionally pinned) copies; the packed bases are NOT copied,
 * the batch borrows recs->seq4.  Returns PTL_ERR_INPUT where the reference panics (ptl_pack_last_error). */
typedef struct ptl_packed_batch ptl_packed_batch;
int ptl_pack_batch(const ptl_read_records* recs, uint32_t first, uint32_t count, uint32_t n_contigs,
                   const char* const* contig_names, int pinned, ptl_packed_batch** out);
/* As ptl_pack_batch, plus indel windows (ptl_batch.indel_win):
 *   PTL_WIN_NONE           no windows (= ptl_pack_batch)
 *   PTL_WIN_ALL            for every read segment
 *   PTL_WIN_REVERSE_PAIRS  only for read segments that pair (get_contig_split_segments_from_read_mapping,
 *                          src/read_alignment_scanner.rs:80-103) with a reverse-strand contig segment of `segs`
 *                          (ptl_get_contig_segments) - the only ones that go through left_shift_indels. */
enum { PTL_WIN_NONE = 0, PTL_WIN_ALL = 1, PTL_WIN_REVERSE_PAIRS = 2 };
int ptl_pack_batch_ex(const ptl_read_records* recs, uint32_t first, uint32_t count, uint32_t n_contigs,
                      const char* const* contig_names, int pinned, int window_mode, const ptl_contig_segments* segs,
                      ptl_packed_batch** out);
/* As ptl_pack_batch_ex, into an EXISTING packed batch whose arena is reused when it is large enough (same pinned-ness):
 * a streaming host keeps a few packed batches and repacks them, instead of allocating pinned memory per batch.  Distinct
 * packed batches may be packed from distinct host threads concurrently (one packer thread per slot or window). */
int ptl_pack_batch_into(ptl_packed_batch* reuse, const ptl_read_records* recs, uint32_t first, uint32_t count, uint32_t n_contigs,
                        const char* const* contig_names, int window_mode, const ptl_contig_segments* segs);
void ptl_packed_batch_view(const ptl_packed_batch* p, ptl_batch* out);
/* [view.n_reads] index of each batch read in `recs` (supplementary records are dropped by the packer). */
const uint32_t* ptl_packed_batch_record_index(const ptl_packed_batch* p);
void ptl_packed_batch_free(ptl_packed_batch* p);

/* Parse an SA:Z aux value into split segments = parse_sa_aux_val (lib/rust-vc-utils/src/bam_utils/aux/sa_tag_parser.rs:25-59)
 * + the segment construction / stable ordering of get_seq_order_read_split_segments (split_read.rs:56-155) for one
 * primary record.  `contig_names` resolves rname -> index (ChromList.label_to_index).
 * Outputs (caller-allocated, capacity `cap_segments` / `cap_cigar`): segments in sequencing order incl. the primary.
 * Returns PTL_OK, PTL_ERR_INPUT where the reference panics (message via ptl_pack_last_error), or
 * PTL_ERR_INVALID_ARG if capacities are too small (*n_segments / *n_cigar then hold the needed sizes). */
typedef struct ptl_split_segments {
    uint32_t* seq_order_start;
    uint32_t* seq_order_end;
    uint32_t* contig;
    int64_t* pos;
    uint8_t* is_fwd;
    uint8_t* mapq;
    uint8_t* from_primary;
    uint32_t* cigar_begin;  /* [cap_segments+1] */
    uint32_t* cigar;
} ptl_split_segments;
int ptl_pack_split_segments(uint32_t n_contig_names, const char* const* contig_names,
                            int32_t tid, int64_t pos, uint16_t flag, uint8_t mapq,
                            const uint32_t* cigar, uint32_t n_cigar, const char* sa_tag /* NULL if none */,
                            uint32_t cap_segments, uint32_t cap_cigar,
                            ptl_split_segments* out, uint32_t* n_segments, uint32_t* n_cigar_out);
const char* ptl_pack_last_error(void);

/* SA:Z text for every record of a result = get_sa_tag_segment + the concatenation loop
 * (src/read_alignment_scanner.rs:292-301,349-363).  Writes one NUL-terminated string per record into `buf`
 * (record i at sa_begin[i]; empty string where the reference pushes no SA tag). Returns needed bytes in *need. */
int ptl_format_sa_tags(const ptl_result* res, uint32_t n_chrom, const char* const* chrom_names,
                       char* buf, uint64_t cap, uint64_t* sa_begin /* [n_records+1] */, uint64_t* need);

/* Work-unit sharding (multi-GPU): the reference's units are (contig x <=20 Mb window), get_region_segments
 * (lib/rust-vc-utils/src/util.rs:50-67) as used at src/read_alignment_scanner.rs:508,574-576. */
uint32_t ptl_region_segment_count(uint64_t size, uint64_t segment_size);
void ptl_region_segments(uint64_t size, uint64_t segment_size, uint64_t* begin, uint64_t* end);
/* Greedy LPT bin-packing of units (weights = read counts) onto n_ranks; owner[i] = rank of unit i. Deterministic. */
void ptl_shard_units(uint32_t n_units, const uint64_t* weight, uint32_t n_ranks, uint32_t* owner);

/* ---------------------------------------------------------------- BAM / BGZF / BAI input without htslib (input half; host C++ + zlib)
 *
 * What the reference gets from rust-htslib on its input side: bam::IndexedReader::from_path, fetch(Region) / fetch(Unmapped)
 * and the read loop (src/read_alignment_scanner.rs:382-393,537-559; src/contig_alignment_scanner/mod.rs:196-203), the
 * header as a ChromList (lib/rust-vc-utils/src/chrom_list.rs:26-44) and the EOF-marker check (bam_reader_utils.rs:29).
 * Written from the SAM specification (4.1 BGZF, 4.2 BAM, 5.2 BAI).  A ptl_bam_file is immutable once opened: any number of
 * threads may fetch from it concurrently, which replaces the per-thread reader set of src/worker_thread_data.rs:8-30.
 * BAM with a .bai or .csi index (CRAM is not supported). */
typedef struct ptl_bam_file ptl_bam_file;
typedef struct ptl_decoded_batch ptl_decoded_batch;
enum {
    PTL_FETCH_ALL = -2,        /* every record of the file, in file order */
    PTL_FETCH_UNMAPPED = -1    /* FetchDefinition::Unmapped: the unplaced reads behind the last mapped one (:546) */
};
enum {
    PTL_BAM_START_IN_REGION = 1,        /* keep only records whose pos lies in [begin, end) (:403-406; mod.rs:213-217) */
    PTL_BAM_SKIP_SUPPLEMENTARY = 2,     /* (:404) */
    PTL_BAM_SKIP_UNMAPPED_SECONDARY = 4, /* (mod.rs:208) */
    PTL_BAM_ONLY_UNMAPPED = 8,          /* (:550-552) */
    PTL_BAM_KEEP_RAW = 16               /* also keep the undecoded record bytes (unmapped pass-through, :554) */
};
int ptl_bam_open(const char* path, ptl_bam_file** out);
void ptl_bam_close(ptl_bam_file* f);
const char* ptl_bam_last_error(void);                       /* thread-local message of the last failing ptl_bam_* / ptl_fasta_* call */
uint32_t ptl_bam_n_ref(const ptl_bam_file* f);
const char* ptl_bam_ref_name(const ptl_bam_file* f, uint32_t i);
uint64_t ptl_bam_ref_len(const ptl_bam_file* f, uint32_t i);
const char* ptl_bam_header_text(const ptl_bam_file* f);
int ptl_bam_has_index(const ptl_bam_file* f);
int ptl_bam_has_eof_marker(const ptl_bam_file* f);          /* assert_bam_eof */
/* Records overlapping [begin, end) of reference `tid` in file order (= fetch + read loop), or PTL_FETCH_*; decoded into
 * the SoA the packer and the record assembly take.  A CIGAR with more than 65535 ops is restored from its CG:B,I tag. */
int ptl_bam_fetch(const ptl_bam_file* f, int32_t tid, int64_t begin, int64_t end, uint32_t filter, ptl_decoded_batch** out);
void ptl_decoded_view(const ptl_decoded_batch* d, ptl_read_records* recs, ptl_read_extras* extras);
/* PTL_BAM_KEEP_RAW: record k = bytes[rec_off[k] .. rec_off[k+1]) (block_size + record), ready for a BGZF writer. */
const uint8_t* ptl_decoded_raw(const ptl_decoded_batch* d, const uint64_t** rec_off, uint64_t* n_bytes);
void ptl_decoded_free(ptl_decoded_batch* d);
/* samtools-index equivalent (coordinate-sorted BAM -> .bai with the linear index and the metadata pseudo-bins). */
int ptl_bam_index_build(const char* bam_path, const char* bai_path);
/* The same as a CSIv1 index (`.csi`: BGZF-compressed; bins of 2^min_shift bases at the finest of depth + 1 levels; what
 * htslib writes for references longer than 2^29 bases).  depth <= 0: the smallest depth that covers the longest reference.
 * ptl_bam_open looks for <bam>.bai, <stem>.bai, <bam>.csi, <stem>.csi in that order. */
int ptl_bam_index_build_csi(const char* bam_path, const char* csi_path, int min_shift, int depth);

/* ---------------------------------------------------------------- Phase A on real input
 *
 * get_genome_ref_from_fasta (lib/rust-vc-utils/src/genome_ref.rs:43-79): record id = header up to the first white space,
 * bases upper-cased, nothing else changed. */
typedef struct ptl_fasta ptl_fasta;
int ptl_fasta_load(const char* path, int n_threads, ptl_fasta** out);
uint32_t ptl_fasta_n(const ptl_fasta* f);
const char* ptl_fasta_name(const ptl_fasta* f, uint32_t i);
const uint8_t* ptl_fasta_seq(const ptl_fasta* f, uint32_t i, uint64_t* len);
void ptl_fasta_free(ptl_fasta* f);
/* The record loop of scan_contig_bam (src/contig_alignment_scanner/mod.rs:186-240,290-354): every reference chromosome of
 * the contig->reference BAM in <= 20 Mb windows on n_threads threads, records that START in their window, unmapped and
 * secondary records skipped, qname -> assembly contig index through `contig_names` (the read->assembly BAM header; an
 * unknown name is PTL_ERR_INPUT where the reference panics), bases of primary records decoded to ASCII.  The result views
 * as the ptl_contig_records that ptl_set_contig_records / ptl_prepare_contig_records take. */
typedef struct ptl_contig_scan ptl_contig_scan;
int ptl_scan_contig_bam(const ptl_bam_file* contig_bam, uint32_t n_contigs, const char* const* contig_names, const uint64_t* contig_len,
                        int n_threads, ptl_contig_scan** out);
void ptl_contig_scan_view(const ptl_contig_scan* s, ptl_contig_records* out);
void ptl_contig_scan_free(ptl_contig_scan* s);

/* BAM bin, = bam_reg2bin (lib/rust-vc-utils/src/bam_utils/util.rs:10-35). */
uint16_t ptl_reg2bin(int64_t begin, int64_t end);

#ifdef __cplusplus
}
#endif
#endif /* PORTELLO_B200_H */
