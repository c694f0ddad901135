/*
 * ptl_synth.h — deterministic synthetic data for the liftover path (bench + test infrastructure, not product code).
 *
 * Emits data at the level a BAM decoder would hand to portello: an upper-case ASCII reference, the
 * contig->reference alignment records (primary + supplementary + approximate SA tags, =/X CIGARs as minimap2 --eqx
 * writes them) and the read->contig records (pbmm2-like =/X CIGARs, 4-bit packed bases, optional SA tags), all
 * ground-truth by construction.  Every invariant the reference asserts on its inputs holds (SURVEY.md §8d list).
 */
#ifndef PTL_SYNTH_H
#define PTL_SYNTH_H
#include <stdint.h>

#include "../../../include/portello_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ptl_synth_params {
    uint64_t seed;
    uint32_t n_chrom;
    uint64_t chrom_len;            /* every chromosome has this length (+- nothing); use n_chrom to scale */
    uint32_t haplotypes;           /* contig sets tiling every chromosome (2 = diploid assembly) */
    uint32_t contigs_per_chrom;    /* per haplotype */
    double rev_contig_frac;        /* fraction of contigs whose primary alignment is reverse-strand */
    double unmapped_contig_frac;   /* contigs with no record in the contig BAM (reads fall back to unmapped) */
    double contig_snv_rate, contig_indel_rate; /* per reference base */
    double sv_per_mb;              /* 50 bp - 10 kb I/D kept inside one alignment */
    double junction_per_mb;        /* split-alignment junctions (z-drop joinable / gap / inversion / overlap) */
    uint64_t n_reads;
    double read_len_mean, read_len_sd;
    uint32_t read_len_min, read_len_max;
    double read_sub_rate, read_indel_rate; /* per read base */
    double read_cluster_frac;      /* fraction of read indels followed by an adjacent opposite indel (a9 work) */
    double read_clip_frac;         /* reads with a soft-clipped end */
    double read_sa_frac;           /* reads carrying 1-2 SA segments */
    uint32_t n_threads;            /* 0 = hardware concurrency */
    uint32_t defer_reads;          /* 1: only PLAN the reads at creation (ptl_synth_plan_*); ptl_synth_generate_reads makes them */
} ptl_synth_params;

typedef struct ptl_synth ptl_synth;

void ptl_synth_default_params(ptl_synth_params* p);   /* BASELINE.json configs[0]-like small case */
ptl_synth* ptl_synth_create(const ptl_synth_params* p);
void ptl_synth_destroy(ptl_synth* s);

/* Borrowed views, valid until destroy. */
uint32_t ptl_synth_n_chrom(const ptl_synth* s);
const uint64_t* ptl_synth_chrom_len(const ptl_synth* s);
const uint8_t* const* ptl_synth_chrom_seq(const ptl_synth* s);
const char* const* ptl_synth_chrom_names(const ptl_synth* s);
uint32_t ptl_synth_n_contigs(const ptl_synth* s);
const char* const* ptl_synth_contig_names(const ptl_synth* s);
void ptl_synth_contig_records(const ptl_synth* s, ptl_contig_records* out);
void ptl_synth_read_records(const ptl_synth* s, ptl_read_records* out);
/* As ptl_synth_create, but the big packed-bases pool is allocated with `alloc` (e.g. ptl_host_alloc for pinned memory). */
ptl_synth* ptl_synth_create_into(const ptl_synth_params* p, void* (*alloc)(size_t), void (*dealloc)(void*));


/* The read set is planned up front and generated on demand, so that a whole-genome set (6 M reads, 45 GB of packed bases)
 * can be produced shard by shard: read i of the BAM order depends only on (seed, plan[i]).
 * Plan = the reads in coordinate-sorted BAM order: contig index and record position (what assigns a read to one of the
 * reference's (contig x <= 20 Mb window) work units, src/read_alignment_scanner.rs:403-406,508-534). */
uint64_t ptl_synth_n_planned(const ptl_synth* s);
const uint32_t* ptl_synth_plan_contig(const ptl_synth* s);
const int64_t* ptl_synth_plan_pos(const ptl_synth* s);
/* (Re)generate the read records of `n_ranges` ranges [first[i], first[i] + count[i]) of the BAM order, concatenated in
 * the order given; replaces what ptl_synth_read_records returns (earlier views become invalid).  Returns 0 on success. */
int ptl_synth_generate_reads(ptl_synth* s, uint32_t n_ranges, const uint64_t* first, const uint64_t* count);
/* The data as an UNCOMPRESSED BAM byte stream (header + records; frame it with ptl_bgzf_compress, index it with
 * ptl_bam_index_build): which = 0 the contig->reference alignments (minimap2 --eqx style; CIGARs over 65535 ops use the
 * CG:B,I placeholder form), which = 1 the read->contig alignments currently generated (names "synth/<index>/ccs", random
 * qualities, NM / np / SA / RG tags) followed by n_unmapped unplaced reads.  Free with ptl_synth_free_bytes. */
uint8_t* ptl_synth_bam_stream(const ptl_synth* s, int which, uint32_t n_unmapped, uint64_t* n_bytes);
void ptl_synth_free_bytes(uint8_t* p);
/* Contig lengths (the read->assembly BAM header). */
const uint64_t* ptl_synth_contig_len(const ptl_synth* s);

#ifdef __cplusplus
}
#endif
#endif
