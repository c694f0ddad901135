// Record assembly, whole BAM records: every output record of a batch as the bytes bam_write1
// would emit, i.e. what clone_record + the field updates + push_aux calls of
// get_liftover_alignment_for_read_and_contig_segment (src/read_alignment_scanner.rs:105-117,245-282) and
// finish_remapped_alignment_set (:310-366) leave behind.
//
// Five kernels:
//   bam_read_prep   1 thread / read    walk the aux block once, find the FIRST NM, SA, PS, ZM field (remove_aux_if_found):
//                                      the surviving aux is <= 5 pieces of the input block
//   bam_rec_prep    1 thread / record  byte length of the record's own SA entry "{chrom},{pos+1},{+|-},{CIGAR},{mapq},0;"
//   bam_rec_size    1 thread / record  record size (needs the SA entries of the read's other records) -> exclusive scan
//   bam_write_meta  1 warp   / record  block_size + core, name, CIGAR, surviving aux, PS / ZM / SA: a few hundred bytes, byte-parallel
//   bam_write_bases 1 block  / record  HBM-bound streaming: ~22.5 KB in, ~22.5 KB out for a 15 kb read.  A BAM record is a
//                                      byte stream without alignment, so bases and qualities are produced in 16-byte chunks
//                                      aligned to the DESTINATION; the source window (arbitrary alignment, mirrored for a
//                                      flipped record) comes from five aligned 32-bit loads + funnel shifts as in
//                                      assemble.cuh; only the < 16 + 36 bytes at the two ends of a field go byte by byte.
#pragma once
#include <cstdint>

#include "assemble.cuh"
#include "device_types.hpp"

namespace ptl {

struct BamAsmArgs {
    uint32_t n_reads, n_records;
    // static
    const uint8_t* seg_is_fwd;
    const uint32_t* contig_seg_begin;
    const uint64_t* contig_name_off;  // [n_contigs+1]
    const uint8_t* contig_names;
    uint32_t n_contig_names;
    const uint64_t* chrom_name_off;   // [n_chrom+1]
    const uint8_t* chrom_names;
    uint32_t n_chrom_names;
    // batch
    const uint8_t* read_mapq;
    const uint32_t* read_seq_len;
    const uint64_t* read_seq_off;
    const uint8_t* seq4;
    const uint32_t* rseg_contig;
    const uint32_t* rseg_read;
    // extras
    const uint64_t* name_off;
    const uint8_t* names;
    const uint64_t* aux_off;
    const uint8_t* aux;
    const int32_t* mate_tid;
    const int32_t* mate_pos;
    const int32_t* tlen;
    const uint64_t* qual_off;
    const uint8_t* qual;
    uint64_t names_bytes, aux_bytes, qual_bytes;  // pool sizes: the offsets above are validated against them (bam_read_prep)
    // result of the liftover (views into the slot's result arena)
    const uint32_t* read_rec_begin;
    const int8_t* rec_status;
    const uint32_t* rec_read_segment;
    const uint32_t* rec_contig_segment;
    const int32_t* rec_tid;
    const int64_t* rec_pos;
    const uint8_t* rec_mapq;
    const uint16_t* rec_flag;
    const uint16_t* rec_bin;
    const uint8_t* rec_need_flip;
    const uint64_t* rec_cigar_begin;
    const uint32_t* cigar;
    // work
    uint32_t* read_keep;     // [n_reads][10]: 5 x (offset, length) pieces of the aux block that survive clone_record
    uint32_t* rec_sa_len;    // [n_records] bytes of the record's own SA entry (0 for the unmapped fallback)
    uint64_t* rec_begin;     // [n_records+1] record sizes, then their exclusive scan
    uint4* rec_desc;         // [n_records][2] BamRecLayout of the record, computed once by bam_rec_size (the writer's prologue
                             //                is then two dependent loads deep instead of five)
    uint8_t* out;
    unsigned int* error;     // bit 0: a name is missing, 1: a CIGAR of more than 65535 ops spans >= 2^28 reference bases, 2: descriptor field overflow, 3: qname > 254 bytes,
                             // bit 4: an offset of the extras (names / aux / qualities) lies outside its pool
};

__device__ __forceinline__ uint32_t dec_digits(uint64_t v) {
    uint32_t n = 1;
    while (v >= 10) { v /= 10; ++n; }
    return n;
}
// decimal text of v at p (no terminator); returns the number of bytes
__device__ __forceinline__ uint32_t put_dec(uint8_t* p, uint64_t v) {
    const uint32_t n = dec_digits(v);
    for (uint32_t i = n; i-- > 0;) { p[i] = uint8_t('0' + v % 10); v /= 10; }
    return n;
}

// -------------------------------------------------------------------------------------------------------------- aux
// Size of the aux field whose type byte is aux[i + 2] (SAM spec 4.2.4), or 0 when malformed / running past n.
__device__ __forceinline__ uint32_t aux_field_size(const uint8_t* aux, uint32_t i, uint32_t n) {
    if (i + 3u > n) return 0;
    const uint8_t type = aux[i + 2];
    uint32_t sz;
    switch (type) {
        case 'A': case 'c': case 'C': sz = 1; break;
        case 's': case 'S': sz = 2; break;
        case 'i': case 'I': case 'f': sz = 4; break;
        case 'd': sz = 8; break;
        case 'Z': case 'H': {
            uint32_t j = i + 3u;
            while (j < n && aux[j] != 0) ++j;
            if (j >= n) return 0;
            sz = j - (i + 3u) + 1u;
            break;
        }
        case 'B': {
            if (i + 8u > n) return 0;
            const uint8_t sub = aux[i + 3];
            const uint32_t cnt = uint32_t(aux[i + 4]) | (uint32_t(aux[i + 5]) << 8) | (uint32_t(aux[i + 6]) << 16) | (uint32_t(aux[i + 7]) << 24);
            uint32_t es;
            switch (sub) {
                case 'c': case 'C': es = 1; break;
                case 's': case 'S': es = 2; break;
                case 'i': case 'I': case 'f': es = 4; break;
                default: return 0;
            }
            const uint64_t big = 5ull + uint64_t(cnt) * es;
            if (big > n) return 0;
            sz = uint32_t(big);
            break;
        }
        default: return 0;
    }
    if (uint64_t(i) + 3u + sz > n) return 0;
    return 3u + sz;
}

// clone_record (:105-117): remove_aux_if_found(NM), (SA), (PS), (ZM), each the FIRST field of that name.  One walk finds
// all four (removing one field does not change which field of another name comes first).
__device__ __forceinline__ void bam_read_prep_body(const BamAsmArgs& A, uint32_t r) {
    // The extras are caller memory: an offset outside its pool must be reported, never dereferenced (the writers only run
    // after the host has seen a clean error word).
    {
        const uint64_t n0 = A.name_off[r], n1 = A.name_off[r + 1], a0 = A.aux_off[r], a1 = A.aux_off[r + 1], q0 = A.qual_off[r];
        const bool bad = n0 > n1 || n1 > A.names_bytes || a0 > a1 || a1 > A.aux_bytes || a1 - a0 > 0x7fffffffull ||
                         q0 > A.qual_bytes || uint64_t(A.read_seq_len[r]) > A.qual_bytes - q0;
        if (bad) {
            atomicOr(A.error, 16u);
            uint32_t* keep = A.read_keep + size_t(r) * 10;
            for (int j = 0; j < 10; ++j) keep[j] = 0u;
            return;
        }
    }
    const uint8_t* aux = A.aux + A.aux_off[r];
    const uint32_t n = uint32_t(A.aux_off[r + 1] - A.aux_off[r]);
    uint32_t cut_a[4], cut_e[4];
    bool found[4] = {false, false, false, false};
    uint32_t i = 0;
    for (;;) {
        const uint32_t sz = aux_field_size(aux, i, n);
        if (sz == 0) break;
        const uint32_t t = (uint32_t(aux[i]) << 8) | aux[i + 1];
        const int w = (t == (('N' << 8) | 'M')) ? 0 : (t == (('S' << 8) | 'A')) ? 1 : (t == (('P' << 8) | 'S')) ? 2 : (t == (('Z' << 8) | 'M')) ? 3 : -1;
        if (w >= 0 && !found[w]) { found[w] = true; cut_a[w] = i; cut_e[w] = i + sz; }
        i += sz;
    }
    // pieces between the cuts, in input order (the cuts are disjoint fields: sort the <= 4 of them by offset)
    uint32_t ca[4], ce[4], nc = 0;
    for (int w = 0; w < 4; ++w)
        if (found[w]) {
            uint32_t j = nc++;
            while (j > 0 && ca[j - 1] > cut_a[w]) { ca[j] = ca[j - 1]; ce[j] = ce[j - 1]; --j; }
            ca[j] = cut_a[w]; ce[j] = cut_e[w];
        }
    uint32_t* keep = A.read_keep + size_t(r) * 10;
    uint32_t at = 0;
    for (uint32_t j = 0; j < 5; ++j) {
        const uint32_t end = (j < nc) ? ca[j] : n;
        keep[2 * j] = at;
        keep[2 * j + 1] = (j <= nc) ? end - at : 0u;
        at = (j < nc) ? ce[j] : n;
    }
}

// ---------------------------------------------------------------------------------------------------- SA / PS text
// get_sa_tag_segment (:292-301) of record k: its length, or its bytes at p (p == nullptr: count only)
__device__ __forceinline__ uint32_t sa_entry(const BamAsmArgs& A, uint32_t k, uint8_t* p) {
    const uint32_t tid = uint32_t(A.rec_tid[k]);
    const uint64_t c0 = A.chrom_name_off[tid], c1 = A.chrom_name_off[tid + 1];
    uint32_t n = 0;
    auto put = [&](uint8_t c) { if (p) p[n] = c; ++n; };
    for (uint64_t i = c0; i < c1; ++i) put(A.chrom_names[i]);
    put(',');
    const uint64_t pos1 = uint64_t(A.rec_pos[k] + 1);
    if (p) n += put_dec(p + n, pos1); else n += dec_digits(pos1);
    put(',');
    put((A.rec_flag[k] & 0x10) ? '-' : '+');
    put(',');
    for (uint64_t i = A.rec_cigar_begin[k]; i < A.rec_cigar_begin[k + 1]; ++i) {
        const uint32_t c = A.cigar[i];
        if (p) n += put_dec(p + n, c >> 4); else n += dec_digits(c >> 4);
        put("MIDNSHP=X"[min(c & 0xfu, 8u)]);
    }
    put(',');
    if (p) n += put_dec(p + n, A.rec_mapq[k]); else n += dec_digits(A.rec_mapq[k]);
    put(',');
    put('0');
    put(';');
    return n;
}

// bam_cigar2rlen: reference bases of M, D, N, = and X ops
__device__ __forceinline__ uint64_t cigar_ref_len(const uint32_t* c, uint64_t n) {
    uint64_t len = 0;
    for (uint64_t i = 0; i < n; ++i)
        if ((0x18du >> (c[i] & 0xfu)) & 1u) len += c[i] >> 4;
    return len;
}

__device__ __forceinline__ void bam_rec_prep_body(const BamAsmArgs& A, uint32_t k) {
    uint32_t n = 0;
    if (A.rec_status[k] == 1) {
        const int32_t tid = A.rec_tid[k];
        const uint32_t ctg = A.rseg_contig[A.rec_read_segment[k]];
        if (tid < 0 || uint32_t(tid) >= A.n_chrom_names || ctg >= A.n_contig_names) atomicOr(A.error, 1u);
        else n = sa_entry(A, k, nullptr);
        // more than 65535 ops: the record carries a placeholder CIGAR and a CG:B,I tag (BamRecLayout); bam_write1 refuses it
        // only when the placeholder's N op cannot hold the reference length
        const uint64_t c0 = A.rec_cigar_begin[k], c1 = A.rec_cigar_begin[k + 1];
        if (c1 - c0 > 65535ull && cigar_ref_len(A.cigar + c0, c1 - c0) >= (1ull << 28)) atomicOr(A.error, 2u);
    }
    A.rec_sa_len[k] = n;
}

// A CIGAR of more than 65535 ops does not fit n_cigar_op: as bam_write1 (htslib sam.c) does, the CIGAR field then holds the
// two-op placeholder <l_seq>S<ref_len>N and the real ops follow every other tag as CG:B,I (SAM spec 4.2.2).
struct BamRecLayout {
    uint32_t r, k0, k1;
    uint32_t name_n, n_cigar, l_seq, seq_bytes, keep_n, ps_n, sa_n;  // ps_n / sa_n: payload bytes without tag, type and NUL
    bool lifted;
    uint64_t total;  // block_size + 4
    __device__ __forceinline__ bool long_cigar() const { return n_cigar > 65535u; }
    __device__ __forceinline__ uint64_t cigar_field() const { return long_cigar() ? 8ull : 4ull * n_cigar; }          // bytes of the CIGAR field
    __device__ __forceinline__ uint64_t cg_tail() const { return long_cigar() ? 8ull + 4ull * n_cigar : 0ull; }       // "CGBI" + count + ops
    __device__ __forceinline__ uint64_t o_seq() const { return 36ull + name_n + 1ull + cigar_field(); }
    __device__ __forceinline__ uint64_t size() const {
        uint64_t t = 4 + 32 + uint64_t(name_n) + 1 + cigar_field() + seq_bytes + l_seq + keep_n + cg_tail();
        if (lifted) t += 3ull + ps_n + 1 + 4 + (sa_n ? 3ull + sa_n + 1 : 0ull);
        return t;
    }
    __device__ __forceinline__ void pack(uint4* d) const {
        d[0] = make_uint4(r, k0, k1 - k0, name_n | (lifted ? 0x80000000u : 0u));
        d[1] = make_uint4(n_cigar, l_seq, keep_n, ps_n | (sa_n << 8));
    }
    static __device__ __forceinline__ BamRecLayout unpack(const uint4* d) {
        const uint4 a = d[0], b = d[1];
        BamRecLayout L;
        L.r = a.x; L.k0 = a.y; L.k1 = a.y + a.z;
        L.name_n = a.w & 0x7fffffffu; L.lifted = (a.w >> 31) != 0;
        L.n_cigar = b.x; L.l_seq = b.y; L.seq_bytes = (b.y + 1u) >> 1; L.keep_n = b.z;
        L.ps_n = b.w & 0xffu; L.sa_n = b.w >> 8;
        L.total = L.size();
        return L;
    }
};

__device__ __forceinline__ BamRecLayout bam_rec_layout(const BamAsmArgs& A, uint32_t k) {
    BamRecLayout L;
    {   // the read that owns record k: last r with read_rec_begin[r] <= k (a fallback record of a read without segments has
        // no valid rec_read_segment to go through)
        uint32_t lo = 0, hi = A.n_reads;
        while (hi - lo > 1u) {
            const uint32_t mid = (lo + hi) >> 1;
            if (A.read_rec_begin[mid] <= k) lo = mid; else hi = mid;
        }
        L.r = lo;
    }
    L.k0 = A.read_rec_begin[L.r];
    L.k1 = A.read_rec_begin[L.r + 1];
    L.lifted = A.rec_status[k] == 1;
    L.name_n = uint32_t(A.name_off[L.r + 1] - A.name_off[L.r]);
    L.n_cigar = L.lifted ? uint32_t(A.rec_cigar_begin[k + 1] - A.rec_cigar_begin[k]) : 0u;
    L.l_seq = A.read_seq_len[L.r];
    L.seq_bytes = (L.l_seq + 1u) >> 1;
    const uint32_t* keep = A.read_keep + size_t(L.r) * 10;
    L.keep_n = keep[1] + keep[3] + keep[5] + keep[7] + keep[9];
    L.ps_n = 0;
    L.sa_n = 0;
    if (L.lifted) {
        const uint32_t ctg = A.rseg_contig[A.rec_read_segment[k]];
        if (ctg < A.n_contig_names)
            L.ps_n = uint32_t(A.contig_name_off[ctg + 1] - A.contig_name_off[ctg]) + 6u + dec_digits(A.rec_contig_segment[k]) + 1u;
        for (uint32_t j = L.k0; j < L.k1; ++j)
            if (j != k) L.sa_n += A.rec_sa_len[j];
    }
    L.total = L.size();
    return L;
}

// ------------------------------------------------------------------------------------------------------- the writer
// 16 bytes `v` of a field at field offset o (dst + o is 16-byte aligned); bytes outside [0, flen) are not stored.
// (the rare paths are kept out of line: inlined, they cost the streaming loop 120 registers and its occupancy)
static __device__ __noinline__ void put16_partial(uint8_t* dst, int64_t o, int64_t flen, uint4 v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int b = 0; b < 16; ++b)
        if (o + b >= 0 && o + b < flen) dst[o + b] = uint8_t(w[b >> 2] >> (8 * (b & 3)));
}
__device__ __forceinline__ void put16(uint8_t* dst, int64_t o, int64_t flen, const uint4& v) {
    if (o >= 0 && o + 16 <= flen) *reinterpret_cast<uint4*>(dst + o) = v;
    else put16_partial(dst, o, flen, v);
}
static __device__ __noinline__ uint4 window128_guarded(const uint8_t* base, int64_t off, int64_t len) {
    uint4 v;
    v.x = window32(base, off, len);
    v.y = window32(base, off + 4, len);
    v.z = window32(base, off + 8, len);
    v.w = window32(base, off + 12, len);
    return v;
}
// window128 (assemble.cuh) with the guarded path out of line
__device__ __forceinline__ uint4 window128_lean(const uint8_t* __restrict__ base, int64_t off, int64_t len) {
    const uint64_t addr = reinterpret_cast<uint64_t>(base) + uint64_t(off);
    const int64_t a0 = off - int64_t(addr & 3ull);
    if (a0 >= 0 && a0 + 20 <= len) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(addr & ~3ull);
        const uint32_t sh = uint32_t(addr & 3ull) * 8u;
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
        uint4 v;
        v.x = __funnelshift_r(w0, w1, sh);
        v.y = __funnelshift_r(w1, w2, sh);
        v.z = __funnelshift_r(w2, w3, sh);
        v.w = __funnelshift_r(w3, w4, sh);
        return v;
    }
    return window128_guarded(base, off, len);
}
template <class F>
__device__ __forceinline__ void emit_field(uint8_t* dst, int64_t flen, uint32_t tid, uint32_t n_threads, F produce) {
    if (flen <= 0) return;
    const int64_t o0 = -int64_t(reinterpret_cast<uint64_t>(dst) & 15ull);
    const uint32_t n_chunks = uint32_t((flen - o0 + 15) >> 4);
    for (uint32_t j = tid; j < n_chunks; j += n_threads) {
        const int64_t o = o0 + 16 * int64_t(j);
        put16(dst, o, flen, produce(o));
    }
}
// Plain copy of a field.  [head: bytes up to the first 16-byte boundary of the destination | body: aligned 16-byte stores,
// each fed by five aligned 32-bit loads + funnel shifts, two chunks in flight per thread | tail: < 36 bytes].  Head and tail go
// byte by byte (a byte per thread): the guarded partial-chunk paths cost ~190 instructions per chunk at 1-2 active lanes and
// kept warp 0 of every block busy three times longer than the streaming itself (ncu s5g).
// The body may read up to 3 bytes in front of `src` (inside the same pool) but never past src + flen.
__device__ __forceinline__ uint4 window128_body(const uint8_t* __restrict__ p) {
    const uint64_t addr = reinterpret_cast<uint64_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(addr & ~3ull);
    const uint32_t sh = uint32_t(addr & 3ull) * 8u;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
    uint4 v;
    v.x = __funnelshift_r(w0, w1, sh);
    v.y = __funnelshift_r(w1, w2, sh);
    v.z = __funnelshift_r(w2, w3, sh);
    v.w = __funnelshift_r(w3, w4, sh);
    return v;
}
// (one out-of-line copy of the loop for all nine call sites: inlined, the kernel was 5.7 k instructions and stalled on
// instruction fetch; 32-bit index arithmetic: a BAM record is < 2 GB)
static __device__ __noinline__ void copy_field(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t flen, uint32_t tid, uint32_t n_threads) {
    if (flen == 0) return;
    const uint32_t head = min((16u - uint32_t(reinterpret_cast<uint64_t>(dst) & 15ull)) & 15u, flen);
    const uint32_t n_body = (flen - head >= 20u) ? (flen - head - 20u) / 16u + 1u : 0u;  // chunk c reads src[head + 16 c - 3, head + 16 c + 20)
    const uint32_t tail = head + 16u * n_body;
    if (tid < head) dst[tid] = src[tid];                                         // (head < 16 <= n_threads)
    for (uint32_t i = tail + tid; i < flen; i += n_threads) dst[i] = src[i];     // (< 36 bytes)
    uint4* d = reinterpret_cast<uint4*>(dst + head);
    const uint8_t* sb = src + head;
    for (uint32_t c = tid; c < n_body; c += 2u * n_threads) {
        const uint32_t c2 = c + n_threads;
        const uint4 a = window128_body(sb + 16u * c);
        if (c2 < n_body) {
            const uint4 b = window128_body(sb + 16u * c2);
            d[c] = a;
            d[c2] = b;
        } else {
            d[c] = a;
        }
    }
}

// 16 output bytes at byte offset o of the reverse-complemented packed bases of a read of `len` bases (assemble.cuh)
static __device__ __noinline__ uint4 revcomp_chunk(const uint8_t* src_s, int64_t len, int64_t o) {
    const int64_t seq_bytes = (len + 1) >> 1;
    const int64_t a = len - 2 * o - 32;  // source nibbles [a, a + 32) in nibble-monotonic form
    const int64_t byte0 = a >> 1;        // floor: a may be negative
    uint4 s = window128(src_s, byte0, seq_bytes);
    s.x = swap_nibbles(s.x); s.y = swap_nibbles(s.y); s.z = swap_nibbles(s.z); s.w = swap_nibbles(s.w);
    if (a & 1) {
        const uint32_t t = swap_nibbles(window32(src_s, byte0 + 16, seq_bytes));
        s.x = __funnelshift_r(s.x, s.y, 4u);
        s.y = __funnelshift_r(s.y, s.z, 4u);
        s.z = __funnelshift_r(s.z, s.w, 4u);
        s.w = __funnelshift_r(s.w, t, 4u);
    }
    uint32_t w[4] = {revcomp8(s.w), revcomp8(s.z), revcomp8(s.y), revcomp8(s.x)};
    // nibbles outside the read (output index < 0 or >= len) are padding: zero
    const int64_t first_out = 2 * o;
    if (first_out < 0 || first_out + 32 > len) {
#pragma unroll
        for (int t = 0; t < 32; ++t)
            if (first_out + t < 0 || first_out + t >= len) w[t >> 3] &= ~(0xfu << (4 * (t & 7)));
    }
    uint4 v;
    v.x = swap_nibbles(w[0]); v.y = swap_nibbles(w[1]); v.z = swap_nibbles(w[2]); v.w = swap_nibbles(w[3]);
    return v;
}

// Everything of record k except bases and qualities (a few hundred bytes): block_size + core, name, CIGAR, surviving aux
// pieces, PS / ZM / SA tags.  `n_threads` threads (one warp on the device) copy byte-parallel; text goes through single lanes.
__device__ __forceinline__ void bam_write_meta_body(const BamAsmArgs& A, uint32_t k, uint32_t tid, uint32_t n_threads) {
    const BamRecLayout L = BamRecLayout::unpack(A.rec_desc + 2 * size_t(k));
    uint8_t* out = A.out + A.rec_begin[k];
    const uint32_t r = L.r;
    // ---- field offsets
    const uint64_t o_name = 36, o_cigar = o_name + L.name_n + 1, o_seq = L.o_seq(), o_qual = o_seq + L.seq_bytes,
                   o_aux = o_qual + L.l_seq, o_ps = o_aux + L.keep_n, o_cg = L.total - L.cg_tail();
    const bool long_cigar = L.long_cigar();
    // ---- block_size + core (36 bytes) and the NUL of the name: one thread, byte stores (unaligned destination)
    if (tid == 0) {
        const uint32_t h[9] = {uint32_t(L.total - 4),
                               uint32_t(A.rec_tid[k]),
                               uint32_t(int32_t(A.rec_pos[k])),
                               ((L.name_n + 1u) & 0xffu) | (uint32_t(A.rec_mapq[k]) << 8) | (uint32_t(A.rec_bin[k]) << 16),
                               (long_cigar ? 2u : L.n_cigar) | (uint32_t(A.rec_flag[k]) << 16),
                               L.l_seq,
                               uint32_t(A.mate_tid[r]),
                               uint32_t(A.mate_pos[r]),
                               uint32_t(A.tlen[r])};
        for (int i = 0; i < 36; ++i) out[i] = uint8_t(h[i >> 2] >> (8 * (i & 3)));
        out[o_name + L.name_n] = 0;
        if (long_cigar) {  // placeholder CIGAR, and the head of the CG tag
            const uint32_t ref_len = uint32_t(cigar_ref_len(A.cigar + A.rec_cigar_begin[k], L.n_cigar));  // < 2^28 (bam_rec_prep_body)
            const uint32_t w[4] = {(L.l_seq << 4) | 4u, (ref_len << 4) | 3u, 0x49424743u /* "CGBI" */, L.n_cigar};
            for (int i = 0; i < 8; ++i) out[o_cigar + i] = uint8_t(w[i >> 2] >> (8 * (i & 3)));
            for (int i = 0; i < 8; ++i) out[o_cg + i] = uint8_t(w[2 + (i >> 2)] >> (8 * (i & 3)));
        }
    }
    // ---- PS:Z, ZM:C (:255-269) and SA:Z (:349-363): short text, one thread each
    if (L.lifted && tid == 1 % n_threads) {
        uint8_t* p = out + o_ps;
        const uint32_t ctg = A.rseg_contig[A.rec_read_segment[k]];
        *p++ = 'P'; *p++ = 'S'; *p++ = 'Z';
        if (ctg < A.n_contig_names)
            for (uint64_t i = A.contig_name_off[ctg]; i < A.contig_name_off[ctg + 1]; ++i) *p++ = A.contig_names[i];
        const char* sp = "_split";
        for (int i = 0; i < 6; ++i) *p++ = uint8_t(sp[i]);
        p += put_dec(p, A.rec_contig_segment[k]);
        *p++ = A.seg_is_fwd[A.contig_seg_begin[ctg < A.n_contig_names ? ctg : 0] + A.rec_contig_segment[k]] ? '+' : '-';
        *p++ = 0;
        *p++ = 'Z'; *p++ = 'M'; *p++ = 'C'; *p++ = A.read_mapq[r];
        if (L.sa_n) {
            *p++ = 'S'; *p++ = 'A'; *p++ = 'Z';
            p[L.sa_n] = 0;
        }
    }
    if (L.lifted && L.sa_n) {
        // entry of the j-th other record: one thread each (reads with several records are few and have few records)
        const uint64_t o_sa = o_ps + 3ull + L.ps_n + 1 + 4 + 3;
        const uint32_t n_other = L.k1 - L.k0;
        for (uint32_t j = (tid + n_threads - 2 % n_threads) % n_threads; j < n_other; j += n_threads) {
            const uint32_t kj = L.k0 + j;
            if (kj == k) continue;
            uint64_t at = o_sa;
            for (uint32_t i = L.k0; i < kj; ++i)
                if (i != k) at += A.rec_sa_len[i];
            sa_entry(A, kj, out + at);
        }
    }
    // ---- name, CIGAR, surviving aux pieces: byte-parallel copies
    auto copy_small = [&](uint8_t* d, const uint8_t* src, uint32_t n) {
        for (uint32_t i = tid; i < n; i += n_threads) d[i] = src[i];
    };
    copy_small(out + o_name, A.names + A.name_off[r], L.name_n);
    copy_small(out + (long_cigar ? o_cg + 8 : o_cigar), reinterpret_cast<const uint8_t*>(A.cigar + A.rec_cigar_begin[k]), 4u * L.n_cigar);
    {
        const uint32_t* keep = A.read_keep + size_t(r) * 10;
        const uint8_t* aux = A.aux + A.aux_off[r];
        uint64_t at = o_aux;
        for (int j = 0; j < 5; ++j) {
            const uint32_t len = keep[2 * j + 1];
            copy_small(out + at, aux + keep[2 * j], len);
            at += len;
        }
    }
}

// Bases and qualities of record k (99 % of its bytes): the streaming part, one block per record
// (reverse_alignment_seq_and_qual when flipped, :125-133).
__device__ __forceinline__ void bam_write_bases_body(const BamAsmArgs& A, uint32_t k, uint32_t tid, uint32_t n_threads) {
    const BamRecLayout L = BamRecLayout::unpack(A.rec_desc + 2 * size_t(k));
    uint8_t* out = A.out + A.rec_begin[k];
    const uint32_t r = L.r;
    const bool flip = A.rec_need_flip[k] != 0;
    const uint64_t o_seq = L.o_seq(), o_qual = o_seq + L.seq_bytes;
    const uint8_t* src_s = A.seq4 + A.read_seq_off[r];
    const uint8_t* src_q = A.qual + A.qual_off[r];
    if (!flip) {
        copy_field(out + o_seq, src_s, L.seq_bytes, tid, n_threads);
        copy_field(out + o_qual, src_q, L.l_seq, tid, n_threads);
    } else {
        const int64_t len = L.l_seq;
        emit_field(out + o_seq, L.seq_bytes, tid, n_threads, [&](int64_t o) { return revcomp_chunk(src_s, len, o); });
        emit_field(out + o_qual, len, tid, n_threads, [&](int64_t o) {
            const uint4 s = window128_lean(src_q, len - 16 - o, len);  // the 16 bytes in front of the mirrored position
            uint4 v;
            v.x = __byte_perm(s.w, 0u, 0x0123u);
            v.y = __byte_perm(s.z, 0u, 0x0123u);
            v.z = __byte_perm(s.y, 0u, 0x0123u);
            v.w = __byte_perm(s.x, 0u, 0x0123u);
            return v;
        });
    }
}

}  // namespace ptl
