// Synthetic data generator for the liftover path (bench + test infrastructure; see ptl_synth.h).
//
// Model: a random reference with planted homopolymer/STR runs; per chromosome and haplotype a tiling
// of contigs, each aligned to the reference as 1..k split segments (z-drop gaps that portello joins, larger gaps,
// inversions, overlapping "repeated match" junctions that portello trims), =/X CIGARs with SNVs, small indels and
// SV-sized I/D; HiFi-like reads sampled from the contigs with substitutions, left-normalised indels (as an aligner
// reports them on the contig's forward strand), occasional adjacent I/D clusters, soft clips and chimeric reads with SA.
#include "ptl_synth.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------ RNG
struct Rng {
    uint64_t s[2];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    Rng(uint64_t seed, uint64_t stream) {
        uint64_t x = seed * 0xD1342543DE82EF95ull + stream * 0x9E3779B97F4A7C15ull + 0x1234567;
        s[0] = splitmix(x);
        s[1] = splitmix(x);
        if (!(s[0] | s[1])) s[0] = 1;
    }
    uint64_t next() {  // xoroshiro128+
        const uint64_t a = s[0];
        uint64_t b = s[1];
        const uint64_t r = a + b;
        b ^= a;
        s[0] = ((a << 24) | (a >> 40)) ^ b ^ (b << 16);
        s[1] = (b << 37) | (b >> 27);
        return r;
    }
    double uni() { return double(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return n ? next() % n : 0; }
    int64_t range(int64_t lo, int64_t hi) { return lo + int64_t(below(uint64_t(hi - lo + 1))); }  // inclusive
    bool chance(double p) { return uni() < p; }
    // number of failures before the first success, success prob p (mean (1-p)/p)
    uint64_t geometric(double p) {
        if (p >= 1.0) return 0;
        if (p <= 0.0) return UINT64_MAX / 4;
        const double u = std::max(uni(), 1e-300);
        return uint64_t(std::log(u) / std::log1p(-p));
    }
    double normal() {
        const double u1 = std::max(uni(), 1e-300), u2 = uni();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

const char kBases[] = "ACGT";
inline uint8_t rand_base(Rng& r) { return uint8_t(kBases[r.next() & 3]); }
inline uint8_t other_base(Rng& r, uint8_t b) {
    uint8_t x;
    do x = rand_base(r); while (x == b);
    return x;
}
inline uint8_t comp(uint8_t b) {
    switch (b) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return 'N'; }
}
inline uint8_t code4(uint8_t b) {
    switch (b) { case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8; default: return 15; }
}

enum : uint8_t { M = 0, I = 1, D = 2, N = 3, S = 4, H = 5, P = 6, EQ = 7, X = 8 };
using Ops = std::vector<uint32_t>;
inline void push_op(Ops& v, uint8_t op, uint64_t len) {
    if (!len) return;
    if (!v.empty() && (v.back() & 0xf) == op) v.back() += uint32_t(len << 4);
    else v.push_back(uint32_t(len << 4) | op);
}
inline uint64_t ref_len_of(const Ops& v) {
    uint64_t r = 0;
    for (uint32_t o : v) { const uint8_t op = o & 0xf; if (op == M || op == D || op == N || op == EQ || op == X) r += o >> 4; }
    return r;
}
inline uint64_t query_len_of(const Ops& v) {
    uint64_t r = 0;
    for (uint32_t o : v) { const uint8_t op = o & 0xf; if (op == M || op == I || op == S || op == H || op == EQ || op == X) r += o >> 4; }
    return r;
}
std::string ops_to_string(const Ops& v) {
    std::string s;
    for (uint32_t o : v) { s += std::to_string(o >> 4); s += "MIDNSHP=X"[o & 0xf]; }
    return s;
}
uint16_t reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return uint16_t(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return uint16_t(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return uint16_t(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return uint16_t(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return uint16_t(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

// ------------------------------------------------------------------------------------------------ contig model
struct SegPlan {
    int64_t ra = 0, rb = 0;        // reference interval
    bool inverted = false;         // aligned on the strand opposite to the contig's base orientation
    uint8_t mapq = 60;
    std::vector<uint8_t> head_forced, tail_forced;  // planted overlap bases (query == these, in ref orientation)
    uint32_t q_gap_after = 0;      // novel query bases after this segment (z-drop / gap junction)
    uint32_t q_overlap_next = 0;   // next segment re-uses this many of my last query bases (overlap junction)
    // built:
    Ops ops;                       // =/X/I/D in reference orientation
    int64_t qs = 0, qe = 0;        // query interval in the Q frame
};
struct Contig {
    uint32_t chrom = 0;
    bool reverse = false;          // base orientation of the contig relative to the reference
    bool unmapped = false;
    std::vector<SegPlan> segs;
    std::vector<uint8_t> C;        // true contig sequence, forward strand
    std::string name;
};
struct Plant { uint32_t chrom; int64_t pos; std::vector<uint8_t> bases; };

void revcomp_inplace(std::vector<uint8_t>& v) {
    std::reverse(v.begin(), v.end());
    for (auto& b : v) b = comp(b);
}

// Align a fresh haplotype sequence to ref[ra,rb): appends query bases to q, fills ops.
void build_segment(const std::vector<uint8_t>& ref, SegPlan& sp, const ptl_synth_params& P, Rng& rng, std::vector<uint8_t>& q) {
    Ops& ops = sp.ops;
    ops.clear();
    int64_t r = sp.ra;
    const int64_t head_end = sp.ra + int64_t(sp.head_forced.size());
    const int64_t tail_start = sp.rb - int64_t(sp.tail_forced.size());
    for (size_t k = 0; k < sp.head_forced.size(); ++k, ++r) {
        q.push_back(sp.head_forced[k]);
        push_op(ops, sp.head_forced[k] == ref[size_t(r)] ? EQ : X, 1);
    }
    const int64_t anchor = 30;
    const int64_t lo = std::max(head_end, sp.ra + anchor), hi = std::min(tail_start, sp.rb - anchor);
    const double sv_rate = P.sv_per_mb * 1e-6;
    const double total = P.contig_snv_rate + P.contig_indel_rate + sv_rate;
    while (r < tail_start) {
        int64_t next_ev = (total > 0) ? r + 1 + int64_t(rng.geometric(total)) : tail_start;
        if (next_ev < lo) next_ev = lo;
        if (next_ev >= hi || r >= hi) next_ev = tail_start;
        const int64_t copy_to = std::min(next_ev, tail_start);
        if (copy_to > r) {
            q.insert(q.end(), ref.begin() + r, ref.begin() + copy_to);
            push_op(ops, EQ, uint64_t(copy_to - r));
            r = copy_to;
        }
        if (r >= tail_start) break;
        const double u = rng.uni() * total;
        if (u < P.contig_snv_rate) {
            q.push_back(other_base(rng, ref[size_t(r)]));
            push_op(ops, X, 1);
            ++r;
        } else {
            uint64_t len;
            if (u < P.contig_snv_rate + P.contig_indel_rate) len = 1 + std::min<uint64_t>(rng.geometric(0.5), 9);
            else len = uint64_t(50.0 * std::pow(200.0, rng.uni()));  // 50 .. 10k, log-uniform
            if (rng.chance(0.5)) {  // insertion
                const bool hp = len <= 3 && rng.chance(0.5) && !q.empty();
                for (uint64_t k = 0; k < len; ++k) q.push_back(hp ? q[q.size() - 1] : rand_base(rng));
                push_op(ops, I, len);
            } else {  // deletion
                len = std::min<uint64_t>(len, uint64_t(std::max<int64_t>(hi - r - 1, 0)));
                push_op(ops, D, len);
                r += int64_t(len);
            }
            // at least one matching base after an indel so I and D are never adjacent in contig alignments
            if (r < tail_start) {
                q.push_back(ref[size_t(r)]);
                push_op(ops, EQ, 1);
                ++r;
            }
        }
    }
    for (size_t k = 0; k < sp.tail_forced.size(); ++k, ++r) {
        q.push_back(sp.tail_forced[k]);
        push_op(ops, sp.tail_forced[k] == ref[size_t(r)] ? EQ : X, 1);
    }
}

struct ContigRecordOut {
    uint32_t contig_id;
    uint16_t flag;
    int32_t tid;
    int64_t pos;
    uint8_t mapq;
    Ops cigar;
    std::string sa;
    bool has_sa = false;
    bool has_seq = false;
    std::vector<uint8_t> seq;
};

// ------------------------------------------------------------------------------------------------ read model
struct PartOut {
    std::vector<uint8_t> bases;  // read bases, contig-forward orientation
    Ops ops;                     // =/X/I/D vs the contig
    int64_t pos = 0;
};

// One read part sampled from contig C at `start` with target read length L (may come out shorter at the contig end).
void gen_part(const std::vector<uint8_t>& C, int64_t start, uint32_t L, const ptl_synth_params& P, Rng& rng, PartOut& out) {
    out.bases.clear();
    out.ops.clear();
    out.pos = start;
    const int64_t clen = int64_t(C.size());
    int64_t c = start;
    const double rate = P.read_sub_rate + P.read_indel_rate;
    const int64_t anchor = 20;
    int64_t prev_event_end = start + anchor;  // events may not be normalised to the left of this
    auto copy_eq = [&](int64_t to) {
        to = std::min<int64_t>(to, clen);
        const int64_t room = int64_t(L) - int64_t(out.bases.size());
        const int64_t n = std::min(to - c, room);
        if (n > 0) {
            out.bases.insert(out.bases.end(), C.begin() + c, C.begin() + c + n);
            push_op(out.ops, EQ, uint64_t(n));
            c += n;
        }
    };
    while (out.bases.size() < L && c < clen) {
        int64_t ev = (rate > 0) ? c + 1 + int64_t(rng.geometric(rate)) : clen;
        ev = std::max(ev, start + anchor);
        // need room for the event and a closing anchor
        if (ev + 40 >= clen || int64_t(out.bases.size()) + (ev - c) + 40 >= int64_t(L)) {
            copy_eq(clen);
            break;
        }
        const double u = rng.uni() * rate;
        if (u < P.read_sub_rate) {
            copy_eq(ev);
            out.bases.push_back(other_base(rng, C[size_t(c)]));
            push_op(out.ops, X, 1);
            ++c;
            prev_event_end = c;
            continue;
        }
        // indel, possibly steered into a homopolymer run ahead
        int64_t p = ev;
        bool is_ins = rng.chance(0.5);
        uint32_t len = 1 + uint32_t(std::min<uint64_t>(rng.geometric(0.6), 7));
        std::vector<uint8_t> ins;
        if (rng.chance(0.5)) {
            for (int64_t k = p; k < std::min<int64_t>(p + 30, clen - 45); ++k) {
                if (C[size_t(k)] == C[size_t(k + 1)] && C[size_t(k)] == C[size_t(k + 2)]) {
                    p = k + 1;
                    len = 1;
                    if (is_ins) ins.assign(1, C[size_t(k)]);
                    break;
                }
            }
        }
        if (is_ins && ins.empty()) for (uint32_t k = 0; k < len; ++k) ins.push_back(rand_base(rng));
        // left-normalise on the contig forward strand (what an aligner reports)
        if (is_ins) {
            while (p - 1 > prev_event_end && C[size_t(p - 1)] == ins.back()) {
                ins.insert(ins.begin(), ins.back());
                ins.pop_back();
                --p;
            }
        } else {
            while (p - 1 > prev_event_end && C[size_t(p - 1)] == C[size_t(p + len - 1)]) --p;
        }
        if (int64_t(out.bases.size()) + (p - c) + int64_t(len) + 40 >= int64_t(L) || p + int64_t(len) + 40 >= clen || p <= c) {
            copy_eq(clen);
            break;
        }
        copy_eq(p);
        const bool cluster = rng.chance(P.read_cluster_frac);
        // An adjacent I/D cluster is only emitted in a form an aligner could report: the first and the last inserted
        // base differ from the first / last deleted base (otherwise the match would simply extend into the cluster).
        auto cluster_ins = [&](uint32_t n, int64_t del_at, uint32_t d) {
            std::vector<uint8_t> v(n);
            for (auto& x : v) x = rand_base(rng);
            v.front() = other_base(rng, C[size_t(del_at)]);
            if (n == 1) { while (v[0] == C[size_t(del_at)] || v[0] == C[size_t(del_at + d - 1)]) v[0] = rand_base(rng); }
            else v.back() = other_base(rng, C[size_t(del_at + d - 1)]);
            return v;
        };
        if (is_ins) {
            if (cluster) {
                const uint32_t d = 1 + uint32_t(rng.below(3));
                ins = cluster_ins(uint32_t(ins.size()), c, d);
                out.bases.insert(out.bases.end(), ins.begin(), ins.end());
                push_op(out.ops, I, ins.size());
                push_op(out.ops, D, d);
                c += d;
            } else {
                out.bases.insert(out.bases.end(), ins.begin(), ins.end());
                push_op(out.ops, I, ins.size());
            }
        } else {
            push_op(out.ops, D, len);
            if (cluster) {
                const uint32_t n = 1 + uint32_t(rng.below(3));
                const std::vector<uint8_t> v = cluster_ins(n, c, len);
                out.bases.insert(out.bases.end(), v.begin(), v.end());
                push_op(out.ops, I, n);
            }
            c += len;
        }
        prev_event_end = c;
        // one matching base after the event keeps separate events separate
        copy_eq(c + 1);
    }
}

struct ReadOut {
    int32_t tid;
    int64_t pos;
    uint16_t flag;
    uint8_t mapq;
    uint16_t bin;
    std::vector<uint8_t> bases;  // stored orientation
    Ops cigar;
    std::string sa;
    bool has_sa = false;
};

struct ReadSeed { uint32_t contig; int64_t start; uint64_t id; };

}  // namespace

struct ptl_synth {
    ptl_synth_params P;
    void* (*alloc)(size_t) = nullptr;
    void (*dealloc)(void*) = nullptr;
    // read plan: every read of the set in BAM order (generated on demand)
    std::vector<ReadSeed> seeds;
    std::vector<uint32_t> plan_contig;
    std::vector<int64_t> plan_pos;
    // reference
    std::vector<std::vector<uint8_t>> chroms;
    std::vector<uint64_t> chrom_len;
    std::vector<const uint8_t*> chrom_ptr;
    std::vector<std::string> chrom_name_s;
    std::vector<const char*> chrom_name;
    // contigs
    std::vector<Contig> contigs;
    std::vector<uint64_t> contig_len;
    std::vector<const char*> contig_name;
    // contig records (flat)
    std::vector<uint32_t> cr_contig;
    std::vector<uint16_t> cr_flag;
    std::vector<int32_t> cr_tid;
    std::vector<int64_t> cr_pos;
    std::vector<uint8_t> cr_mapq;
    std::vector<uint64_t> cr_cigar_begin;
    std::vector<uint32_t> cr_cigar;
    std::vector<std::string> cr_sa_s;
    std::vector<const char*> cr_sa;
    std::vector<std::vector<uint8_t>> cr_seq_s;
    std::vector<const uint8_t*> cr_seq;
    // read records (flat)
    std::vector<int32_t> rr_tid;
    std::vector<int64_t> rr_pos;
    std::vector<uint16_t> rr_flag, rr_bin;
    std::vector<uint8_t> rr_mapq;
    std::vector<uint32_t> rr_seq_len;
    std::vector<uint64_t> rr_seq_off;
    uint8_t* rr_seq4 = nullptr;
    uint64_t rr_seq4_bytes = 0;
    std::vector<uint64_t> rr_cigar_begin;
    std::vector<uint32_t> rr_cigar;
    std::vector<std::string> rr_sa_s;
    std::vector<const char*> rr_sa;
    std::vector<uint64_t> rr_plan_index;  // index of each generated read in the planned BAM order (its name)
};

namespace {

template <class F>
void parallel_for(uint64_t n, uint32_t n_threads, F f) {
    if (n_threads <= 1 || n <= 1) {
        for (uint64_t i = 0; i < n; ++i) f(i);
        return;
    }
    std::atomic<uint64_t> next{0};
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; ++t)
        th.emplace_back([&]() {
            for (;;) {
                const uint64_t i = next.fetch_add(1);
                if (i >= n) break;
                f(i);
            }
        });
    for (auto& t : th) t.join();
}

void make_reference(ptl_synth& s, uint32_t nt) {
    const auto& P = s.P;
    s.chroms.resize(P.n_chrom);
    for (uint32_t c = 0; c < P.n_chrom; ++c) s.chroms[c].resize(P.chrom_len);
    const uint64_t block = 1 << 20;
    const uint64_t blocks_per = (P.chrom_len + block - 1) / block;
    parallel_for(uint64_t(P.n_chrom) * blocks_per, nt, [&](uint64_t i) {
        const uint32_t c = uint32_t(i / blocks_per);
        const uint64_t b0 = (i % blocks_per) * block, b1 = std::min<uint64_t>(b0 + block, P.chrom_len);
        Rng rng(P.seed, 0x100000000ull + i);
        auto& v = s.chroms[c];
        for (uint64_t k = b0; k < b1;) {
            uint64_t w = rng.next();
            for (int j = 0; j < 32 && k < b1; ++j, ++k, w >>= 2) v[k] = uint8_t(kBases[w & 3]);
        }
        // plant homopolymers / dinucleotide repeats every ~1.5 kb so indel shifting has work to do
        for (uint64_t k = b0 + rng.below(1500); k + 64 < b1; k += 200 + rng.below(2600)) {
            if (rng.chance(0.6)) {
                const uint64_t n = 4 + rng.below(14);
                const uint8_t b = rand_base(rng);
                for (uint64_t j = 0; j < n; ++j) v[k + j] = b;
            } else {
                const uint64_t n = 3 + rng.below(10);
                const uint8_t b1c = rand_base(rng), b2c = other_base(rng, b1c);
                for (uint64_t j = 0; j < n; ++j) { v[k + 2 * j] = b1c; v[k + 2 * j + 1] = b2c; }
            }
        }
    });
}

void plan_contigs(ptl_synth& s, std::vector<Plant>& plants) {
    const auto& P = s.P;
    uint32_t id = 0;
    for (uint32_t c = 0; c < P.n_chrom; ++c) {
        for (uint32_t h = 0; h < P.haplotypes; ++h) {
            Rng rng(P.seed, 0x200000000ull + c * 64 + h);
            // tile [0, L) with contigs_per_chrom contigs, random breakpoints, small gaps between contigs
            const int64_t L = int64_t(P.chrom_len);
            std::vector<int64_t> cuts{0, L};
            for (uint32_t k = 1; k < P.contigs_per_chrom; ++k) cuts.push_back(rng.range(L / 20, L - L / 20));
            std::sort(cuts.begin(), cuts.end());
            for (size_t k = 0; k + 1 < cuts.size(); ++k) {
                int64_t a = cuts[k] + (k ? rng.range(0, 2000) : 0), b = cuts[k + 1] - rng.range(0, 2000);
                if (b - a < 20000) continue;
                Contig ct;
                ct.chrom = c;
                ct.reverse = rng.chance(P.rev_contig_frac);
                ct.unmapped = rng.chance(P.unmapped_contig_frac);
                ct.name = "ctg" + std::to_string(id++);
                // junctions
                std::vector<int64_t> jpos;
                const double expect = P.junction_per_mb * double(b - a) * 1e-6;
                uint32_t nj = uint32_t(expect);
                if (rng.chance(expect - nj)) ++nj;
                for (uint32_t j = 0; j < nj; ++j) jpos.push_back(rng.range(a + 8000, b - 8000));
                std::sort(jpos.begin(), jpos.end());
                SegPlan cur;
                cur.ra = a;
                auto seg_mapq = [&]() { return uint8_t(rng.chance(0.2) ? rng.range(1, 59) : 60); };
                cur.mapq = seg_mapq();
                int64_t last = a;
                for (int64_t r : jpos) {
                    if (r - last < 6000 || b - r < 70000) continue;
                    const double u = rng.uni();
                    if (u < 0.35 || u >= 0.70) {  // z-drop (joinable when mapq equal) or larger gap
                        const bool zdrop = u < 0.35;
                        cur.rb = r;
                        cur.q_gap_after = uint32_t(zdrop ? rng.range(0, 300) : rng.range(0, 5000));
                        const int64_t r_gap = zdrop ? rng.range(0, 1000) : rng.range(1001, 20000);
                        ct.segs.push_back(cur);
                        SegPlan nx;
                        nx.ra = r + r_gap;
                        nx.mapq = zdrop && rng.chance(0.8) ? cur.mapq : seg_mapq();
                        cur = nx;
                        last = nx.ra;
                    } else if (u < 0.50) {  // inversion
                        const int64_t inv = rng.range(2000, 40000);
                        cur.rb = r;
                        ct.segs.push_back(cur);
                        SegPlan bseg;
                        bseg.ra = r;
                        bseg.rb = r + inv;
                        bseg.inverted = true;
                        bseg.mapq = seg_mapq();
                        ct.segs.push_back(bseg);
                        SegPlan nx;
                        nx.ra = r + inv;
                        nx.mapq = seg_mapq();
                        cur = nx;
                        last = nx.ra;
                    } else {  // overlapping junction (repeated match): same query bases aligned by both segments
                        const uint32_t ov = uint32_t(rng.range(50, 500));
                        const int64_t delta = rng.range(0, 5000);
                        std::vector<uint8_t> qov(ov);
                        for (auto& x : qov) x = rand_base(rng);
                        const int k1 = int(rng.below(3)), k2 = int(rng.below(3));
                        auto mutated = [&](int k) {
                            std::vector<uint8_t> v = qov;
                            for (int j = 0; j < k; ++j) { const size_t p = size_t(rng.below(ov)); v[p] = other_base(rng, v[p]); }
                            return v;
                        };
                        plants.push_back({c, r - int64_t(ov), mutated(k1)});
                        plants.push_back({c, r + delta, mutated(k2)});
                        cur.rb = r;
                        cur.tail_forced = qov;
                        cur.q_overlap_next = ov;
                        ct.segs.push_back(cur);
                        SegPlan nx;
                        nx.ra = r + delta;
                        nx.head_forced = qov;
                        nx.mapq = rng.chance(0.5) ? cur.mapq : seg_mapq();
                        cur = nx;
                        last = nx.ra + ov;
                    }
                }
                cur.rb = b;
                ct.segs.push_back(cur);
                s.contigs.push_back(std::move(ct));
            }
        }
    }
}

void build_contig(ptl_synth& s, uint32_t ci, std::vector<ContigRecordOut>& recs) {
    const auto& P = s.P;
    Contig& ct = s.contigs[ci];
    const auto& ref = s.chroms[ct.chrom];
    Rng rng(P.seed, 0x300000000ull + ci);
    std::vector<uint8_t> Q;  // contig in its base (reference-parallel) orientation
    for (size_t k = 0; k < ct.segs.size(); ++k) {
        SegPlan& sp = ct.segs[k];
        if (k > 0 && ct.segs[k - 1].q_overlap_next) {
            // re-use the previous segment's last `ov` query bases: build into a scratch, drop the duplicated head
            std::vector<uint8_t> tmp;
            build_segment(ref, sp, P, rng, tmp);
            const uint32_t ov = ct.segs[k - 1].q_overlap_next;
            sp.qs = int64_t(Q.size()) - ov;
            Q.insert(Q.end(), tmp.begin() + ov, tmp.end());
            sp.qe = int64_t(Q.size());
        } else if (sp.inverted) {
            std::vector<uint8_t> tmp;
            build_segment(ref, sp, P, rng, tmp);
            revcomp_inplace(tmp);
            sp.qs = int64_t(Q.size());
            Q.insert(Q.end(), tmp.begin(), tmp.end());
            sp.qe = int64_t(Q.size());
        } else {
            sp.qs = int64_t(Q.size());
            build_segment(ref, sp, P, rng, Q);
            sp.qe = int64_t(Q.size());
        }
        for (uint32_t g = 0; g < sp.q_gap_after; ++g) Q.push_back(rand_base(rng));
    }
    const int64_t qlen = int64_t(Q.size());
    ct.C = Q;
    if (ct.reverse) revcomp_inplace(ct.C);
    if (ct.unmapped) return;

    // records: segment strand = contig.reverse XOR seg.inverted
    size_t primary = 0;
    for (size_t k = 1; k < ct.segs.size(); ++k)
        if (ct.segs[k].qe - ct.segs[k].qs > ct.segs[primary].qe - ct.segs[primary].qs) primary = k;
    struct Rec { bool rev; int64_t lead, trail; };
    std::vector<Rec> meta(ct.segs.size());
    for (size_t k = 0; k < ct.segs.size(); ++k) {
        const SegPlan& sp = ct.segs[k];
        Rec m;
        m.rev = ct.reverse ^ sp.inverted;
        // CIGAR is written in the segment's own reference-forward orientation:
        // non-inverted segments read Q left to right; inverted ones read revcomp(Q).
        if (!sp.inverted) { m.lead = sp.qs; m.trail = qlen - sp.qe; }
        else { m.lead = qlen - sp.qe; m.trail = sp.qs; }
        meta[k] = m;
    }
    std::string sa;
    for (size_t k = 0; k < ct.segs.size(); ++k) {
        if (k == primary) continue;
        const SegPlan& sp = ct.segs[k];
        const uint64_t ql = uint64_t(sp.qe - sp.qs), rl = uint64_t(sp.rb - sp.ra);
        // minimap2-style approximate SA CIGAR: clips + matched length + net indel
        Ops approx;
        push_op(approx, S, uint64_t(meta[k].lead));
        push_op(approx, M, std::min(ql, rl));
        if (rl > ql) push_op(approx, D, rl - ql);
        else if (ql > rl) push_op(approx, I, ql - rl);
        push_op(approx, S, uint64_t(meta[k].trail));
        sa += s.chrom_name_s[ct.chrom] + "," + std::to_string(sp.ra + 1) + "," + (meta[k].rev ? "-" : "+") + "," +
              ops_to_string(approx) + "," + std::to_string(unsigned(sp.mapq)) + ",0;";
    }
    for (size_t k = 0; k < ct.segs.size(); ++k) {
        const SegPlan& sp = ct.segs[k];
        ContigRecordOut r;
        r.contig_id = ci;
        r.tid = int32_t(ct.chrom);
        r.pos = sp.ra;
        r.mapq = sp.mapq;
        const bool is_primary = (k == primary);
        r.flag = uint16_t((meta[k].rev ? 0x10 : 0) | (is_primary ? 0 : 0x800));
        const uint8_t clip = is_primary ? S : H;
        push_op(r.cigar, clip, uint64_t(meta[k].lead));
        for (uint32_t o : sp.ops) r.cigar.push_back(o);  // never merges with a clip
        push_op(r.cigar, clip, uint64_t(meta[k].trail));
        if (is_primary) {
            r.has_sa = !sa.empty();
            r.sa = sa;
            r.has_seq = true;
            r.seq = ct.C;                     // stored bases are in the record's alignment orientation
            if (meta[k].rev) revcomp_inplace(r.seq);
        }
        recs.push_back(std::move(r));
    }
}

void gen_read(const ptl_synth& s, const ReadSeed& seed, ReadOut& out, PartOut& p0, PartOut& p1) {
    const auto& P = s.P;
    Rng rng(P.seed, 0x400000000ull + seed.id);
    const Contig& ct = s.contigs[seed.contig];
    double lf = P.read_len_mean + P.read_len_sd * rng.normal();
    uint32_t L = uint32_t(std::min<double>(std::max<double>(lf, P.read_len_min), P.read_len_max));
    L = uint32_t(std::min<uint64_t>(L, ct.C.size() - uint64_t(seed.start)));
    gen_part(ct.C, seed.start, L, P, rng, p0);
    const bool rev = rng.chance(0.5);
    out.tid = int32_t(seed.contig);
    out.pos = p0.pos;
    out.flag = rev ? 0x10 : 0;
    out.mapq = uint8_t(rng.chance(0.7) ? 60 : rng.range(0, 59));
    out.bases.clear();
    out.cigar.clear();
    out.has_sa = false;
    out.sa.clear();
    const bool chimeric = rng.chance(P.read_sa_frac) && p0.bases.size() > 2000;
    if (!chimeric) {
        uint32_t lead = 0, trail = 0;
        if (rng.chance(P.read_clip_frac)) {
            if (rng.chance(0.5)) lead = uint32_t(rng.range(1, 200)); else trail = uint32_t(rng.range(1, 200));
        }
        push_op(out.cigar, S, lead);
        for (uint32_t k = 0; k < lead; ++k) out.bases.push_back(rand_base(rng));
        out.bases.insert(out.bases.end(), p0.bases.begin(), p0.bases.end());
        for (uint32_t o : p0.ops) out.cigar.push_back(o);
        push_op(out.cigar, S, trail);
        for (uint32_t k = 0; k < trail; ++k) out.bases.push_back(rand_base(rng));
    } else {
        // second part from another locus (any contig), same or opposite strand relative to the primary record
        const uint32_t c2 = uint32_t(rng.below(s.contigs.size()));
        const Contig& ct2 = s.contigs[c2];
        const uint32_t L2 = uint32_t(std::min<uint64_t>(uint64_t(rng.range(1500, 9000)), ct2.C.size() / 2));
        const int64_t start2 = rng.range(0, int64_t(ct2.C.size()) - int64_t(L2) - 1);
        gen_part(ct2.C, start2, L2, P, rng, p1);
        const bool opposite = rng.chance(0.5);
        const bool primary_first = rng.chance(0.5);
        std::vector<uint8_t> b1 = p1.bases;  // part-2 bases as they appear in the stored (primary) orientation
        if (opposite) revcomp_inplace(b1);
        const uint64_t n0 = p0.bases.size(), n1 = b1.size();
        Ops sa_cigar;
        if (primary_first) {
            out.bases = p0.bases;
            out.bases.insert(out.bases.end(), b1.begin(), b1.end());
            for (uint32_t o : p0.ops) out.cigar.push_back(o);
            push_op(out.cigar, S, n1);
            // SA segment: in its own orientation the clipped primary part lies before (same strand) or after it
            if (!opposite) { push_op(sa_cigar, S, n0); for (uint32_t o : p1.ops) sa_cigar.push_back(o); }
            else { for (uint32_t o : p1.ops) sa_cigar.push_back(o); push_op(sa_cigar, S, n0); }
        } else {
            out.bases = b1;
            out.bases.insert(out.bases.end(), p0.bases.begin(), p0.bases.end());
            push_op(out.cigar, S, n1);
            for (uint32_t o : p0.ops) out.cigar.push_back(o);
            if (!opposite) { for (uint32_t o : p1.ops) sa_cigar.push_back(o); push_op(sa_cigar, S, n0); }
            else { push_op(sa_cigar, S, n0); for (uint32_t o : p1.ops) sa_cigar.push_back(o); }
        }
        const bool sa_rev = rev ^ opposite;
        out.has_sa = true;
        out.sa = s.contigs[c2].name + "," + std::to_string(p1.pos + 1) + "," + (sa_rev ? "-" : "+") + "," +
                 ops_to_string(sa_cigar) + "," + std::to_string(unsigned(rng.chance(0.7) ? 60 : rng.range(0, 59))) + ",0;";
    }
    const int64_t end = out.pos + int64_t(ref_len_of(out.cigar));
    out.bin = reg2bin(out.pos, std::max(end, out.pos + 1));
}

// The whole read set in coordinate-sorted BAM order: sample loci, sort (a read's record position is its sampled start).
void plan_reads(ptl_synth& s) {
    const auto& P = s.P;
    const uint64_t n = P.n_reads;
    std::vector<double> cum(s.contigs.size());
    double tot = 0;
    for (size_t i = 0; i < s.contigs.size(); ++i) { tot += double(s.contigs[i].C.size()); cum[i] = tot; }
    std::vector<ReadSeed>& seeds = s.seeds;
    seeds.resize(n);
    {
        Rng rng(P.seed, 0x500000000ull);
        for (uint64_t i = 0; i < n; ++i) {
            const double u = rng.uni() * tot;
            const uint32_t c = uint32_t(std::lower_bound(cum.begin(), cum.end(), u) - cum.begin());
            const int64_t clen = int64_t(s.contigs[c].C.size());
            const int64_t max_start = std::max<int64_t>(clen - int64_t(P.read_len_min), 0);
            seeds[i] = ReadSeed{c, rng.range(0, max_start), i};
        }
    }
    std::sort(seeds.begin(), seeds.end(), [](const ReadSeed& a, const ReadSeed& b) {
        return a.contig != b.contig ? a.contig < b.contig : (a.start != b.start ? a.start < b.start : a.id < b.id);
    });
    s.plan_contig.resize(n);
    s.plan_pos.resize(n);
    for (uint64_t i = 0; i < n; ++i) { s.plan_contig[i] = seeds[i].contig; s.plan_pos[i] = seeds[i].start; }
}

// Generate the reads of the given ranges of the plan (concatenated) into the flat read-record arrays.
void make_reads(ptl_synth& s, uint32_t nt, const std::vector<std::pair<uint64_t, uint64_t>>& ranges) {
    std::vector<uint64_t> pick;  // plan index of every generated read
    for (const auto& r : ranges)
        for (uint64_t i = r.first; i < r.first + r.second; ++i) pick.push_back(i);
    const uint64_t n = pick.size();
    s.rr_plan_index = pick;
    const std::vector<ReadSeed>& seeds = s.seeds;
    // generate in chunks
    const uint64_t chunk = 2048;
    const uint64_t n_chunks = (n + chunk - 1) / chunk;
    struct Chunk {
        std::vector<uint8_t> seq4;
        std::vector<uint32_t> cigar;
        std::vector<uint32_t> seq_len, n_ops;
        std::vector<uint64_t> seq_bytes;
    };
    std::vector<Chunk> chunks(n_chunks);
    if (s.rr_seq4 && s.dealloc) s.dealloc(s.rr_seq4);
    s.rr_seq4 = nullptr;
    s.rr_tid.assign(n, 0); s.rr_pos.assign(n, 0); s.rr_flag.assign(n, 0); s.rr_bin.assign(n, 0); s.rr_mapq.assign(n, 0);
    s.rr_seq_len.assign(n, 0); s.rr_seq_off.assign(n, 0); s.rr_cigar_begin.assign(n + 1, 0); s.rr_sa_s.assign(n, std::string()); s.rr_sa.assign(n, nullptr);
    parallel_for(n_chunks, nt, [&](uint64_t ci) {
        Chunk& ch = chunks[ci];
        ReadOut ro;
        PartOut p0, p1;
        const uint64_t r0 = ci * chunk, r1 = std::min(n, r0 + chunk);
        for (uint64_t r = r0; r < r1; ++r) {
            gen_read(s, seeds[pick[r]], ro, p0, p1);
            s.rr_tid[r] = ro.tid; s.rr_pos[r] = ro.pos; s.rr_flag[r] = ro.flag; s.rr_bin[r] = ro.bin; s.rr_mapq[r] = ro.mapq;
            const uint32_t L = uint32_t(ro.bases.size());
            s.rr_seq_len[r] = L;
            ch.seq_len.push_back(L);
            ch.n_ops.push_back(uint32_t(ro.cigar.size()));
            const size_t nb = (L + 1) / 2;
            const size_t at = ch.seq4.size();
            ch.seq4.resize(at + nb);
            uint8_t* dst = ch.seq4.data() + at;
            for (uint32_t k = 0; k + 1 < L; k += 2) dst[k >> 1] = uint8_t((code4(ro.bases[k]) << 4) | code4(ro.bases[k + 1]));
            if (L & 1) dst[L >> 1] = uint8_t(code4(ro.bases[L - 1]) << 4);
            ch.cigar.insert(ch.cigar.end(), ro.cigar.begin(), ro.cigar.end());
            if (ro.has_sa) s.rr_sa_s[r] = ro.sa;
        }
    });
    // concatenate
    std::vector<uint64_t> seq_base(n_chunks + 1, 0), cig_base(n_chunks + 1, 0);
    for (uint64_t c = 0; c < n_chunks; ++c) {
        seq_base[c + 1] = seq_base[c] + chunks[c].seq4.size();
        cig_base[c + 1] = cig_base[c] + chunks[c].cigar.size();
    }
    s.rr_seq4_bytes = seq_base[n_chunks];
    s.rr_seq4 = static_cast<uint8_t*>(s.alloc(std::max<uint64_t>(s.rr_seq4_bytes, 16)));
    s.rr_cigar.resize(cig_base[n_chunks]);
    parallel_for(n_chunks, nt, [&](uint64_t c) {
        Chunk& ch = chunks[c];
        if (!ch.seq4.empty()) std::memcpy(s.rr_seq4 + seq_base[c], ch.seq4.data(), ch.seq4.size());
        if (!ch.cigar.empty()) std::memcpy(s.rr_cigar.data() + cig_base[c], ch.cigar.data(), ch.cigar.size() * 4);
        uint64_t so = seq_base[c], co = cig_base[c];
        const uint64_t r0 = c * chunk;
        for (size_t k = 0; k < ch.seq_len.size(); ++k) {
            s.rr_seq_off[r0 + k] = so;
            s.rr_cigar_begin[r0 + k] = co;
            so += (ch.seq_len[k] + 1) / 2;
            co += ch.n_ops[k];
        }
        std::vector<uint8_t>().swap(ch.seq4);
        std::vector<uint32_t>().swap(ch.cigar);
    });
    s.rr_cigar_begin[n] = cig_base[n_chunks];
    for (uint64_t r = 0; r < n; ++r) if (!s.rr_sa_s[r].empty()) s.rr_sa[r] = s.rr_sa_s[r].c_str();
}

uint32_t thread_count(const ptl_synth_params& P) { return P.n_threads ? P.n_threads : std::max(1u, std::thread::hardware_concurrency()); }

}  // namespace

extern "C" {

void ptl_synth_default_params(ptl_synth_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->seed = 1001;
    p->n_chrom = 1;
    p->chrom_len = 1000000;
    p->haplotypes = 2;
    p->contigs_per_chrom = 1;
    p->rev_contig_frac = 0.5;
    p->unmapped_contig_frac = 0.0;
    p->contig_snv_rate = 1e-3;
    p->contig_indel_rate = 2e-4;
    p->sv_per_mb = 3.0;
    p->junction_per_mb = 2.0;
    p->n_reads = 20000;
    p->read_len_mean = 15000;
    p->read_len_sd = 2000;
    p->read_len_min = 5000;
    p->read_len_max = 30000;
    p->read_sub_rate = 2e-4;
    p->read_indel_rate = 5e-4;
    p->read_cluster_frac = 0.05;
    p->read_clip_frac = 0.01;
    p->read_sa_frac = 0.02;
    p->n_threads = 0;
}

ptl_synth* ptl_synth_create_into(const ptl_synth_params* p, void* (*alloc)(size_t), void (*dealloc)(void*)) {
    if (!p || !p->n_chrom || p->chrom_len < 200000 || !p->haplotypes || !p->contigs_per_chrom) return nullptr;
    auto* s = new ptl_synth();
    s->P = *p;
    s->alloc = alloc;
    s->dealloc = dealloc;
    uint32_t nt = p->n_threads ? p->n_threads : std::max(1u, std::thread::hardware_concurrency());
    for (uint32_t c = 0; c < p->n_chrom; ++c) s->chrom_name_s.push_back("chr" + std::to_string(c + 1));
    make_reference(*s, nt);
    std::vector<Plant> plants;
    plan_contigs(*s, plants);
    for (const auto& pl : plants) {
        auto& v = s->chroms[pl.chrom];
        if (pl.pos < 0 || size_t(pl.pos) + pl.bases.size() > v.size()) continue;
        std::memcpy(v.data() + pl.pos, pl.bases.data(), pl.bases.size());
    }
    std::vector<std::vector<ContigRecordOut>> per(s->contigs.size());
    parallel_for(s->contigs.size(), nt, [&](uint64_t ci) { build_contig(*s, uint32_t(ci), per[ci]); });
    // flatten contig records in "BAM order": by (tid, pos)
    std::vector<ContigRecordOut*> all;
    for (auto& v : per) for (auto& r : v) all.push_back(&r);
    std::stable_sort(all.begin(), all.end(), [](const ContigRecordOut* a, const ContigRecordOut* b) {
        return a->tid != b->tid ? a->tid < b->tid : a->pos < b->pos;
    });
    s->cr_cigar_begin.push_back(0);
    s->cr_sa_s.reserve(all.size());
    s->cr_seq_s.reserve(all.size());
    for (auto* r : all) {
        s->cr_contig.push_back(r->contig_id);
        s->cr_flag.push_back(r->flag);
        s->cr_tid.push_back(r->tid);
        s->cr_pos.push_back(r->pos);
        s->cr_mapq.push_back(r->mapq);
        s->cr_cigar.insert(s->cr_cigar.end(), r->cigar.begin(), r->cigar.end());
        s->cr_cigar_begin.push_back(s->cr_cigar.size());
        s->cr_sa_s.push_back(r->has_sa ? r->sa : std::string());
        s->cr_seq_s.push_back(r->has_seq ? std::move(r->seq) : std::vector<uint8_t>());
    }
    for (size_t i = 0; i < all.size(); ++i) {
        s->cr_sa.push_back(all[i]->has_sa ? s->cr_sa_s[i].c_str() : nullptr);
        s->cr_seq.push_back(all[i]->has_seq ? s->cr_seq_s[i].data() : nullptr);
    }
    for (auto& c : s->contigs) {
        s->contig_len.push_back(c.C.size());
        s->contig_name.push_back(c.name.c_str());
    }
    for (uint32_t c = 0; c < p->n_chrom; ++c) {
        s->chrom_len.push_back(s->chroms[c].size());
        s->chrom_ptr.push_back(s->chroms[c].data());
        s->chrom_name.push_back(s->chrom_name_s[c].c_str());
    }
    plan_reads(*s);
    if (!p->defer_reads) make_reads(*s, nt, {{0, p->n_reads}});
    return s;
}
ptl_synth* ptl_synth_create(const ptl_synth_params* p) { return ptl_synth_create_into(p, std::malloc, std::free); }
void ptl_synth_destroy(ptl_synth* s) {
    if (!s) return;
    if (s->rr_seq4 && s->dealloc) s->dealloc(s->rr_seq4);
    delete s;
}
uint32_t ptl_synth_n_chrom(const ptl_synth* s) { return uint32_t(s->chroms.size()); }
const uint64_t* ptl_synth_chrom_len(const ptl_synth* s) { return s->chrom_len.data(); }
const uint8_t* const* ptl_synth_chrom_seq(const ptl_synth* s) { return s->chrom_ptr.data(); }
const char* const* ptl_synth_chrom_names(const ptl_synth* s) { return s->chrom_name.data(); }
uint32_t ptl_synth_n_contigs(const ptl_synth* s) { return uint32_t(s->contigs.size()); }
const char* const* ptl_synth_contig_names(const ptl_synth* s) { return s->contig_name.data(); }
void ptl_synth_contig_records(const ptl_synth* s, ptl_contig_records* o) {
    o->n_records = uint32_t(s->cr_contig.size());
    o->contig_id = s->cr_contig.data();
    o->flag = s->cr_flag.data();
    o->tid = s->cr_tid.data();
    o->pos = s->cr_pos.data();
    o->mapq = s->cr_mapq.data();
    o->cigar_begin = s->cr_cigar_begin.data();
    o->cigar = s->cr_cigar.data();
    o->sa_tag = s->cr_sa.data();
    o->seq = s->cr_seq.data();
    o->n_contigs = uint32_t(s->contigs.size());
    o->contig_len = s->contig_len.data();
    o->contig_names = s->contig_name.data();
    o->n_ref_chrom = uint32_t(s->chroms.size());
    o->ref_chrom_names = s->chrom_name.data();
}
uint64_t ptl_synth_n_planned(const ptl_synth* s) { return s->seeds.size(); }
const uint32_t* ptl_synth_plan_contig(const ptl_synth* s) { return s->plan_contig.data(); }
const int64_t* ptl_synth_plan_pos(const ptl_synth* s) { return s->plan_pos.data(); }
const uint64_t* ptl_synth_contig_len(const ptl_synth* s) { return s->contig_len.data(); }
int ptl_synth_generate_reads(ptl_synth* s, uint32_t n_ranges, const uint64_t* first, const uint64_t* count) {
    if (!s || (n_ranges && (!first || !count))) return 1;
    std::vector<std::pair<uint64_t, uint64_t>> ranges;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_ranges; ++i) {
        if (first[i] > s->seeds.size() || count[i] > s->seeds.size() - first[i]) return 1;
        ranges.emplace_back(first[i], count[i]);
        total += count[i];
    }
    if (total > 0xffffffffull) return 1;
    make_reads(*s, thread_count(s->P), ranges);
    return 0;
}
// ---- the same data as BAM byte streams (uncompressed; the caller frames them as BGZF and indexes them), so that the
//      file-level path (BGZF inflate, BAM decode, index fetch, the command-line tool) runs on exactly the records the
//      in-memory path gets, and so that the real portello can be run on the same files elsewhere.
namespace {
struct ByteSink {
    std::vector<uint8_t> v;
    void u8(uint8_t x) { v.push_back(x); }
    void u16(uint16_t x) { v.push_back(uint8_t(x)); v.push_back(uint8_t(x >> 8)); }
    void u32(uint32_t x) { for (int b = 0; b < 4; ++b) v.push_back(uint8_t(x >> (8 * b))); }
    void bytes(const void* p, size_t n) { const uint8_t* q = static_cast<const uint8_t*>(p); v.insert(v.end(), q, q + n); }
    void str(const std::string& t) { bytes(t.data(), t.size()); }
};
void bam_header(ByteSink& o, const std::string& text, const std::vector<std::string>& names, const std::vector<uint64_t>& lens) {
    o.bytes("BAM\1", 4);
    o.u32(uint32_t(text.size()));
    o.str(text);
    o.u32(uint32_t(names.size()));
    for (size_t i = 0; i < names.size(); ++i) {
        o.u32(uint32_t(names[i].size() + 1));
        o.str(names[i]);
        o.u8(0);
        o.u32(uint32_t(lens[i]));
    }
}
// one record; `seq4` = packed bases (l_seq of them), `qual` = l_seq bytes or nullptr (0xff), `aux` = raw aux block.
// A CIGAR with more than 65535 ops is written the way htslib writes it: a `<l_seq>S<ref_len>N` placeholder + CG:B,I.
void bam_record(ByteSink& o, int32_t tid, int64_t pos, uint8_t mapq, uint16_t flag, const std::string& name, const uint32_t* cigar, size_t n_cigar,
                const uint8_t* seq4, uint32_t l_seq, const uint8_t* qual, const std::vector<uint8_t>& aux) {
    uint64_t ref_len = 0;
    for (size_t i = 0; i < n_cigar; ++i) if ((0x18dU >> (cigar[i] & 0xf)) & 1u) ref_len += cigar[i] >> 4;
    const bool big = n_cigar > 65535;
    const size_t n_cig_field = big ? 2 : n_cigar;
    const size_t size = 32 + name.size() + 1 + 4 * n_cig_field + (size_t(l_seq) + 1) / 2 + l_seq + aux.size() + (big ? 8 + 4 * n_cigar : 0);
    o.u32(uint32_t(size));
    o.u32(uint32_t(tid));
    o.u32(uint32_t(int32_t(pos)));
    o.u8(uint8_t(name.size() + 1));
    o.u8(mapq);
    o.u16((flag & 0x4) ? 4680 : reg2bin(pos, pos + int64_t(std::max<uint64_t>(ref_len, 1))));
    o.u16(uint16_t(n_cig_field));
    o.u16(flag);
    o.u32(l_seq);
    o.u32(uint32_t(-1));
    o.u32(uint32_t(-1));
    o.u32(0);
    o.str(name);
    o.u8(0);
    if (big) { o.u32((l_seq << 4) | S); o.u32(uint32_t(ref_len << 4) | 3u); }
    else for (size_t i = 0; i < n_cigar; ++i) o.u32(cigar[i]);
    o.bytes(seq4, (size_t(l_seq) + 1) / 2);
    if (qual) o.bytes(qual, l_seq);
    else o.v.insert(o.v.end(), l_seq, uint8_t(0xff));
    o.bytes(aux.data(), aux.size());
    if (big) {
        o.bytes("CGBI", 4);
        o.u32(uint32_t(n_cigar));
        for (size_t i = 0; i < n_cigar; ++i) o.u32(cigar[i]);
    }
}
void aux_z(std::vector<uint8_t>& a, const char* tag, const std::string& v) {
    a.push_back(uint8_t(tag[0])); a.push_back(uint8_t(tag[1])); a.push_back('Z');
    a.insert(a.end(), v.begin(), v.end());
    a.push_back(0);
}
void aux_i(std::vector<uint8_t>& a, const char* tag, int32_t v) {
    a.push_back(uint8_t(tag[0])); a.push_back(uint8_t(tag[1])); a.push_back('i');
    for (int b = 0; b < 4; ++b) a.push_back(uint8_t(uint32_t(v) >> (8 * b)));
}
}  // namespace

uint8_t* ptl_synth_bam_stream(const ptl_synth* s, int which, uint32_t n_unmapped, uint64_t* n_bytes) {
    if (!s || !n_bytes) return nullptr;
    ByteSink o;
    if (which == 0) {  // contig -> reference (minimap2 --eqx style: primary soft-clipped with SA + bases, supplementary hard-clipped)
        std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
        std::vector<std::string> names;
        for (size_t c = 0; c < s->chroms.size(); ++c) {
            names.push_back(s->chrom_name_s[c]);
            text += "@SQ\tSN:" + s->chrom_name_s[c] + "\tLN:" + std::to_string(s->chrom_len[c]) + "\n";
        }
        bam_header(o, text, names, s->chrom_len);
        for (size_t i = 0; i < s->cr_contig.size(); ++i) {
            std::vector<uint8_t> aux, seq4;
            aux_i(aux, "NM", 0);
            if (s->cr_sa[i]) aux_z(aux, "SA", s->cr_sa[i]);
            uint32_t l_seq = 0;
            if (s->cr_seq[i]) {
                const std::vector<uint8_t>& q = s->cr_seq_s[i];
                l_seq = uint32_t(q.size());
                seq4.assign((q.size() + 1) / 2, 0);
                for (size_t k = 0; k < q.size(); ++k) seq4[k >> 1] |= uint8_t(code4(q[k]) << ((k & 1) ? 0 : 4));
            }
            bam_record(o, s->cr_tid[i], s->cr_pos[i], s->cr_mapq[i], s->cr_flag[i], s->contigs[s->cr_contig[i]].name,
                       s->cr_cigar.data() + s->cr_cigar_begin[i], size_t(s->cr_cigar_begin[i + 1] - s->cr_cigar_begin[i]), seq4.data(), l_seq, nullptr, aux);
        }
    } else {  // read -> contig (pbmm2 style), the reads currently generated, then n_unmapped unplaced reads
        std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
        std::vector<std::string> names;
        for (size_t c = 0; c < s->contigs.size(); ++c) {
            names.push_back(s->contigs[c].name);
            text += "@SQ\tSN:" + s->contigs[c].name + "\tLN:" + std::to_string(s->contig_len[c]) + "\n";
        }
        bam_header(o, text, names, s->contig_len);
        const size_t n = s->rr_tid.size();
        std::vector<uint8_t> qual;
        for (size_t r = 0; r < n; ++r) {
            const uint32_t L = s->rr_seq_len[r];
            Rng rng(s->P.seed, 0x600000000ull + s->rr_plan_index[r]);
            qual.resize(L);
            for (uint32_t k = 0; k < L;) {
                uint64_t w = rng.next();
                for (int j = 0; j < 10 && k < L; ++j, ++k, w >>= 6) qual[k] = uint8_t((w & 63) % 94);
            }
            std::vector<uint8_t> aux;
            aux_i(aux, "NM", int32_t(rng.below(40)));
            aux_i(aux, "np", int32_t(4 + rng.below(20)));
            if (s->rr_sa[r]) aux_z(aux, "SA", s->rr_sa[r]);
            aux_z(aux, "RG", "synth");
            bam_record(o, s->rr_tid[r], s->rr_pos[r], s->rr_mapq[r], s->rr_flag[r], "synth/" + std::to_string(s->rr_plan_index[r]) + "/ccs",
                       s->rr_cigar.data() + s->rr_cigar_begin[r], size_t(s->rr_cigar_begin[r + 1] - s->rr_cigar_begin[r]),
                       s->rr_seq4 + s->rr_seq_off[r], L, qual.data(), aux);
        }
        for (uint32_t u = 0; u < n_unmapped; ++u) {
            Rng rng(s->P.seed, 0x700000000ull + u);
            const uint32_t L = 500 + uint32_t(rng.below(3000));
            std::vector<uint8_t> seq4((L + 1) / 2, 0);
            qual.resize(L);
            for (uint32_t k = 0; k < L; ++k) {
                seq4[k >> 1] |= uint8_t(code4(rand_base(rng)) << ((k & 1) ? 0 : 4));
                qual[k] = uint8_t(rng.below(94));
            }
            std::vector<uint8_t> aux;
            aux_i(aux, "np", int32_t(3 + rng.below(9)));
            aux_z(aux, "RG", "synth");
            bam_record(o, -1, -1, 0, 4, "synth/unmapped" + std::to_string(u) + "/ccs", nullptr, 0, seq4.data(), L, qual.data(), aux);
        }
    }
    uint8_t* out = static_cast<uint8_t*>(std::malloc(std::max<size_t>(o.v.size(), 1)));
    if (!out) return nullptr;
    std::memcpy(out, o.v.data(), o.v.size());
    *n_bytes = o.v.size();
    return out;
}
void ptl_synth_free_bytes(uint8_t* p) { std::free(p); }

void ptl_synth_read_records(const ptl_synth* s, ptl_read_records* o) {
    o->n_reads = uint32_t(s->rr_tid.size());
    o->tid = s->rr_tid.data();
    o->pos = s->rr_pos.data();
    o->flag = s->rr_flag.data();
    o->mapq = s->rr_mapq.data();
    o->bin = s->rr_bin.data();
    o->seq_len = s->rr_seq_len.data();
    o->seq_off = s->rr_seq_off.data();
    o->seq4 = s->rr_seq4;
    o->seq4_bytes = s->rr_seq4_bytes;
    o->cigar_begin = s->rr_cigar_begin.data();
    o->cigar = s->rr_cigar.data();
    o->sa_tag = s->rr_sa.data();
}

}  // extern "C"
