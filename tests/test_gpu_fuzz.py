"""Adversarial parity: random CIGARs with every op kind (M I D N S H P = X, zero lengths, leading/trailing indels,
adjacent I/D runs), random segment tables with overlapping sequencing-order ranges, multi-segment reads on both strands,
random bases incl. IUPAC codes.  Nothing here looks like aligner output; the CUDA path must still equal the oracle
bit for bit (or report the same reference panic)."""
import numpy as np
import pytest

import helpers
from portello_b200 import abi, lib

pytestmark = pytest.mark.gpu
OPS = {c: i for i, c in enumerate("MIDNSHP=X")}


def rand_cigar(rng, n_ops, weights, max_len=40, zero_prob=0.05):
    kinds = rng.choice(len(weights), size=n_ops, p=np.asarray(weights) / np.sum(weights))
    lens = rng.integers(1, max_len, size=n_ops)
    lens[rng.random(n_ops) < zero_prob] = 0
    return ((lens.astype(np.uint32) << 4) | kinds.astype(np.uint32)).astype(np.uint32)


def read_len(c):
    return int(sum(int(v) >> 4 for v in c if (int(v) & 15) in (0, 1, 4, 5, 7, 8)))


def ref_len(c):
    return int(sum(int(v) >> 4 for v in c if (int(v) & 15) in (0, 2, 3, 7, 8)))


def make_case(seed, n_contigs=6, n_reads=400, zero_prob=0.05):
    rng = np.random.default_rng(seed)
    n_chrom = 2
    chrom_len = 60_000
    alphabet = np.frombuffer(b"ACGTACGTACGTNRYM", dtype=np.uint8)
    chroms = [alphabet[rng.integers(0, 16, chrom_len)].copy() for _ in range(n_chrom)]
    # contigs: 1-3 segments each, arbitrary (possibly overlapping / unordered-looking) seq-order ranges, both strands
    contig_len, seg_begin = [], [0]
    so_s, so_e, chrom, pos, fwd, mapq, cb, cig, rev = [], [], [], [], [], [], [0], [], []
    #                 M   I   D   N  S  H  P  =   X
    w_contig = [1.0, 2.0, 2.0, 0.2, 0, 0, 0.1, 8.0, 2.0]
    for c in range(n_contigs):
        nseg = int(rng.integers(1, 4))
        segs = []
        for _ in range(nseg):
            body = rand_cigar(rng, int(rng.integers(3, 40)), w_contig, max_len=60)
            segs.append(body)
        clen = max(read_len(b) for b in segs) + int(rng.integers(0, 50))
        contig_len.append(clen)
        for b in segs:
            lead = int(rng.integers(0, clen - read_len(b) + 1))
            trail = clen - read_len(b) - lead
            clip = int(rng.choice([4, 5]))
            full = np.concatenate([np.array([(lead << 4) | clip], np.uint32), b, np.array([(trail << 4) | clip], np.uint32)])
            is_fwd = bool(rng.random() < 0.5)
            # sequencing-order range as split_read.rs computes it, then sometimes perturbed (pairing is a pure range test)
            s0, s1 = (lead, clen - trail) if is_fwd else (trail, clen - lead)
            if rng.random() < 0.2:
                s0 = max(0, s0 - int(rng.integers(0, 20)))
            if s1 <= s0:
                s1 = s0 + 1
            so_s.append(s0); so_e.append(s1); chrom.append(int(rng.integers(0, n_chrom)))
            pos.append(int(rng.integers(0, chrom_len - ref_len(full) - 1))); fwd.append(int(is_fwd)); mapq.append(int(rng.integers(0, 61)))
            cig.append(full); cb.append(cb[-1] + len(full))
        seg_begin.append(len(so_s))
        rev.append(alphabet[rng.integers(0, 16, clen)].copy())
    segs = abi.ContigSegments(np.array(contig_len, np.uint64), np.array(seg_begin, np.uint32), rev, np.array(so_s, np.uint32),
                              np.array(so_e, np.uint32), np.array(chrom, np.int32), np.array(pos, np.int64), np.array(fwd, np.uint8),
                              np.array(mapq, np.uint8), np.array(cb, np.uint64), np.concatenate(cig))
    # reads: 1-3 segments, every op kind, read length consistent across the read's segments (as the packer guarantees)
    #          M   I    D    N    S    H    P    =    X
    w_read = [2.0, 2.5, 2.5, 0.3, 0.6, 0.2, 0.2, 6.0, 2.0]
    flag, rmapq, rbin, slen, soff, rsb = [], [], [], [], [], [0]
    rc, rp, rf, rcb, rcl, pool, seq = [], [], [], [], [], [], []
    off = 0
    for _ in range(n_reads):
        nseg = int(rng.choice([1, 1, 1, 2, 3]))
        body = rand_cigar(rng, int(rng.integers(1, 30)), w_read, zero_prob=zero_prob)
        L = read_len(body)
        if L == 0:
            body = np.concatenate([body, np.array([(5 << 4) | 0], np.uint32)])
            L = 5
        flag.append(int(rng.choice([0, 16, 1, 17 + 64])))
        rmapq.append(int(rng.integers(0, 61))); rbin.append(int(rng.integers(0, 37450)))
        slen.append(L); soff.append(off)
        codes = rng.integers(0, 16, L + (L & 1)).astype(np.uint8)
        seq.append(((codes[0::2] << 4) | codes[1::2]).astype(np.uint8))
        off += len(seq[-1])
        for k in range(nseg):
            if k == 0:
                cg = body
            else:  # another segment CIGAR with the same read length
                cg = rand_cigar(rng, int(rng.integers(1, 20)), w_read, zero_prob=zero_prob)
                d = L - read_len(cg)
                if d > 0:
                    cg = np.concatenate([cg, np.array([(d << 4) | 4], np.uint32)])
                elif d < 0:
                    cg = np.array([(L << 4) | 0], np.uint32)
            c = int(rng.integers(0, n_contigs))
            span = ref_len(cg)
            hi = contig_len[c] - span
            p = int(rng.integers(0, max(hi, 0) + 1)) if hi >= 0 or rng.random() < 0.7 else 0
            rc.append(c); rp.append(min(p, 2**31 - 1)); rf.append(int(rng.random() < 0.5))
            rcb.append(sum(len(x) for x in pool)); rcl.append(len(cg)); pool.append(cg)
        rsb.append(len(rc))
    batch = abi.Batch(np.array(flag, np.uint16), np.array(rmapq, np.uint8), np.array(rbin, np.uint16), np.array(slen, np.uint32),
                      np.array(soff, np.uint64), np.array(rsb, np.uint32), np.array(rc, np.uint32), np.array(rp, np.int64),
                      np.array(rf, np.uint8), np.array(rcb, np.uint64), np.array(rcl, np.uint32), np.concatenate(pool), np.concatenate(seq))
    return chroms, segs, batch


@pytest.mark.parametrize("long_ops", [64, 0], ids=["default", "warp-all"])
@pytest.mark.parametrize("seed", range(12))
def test_fuzz_full_path(seed, oracle, long_ops):
    chroms, segs, batch = make_case(1000 + seed)
    octx = abi.Context(oracle, 0, 1)
    gctx = lib.GpuContext(0, 1)
    gctx.set_long_pair_ops(long_ops)
    for ctx in (octx, gctx):
        ctx.set_reference(chroms)
        ctx.set_contig_segments(segs)
    ro = octx.lift(batch, allow_panic=True)
    rg = gctx.lift(batch, allow_panic=True)
    d = rg.diff(ro)
    assert d is None, d
    assert rg.first_error_read == ro.first_error_read and rg.first_error_status == ro.first_error_status
    assert ro.n_pairs > 50
    gctx.close()


@pytest.mark.parametrize("long_ops", [64, 0], ids=["default", "warp-all"])
@pytest.mark.parametrize("mask", [1, 2, 3, 6])
def test_fuzz_stage_masks(mask, oracle, long_ops):
    chroms, segs, batch = make_case(77 + mask, n_reads=250)
    octx = abi.Context(oracle, 0, 1)
    gctx = lib.GpuContext(0, 1)
    gctx.set_long_pair_ops(long_ops)
    for ctx in (octx, gctx):
        ctx.set_reference(chroms)
        ctx.set_contig_segments(segs)
    ro = octx.lift(batch, stage_mask=mask, allow_panic=True)
    rg = gctx.lift(batch, stage_mask=mask, allow_panic=True)
    d = rg.diff(ro)
    assert d is None, d
    gctx.close()
