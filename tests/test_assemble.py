"""Record assembly, bases (ptl_assemble_bases): seq + qual of every output record as
reverse_alignment_seq_and_qual (src/read_alignment_scanner.rs:125-133) leaves them.  The oracle restates the reference
(decode -> rev_comp_in_place -> Record::set re-encode); an independent numpy model pins the oracle on CPU (rev_comp itself is pinned by the reference vector in
test_oracle_golden.py), and the CUDA
kernel (no decode: bit reversal of nibble windows) must equal the oracle byte for byte."""
import numpy as np
import pytest

import helpers
from portello_b200 import abi, synth

COMP = np.full(16, 15, np.uint8)
COMP[[1, 2, 4, 8]] = [8, 4, 2, 1]  # A<->T, C<->G; '=', IUPAC codes and N all come back as N (comp_base -> seq_nt16_table)


def model(batch_c, res, qual, qual_off):
    """numpy restatement: per record (seq4 bytes, qual bytes)."""
    n_rsegs = batch_c.n_read_segments
    seg_begin = np.ctypeslib.as_array(batch_c.read_seg_begin, (batch_c.n_reads + 1,))
    rseg_read = np.repeat(np.arange(batch_c.n_reads), np.diff(seg_begin.astype(np.int64)))
    assert len(rseg_read) == n_rsegs
    seq_len = np.ctypeslib.as_array(batch_c.read_seq_len, (batch_c.n_reads,))
    seq_off = np.ctypeslib.as_array(batch_c.read_seq_off, (batch_c.n_reads,))
    pool = np.ctypeslib.as_array(batch_c.seq4, (int(batch_c.seq4_bytes),))
    out = []
    for k in range(res.n_records):
        r = int(rseg_read[int(res.rec_read_segment[k])])
        n = int(seq_len[r])
        raw = pool[int(seq_off[r]): int(seq_off[r]) + (n + 1) // 2]
        q = qual[int(qual_off[r]): int(qual_off[r]) + n]
        if not res.rec_need_flip[k]:
            out.append((raw.copy(), q.copy()))
            continue
        nib = np.empty(2 * len(raw), np.uint8)
        nib[0::2], nib[1::2] = raw >> 4, raw & 15
        rc = COMP[nib[:n][::-1]]
        if n & 1:
            rc = np.append(rc, np.uint8(0))
        out.append(((rc[0::2] << 4 | rc[1::2]).astype(np.uint8), q[::-1].copy()))
    return out


def make_quals(batch_c, seed=5, gap=3):
    rng = np.random.default_rng(seed)
    seq_len = np.ctypeslib.as_array(batch_c.read_seq_len, (batch_c.n_reads,)).astype(np.int64)
    off = np.zeros(batch_c.n_reads, np.uint64)
    off[1:] = np.cumsum(seq_len[:-1] + gap)  # deliberately unaligned offsets
    total = int(seq_len.sum() + gap * batch_c.n_reads) + 8
    return rng.integers(0, 94, total, dtype=np.uint8), off


def split(packed, k):
    sb, seq, qb, ql = packed
    return seq[int(sb[k]): int(sb[k + 1])], ql[int(qb[k]): int(qb[k + 1])]


def check_against_model(ctx, pb, res, qual, off):
    _, packed = ctx.assemble_bases(qual, off)
    want = model(pb.c, res, qual, off)
    assert len(want) == res.n_records
    for k, (ws, wq) in enumerate(want):
        gs, gq = split(packed, k)
        assert np.array_equal(gs[:len(ws)], ws), f"record {k}: bases differ"
        assert np.array_equal(gq[:len(wq)], wq), f"record {k}: qualities differ"
        assert not gs[len(ws):].any() and not gq[len(wq):].any()  # padding to 4 bytes is zero
    return packed


def iupac_case(seed, n_reads=400):
    """Reads with '=' / IUPAC / N codes and odd lengths; some reverse reads so that both orientations appear."""
    s = synth.make("tiny", seed=seed, n_reads=n_reads, read_len_mean=301, read_len_sd=120, read_len_min=1, read_len_max=900, rev_contig_frac=0.6)
    pb = helpers.pack(s)
    pool = np.ctypeslib.as_array(pb.c.seq4, (int(pb.c.seq4_bytes),))
    rng = np.random.default_rng(seed)
    hit = rng.random(pool.size) < 0.2
    pool[hit] = rng.integers(0, 256, int(hit.sum()), dtype=np.uint8)  # arbitrary nibbles (the liftover result may change: irrelevant here)
    return s, pb


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_assemble_matches_numpy_model(seed):
    s, pb = iupac_case(seed)
    octx = helpers.oracle_context(s)
    res = helpers.lift_c(octx, pb.c, allow_panic=True)
    assert res.rec_need_flip.any() and not res.rec_need_flip.all()
    qual, off = make_quals(pb.c, seed)
    check_against_model(octx, pb, res, qual, off)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["iupac-odd", "tiny", "config1", "zero-copy"])
def test_gpu_assemble_matches_oracle(case):
    from portello_b200 import lib
    if case == "iupac-odd":
        s, pb = iupac_case(11, n_reads=3000)
    elif case == "zero-copy":  # the packed bases stay in pinned, mapped host memory: the kernel reads them over PCIe
        import ctypes as C
        L = lib.load()
        s = synth.make("tiny", host_alloc=C.cast(L.dll.ptl_host_alloc, C.c_void_p), host_free=C.cast(L.dll.ptl_host_free, C.c_void_p), seed=4, n_reads=1500)
        pb = helpers.pack(s, pinned=True)
    else:
        s = synth.make(case, **({"n_reads": 3000} if case == "config1" else {}))
        pb = helpers.pack(s)
    octx, gctx = helpers.oracle_context(s), helpers.gpu_context(s)
    if case == "zero-copy":
        gctx.set_seq_zero_copy(True)
    ro = helpers.lift_c(octx, pb.c, allow_panic=True)
    rg = helpers.lift_c(gctx, pb.c, allow_panic=True)
    assert rg.diff(ro) is None
    qual, off = make_quals(pb.c, 9)
    _, po = octx.assemble_bases(qual, off)
    og, pg = gctx.assemble_bases(qual, off)
    for a, b, what in zip(po, pg, ("rec_seq_begin", "seq4", "rec_qual_begin", "qual")):
        assert np.array_equal(a, b), what
    assert og.kernel_ms > 0 and og.bytes_read == og.bytes_written > 0
    # resident qualities + no download: the timing-only mode the bench uses
    o2, none = gctx.assemble_bases(None, None, flags=abi.ASM_RESIDENT_QUAL | abi.ASM_NO_DOWNLOAD)
    assert none is None and o2.n_records == og.n_records and o2.kernel_ms > 0
    if case == "iupac-odd":
        check_against_model(gctx, pb, rg, qual, off)


def _with_segmentless_reads(pb, every=7):
    """The packed batch as numpy with the segments of every `every`-th read removed (such a read yields only the unmapped
    fallback record; its rec_read_segment cannot identify it)."""
    b = pb.c
    n, ns = b.n_reads, b.n_read_segments
    g = lambda p, k: np.ctypeslib.as_array(p, (k,)).copy()
    seg_begin = g(b.read_seg_begin, n + 1).astype(np.int64)
    drop_read = (np.arange(n) % every) == 3
    keep_seg = ~np.repeat(drop_read, np.diff(seg_begin))
    cnt = np.where(drop_read, 0, np.diff(seg_begin))
    new_begin = np.zeros(n + 1, np.uint32)
    new_begin[1:] = np.cumsum(cnt)
    return abi.Batch(read_flag=g(b.read_flag, n), read_mapq=g(b.read_mapq, n), read_bin=g(b.read_bin, n), read_seq_len=g(b.read_seq_len, n),
                     read_seq_off=g(b.read_seq_off, n), read_seg_begin=new_begin, rseg_contig=g(b.rseg_contig, ns)[keep_seg],
                     rseg_pos=g(b.rseg_pos, ns)[keep_seg], rseg_is_fwd=g(b.rseg_is_fwd, ns)[keep_seg],
                     rseg_cigar_begin=g(b.rseg_cigar_begin, ns)[keep_seg], rseg_cigar_len=g(b.rseg_cigar_len, ns)[keep_seg],
                     cigar=g(b.cigar, int(b.n_cigar)), seq4=g(b.seq4, int(b.seq4_bytes))), drop_read


def _check_segmentless(ctx, batch, drop_read, qual, off):
    res = ctx.lift(batch, allow_panic=True)
    _, (sb, seq, qb, ql) = ctx.assemble_bases(qual, off)
    rec_read = np.repeat(np.arange(batch.n_reads), np.diff(res.read_rec_begin.astype(np.int64)))
    assert (np.diff(res.read_rec_begin.astype(np.int64))[drop_read] == 1).all()
    for k in np.flatnonzero(drop_read[rec_read]):
        r = int(rec_read[k])
        n = int(batch.read_seq_len[r])
        assert int(qb[k + 1] - qb[k]) == (n + 15) // 16 * 16, "fallback record sized from another read"
        q = qual[int(off[r]): int(off[r]) + n]
        got = ql[int(qb[k]): int(qb[k]) + n]
        assert np.array_equal(got, q[::-1] if res.rec_need_flip[k] else q), f"record {k}: qualities of another read"
    return res, (sb, seq, qb, ql)


def test_oracle_assemble_read_without_segments():
    s = synth.make("tiny", seed=6, n_reads=300)
    pb = helpers.pack(s)
    batch, drop_read = _with_segmentless_reads(pb)
    qual, off = make_quals(pb.c, 3)
    _check_segmentless(helpers.oracle_context(s), batch, drop_read, qual, off)


@pytest.mark.gpu
def test_gpu_assemble_read_without_segments():
    s = synth.make("tiny", seed=6, n_reads=300)
    pb = helpers.pack(s)
    batch, drop_read = _with_segmentless_reads(pb)
    qual, off = make_quals(pb.c, 3)
    ro, po = _check_segmentless(helpers.oracle_context(s), batch, drop_read, qual, off)
    rg, pg = _check_segmentless(helpers.gpu_context(s), batch, drop_read, qual, off)
    assert rg.diff(ro) is None
    for a, b2 in zip(po, pg):
        assert np.array_equal(a, b2)
