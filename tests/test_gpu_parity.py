"""Bit-exact parity of the CUDA path against the CPU oracle on seeded synthetic data, through the C-ABI."""
import ctypes as C

import numpy as np
import pytest

import helpers
from portello_b200 import abi, lib, synth

pytestmark = pytest.mark.gpu


def both(s, stage_mask=abi.STAGE_ALL, first=0, count=None, long_ops=None):
    pb = helpers.pack(s, first, count)
    octx = helpers.oracle_context(s)
    gctx = helpers.gpu_context(s)
    if long_ops is not None:
        gctx.set_long_pair_ops(long_ops)
    ro = helpers.lift_c(octx, pb.c, stage_mask)
    rg = helpers.lift_c(gctx, pb.c, stage_mask)
    return ro, rg, octx, gctx, pb


@pytest.mark.parametrize("name,kw", [
    ("tiny", {}),
    ("tiny", dict(seed=11, chrom_len=3_000_000, contigs_per_chrom=3, junction_per_mb=8, n_reads=20000)),
    ("config1", {}),
    ("stress", dict(n_reads=1500)),
    # ~190 ops per read: the warp-per-read pair count with 4 reads per warp in the record emission (96 < mean ops <= 256)
    ("stress", dict(n_reads=1200, read_len_mean=17000, read_len_sd=2000, read_len_min=10000, read_len_max=24000)),
], ids=["tiny", "tiny-junctions", "config1", "stress", "stress-17kb"])
@pytest.mark.parametrize("long_ops", [None, 0, 1 << 30], ids=["default", "warp-all", "thread-all"])
def test_full_path_parity(name, kw, long_ops):
    """long_ops: which pairs take the warp-cooperative liftover (default: CIGARs over 64 ops; 0: all; 2^30: none)."""
    s = synth.make(name, **kw)
    ro, rg, octx, gctx, pb = both(s, long_ops=long_ops)
    assert ro.n_errors == 0
    d = rg.diff(ro)
    assert d is None, d
    assert rg.n_records >= pb.c.n_reads and rg.n_lifted > 0
    # SA:Z text built by the product's host helper from the GPU result == oracle's
    r = abi.ResultC()
    gctx._check(gctx.lib._lift_wait(gctx.h, 0, C.byref(r)))
    import oracle_lib
    O = oracle_lib.load()
    sa_g = lib.load().format_sa_tags(r, s.chrom_names)
    sa_o = lib.format_sa_tags_call(O.dll.ptl_oracle_format_sa_tags, r, s.chrom_names)
    assert sa_g == sa_o
    assert any(x for x in sa_g)


@pytest.mark.parametrize("mask", [abi.STAGE_LEFT_SHIFT, abi.STAGE_LIFTOVER, abi.STAGE_LEFT_SHIFT | abi.STAGE_LIFTOVER,
                                  abi.STAGE_LIFTOVER | abi.STAGE_SIMPLIFY])
@pytest.mark.parametrize("long_ops", [None, 0], ids=["default", "warp-all"])
def test_stage_parity(mask, long_ops):  # (mask 1 = left shift only also runs on the warp path when long_ops = 0)
    s = synth.make("tiny", seed=23, n_reads=4000)
    ro, rg, *_ = both(s, mask, long_ops=long_ops)
    d = rg.diff(ro)
    assert d is None, d


def test_tables_match_oracle():
    s = synth.make("tiny", seed=5, chrom_len=2_000_000, junction_per_mb=10)
    octx = helpers.oracle_context(s)
    gctx = helpers.gpu_context(s)
    so, sg = octx.get_contig_segments(), gctx.get_contig_segments()
    assert np.array_equal(so.cigar, sg.cigar) and np.array_equal(so.seg_pos, sg.seg_pos)
    for g in range(len(so.seg_pos)):
        ko, vo = octx.get_segment_table(g)
        kg, vg = gctx.get_segment_table(g)
        assert np.array_equal(ko, kg) and np.array_equal(vo, vg), g


def test_batches_slots_and_order():
    """Several batches in flight on different slots give the same records as one big batch, in submission order."""
    s = synth.make("tiny", seed=31, n_reads=6000)
    gctx = helpers.gpu_context(s, n_slots=3)
    whole = helpers.lift_c(gctx, helpers.pack(s).c)
    n = s.read_records.n_reads
    cuts = [0, n // 3, 2 * n // 3, n]
    packs = [helpers.pack(s, cuts[i], cuts[i + 1] - cuts[i]) for i in range(3)]
    for i, p in enumerate(packs):
        gctx._check(gctx.lib._lift_submit(gctx.h, i, C.byref(p.c)))
    parts = []
    for i in range(3):
        r = abi.ResultC()
        gctx._check(gctx.lib._lift_wait(gctx.h, i, C.byref(r)))
        parts.append(abi.Result.from_c(r))
    assert sum(p.n_records for p in parts) == whole.n_records
    assert np.array_equal(np.concatenate([p.rec_pos for p in parts]), whole.rec_pos)
    assert np.array_equal(np.concatenate([p.cigar for p in parts]), whole.cigar)
    assert np.array_equal(np.concatenate([p.rec_flag for p in parts]), whole.rec_flag)


def test_empty_and_degenerate_batches():
    s = synth.make("tiny", seed=3, n_reads=500)
    gctx = helpers.gpu_context(s)
    octx = helpers.oracle_context(s)
    pb = helpers.pack(s, 0, 0)
    rg = helpers.lift_c(gctx, pb.c)
    assert rg.n_records == 0 and rg.n_pairs == 0 and list(rg.read_rec_begin) == [0]
    one = helpers.pack(s, 17, 1)
    d = helpers.lift_c(gctx, one.c).diff(helpers.lift_c(octx, one.c))
    assert d is None, d


def test_reference_panics_are_reported_not_hidden():
    """A read whose seq_len disagrees with its CIGAR makes the reference panic (read_alignment_scanner.rs:204-229):
    both implementations must flag the same read with PTL_PAIR_ERR_LENGTH and keep the rest of the batch intact."""
    s = synth.make("tiny", seed=3, n_reads=300, unmapped_contig_frac=0.0)
    pb = helpers.pack(s)
    b = pb.c
    lens = np.ctypeslib.as_array(b.read_seq_len, (b.n_reads,)).copy()
    lens[100] += 1
    b2 = abi.BatchC.from_buffer_copy(b)
    b2.read_seq_len = lens.ctypes.data_as(abi.u32p)
    octx, gctx = helpers.oracle_context(s), helpers.gpu_context(s)
    ro = helpers.lift_c(octx, b2, allow_panic=True)
    rg = helpers.lift_c(gctx, b2, allow_panic=True)
    assert ro.n_errors == 1 and rg.n_errors == 1
    assert rg.first_error_read == ro.first_error_read == 100
    assert rg.first_error_status == ro.first_error_status == -1
    assert rg.diff(ro) is None


def test_determinism_and_rerun_after_capacity_growth():
    s = synth.make("tiny", seed=41, n_reads=5000)
    gctx = helpers.gpu_context(s, n_slots=1)
    small = helpers.lift_c(gctx, helpers.pack(s, 0, 50).c)     # sizes the work buffers for a tiny batch
    big1 = helpers.lift_c(gctx, helpers.pack(s).c)             # forces growth
    big2 = helpers.lift_c(gctx, helpers.pack(s).c)
    assert big1.diff(big2) is None and small.n_records >= 50


def test_size_independent_properties_at_scale():
    """configs[1]-shaped data at 1/5 scale: properties that must hold at any size."""
    s = synth.make("chr20", n_reads=200_000)
    gctx = helpers.gpu_context(s)
    pb = helpers.pack(s)
    r = helpers.lift_c(gctx, pb.c)
    b = pb.c
    assert r.n_errors == 0                                  # lifted read length == seq_len for every pair (kernel check)
    assert np.all(np.diff(r.read_rec_begin.astype(np.int64)) >= 1)
    assert np.all(np.diff(r.rec_cigar_begin.astype(np.int64)) >= 0) and int(r.rec_cigar_begin[-1]) == len(r.cigar)
    lifted = r.rec_status == 1
    # exactly one non-supplementary record per read
    prim = ((r.rec_flag & 0x800) == 0)
    assert np.array_equal(np.add.reduceat(prim.astype(np.int64), r.read_rec_begin[:-1].astype(np.int64)), np.ones(b.n_reads, np.int64))
    # read length of every lifted CIGAR equals seq_len; no =/X survive; first/last aligned op is M
    ops, lens = r.cigar & 15, r.cigar >> 4
    assert not np.any((ops == 7) | (ops == 8))
    consumes_read = np.isin(ops, [0, 1, 4, 5])
    csum = np.concatenate([[0], np.cumsum(np.where(consumes_read, lens, 0).astype(np.int64))])
    rec_read_len = csum[r.rec_cigar_begin[1:].astype(np.int64)] - csum[r.rec_cigar_begin[:-1].astype(np.int64)]
    seq_len = np.ctypeslib.as_array(b.read_seq_len, (b.n_reads,))
    rec_read = np.repeat(np.arange(b.n_reads), np.diff(r.read_rec_begin.astype(np.int64)))
    assert np.array_equal(rec_read_len[lifted], seq_len[rec_read][lifted].astype(np.int64))
    # oracle agrees on a random window of the batch
    octx = helpers.oracle_context(s)
    sub = helpers.pack(s, 77_000, 5_000)
    assert helpers.lift_c(gctx, sub.c).diff(helpers.lift_c(octx, sub.c)) is None


def test_zero_copy_bases_give_identical_results():
    """Zero-copy mode (packed bases stay in pinned, mapped host memory; kernels fetch the few bytes they need over PCIe)
    must equal the bulk-upload mode and the oracle bit for bit."""
    L = lib.load()
    alloc = C.cast(L.dll.ptl_host_alloc, C.c_void_p)
    free = C.cast(L.dll.ptl_host_free, C.c_void_p)
    s = synth.make("tiny", host_alloc=alloc, host_free=free, seed=53, n_reads=6000, rev_contig_frac=0.7, read_cluster_frac=0.2)
    pb = helpers.pack(s, pinned=True)
    gctx = helpers.gpu_context(s)
    bulk = helpers.lift_c(gctx, pb.c)
    gctx.set_seq_zero_copy(True)
    zc = helpers.lift_c(gctx, pb.c)
    gctx.set_seq_zero_copy(False)
    assert zc.diff(bulk) is None
    ro = helpers.lift_c(helpers.oracle_context(s), pb.c)
    assert zc.diff(ro) is None
    # chunked zero-copy submissions over several slots (the e2e pipeline of bench.py)
    gctx.set_seq_zero_copy(True)
    n = s.read_records.n_reads
    packs = [helpers.pack(s, a, min(2048, n - a), pinned=True) for a in range(0, n, 2048)]
    recs = []
    for i, p in enumerate(packs):
        sl = i % 2
        if i >= 2:
            recs.append(abi.Result.from_c(gctx.wait_c(sl)))
        gctx.submit_c(p.c, sl)
    for i in range(len(packs), len(packs) + 2):
        if i - 2 >= 0 and i - 2 < len(packs):
            recs.append(abi.Result.from_c(gctx.wait_c(i % 2)))
    assert np.array_equal(np.concatenate([r.cigar for r in recs]), bulk.cigar)
    assert np.array_equal(np.concatenate([r.rec_pos for r in recs]), bulk.rec_pos)
    assert np.array_equal(np.concatenate([r.rec_bin for r in recs]), bulk.rec_bin)


def test_indel_windows_give_identical_results():
    """Batches packed with indel windows (ptl_pack_batch_ex; the first 16 read bases of every homology walk travel with
    the batch, so zero-copy runs stop fetching them over PCIe) must equal the plain batches and the oracle bit for bit."""
    L = lib.load()
    alloc = C.cast(L.dll.ptl_host_alloc, C.c_void_p)
    free = C.cast(L.dll.ptl_host_free, C.c_void_p)
    s = synth.make("tiny", host_alloc=alloc, host_free=free, seed=59, n_reads=6000, rev_contig_frac=0.7, read_cluster_frac=0.2)
    gctx = helpers.gpu_context(s)
    plain = helpers.pack(s, pinned=True)
    win = helpers.pack(s, pinned=True, windows=gctx.get_contig_segments())
    assert 0 < win.c.n_indel_win < helpers.pack(s, windows=True).c.n_indel_win
    ro = helpers.lift_c(helpers.oracle_context(s), plain.c)
    for zero_copy in (False, True):
        gctx.set_seq_zero_copy(zero_copy)
        assert helpers.lift_c(gctx, win.c).diff(ro) is None
        assert helpers.lift_c(gctx, plain.c).diff(ro) is None
    # windows for every read segment and a stress-shaped workload with long walks (> 16 bases fall back to seq4)
    s2 = synth.make("stress", host_alloc=alloc, host_free=free, n_reads=800)
    g2 = helpers.gpu_context(s2)
    w2 = helpers.pack(s2, pinned=True, windows=True)
    g2.set_seq_zero_copy(True)
    assert helpers.lift_c(g2, w2.c).diff(helpers.lift_c(helpers.oracle_context(s2), w2.c)) is None


def test_malformed_batches_are_rejected_not_dereferenced():
    """A batch whose indices leave their pools must come back as PTL_ERR_INVALID_ARG from the wait (validation runs on
    the device, inside the pair enumeration) and leave the context usable."""
    s = synth.make("tiny", seed=3, n_reads=200)
    gctx = helpers.gpu_context(s)
    for what, b2 in helpers.malformed_batches(s):
        with pytest.raises(abi.PtlError) as e:
            helpers.lift_c(gctx, b2)
        assert e.value.code == abi.PTL_ERR_INVALID_ARG, what
    ok = helpers.lift_c(gctx, helpers.pack(s).c)
    assert ok.diff(helpers.lift_c(helpers.oracle_context(s), helpers.pack(s).c)) is None


def _full_digest_parity(s, chunk, threads=16, windows=True):
    """100 % of the records of `s`, chunk by chunk over two slots: digest of the CUDA result == digest of the oracle's
    (the reference's own whole-run assertion is per pair, src/read_alignment_scanner.rs:204-229; this is per record)."""
    from portello_b200.digest import Digest
    n = s.read_records.n_reads
    gctx = helpers.gpu_context(s, n_slots=2)
    octx = helpers.oracle_context(s, threads=threads)
    segs = gctx.get_contig_segments() if windows else None
    dg, do = Digest(), Digest()
    n_pairs = n_lifted = 0
    for a in range(0, n, chunk):
        pb = helpers.pack(s, a, min(chunk, n - a), windows=segs)
        sb = np.ctypeslib.as_array(pb.c.read_seg_begin, (pb.c.n_reads + 1,))
        gri = np.arange(a, a + pb.c.n_reads)
        gctx.submit_c(pb.c, 0)
        ro = helpers.lift_c(octx, pb.c)
        rg = abi.Result.from_c(gctx.wait_c(0), copy=False)
        assert rg.n_errors == 0 and ro.n_errors == 0
        dg.add(rg, sb, gri)
        do.add(ro, sb, gri)
        assert dg == do, f"chunk at read {a}: CUDA {dg.hex()} != oracle {do.hex()}"
        n_pairs += rg.n_pairs
        n_lifted += rg.n_lifted
    assert dg.n_reads == n and n_lifted > 0.9 * n
    return dg, n_pairs


def test_whole_genome_config_full_digest_parity():
    """BASELINE.json configs[2] shape (24 x 130 Mb, ~1000 contigs -> ~2200 segments, 45 % reverse-strand), 2,000,000 reads:
    every record of every read against the oracle."""
    s = synth.make("wg", n_reads=2_000_000)
    dg, n_pairs = _full_digest_parity(s, 250_000)
    assert n_pairs >= 2_000_000


def test_stress_config_full_scale_full_digest_parity():
    """BASELINE.json configs[4] at the stated size: 64 Mb, ~2000 contigs, 200,000 reads of up to 100 kb with
    dense clustered indels and SA segments: every record against the oracle."""
    s = synth.make("stress_full")
    assert 1500 <= s.n_contigs <= 2200
    dg, n_pairs = _full_digest_parity(s, 50_000)
    assert n_pairs >= 200_000


def test_two_host_threads_drive_two_slots():
    """include/portello_b200.h: distinct slots may be driven from distinct host threads.  Two threads, one slot each,
    interleaved batches of different shapes; every result must equal the single-threaded one."""
    import threading
    s = synth.make("tiny", seed=31, n_reads=8000)
    gctx = helpers.gpu_context(s, n_slots=2)
    packs = [helpers.pack(s, a, n) for a, n in ((0, 3000), (3000, 500), (3500, 2500), (6000, 2000), (100, 4000), (7000, 1000))]
    want = [helpers.lift_c(gctx, p.c) for p in packs]
    errors, got = [], {}

    def worker(slot, reps=4):
        try:
            for rep in range(reps):
                for i in range(slot, len(packs), 2):
                    got[(slot, rep, i)] = helpers.lift_c(gctx, packs[i].c, slot=slot)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    for k in (0, 1):  # the same schedule serially: warms the buffers of both slots, then counts its launches
        worker(k, 1)
    l0 = gctx.launch_count()
    for k in (0, 1):
        worker(k)
    serial = gctx.launch_count() - l0
    got.clear()
    l0 = gctx.launch_count()
    th = [threading.Thread(target=worker, args=(k,)) for k in (0, 1)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors
    assert len(got) == 4 * len(packs)
    for (slot, rep, i), r in got.items():
        assert r.diff(want[i]) is None, (slot, rep, i)
    assert gctx.launch_count() - l0 == serial > 0  # the launch counter is atomic: no lost increments


def test_reference_set_after_the_segments_must_cover_their_chromosomes():
    """ptl_set_reference after ptl_set_contig_records: a reference with fewer chromosomes than the segments use is refused."""
    s = synth.make("tiny", seed=4, n_reads=10)
    ref = helpers.reference_arrays(s)
    ctx = lib.GpuContext(0, 1)
    ctx.set_contig_records(s.contig_records)
    with pytest.raises(abi.PtlError):
        ctx.set_reference(ref[:1])
    ctx.set_reference(ref)
    pb = helpers.pack(s)
    assert helpers.lift_c(ctx, pb.c).diff(helpers.lift_c(helpers.oracle_context(s), pb.c)) is None
