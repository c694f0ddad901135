"""CPU parity of the KERNEL LOGIC: the product's per-index device bodies (pair_bodies.cuh, lift_device.cuh), compiled for
the host under a one-lane SIMT shim (tests/emul), against the oracle — same C-ABI, same seeded inputs, bit-exact.
The GPU itself (32 lanes in lock step, scans, record emission, streams) is covered by the -m gpu tests."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import emul_lib
import helpers
from portello_b200 import abi, synth
from test_gpu_fuzz import make_case

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_unit_vectors.json")))


@pytest.fixture(scope="module")
def emul():
    return emul_lib.load()


def emul_context(E, s):
    ctx = abi.Context(E, 0, 1)
    ctx.set_reference(helpers.reference_arrays(s))
    ctx.set_contig_records(s.contig_records)
    return ctx


@pytest.mark.parametrize("name,kw", [
    ("tiny", dict(n_reads=3000)),
    ("tiny", dict(seed=11, chrom_len=3_000_000, contigs_per_chrom=3, junction_per_mb=8, n_reads=4000)),
    ("tiny", dict(seed=53, n_reads=3000, rev_contig_frac=0.7, read_cluster_frac=0.2)),
    ("config1", dict(n_reads=3000)),
    ("stress", dict(n_reads=300)),
], ids=["tiny", "tiny-junctions", "tiny-reverse-clusters", "config1", "stress"])
def test_emulated_kernels_match_oracle(emul, name, kw):
    s = synth.make(name, **kw)
    pb = helpers.pack(s)
    ro = helpers.lift_c(helpers.oracle_context(s), pb.c)
    re = helpers.lift_c(emul_context(emul, s), pb.c)
    assert ro.n_errors == 0
    d = re.diff(ro)
    assert d is None, d
    assert re.n_lifted > 0


@pytest.mark.parametrize("mask", [1, 2, 3, 4, 5, 6])
def test_emulated_stage_masks(emul, mask):
    s = synth.make("tiny", seed=23, n_reads=1500)
    pb = helpers.pack(s)
    ro = helpers.lift_c(helpers.oracle_context(s), pb.c, mask)
    re = helpers.lift_c(emul_context(emul, s), pb.c, mask)
    d = re.diff(ro)
    assert d is None, d


@pytest.mark.parametrize("seed", range(40))
def test_emulated_fuzz_full_path(emul, oracle, seed):
    chroms, segs, batch = make_case(5000 + seed, n_reads=300)
    octx, ectx = abi.Context(oracle, 0, 1), abi.Context(emul, 0, 1)
    for ctx in (octx, ectx):
        ctx.set_reference(chroms)
        ctx.set_contig_segments(segs)
    ro = octx.lift(batch, allow_panic=True)
    re = ectx.lift(batch, allow_panic=True)
    d = re.diff(ro)
    assert d is None, d
    assert re.first_error_read == ro.first_error_read and re.first_error_status == ro.first_error_status


@pytest.mark.parametrize("mask", [1, 2, 3, 4, 5, 6])
def test_emulated_fuzz_stage_masks(emul, oracle, mask):
    # Stage-test modes 4 and 5 feed RAW read CIGARs to simplify.  A zero-length M/=/X op there ends the reference's edge
    # clean-up (is_alignment_match ignores the length, cigar/mod.rs:278-288); the streaming sink drops empty ops first.
    # The product path (mask 7) never hands simplify such an op (the liftover emits none), so those two modes are fuzzed
    # without empty ops.
    for seed in range(6):
        chroms, segs, batch = make_case(900 + 10 * mask + seed, n_reads=200, zero_prob=0.0 if mask in (4, 5) else 0.05)
        octx, ectx = abi.Context(oracle, 0, 1), abi.Context(emul, 0, 1)
        for ctx in (octx, ectx):
            ctx.set_reference(chroms)
            ctx.set_contig_segments(segs)
        ro = octx.lift(batch, stage_mask=mask, allow_panic=True)
        re = ectx.lift(batch, stage_mask=mask, allow_panic=True)
        d = re.diff(ro)
        assert d is None, (seed, d)


def test_emulated_tables_and_gaps(emul, oracle):
    """Tables entry by entry, and TabEntry::gap == the deletion the reference would push when the walk enters the block
    (distance from the end of the previous aligned run, liftover_read_alignment.rs:91-96)."""
    s = synth.make("tiny", seed=5, chrom_len=2_000_000, junction_per_mb=10)
    octx, ectx = helpers.oracle_context(s), emul_context(emul, s)
    so = octx.get_contig_segments()
    for g in range(len(so.seg_pos)):
        ko, vo = octx.get_segment_table(g)
        ke, ve = ectx.get_segment_table(g)
        assert np.array_equal(ko, ke) and np.array_equal(vo, ve), g
        gaps = np.zeros(len(ke), np.uint32)
        assert emul.dll.ptl_emul_get_segment_table_gaps(ectx.h, g, len(gaps), gaps.ctypes.data_as(abi.u32p)) == 0
        prev_end = None
        for i in range(len(ke)):
            if ve[i] < 0:
                assert gaps[i] == 0
                continue
            run_len = int(ke[i + 1]) - int(ke[i])  # the next key is the run end (None) or the next run start
            want = 0 if prev_end is None else max(0, int(ve[i]) - prev_end)
            assert int(gaps[i]) == want, (g, i)
            prev_end = int(ve[i]) + run_len


@pytest.mark.parametrize("v", G["liftover"], ids=[f"liftover{i}" for i in range(len(G["liftover"]))])
def test_emulated_liftover_vectors(emul, v):
    ctx = abi.Context(emul, 0, 1)
    c2r = v["c2r"] if v["c2r"] is not None else "100S"
    contig_len = helpers.cigar_read_len(c2r)
    segs, batch = helpers.single_pair_case(c2r, v["c2r_pos"], True, contig_len, None, v["pos"], v["cigar"], [], helpers.cigar_read_len(v["cigar"]))
    ctx.set_contig_segments(segs)
    res = ctx.lift(batch, stage_mask=abi.STAGE_LIFTOVER)
    if v["expect"] is None:
        assert res.n_records == 0 and res.n_pairs == 1 and res.n_lifted == 0, v["src"]
    else:
        assert res.n_records == 1, v["src"]
        assert (int(res.rec_pos[0]), res.record_cigar(0)) == (v["expect"]["pos"], v["expect"]["cigar"]), v["src"]


@pytest.mark.parametrize("v", G["simplify"], ids=[f"simplify{i}" for i in range(len(G["simplify"]))])
def test_emulated_simplify_vectors(emul, v):
    ctx = abi.Context(emul, 0, 1)
    ref, read = helpers.relabel(v["ref"]), helpers.relabel(v["read"])
    ctx.set_reference([np.frombuffer(ref.encode(), dtype=np.uint8)])
    segs, batch = helpers.single_pair_case(f"{len(ref)}=", 0, True, len(ref), None, v["pos"], v["cigar"], abi.pack_seq4(read), len(read))
    ctx.set_contig_segments(segs)
    res = ctx.lift(batch, stage_mask=abi.STAGE_SIMPLIFY)
    assert res.n_records == 1
    assert (int(res.rec_pos[0]), res.record_cigar(0)) == (v["expect"]["pos"], v["expect"]["cigar"]), v["src"]


@pytest.mark.parametrize("v", [x for x in G["shift"] if x["dir"] == "left"], ids=lambda v: v["cigar"])
def test_emulated_left_shift_vectors(emul, v):
    ctx = abi.Context(emul, 0, 1)
    ref, read = helpers.relabel(v["ref"]), helpers.relabel(v["read"])
    cig = abi.cigar_from_string(v["cigar"])
    contig_len = len(ref)
    pos_fwd = contig_len - (v["pos"] + helpers.cigar_ref_len(v["cigar"]))
    segs, batch = helpers.single_pair_case(f"{contig_len}=", 0, False, contig_len, ref, pos_fwd, cig[::-1].copy(), abi.pack_seq4(read), len(read),
                                           read_flag=0, rseg_fwd=0)
    ctx.set_contig_segments(segs)
    res = ctx.lift(batch, stage_mask=abi.STAGE_LEFT_SHIFT)
    assert res.n_records == 1 and int(res.rec_need_flip[0]) == 0
    assert (int(res.rec_pos[0]), res.record_cigar(0)) == (v["expect"]["pos"], v["expect"]["cigar"]), v["src"]


@pytest.mark.parametrize("name,kw", [
    ("tiny", dict(seed=53, n_reads=3000, rev_contig_frac=0.7, read_cluster_frac=0.2)),
    ("config1", dict(n_reads=3000)),
    ("stress", dict(n_reads=120)),
], ids=["tiny-reverse-clusters", "config1", "stress"])
def test_emulated_indel_windows_change_nothing(emul, name, kw):
    """Batches packed WITH indel windows (ptl_pack_batch_ex: the first 16 read bases of every homology walk travel with
    the batch) must give the oracle's records bit for bit, and the windows must actually be present."""
    s = synth.make(name, **kw)
    pw = helpers.pack(s, windows=True)
    assert pw.c.n_indel_win > 0 and bool(pw.c.indel_win) and bool(pw.c.rseg_win_begin)
    wb = np.ctypeslib.as_array(pw.c.rseg_win_begin, (pw.c.n_read_segments + 1,))
    assert int(wb[-1]) == pw.c.n_indel_win and np.all(np.diff(wb.astype(np.int64)) >= 0)
    ro = helpers.lift_c(helpers.oracle_context(s), pw.c)
    ectx = emul_context(emul, s)
    re = helpers.lift_c(ectx, pw.c)
    d = re.diff(ro)
    assert d is None, d
    # same counters (base bytes compared) with and without windows: the windows only change where the bytes come from
    cw = np.zeros(6, np.uint64)
    assert emul.dll.ptl_emul_slot_counters(ectx.h, 0, cw.ctypes.data_as(abi.u64p)) == 0
    helpers.lift_c(ectx, helpers.pack(s).c)
    c0 = np.zeros(6, np.uint64)
    assert emul.dll.ptl_emul_slot_counters(ectx.h, 0, c0.ctypes.data_as(abi.u64p)) == 0
    assert np.array_equal(cw, c0) and int(cw[4]) > 0
    # PTL_WIN_REVERSE_PAIRS: only the read segments that pair with a reverse-strand contig segment carry windows
    px = helpers.pack(s, windows=ectx.get_contig_segments())
    assert 0 < px.c.n_indel_win <= pw.c.n_indel_win
    assert helpers.lift_c(ectx, px.c).diff(ro) is None


def test_emulated_windows_are_used_not_the_bases(emul):
    """With windows present, walks of <= 16 bases never touch seq4: zeroing the packed bases of a workload without
    clustered indels (no simplify base access, short walks) must not change the result when windows are on."""
    s = synth.make("tiny", seed=7, n_reads=1500, rev_contig_frac=1.0, read_cluster_frac=0.0)
    pw = helpers.pack(s, windows=True)
    ectx = emul_context(emul, s)
    good = helpers.lift_c(ectx, pw.c)
    b2 = abi.BatchC.from_buffer_copy(pw.c)
    zeros = np.zeros(int(pw.c.seq4_bytes) + 64, np.uint8)
    b2.seq4 = zeros.ctypes.data_as(abi.u8p)
    blind = helpers.lift_c(ectx, b2, stage_mask=abi.STAGE_LEFT_SHIFT | abi.STAGE_LIFTOVER)
    ref = helpers.lift_c(ectx, pw.c, stage_mask=abi.STAGE_LEFT_SHIFT | abi.STAGE_LIFTOVER)
    diff = blind.diff(ref)
    # long homopolymer walks (> 16 bases) may still read seq4: allow a handful of differing records, not a systematic one
    if diff is not None:
        same = np.sum(blind.rec_pos == ref.rec_pos) if len(blind.rec_pos) == len(ref.rec_pos) else 0
        assert same >= 0.98 * len(ref.rec_pos), diff
    assert good.n_lifted > 0


def test_emulated_malformed_batches_are_rejected(emul):
    """Index validation runs inside pair_count_body (no host pass over the batch): the wait reports PTL_ERR_INVALID_ARG."""
    s = synth.make("tiny", seed=3, n_reads=200)
    ectx = emul_context(emul, s)
    for what, b2 in helpers.malformed_batches(s):
        with pytest.raises(abi.PtlError) as e:
            helpers.lift_c(ectx, b2)
        assert e.value.code == abi.PTL_ERR_INVALID_ARG, what
    assert helpers.lift_c(ectx, helpers.pack(s).c).n_lifted > 0  # the context stays usable


# ------------------------------------------------------------------------------------------------------------------
# Warp-cooperative liftover (lift_warp.cuh): 32 host threads in lock step (tests/emul/cuda_shim_warp.hpp).  Slow
# (~13 us per warp collective), so the cases are small; ptl_set_long_pair_ops(0) sends EVERY pair down that path.
def long_pairs(emul, ctx):
    out = C.c_uint64(0)
    assert emul.dll.ptl_emul_slot_long_pairs(ctx.h, 0, C.byref(out)) == 0
    return int(out.value)


@pytest.mark.parametrize("v", G["liftover"], ids=[f"warp-liftover{i}" for i in range(len(G["liftover"]))])
def test_emulated_warp_liftover_vectors(emul, v):
    ctx = abi.Context(emul, 0, 1)
    ctx.set_long_pair_ops(0)
    c2r = v["c2r"] if v["c2r"] is not None else "100S"
    contig_len = helpers.cigar_read_len(c2r)
    segs, batch = helpers.single_pair_case(c2r, v["c2r_pos"], True, contig_len, None, v["pos"], v["cigar"], [], helpers.cigar_read_len(v["cigar"]))
    ctx.set_contig_segments(segs)
    res = ctx.lift(batch, stage_mask=abi.STAGE_LIFTOVER)
    assert long_pairs(emul, ctx) == 1
    if v["expect"] is None:
        assert res.n_records == 0 and res.n_pairs == 1 and res.n_lifted == 0, v["src"]
    else:
        assert res.n_records == 1, v["src"]
        assert (int(res.rec_pos[0]), res.record_cigar(0)) == (v["expect"]["pos"], v["expect"]["cigar"]), v["src"]


@pytest.mark.parametrize("name,kw,long_ops", [
    ("tiny", dict(n_reads=500), 0),
    ("tiny", dict(seed=53, n_reads=400, rev_contig_frac=0.7, read_cluster_frac=0.2), 0),
    ("tiny", dict(seed=11, chrom_len=3_000_000, contigs_per_chrom=3, junction_per_mb=8, n_reads=400), 0),
    ("stress", dict(n_reads=40), 0),
    ("stress", dict(n_reads=60, seed=77), 64),
], ids=["tiny", "tiny-reverse-clusters", "tiny-junctions", "stress-all", "stress-long-only"])
def test_emulated_warp_liftover_matches_oracle(emul, name, kw, long_ops):
    s = synth.make(name, **kw)
    pb = helpers.pack(s)
    ro = helpers.lift_c(helpers.oracle_context(s), pb.c)
    ectx = emul_context(emul, s)
    ectx.set_long_pair_ops(long_ops)
    re = helpers.lift_c(ectx, pb.c)
    d = re.diff(ro)
    assert d is None, d
    n_long = long_pairs(emul, ectx)
    assert n_long > 0 and (long_ops > 0 or n_long >= re.n_lifted)


@pytest.mark.parametrize("mask", [1, 2, 3, 6])
def test_emulated_warp_liftover_stage_masks(emul, mask):
    s = synth.make("tiny", seed=23, n_reads=300)
    pb = helpers.pack(s)
    ro = helpers.lift_c(helpers.oracle_context(s), pb.c, mask)
    ectx = emul_context(emul, s)
    ectx.set_long_pair_ops(0)
    re = helpers.lift_c(ectx, pb.c, mask)
    d = re.diff(ro)
    assert d is None, d
    assert long_pairs(emul, ectx) > 0


@pytest.mark.parametrize("seed", range(6))
def test_emulated_warp_liftover_fuzz(emul, oracle, seed):
    """Adversarial CIGARs (every op kind, zero lengths, edge indels, Pads) through the warp path."""
    chroms, segs, batch = make_case(7000 + seed, n_reads=120)
    octx, ectx = abi.Context(oracle, 0, 1), abi.Context(emul, 0, 1)
    ectx.set_long_pair_ops(0)
    for ctx in (octx, ectx):
        ctx.set_reference(chroms)
        ctx.set_contig_segments(segs)
    ro = octx.lift(batch, allow_panic=True)
    re = ectx.lift(batch, allow_panic=True)
    d = re.diff(ro)
    assert d is None, d
    assert re.first_error_read == ro.first_error_read and re.first_error_status == ro.first_error_status


@pytest.mark.parametrize("n_indels,read_cigar", [
    (40, "3S5000=2I1000=3D2500="),       # one op crossing > 64 table keys: several piece rounds, a `big` window
    (120, "10=1X4000=5D30=7I4000=20S"),
    (33, "2000=1I2000="),
], ids=["40", "120", "33"])
def test_emulated_warp_liftover_many_keys_per_op(emul, oracle, n_indels, read_cigar):
    c2r = "".join(f"{40 + (k % 7)}={1 + k % 3}{'ID'[k % 2]}" for k in range(n_indels)) + "9000="
    contig_len = helpers.cigar_read_len(c2r)
    rlen = helpers.cigar_read_len(read_cigar)
    ref = np.frombuffer(b"ACGT" * ((helpers.cigar_ref_len(c2r) + 200) // 4 + 1), dtype=np.uint8).copy()
    out = []
    for L, long_ops in ((oracle, 64), (emul, 0), (emul, 1 << 30)):
        segs, batch = helpers.single_pair_case(c2r, 100, True, contig_len, None, 17, read_cigar, [], rlen)
        ctx = abi.Context(L, 0, 1)
        ctx.set_long_pair_ops(long_ops)
        ctx.set_reference([ref])
        ctx.set_contig_segments(segs)
        r = ctx.lift(batch, stage_mask=abi.STAGE_LIFTOVER)
        assert r.n_records == 1
        out.append((int(r.rec_pos[0]), r.record_cigar(0)))
    assert out[0] == out[1] == out[2], out


@pytest.mark.parametrize("v", [x for x in G["shift"] if x["dir"] == "left"], ids=lambda v: "warp-" + v["cigar"])
def test_emulated_warp_left_shift_vectors(emul, v):
    """The reference's left_shift_indels vectors through warp_left_shift (lift_warp.cuh)."""
    ctx = abi.Context(emul, 0, 1)
    ctx.set_long_pair_ops(0)
    ref, read = helpers.relabel(v["ref"]), helpers.relabel(v["read"])
    cig = abi.cigar_from_string(v["cigar"])
    contig_len = len(ref)
    pos_fwd = contig_len - (v["pos"] + helpers.cigar_ref_len(v["cigar"]))
    segs, batch = helpers.single_pair_case(f"{contig_len}=", 0, False, contig_len, ref, pos_fwd, cig[::-1].copy(), abi.pack_seq4(read), len(read),
                                           read_flag=0, rseg_fwd=0)
    ctx.set_contig_segments(segs)
    res = ctx.lift(batch, stage_mask=abi.STAGE_LEFT_SHIFT)
    assert long_pairs(emul, ctx) == 1
    assert res.n_records == 1 and int(res.rec_need_flip[0]) == 0
    assert (int(res.rec_pos[0]), res.record_cigar(0)) == (v["expect"]["pos"], v["expect"]["cigar"]), v["src"]


@pytest.mark.parametrize("name,kw", [("tiny", {}), ("config1", {"n_reads": 50}), ("stress", {"n_reads": 50})])
def test_emulated_warp_table_build_matches_scalar(emul, name, kw):
    """table_build_kernel's body (one warp per segment, lanes over CIGAR ops, warp scans, overwrite prefix count) run by 32
    host threads in lock step must produce the same counts and entries as the scalar build the rest of the emulation (and
    the oracle comparison of test_emulated_tables_and_gaps) uses."""
    import ctypes as C
    s = synth.make(name, **kw)
    ctx = emul_context(emul, s)
    fn = emul.dll.ptl_emul_table_build_warp_mismatches
    fn.restype, fn.argtypes = C.c_int64, [C.c_void_p]
    assert fn(ctx.h) == 0


@pytest.mark.parametrize("name,kw", [("tiny", dict(n_reads=400)), ("stress", dict(n_reads=60))])
def test_emulated_warp_pair_count_matches_scalar(emul, name, kw):
    """pair_count_warp_kernel's body (a warp per read: long-read batches on the device) against the per-thread body, on valid
    batches and on the malformed ones (the validity flag must agree too)."""
    import ctypes as C
    s = synth.make(name, **kw)
    ctx = emul_context(emul, s)
    fn = emul.dll.ptl_emul_check_warp_pair_count
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_int]
    assert fn(ctx.h, 1) == 0
    pb = helpers.pack(s)
    res = helpers.lift_c(ctx, pb.c, allow_panic=True)   # submit fails with PTL_ERR_STATE if the two counts differ
    assert res.n_pairs > 0
    if name == "tiny":
        for what, b in helpers.malformed_batches(s):
            with pytest.raises(abi.PtlError) as e:
                helpers.lift_c(ctx, b)
            assert "differs" not in str(e.value), what
