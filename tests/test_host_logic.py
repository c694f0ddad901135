"""CPU-only tests of the product's host-side logic (no GPU needed): the C-ABI library loads and exports every symbol
declared in include/portello_b200.h, and its host helpers (a2 packer, a11/a12 contig preparation, SA text, work-unit
sharding, reg2bin) agree with the oracle / the reference's vectors."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import helpers
import oracle_lib
from portello_b200 import abi, lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_unit_vectors.json")))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "portello_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(ptl_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 30
    dll = C.CDLL(lib.build())
    missing = [n for n in sorted(names) if not hasattr(dll, n)]
    assert not missing, missing
    odll = oracle_lib.load().dll
    for n in ("create", "destroy", "last_error", "set_reference", "set_contig_segments", "set_raw_contig_segments", "set_contig_records",
              "get_contig_segments", "get_segment_table", "lift_submit", "lift_submit_ex", "lift_wait"):
        assert hasattr(odll, "ptl_oracle_" + n)


def test_no_device_means_no_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(abi.PtlError) as e:
        lib.GpuContext(0, 1)
    assert e.value.code == abi.PTL_ERR_NO_DEVICE


def test_product_does_not_link_the_oracle():
    """The product tree may mention the oracle in comments, but must not include, import, link or dlopen it."""
    for dp, _, files in os.walk(os.path.join(ROOT, "portello_b200")):
        for f in files:
            if not (f.endswith((".cpp", ".cu", ".hpp", ".cuh", ".h", ".py")) or f == "Makefile"):
                continue
            for line in open(os.path.join(dp, f)):
                st = line.lstrip()
                if st.startswith(("#include", "import ", "from ")) or "CDLL(" in line or "dlopen" in line or f == "Makefile":
                    assert "oracle" not in line, (dp, f, line)


def test_split_segments_vectors_product_packer():
    L = lib.load()
    for v in G["split_segments"]:
        got = oracle_lib.split_segments_call(L.dll.ptl_pack_split_segments, L.dll.ptl_pack_last_error, v["names"], v["tid"], v["pos"],
                                             v["flag"], v["mapq"], v["cigar"], v["sa"])
        assert got == v["expect"], v["src"]


@pytest.mark.parametrize("sa,cigar", [
    ("chr0,20,-,5M15S,60;", "10S5M5S"),                 # 5 fields
    ("chr0,x,-,5M15S,60,0;", "10S5M5S"),                # bad position
    ("chr0,20,-,5Q15S,60,0;", "10S5M5S"),               # bad CIGAR op
    ("chr0,20,-,20S,60,0;", "10S5M5S"),                 # unaligned SA segment  (split_read.rs:107-110)
    ("chr0,20,-,5M14S,60,0;", "10S5M5S"),               # read size mismatch    (split_read.rs:113)
    ("chrZ,20,-,5M15S,60,0;", "10S5M5S"),               # unknown contig        (split_read.rs:116-126)
    ("chr0,20,-,5M15S,300,0;", "10S5M5S"),              # MAPQ does not fit u8
])
def test_split_segments_panics_agree(oracle, sa, cigar):
    L = lib.load()
    names = ["chr0", "chr1", "chr2"]
    a = oracle.split_segments(names, 2, 9, 0, 60, cigar, sa)
    b = oracle_lib.split_segments_call(L.dll.ptl_pack_split_segments, L.dll.ptl_pack_last_error, names, 2, 9, 0, 60, cigar, sa)
    assert isinstance(a, dict) and isinstance(b, dict) and a["error"] == b["error"] == abi.PTL_ERR_INPUT


def _same_segments(a: abi.ContigSegments, b: abi.ContigSegments):
    for f in ("contig_len", "contig_seg_begin", "seg_seq_order_start", "seg_seq_order_end", "seg_chrom_index", "seg_pos", "seg_is_fwd",
              "seg_mapq", "seg_cigar_begin", "cigar"):
        x, y = getattr(a, f), getattr(b, f)
        assert x.shape == y.shape and np.array_equal(x, y), f
    for x, y in zip(a.rev_contig_seq, b.rev_contig_seq):
        assert (x is None) == (y is None) and (x is None or np.array_equal(x, y))


@pytest.mark.parametrize("kw", [dict(seed=5), dict(seed=11, chrom_len=3_000_000, contigs_per_chrom=3, junction_per_mb=8),
                                dict(seed=12, chrom_len=6_000_000, contigs_per_chrom=2, junction_per_mb=6, rev_contig_frac=1.0),
                                dict(seed=13, chrom_len=6_000_000, contigs_per_chrom=2, junction_per_mb=6, rev_contig_frac=0.0)])
def test_contig_preparation_matches_oracle(kw):
    """a11 + a12 (record assembly, supplementary fill-in, trim, join, rev_contig_seq) product host C++ vs oracle."""
    s = synth.make("tiny", n_reads=10, **kw)
    got = lib.load().prepare_contig_records(s.contig_records)
    octx = abi.Context(oracle_lib.load(), 0, 1)
    octx.set_contig_records(s.contig_records)
    want = octx.get_contig_segments()
    _same_segments(got, want)
    if kw.get("junction_per_mb", 0) >= 6:
        assert len(want.seg_pos) < s.contig_records.n_records  # joins/eliminations actually happened
    # raw-segment entry point: feeding the prepared segments again must be idempotent for trim (nothing overlaps any more)
    again = lib.load().prepare_raw_contig_segments(got)
    assert np.array_equal(again.seg_seq_order_start, got.seg_seq_order_start) or len(again.seg_pos) <= len(got.seg_pos)


def test_trim_vectors_through_product_prepare():
    """clip_seg_isec_range vectors (contig_repeated_match_trimmer.rs:311-397) replayed as a two-segment overlap through
    the product's trimmer: the second segment is made the winner (=, higher MAPQ) so the first one is clipped."""
    L = lib.load()
    for v in G["clip_seg_isec_range"]:
        cig = v["cigar"].replace("M", "=")
        a = abi.cigar_from_string(cig)
        so0, so1 = v["so"]
        i0, i1 = v["isec"]
        # seg1 = the vector's segment; seg2 starts at isec.start (sequencing order) and covers the rest
        seg2_cig = abi.cigar_from_string(f"{i0}S{so1 - i0}=")
        if not v["is_fwd"]:
            seg2_cig = seg2_cig[::-1].copy()
        raw = abi.ContigSegments(
            contig_len=np.array([so1], np.uint64), contig_seg_begin=np.array([0, 2], np.uint32), rev_contig_seq=[None],
            seg_seq_order_start=np.array([so0, i0], np.uint32), seg_seq_order_end=np.array([so1, so1], np.uint32),
            seg_chrom_index=np.array([0, 0], np.int32), seg_pos=np.array([v["pos"], 5000], np.int64),
            seg_is_fwd=np.array([int(v["is_fwd"])] * 2, np.uint8), seg_mapq=np.array([20, 60], np.uint8),
            seg_cigar_begin=np.array([0, len(a), len(a) + len(seg2_cig)], np.uint64), cigar=np.concatenate([a, seg2_cig]))
        out = L.prepare_raw_contig_segments(raw)
        e = v["expect"]
        assert int(out.seg_pos[0]) == e["pos"], v["src"]
        assert abi.cigar_to_string(out.segment_cigar(0)) == e["cigar"].replace("M", "="), v["src"]
        assert [int(out.seg_seq_order_start[0]), int(out.seg_seq_order_end[0])] == e["so"], v["src"]


def test_pack_batch_matches_per_record_oracle_split(oracle):
    s = synth.make("tiny", seed=19, n_reads=1500, read_sa_frac=0.3)
    pb = helpers.pack(s)
    b, rr = pb.c, s.read_records
    assert b.n_reads == rr.n_reads
    n_sa = 0
    for r in range(rr.n_reads):
        cg = np.ctypeslib.as_array(rr.cigar, (int(rr.cigar_begin[rr.n_reads]),))[int(rr.cigar_begin[r]):int(rr.cigar_begin[r + 1])]
        sa = rr.sa_tag[r].decode() if rr.sa_tag[r] else None
        n_sa += sa is not None
        want = oracle.split_segments(s.contig_names, rr.tid[r], rr.pos[r], rr.flag[r], rr.mapq[r], cg, sa)
        s0, s1 = b.read_seg_begin[r], b.read_seg_begin[r + 1]
        assert s1 - s0 == len(want)
        for k, w in zip(range(s0, s1), want):
            got_c = abi.cigar_to_string(np.ctypeslib.as_array(b.cigar, (int(b.n_cigar),))[int(b.rseg_cigar_begin[k]):int(b.rseg_cigar_begin[k]) + int(b.rseg_cigar_len[k])])
            assert (b.rseg_contig[k], b.rseg_pos[k], bool(b.rseg_is_fwd[k]), got_c) == (w["contig"], w["pos"], w["is_fwd"], w["cigar"])
    assert n_sa > 100


def test_region_segments_shard_units_reg2bin(oracle):
    L = lib.load()
    for v in G["region_segments"]:
        assert L.region_segments(v["size"], v["segment_size"]) == v["expect"], v["src"]
    rnd = np.random.default_rng(3)
    for _ in range(300):
        size, seg = int(rnd.integers(1, 10**9)), int(rnd.integers(1, 3 * 10**7))
        assert L.region_segments(size, seg) == oracle.region_segments(size, seg)
    for _ in range(5000):
        b = int(rnd.integers(0, 1 << 29))
        e = min(b + int(rnd.choice([1, 2, 100, 15000, 1 << 14, 1 << 17, 1 << 20])), 1 << 29)
        if e > b:
            assert L.reg2bin(b, e) == oracle.reg2bin(b, e)
    w = rnd.integers(1, 10**6, size=200)
    for n in (1, 2, 4, 8):
        owner = L.shard_units(w, n)
        loads = np.bincount(owner, weights=w, minlength=n)
        assert owner.max() < n and loads.max() <= loads.mean() + w.max()  # LPT bound
        assert np.array_equal(owner, L.shard_units(w, n))                 # deterministic: every rank computes the same map


def test_format_sa_tags_matches_oracle_text(oracle):
    """a10 text part: SA:Z strings built by the product's host helper from a result == the oracle's
    (get_sa_tag_segment + concatenation, src/read_alignment_scanner.rs:292-301,349-363)."""
    import ctypes as C

    s = synth.make("tiny", seed=29, n_reads=1200, read_sa_frac=0.3, junction_per_mb=10.0)
    octx = helpers.oracle_context(s)
    pb = helpers.pack(s)
    O = oracle
    assert O._lift_submit(octx.h, 0, C.byref(pb.c)) == 0
    r = abi.ResultC()
    assert O._lift_wait(octx.h, 0, C.byref(r)) == 0
    got = lib.load().format_sa_tags(r, s.chrom_names)
    want = lib.format_sa_tags_call(O.dll.ptl_oracle_format_sa_tags, r, s.chrom_names)
    assert got == want
    res = abi.Result.from_c(r)
    multi = [k for rd in range(len(res.read_rec_begin) - 1) for k in range(res.read_rec_begin[rd], res.read_rec_begin[rd + 1])
             if res.read_rec_begin[rd + 1] - res.read_rec_begin[rd] > 1]
    assert len(multi) > 20 and all(got[k].endswith(",0;") for k in multi)
    single_unmapped = [k for k in range(res.n_records) if res.rec_status[k] == 0]
    assert all(got[k] == "" for k in single_unmapped)
    # spot-check the format of one entry: chrom,pos+1,strand,CIGAR,mapq,0;
    k = multi[0]
    rd = int(np.searchsorted(res.read_rec_begin, k, side="right") - 1)
    others = [j for j in range(res.read_rec_begin[rd], res.read_rec_begin[rd + 1]) if j != k]
    j = others[0]
    first = got[k].split(";")[0].split(",")
    assert first[0] == s.chrom_names[res.rec_tid[j]] and int(first[1]) == int(res.rec_pos[j]) + 1
    assert first[2] == ("-" if res.rec_flag[j] & 0x10 else "+") and first[3] == res.record_cigar(j) and int(first[4]) == int(res.rec_mapq[j])


def _batch_arrays(c):
    g = lambda p, k: np.ctypeslib.as_array(p, (k,)).copy() if k else np.zeros(0)
    n, ns = c.n_reads, c.n_read_segments
    out = {"n": (n, ns, int(c.n_cigar))}
    for f, k in (("read_flag", n), ("read_mapq", n), ("read_bin", n), ("read_seq_len", n), ("read_seq_off", n), ("read_seg_begin", n + 1),
                 ("rseg_contig", ns), ("rseg_pos", ns), ("rseg_is_fwd", ns), ("rseg_cigar_len", ns)):
        out[f] = g(getattr(c, f), k)
    cb, cl = g(c.rseg_cigar_begin, ns), out["rseg_cigar_len"]
    pool = g(c.cigar, int(c.n_cigar))
    out["seg_cigars"] = [pool[int(cb[i]): int(cb[i]) + int(cl[i])].tobytes() for i in range(ns)]  # (the pools may be laid out differently)
    return out


def _same(a, b):
    return a.keys() == b.keys() and all((a[k] == b[k]) if isinstance(a[k], (tuple, list)) else np.array_equal(a[k], b[k]) for k in a)


def test_oracle_packer_equals_product_packer():
    """bench.py --impl reference packs with the oracle's own record-loop head (ptl_oracle_pack_batch) so that it never loads
    the product library: both packers must produce the same batch."""
    import oracle_lib
    s = synth.make("tiny", seed=12, n_reads=2500)
    pb = helpers.pack(s)
    ob = oracle_lib.OraclePackedBatch(s.read_records, 0, s.read_records.n_reads, s.contig_names, threads=3)
    assert _same(_batch_arrays(pb.c), _batch_arrays(ob.c))
    assert np.array_equal(pb.record_index(), ob.record_index())


def test_repacking_a_batch_reuses_its_arena_and_equals_a_fresh_pack():
    """ptl_pack_batch_into (a streaming host repacks a ring of pinned batches; distinct batches from distinct threads)."""
    import threading
    s = synth.make("tiny", seed=13, n_reads=3000)
    L = lib.load()
    segs = L.prepare_contig_records(s.contig_records)
    for windows in (None, segs):
        fresh = lib.PackedBatch(L, s.read_records, 1000, 1500, s.contig_names, windows=windows)
        ring = lib.PackedBatch(L, s.read_records, 0, 2900, s.contig_names, windows=windows)   # a larger arena first
        ring.repack(1000, 1500)
        a, b = _batch_arrays(fresh.c), _batch_arrays(ring.c)
        assert _same(a, b)
        if windows is not None:
            assert fresh.c.n_indel_win == ring.c.n_indel_win > 0
            assert np.array_equal(np.ctypeslib.as_array(fresh.c.indel_win, (int(fresh.c.n_indel_win),)), np.ctypeslib.as_array(ring.c.indel_win, (int(ring.c.n_indel_win),)))
        ring.repack(0, 10)
        assert ring.c.n_reads == 10
        ring.repack(0, 3000)  # grows
        assert _same(_batch_arrays(ring.c), _batch_arrays(lib.PackedBatch(L, s.read_records, 0, 3000, s.contig_names, windows=windows).c))
    bufs = [lib.PackedBatch(L, s.read_records, 0, 10, s.contig_names, windows=segs) for _ in range(4)]
    th = [threading.Thread(target=lambda t=t: [bufs[t].repack(i * 250, 250) for i in range(t, 12, 4)]) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for t in range(4):  # the last chunk each thread packed: 8 + t
        assert _same(_batch_arrays(bufs[t].c), _batch_arrays(lib.PackedBatch(L, s.read_records, (8 + t) * 250, 250, s.contig_names, windows=segs).c))


def _records(rows, pool_pad=0):
    """ptl_read_records from (flag, pos, tid, cigar, seq_len, sa) rows; bases are a fixed pseudo-random nibble stream."""
    n = len(rows)
    cig = [abi.cigar_from_string(r[3]).astype(np.uint32) for r in rows]
    cb = np.zeros(n + 1, np.uint64)
    cb[1:] = np.cumsum([len(c) for c in cig])
    seq_len = np.array([r[4] for r in rows], np.uint32)
    seq_off = np.zeros(n, np.uint64)
    seq_off[1:] = np.cumsum((seq_len[:-1].astype(np.uint64) + 1) // 2)
    nbytes = int(seq_off[-1] + (int(seq_len[-1]) + 1) // 2) + pool_pad
    rng = np.random.default_rng(5)
    seq4 = (rng.integers(1, 9, nbytes, dtype=np.uint8) << 4 | rng.integers(1, 9, nbytes, dtype=np.uint8)).astype(np.uint8)
    keep = dict(tid=np.array([r[2] for r in rows], np.int32), pos=np.array([r[1] for r in rows], np.int64), flag=np.array([r[0] for r in rows], np.uint16),
                mapq=np.full(n, 60, np.uint8), bin=np.zeros(n, np.uint16), seq_len=seq_len, seq_off=seq_off, seq4=seq4, cb=cb,
                cigar=np.concatenate(cig), sa=(C.c_char_p * n)(*[r[5].encode() if r[5] else None for r in rows]))
    p = lambda a, t: a.ctypes.data_as(t)
    rc = lib.ReadRecordsC(n, p(keep["tid"], lib.i32p), p(keep["pos"], lib.i64p), p(keep["flag"], lib.u16p), p(keep["mapq"], lib.u8p), p(keep["bin"], lib.u16p),
                          p(seq_len, lib.u32p), p(seq_off, lib.u64p), p(seq4, lib.u8p), nbytes, p(cb, lib.u64p), p(keep["cigar"], lib.u32p), keep["sa"])
    rc._keep = keep
    return rc


def _walk_windows(ops, nib, seq_len, flip):
    """The definition (read_alignment_scanner.rs:153-167 + cigar_indel_shifter.rs:63-85): clusters of the REVERSED CIGAR; the
    window holds read_view[read_end - 1 - q], q = 0..15, read_view flipped for a need_flipped_read_alignment pair."""
    out, read_head, blk, ins, open_ = [], 0, 0, 0, False

    def close():
        read_end, bits = (blk + ins) & 0xffffffff, 0
        for q in range(min(16, read_end)):
            idx = read_end - 1 - q
            if idx < seq_len:
                bits |= int(nib[seq_len - 1 - idx if flip else idx]) << (4 * q)
        out.append(bits)

    for x in reversed([int(v) for v in ops]):
        op, l = x & 15, x >> 4
        if op in (1, 2):
            if l > 0:
                if not open_:
                    open_, blk = True, read_head
                if op == 1:
                    ins += l
        elif open_:
            close()
            open_, ins = False, 0
        if op in (0, 1, 4, 5, 7, 8):
            read_head += l
    if open_:
        close()
    return out


def test_indel_windows_equal_the_reversed_cigar_walk():
    """ptl_pack_batch_ex windows against the definition, walked in Python: both strands of the pair (primary segments are
    always flipped views, opposite-strand SA segments are not), clusters closer than 16 bases to either read end, leading and
    trailing clips, empty ops inside and between clusters, a CIGAR longer than the stored bases, and the last read of the pool
    (the two-load fast path must not read past it)."""
    rows = [
        (0, 100, 0, "20=1I30=2D10=1I1D5=", 67, None),
        (16, 100, 0, "20=1I30=2D10=1I1D5=", 67, None),
        (0, 50, 1, "3=1I50=", 54, None),                      # cluster 4 bases into the read
        (0, 50, 1, "50=2I3=", 55, None),                      # ... and 3 bases before its end
        (0, 70, 0, "5S10=2I40=3D7=3S", 67, None),
        (16, 70, 0, "4H6S10=2I40=3D7=3S", 72, None),
        (0, 10, 1, "10=0I1I0D20=0M1D5=0D0I", 36, None),       # empty ops: inside a cluster, closing one, and a cluster of only empty ops
        (0, 10, 1, "30=1I30=1D30=", 40, None),                # CIGAR says 91 bases, the record holds 40
        (0, 500, 0, "20=1I19=20S", 60, "ctg1,100,-,20S30=1I9=,60,0;"),   # SA segment on the other strand: not flipped
        (16, 500, 0, "20=1D20=20S", 60, "ctg1,100,-,40S3=1I16=,60,0;ctg0,900,+,1I39=20S,33,1;"),
        (0, 20, 0, "1I16=1D17=", 34, None),                   # last read: windows end exactly at the last byte of the pool
    ]
    for pad in (0, 16):
        rc = _records(rows, pool_pad=pad)
        pb = lib.PackedBatch(lib.load(), rc, 0, len(rows), ["ctg0", "ctg1"], windows=True)
        b = pb.c
        assert b.n_reads == len(rows) and b.n_read_segments == len(rows) + 3
        pool = np.ctypeslib.as_array(b.cigar, (int(b.n_cigar),))
        win = np.ctypeslib.as_array(b.indel_win, (int(b.n_indel_win),))
        seq4 = rc._keep["seq4"]
        n_checked = 0
        for i in range(b.n_reads):
            raw = seq4[int(rc._keep["seq_off"][i]):]
            nib = np.stack([raw >> 4, raw & 15], 1).reshape(-1)
            for k in range(b.read_seg_begin[i], b.read_seg_begin[i + 1]):
                ops = pool[int(b.rseg_cigar_begin[k]): int(b.rseg_cigar_begin[k]) + int(b.rseg_cigar_len[k])]
                flip = not (bool(rows[i][0] & 16) == bool(b.rseg_is_fwd[k]))
                want = _walk_windows(ops, nib, rows[i][4], flip)
                got = [int(v) for v in win[b.rseg_win_begin[k]: b.rseg_win_begin[k + 1]]]
                assert got == want, (i, k, [hex(v) for v in got], [hex(v) for v in want])
                n_checked += len(want)
        assert n_checked == b.n_indel_win and n_checked >= 22


def _random_cigar(rng, max_ops):
    """Any op anywhere (clips and pads in the middle, empty ops, I/D runs), biased towards what shapes the packer's scan."""
    n = int(rng.integers(1, max_ops + 1))
    ops = rng.choice(np.arange(9), n, p=[0.22, 0.17, 0.17, 0.03, 0.06, 0.03, 0.04, 0.2, 0.08])
    lens = np.where(rng.random(n) < 0.12, 0, rng.integers(1, 40, n))
    if rng.random() < 0.5:
        ops[0], lens[0] = 4, int(rng.integers(0, 30))
    if rng.random() < 0.5:
        ops[-1], lens[-1] = int(rng.choice([4, 5])), int(rng.integers(0, 30))
    return "".join(f"{int(l)}{'MIDNSHP=X'[int(o)]}" for o, l in zip(ops, lens))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_packer_on_random_cigars_equals_oracle_packer_and_window_walk(seed):
    """The fused CIGAR walk of ptl_pack_batch_ex (clip positions, spans, cluster offsets in one pass) on CIGARs no aligner
    writes: same batch as the oracle's record-loop head, same errors, and windows equal to the reversed-CIGAR walk."""
    rng = np.random.default_rng(seed)
    rows, n_err = [], 0
    while len(rows) < 400:
        cig = _random_cigar(rng, 40)
        qlen = helpers.cigar_read_len(cig) + (int(rng.integers(-3, 4)) if rng.random() < 0.1 else 0)
        row = (int(rng.choice([0, 16])), int(rng.integers(0, 5000)), int(rng.integers(0, 2)), cig, max(qlen, 1), None)
        rc = _records([row])
        try:
            want = oracle_lib.OraclePackedBatch(rc, 0, 1, ["ctg0", "ctg1"])
        except abi.PtlError as e:
            with pytest.raises(abi.PtlError) as got:      # (no aligned read base between the clips)
                lib.PackedBatch(lib.load(), rc, 0, 1, ["ctg0", "ctg1"], windows=True)
            assert str(got.value).split(": ", 1)[-1] == str(e).split(": ", 1)[-1]
            n_err += 1
            continue
        rows.append(row)
    assert n_err > 0
    rc = _records(rows)
    want = oracle_lib.OraclePackedBatch(rc, 0, len(rows), ["ctg0", "ctg1"])
    pb = lib.PackedBatch(lib.load(), rc, 0, len(rows), ["ctg0", "ctg1"], windows=True)
    assert _same(_batch_arrays(pb.c), _batch_arrays(want.c))
    assert _same(_batch_arrays(lib.PackedBatch(lib.load(), rc, 0, len(rows), ["ctg0", "ctg1"]).c), _batch_arrays(want.c))
    b = pb.c
    pool = np.ctypeslib.as_array(b.cigar, (int(b.n_cigar),))
    win = np.ctypeslib.as_array(b.indel_win, (max(int(b.n_indel_win), 1),))
    seq4 = rc._keep["seq4"]
    n_win = 0
    for i in range(b.n_reads):
        raw = seq4[int(rc._keep["seq_off"][i]):]
        nib = np.stack([raw >> 4, raw & 15], 1).reshape(-1)
        for k in range(b.read_seg_begin[i], b.read_seg_begin[i + 1]):
            ops = pool[int(b.rseg_cigar_begin[k]): int(b.rseg_cigar_begin[k]) + int(b.rseg_cigar_len[k])]
            flip = not (bool(rows[i][0] & 16) == bool(b.rseg_is_fwd[k]))
            exp = _walk_windows(ops, nib, rows[i][4], flip)
            assert [int(v) for v in win[b.rseg_win_begin[k]: b.rseg_win_begin[k + 1]]] == exp, (i, rows[i])
            n_win += len(exp)
    assert n_win == b.n_indel_win and n_win > 1000


def test_vector_and_scalar_cigar_walks_pack_the_same_batch():
    """ptl_pack_batch_ex walks long CIGARs eight ops at a time where the CPU has AVX2 (PTL_PACK_SCALAR=1 switches that off):
    both must produce the same batch, windows included, on reads of 20 to 600 ops and on the random CIGARs of every op kind."""
    import hashlib
    import subprocess
    import sys
    code = r'''
import hashlib, sys
sys.path.insert(0, "tests")
import numpy as np
import test_host_logic as T
from portello_b200 import lib, synth
h = hashlib.sha1()
def add(c):
    for k, v in sorted(T._batch_arrays(c).items()):
        for a in (v if isinstance(v, list) else [v]):
            h.update(np.asarray(a).tobytes() if not isinstance(a, bytes) else a)
    if c.n_indel_win:
        h.update(np.ctypeslib.as_array(c.indel_win, (int(c.n_indel_win),)).tobytes())
        h.update(np.ctypeslib.as_array(c.rseg_win_begin, (int(c.n_read_segments) + 1,)).tobytes())
L = lib.load()
for name, kw in (("tiny", dict(seed=3, n_reads=1500, read_sa_frac=0.2)), ("stress", dict(n_reads=60))):
    s = synth.make(name, **kw)
    for w in (None, True, L.prepare_contig_records(s.contig_records)):
        add(lib.PackedBatch(L, s.read_records, 0, s.read_records.n_reads, s.contig_names, windows=w).c)
rng = np.random.default_rng(9)
rows = []
while len(rows) < 300:
    cig = T._random_cigar(rng, 90)
    row = (int(rng.choice([0, 16])), int(rng.integers(0, 5000)), 0, cig, max(T.helpers.cigar_read_len(cig), 1), None)
    try:
        lib.PackedBatch(L, T._records([row]), 0, 1, ["ctg0", "ctg1"], windows=True)
        rows.append(row)
    except Exception:
        pass
add(lib.PackedBatch(L, T._records(rows), 0, len(rows), ["ctg0", "ctg1"], windows=True).c)
print(h.hexdigest())
'''
    out = []
    for scalar in ("", "1"):
        env = dict(os.environ)
        env.pop("PTL_PACK_SCALAR", None)
        if scalar:
            env["PTL_PACK_SCALAR"] = "1"
        r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        out.append(r.stdout.strip().splitlines()[-1])
    assert out[0] == out[1] and len(out[0]) == 40


def test_malformed_contig_segments_are_refused():
    """ptl_set_contig_segments / ptl_prepare_raw_contig_segments are public entry points: CSR arrays that do not span their pools,
    negative chromosome indices and positions outside the BAM range must be reported before anything is indexed with them
    (found by tools/fuzz/fuzz_segments_through_device_code.py under AddressSanitizer)."""
    import copy
    s = synth.make("tiny", seed=29, n_reads=10, junction_per_mb=15)
    L = lib.load()
    good = L.prepare_contig_records(s.contig_records)
    L.prepare_raw_contig_segments(good)
    n_seg = len(good.seg_pos)

    def broken(field, index, value):
        g = copy.copy(good)
        g._keep = []
        a = np.array(getattr(g, field), copy=True)
        a[index] = value
        setattr(g, field, a)
        return g
    for field, index, value in (("contig_seg_begin", 1, 2878940490), ("contig_seg_begin", len(good.contig_seg_begin) - 1, n_seg + 3), ("contig_seg_begin", 0, 1),
                                ("seg_cigar_begin", n_seg // 2, 2**40), ("seg_chrom_index", 0, -5), ("seg_pos", 1, -1), ("seg_pos", 1, 2**31),
                                ("seg_seq_order_start", 0, 2**32 - 1)):
        with pytest.raises(abi.PtlError):
            L.prepare_raw_contig_segments(broken(field, index, value))
        ectx = abi.Context(__import__("emul_lib").load(), 0, 1)
        with pytest.raises(abi.PtlError):
            ectx.set_contig_segments(broken(field, index, value))


def test_reference_set_after_the_segments_must_cover_their_chromosomes():
    """Segments may be installed before the reference; a reference with fewer chromosomes than they use is refused by whichever
    call comes second (the simplify stage indexes the chromosome table with the segment's chromosome)."""
    import emul_lib
    s = synth.make("tiny", seed=4, n_reads=10)          # two chromosomes
    ref = helpers.reference_arrays(s)
    ctx = abi.Context(emul_lib.load(), 0, 1)
    ctx.set_contig_records(s.contig_records)
    assert int(ctx.get_contig_segments().seg_chrom_index.max()) == 1
    with pytest.raises(abi.PtlError):
        ctx.set_reference(ref[:1])
    ctx.set_reference(ref)
    ctx2 = abi.Context(emul_lib.load(), 0, 1)
    ctx2.set_reference(ref[:1])
    with pytest.raises(abi.PtlError):
        ctx2.set_contig_records(s.contig_records)


def test_malformed_contig_records_are_refused():
    """ptl_prepare_contig_records / ptl_set_contig_records are public entry points: a CIGAR CSR that is not monotone, a contig id
    or reference index outside its table and a position outside the BAM range are reported before they index anything
    (found by tools/fuzz/fuzz_contig_records.py under AddressSanitizer)."""
    s = synth.make("tiny", seed=37, n_reads=10, junction_per_mb=20)
    L = lib.load()
    r0 = s.contig_records
    n = r0.n_records
    L.prepare_contig_records(r0)
    types = dict(abi.ContigRecordsC._fields_)
    for field, count, index, value in (("cigar_begin", n + 1, n // 2, 2**50), ("cigar_begin", n + 1, 1, 0 if int(r0.cigar_begin[1]) else None),
                                       ("contig_id", n, 0, r0.n_contigs + 7), ("tid", n, 0, r0.n_ref_chrom + 3), ("tid", n, 0, -2),
                                       ("pos", n, 0, -1), ("pos", n, 0, 2**31)):
        if value is None:
            continue
        arr = np.ctypeslib.as_array(getattr(r0, field), (count,)).copy()
        if field == "cigar_begin" and index == 1:
            arr[2] = 0                                    # (a later entry below an earlier one)
        else:
            arr[index] = value
        r = abi.ContigRecordsC.from_buffer_copy(r0)
        setattr(r, field, arr.ctypes.data_as(types[field]))
        with pytest.raises(abi.PtlError):
            L.prepare_contig_records(r)


def test_read_records_with_a_broken_cigar_csr_are_refused():
    """ptl_pack_batch* takes caller memory: a cigar_begin that is not a CSR of the CIGAR pool (an entry below its predecessor,
    beyond the pool, 2^64 - 1) is reported before it sizes a copy or a walk (found by tools/fuzz/fuzz_read_records.py under
    AddressSanitizer: the vector CIGAR walk read 32 bytes through a wild pointer)."""
    s = synth.make("tiny", seed=41, n_reads=200)
    L = lib.load()
    r0 = s.read_records
    n = r0.n_reads
    T = dict(lib.ReadRecordsC._fields_)
    lib.PackedBatch(L, r0, 0, n, s.contig_names, windows=True)
    for index, value in ((150, 2**64 - 1), (10, 0), (n - 1, int(r0.cigar_begin[n]) + 5)):
        arr = np.ctypeslib.as_array(r0.cigar_begin, (n + 1,)).copy()
        arr[index] = value
        r = lib.ReadRecordsC.from_buffer_copy(r0)
        r.cigar_begin = arr.ctypes.data_as(T["cigar_begin"])
        for w in (None, True):
            with pytest.raises(abi.PtlError, match="CSR"):
                lib.PackedBatch(L, r, 0, n, s.contig_names, windows=w)
        lib.PackedBatch(L, r, 0, min(index, 5), s.contig_names, windows=True)    # (a slice in front of the damage still packs)
