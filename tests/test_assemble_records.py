"""Record assembly, whole BAM records (ptl_assemble_records): every output record as the bytes
bam_write1 would emit after clone_record + the field updates / push_aux calls of
get_liftover_alignment_for_read_and_contig_segment (src/read_alignment_scanner.rs:105-117,245-282) and
finish_remapped_alignment_set (:310-366).

The reference has no test for record mutation ("parity unpinned"), so the oracle (C++ restatement) is
pinned here by an independent pure-Python model that builds every record from the SAM-spec layout, and the CUDA kernels
must equal the oracle byte for byte."""
import struct

import numpy as np
import pytest

import helpers
from portello_b200 import abi, synth
from test_assemble import COMP, iupac_case, make_quals

OPS = "MIDNSHP=X"


def aux_fields(aux: bytes):
    """[(tag, start, end)] of a well-formed prefix of a BAM aux block (SAM spec 4.2.4)."""
    out, i, n = [], 0, len(aux)
    size = {b"A": 1, b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4, b"d": 8}
    while i + 3 <= n:
        t = aux[i + 2: i + 3]
        if t in size:
            e = i + 3 + size[t]
        elif t in (b"Z", b"H"):
            j = aux.find(b"\0", i + 3)
            if j < 0:
                break
            e = j + 1
        elif t == b"B":
            if i + 8 > n or aux[i + 3: i + 4] not in (b"c", b"C", b"s", b"S", b"i", b"I", b"f"):
                break
            es = {b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4}[aux[i + 3: i + 4]]
            e = i + 8 + es * struct.unpack_from("<I", aux, i + 4)[0]
        else:
            break
        if e > n:
            break
        out.append((aux[i: i + 2], i, e))
        i = e
    return out


def strip_tags(aux: bytes) -> bytes:
    """clone_record: drop the first NM, SA, PS and ZM field."""
    cuts = []
    for tag in (b"NM", b"SA", b"PS", b"ZM"):
        for t, a, e in aux_fields(aux):
            if t == tag:
                cuts.append((a, e))
                break
    keep, at = b"", 0
    for a, e in sorted(cuts):
        keep += aux[at:a]
        at = e
    return keep + aux[at:]


def make_extras(s, pb, seed):
    """Synthetic qnames, aux blocks (every field type, some NM/SA/PS/ZM tags at random places, a few malformed tails),
    mate fields and qualities for the reads of a packed batch."""
    rng = np.random.default_rng(seed)
    n = pb.c.n_reads
    names, auxs = [], []
    for r in range(n):
        names.append(("m64011_%d/%d/ccs" % (int(rng.integers(1, 10**6)), r)).encode()[: int(rng.integers(1, 40))])
        fields = []
        for _ in range(int(rng.integers(0, 9))):
            kind = int(rng.integers(0, 12))
            tag = bytes(rng.choice(list(b"abcdefghijXYZ"), 2).astype(np.uint8))
            if kind == 0: fields.append(tag + b"A" + b"Q")
            elif kind == 1: fields.append(tag + b"c" + struct.pack("<b", -5))
            elif kind == 2: fields.append(tag + b"S" + struct.pack("<H", 4711))
            elif kind == 3: fields.append(tag + b"i" + struct.pack("<i", -123456))
            elif kind == 4: fields.append(tag + b"f" + struct.pack("<f", 0.25))
            elif kind == 5: fields.append(tag + b"Z" + bytes(rng.integers(33, 126, int(rng.integers(0, 60)), dtype=np.uint8)) + b"\0")
            elif kind == 6: fields.append(tag + b"H" + b"1AE301" + b"\0")
            elif kind == 7: fields.append(tag + b"B" + b"C" + struct.pack("<I", 37) + bytes(rng.integers(0, 256, 37, dtype=np.uint8)))
            elif kind == 8: fields.append(tag + b"B" + b"f" + struct.pack("<I", 4) + struct.pack("<4f", 1, 2, 3, 4))
            elif kind == 9: fields.append(b"NM" + b"i" + struct.pack("<i", int(rng.integers(0, 500))))
            elif kind == 10: fields.append(rng.choice([b"SA", b"PS"]).tobytes() + b"Z" + b"ctg1,100,+,50M,60,0;" + b"\0")
            else: fields.append(b"ZM" + b"C" + bytes([int(rng.integers(0, 61))]))
        a = b"".join(fields)
        if rng.random() < 0.03:
            a += b"zz" + b"Z" + b"no terminator"  # malformed tail: nothing behind it is found, everything is kept
        auxs.append(a)
    name_off = np.zeros(n + 1, np.uint64)
    name_off[1:] = np.cumsum([len(x) for x in names])
    aux_off = np.zeros(n + 1, np.uint64)
    aux_off[1:] = np.cumsum([len(x) for x in auxs])
    qual, qual_off = make_quals(pb.c, seed)
    return dict(name_off=name_off, names=np.frombuffer(b"".join(names) + b"\0" * 8, np.uint8).copy(), aux_off=aux_off,
                aux=np.frombuffer(b"".join(auxs) + b"\0" * 8, np.uint8).copy(),
                mate_tid=rng.integers(-1, 3, n).astype(np.int32), mate_pos=rng.integers(-1, 10**6, n).astype(np.int32),
                tlen=rng.integers(-5000, 5000, n).astype(np.int32), qual=qual, qual_off=qual_off), names, auxs


def model(s, pb, res, x, names, auxs, contig_seg_is_fwd):
    """Pure-Python restatement: list of record byte strings (block_size included)."""
    b = pb.c
    seg_begin = np.ctypeslib.as_array(b.read_seg_begin, (b.n_reads + 1,))
    rseg_read = np.repeat(np.arange(b.n_reads), np.diff(seg_begin.astype(np.int64)))
    rseg_contig = np.ctypeslib.as_array(b.rseg_contig, (b.n_read_segments,))
    seq_len = np.ctypeslib.as_array(b.read_seq_len, (b.n_reads,))
    seq_off = np.ctypeslib.as_array(b.read_seq_off, (b.n_reads,))
    read_mapq = np.ctypeslib.as_array(b.read_mapq, (b.n_reads,))
    pool = np.ctypeslib.as_array(b.seq4, (int(b.seq4_bytes),))
    out = []
    for r in range(b.n_reads):
        k0, k1 = int(res.read_rec_begin[r]), int(res.read_rec_begin[r + 1])
        kept = strip_tags(auxs[r])
        sa = {}
        for k in range(k0, k1):
            if res.rec_status[k] == 1:
                cg = res.cigar[int(res.rec_cigar_begin[k]): int(res.rec_cigar_begin[k + 1])]
                sa[k] = "%s,%d,%s,%s,%d,0;" % (s.chrom_names[int(res.rec_tid[k])], int(res.rec_pos[k]) + 1, "-" if int(res.rec_flag[k]) & 16 else "+",
                                              "".join("%d%s" % (int(c) >> 4, OPS[int(c) & 15]) for c in cg), int(res.rec_mapq[k]))
        n = int(seq_len[r])
        raw = pool[int(seq_off[r]): int(seq_off[r]) + (n + 1) // 2]
        q = x["qual"][int(x["qual_off"][r]): int(x["qual_off"][r]) + n]
        for k in range(k0, k1):
            lifted = res.rec_status[k] == 1
            aux = kept
            cg = res.cigar[int(res.rec_cigar_begin[k]): int(res.rec_cigar_begin[k + 1])] if lifted else res.cigar[:0]
            if lifted:
                ctg = int(rseg_contig[int(res.rec_read_segment[k])])
                idx = int(res.rec_contig_segment[k])
                aux += b"PSZ" + ("%s_split%d%s" % (s.contig_names[ctg], idx, "+" if contig_seg_is_fwd(ctg, idx) else "-")).encode() + b"\0"
                aux += b"ZMC" + bytes([int(read_mapq[r])])
                text = "".join(sa[j] for j in range(k0, k1) if j != k)
                if text:
                    aux += b"SAZ" + text.encode() + b"\0"
            if res.rec_need_flip[k]:
                nib = np.empty(2 * len(raw), np.uint8)
                nib[0::2], nib[1::2] = raw >> 4, raw & 15
                rc = COMP[nib[:n][::-1]]
                if n & 1:
                    rc = np.append(rc, np.uint8(0))
                seq, qq = (rc[0::2] << 4 | rc[1::2]).astype(np.uint8).tobytes(), q[::-1].tobytes()
            else:
                seq, qq = raw.tobytes(), q.tobytes()
            core = struct.pack("<iiBBHHHIiii", int(res.rec_tid[k]), int(res.rec_pos[k]), len(names[r]) + 1, int(res.rec_mapq[k]), int(res.rec_bin[k]),
                               len(cg), int(res.rec_flag[k]), n, int(x["mate_tid"][r]), int(x["mate_pos"][r]), int(x["tlen"][r]))
            rec = core + names[r] + b"\0" + np.asarray(cg, "<u4").tobytes() + seq + qq + aux
            out.append(struct.pack("<I", len(rec)) + rec)
    return out


def seg_is_fwd_fn(ctx):
    segs = ctx.get_contig_segments()
    return lambda ctg, idx: bool(segs.seg_is_fwd[int(segs.contig_seg_begin[ctg]) + idx])


def check_against_model(ctx, s, pb, res, seed):
    x, names, auxs = make_extras(s, pb, seed)
    ctx.set_names(s.contig_names, s.chrom_names)
    _, (rb, by) = ctx.assemble_records(x)
    want = model(s, pb, res, x, names, auxs, seg_is_fwd_fn(ctx))
    assert len(want) == res.n_records == len(rb) - 1
    for k, w in enumerate(want):
        got = by[int(rb[k]): int(rb[k + 1])].tobytes()
        assert got == w, f"record {k}: {len(got)} vs {len(w)} bytes, first difference at {next((i for i, (a, c) in enumerate(zip(got, w)) if a != c), None)}"
    return x, (rb, by)


def test_strip_tags_model():
    aux = b"NMi\x05\0\0\0" + b"rqf\0\0\x80?" + b"SAZx,1,+,5M,1,0;\0" + b"NMi\x07\0\0\0" + b"ZMC\x09"
    assert strip_tags(aux) == b"rqf\0\0\x80?" + b"NMi\x07\0\0\0"  # only the FIRST NM goes (bam_aux_get + bam_aux_del)
    assert strip_tags(b"") == b""
    assert strip_tags(b"abZno terminator") == b"abZno terminator"


@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_records_match_python_model(seed):
    seen = np.zeros(3, bool)  # flipped records, unmapped fallbacks, reads with several records (SA tags)
    for s, pb in (iupac_case(seed, n_reads=300), (lambda s: (s, helpers.pack(s)))(synth.make("tiny", seed=seed, n_reads=1200))):
        octx = helpers.oracle_context(s)
        res = helpers.lift_c(octx, pb.c, allow_panic=True)
        seen |= [bool(res.rec_need_flip.any()), bool((res.rec_status == 0).any()), bool((np.diff(res.read_rec_begin.astype(np.int64)) > 1).any())]
        check_against_model(octx, s, pb, res, seed)
    assert seen.all(), seen


def test_oracle_records_need_names():
    s = synth.make("tiny", seed=3, n_reads=50)
    pb = helpers.pack(s)
    octx = helpers.oracle_context(s)
    helpers.lift_c(octx, pb.c)
    x, _, _ = make_extras(s, pb, 1)
    with pytest.raises(abi.PtlError):
        octx.assemble_records(x)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["iupac-odd", "tiny", "config1", "stress"])
def test_gpu_records_match_oracle(case):
    if case == "iupac-odd":
        s, pb = iupac_case(11, n_reads=2000)
    else:
        s = synth.make(case, **({"n_reads": 3000} if case == "config1" else {"n_reads": 300} if case == "stress" else {}))
        pb = helpers.pack(s)
    octx, gctx = helpers.oracle_context(s), helpers.gpu_context(s)
    ro = helpers.lift_c(octx, pb.c, allow_panic=True)
    rg = helpers.lift_c(gctx, pb.c, allow_panic=True)
    assert rg.diff(ro) is None
    x, _, _ = make_extras(s, pb, 9)
    for c in (octx, gctx):
        c.set_names(s.contig_names, s.chrom_names)
    _, (rbo, byo) = octx.assemble_records(x)
    og, (rbg, byg) = gctx.assemble_records(x)
    assert np.array_equal(rbo, rbg), "record offsets"
    if not np.array_equal(byo, byg):
        i = int(np.flatnonzero(byo != byg)[0])
        k = int(np.searchsorted(rbo, i, side="right") - 1)
        raise AssertionError(f"record {k} differs at byte {i - int(rbo[k])} of {int(rbo[k + 1] - rbo[k])}")
    assert og.kernel_ms > 0 and og.bytes_written == int(rbg[-1]) and 0 < og.bytes_read <= og.bytes_written
    o2, none = gctx.assemble_records(None, flags=abi.ASM_RESIDENT_QUAL | abi.ASM_NO_DOWNLOAD)
    assert none is None and o2.n_records == og.n_records and o2.kernel_ms > 0
    if case == "iupac-odd":
        check_against_model(gctx, s, pb, rg, 9)


@pytest.mark.gpu
def test_gpu_records_need_names():
    s = synth.make("tiny", seed=3, n_reads=50)
    pb = helpers.pack(s)
    gctx = helpers.gpu_context(s)
    helpers.lift_c(gctx, pb.c)
    x, _, _ = make_extras(s, pb, 1)
    with pytest.raises(abi.PtlError):
        gctx.assemble_records(x)


@pytest.mark.gpu
def test_gpu_records_properties_at_scale():
    """Size-independent properties on 60 k reads of configs[1] (no oracle involved): the record stream parses back as BAM
    records whose core fields are the liftover result, every CIGAR consumes exactly l_seq read bases, a flipped record's
    qualities are the input's reversed and flipping its bases twice gives the input back, unflipped records carry the
    input bytes, and the aux block is the surviving input tags followed by PS, ZM and (iff the read has several records) SA."""
    s = synth.make("chr20", n_reads=60_000)
    pb = helpers.pack(s)
    gctx = helpers.gpu_context(s)
    res = helpers.lift_c(gctx, pb.c)
    x, names, auxs = make_extras(s, pb, 5)
    gctx.set_names(s.contig_names, s.chrom_names)
    out, (rb, by) = gctx.assemble_records(x)
    assert out.n_records == res.n_records and int(rb[-1]) == by.size == out.bytes_written
    b = pb.c
    seg_begin = np.ctypeslib.as_array(b.read_seg_begin, (b.n_reads + 1,))
    seq_len = np.ctypeslib.as_array(b.read_seq_len, (b.n_reads,))
    seq_off = np.ctypeslib.as_array(b.read_seq_off, (b.n_reads,))
    pool = np.ctypeslib.as_array(b.seq4, (int(b.seq4_bytes),))
    rec_read = np.repeat(np.arange(b.n_reads), np.diff(res.read_rec_begin.astype(np.int64)))
    raw = by.tobytes()
    n_flip = n_sa = 0
    for k in range(0, res.n_records, 7):  # every 7th record (python-side parsing is the slow part)
        r = int(rec_read[k])
        o = int(rb[k])
        (block_size,) = struct.unpack_from("<I", raw, o)
        assert o + 4 + block_size == int(rb[k + 1])
        tid, pos, l_name, mapq, bin_, n_cig, flag, l_seq, mtid, mpos, tlen = struct.unpack_from("<iiBBHHHIiii", raw, o + 4)
        assert (tid, pos, mapq, bin_, flag) == (int(res.rec_tid[k]), int(res.rec_pos[k]), int(res.rec_mapq[k]), int(res.rec_bin[k]), int(res.rec_flag[k]))
        assert (mtid, mpos, tlen, l_seq) == (int(x["mate_tid"][r]), int(x["mate_pos"][r]), int(x["tlen"][r]), int(seq_len[r]))
        p = o + 36
        assert raw[p: p + l_name] == names[r] + b"\0"
        p += l_name
        cig = np.frombuffer(raw, "<u4", n_cig, p)
        assert np.array_equal(cig, res.cigar[int(res.rec_cigar_begin[k]): int(res.rec_cigar_begin[k + 1])])
        if res.rec_status[k] == 1:
            assert int(sum(int(c) >> 4 for c in cig if (int(c) & 15) in (0, 1, 4, 7, 8))) == l_seq  # M I S = X consume the read
        p += 4 * n_cig
        seq = np.frombuffer(raw, np.uint8, (l_seq + 1) // 2, p)
        p += (l_seq + 1) // 2
        qual = np.frombuffer(raw, np.uint8, l_seq, p)
        p += l_seq
        src_seq = pool[int(seq_off[r]): int(seq_off[r]) + (l_seq + 1) // 2]
        src_q = x["qual"][int(x["qual_off"][r]): int(x["qual_off"][r]) + l_seq]
        if res.rec_need_flip[k]:
            n_flip += 1
            assert np.array_equal(qual, src_q[::-1])
            nib = np.empty(2 * len(seq), np.uint8)
            nib[0::2], nib[1::2] = seq >> 4, seq & 15
            back = COMP[nib[:l_seq][::-1]]  # revcomp of the revcomp = the input (the synthetic reads are A/C/G/T)
            src_nib = np.empty(2 * len(src_seq), np.uint8)
            src_nib[0::2], src_nib[1::2] = src_seq >> 4, src_seq & 15
            assert np.array_equal(back, src_nib[:l_seq])
        else:
            assert np.array_equal(seq, src_seq) and np.array_equal(qual, src_q)
        aux = raw[p: int(rb[k + 1])]
        kept = strip_tags(auxs[r])
        assert aux.startswith(kept)
        tail = [t for t, _, _ in aux_fields(aux[len(kept):])]
        many = int(res.read_rec_begin[r + 1]) - int(res.read_rec_begin[r]) > 1
        assert tail == ([b"PS", b"ZM"] + ([b"SA"] if many else []) if res.rec_status[k] == 1 else [])
        n_sa += many
    assert n_flip > 0 and n_sa > 0


def _long_cigar_case():
    """One read whose lifted CIGAR has 70 000 ops (1M1I repeated): more than the u16 n_cigar_op of a BAM record."""
    n_pairs = 35_000
    cig = np.tile(np.array([(1 << 4) | 7, (1 << 4) | 1], np.uint32), n_pairs)  # 1= 1I ...
    seq_len = 2 * n_pairs
    segs, batch = helpers.single_pair_case(f"{n_pairs + 10}=", 100, True, n_pairs + 10, None, 5, cig, np.full(seq_len // 2, 0x12, np.uint8), seq_len)
    x = dict(name_off=np.array([0, 2], np.uint64), names=np.frombuffer(b"r1" + b"\0" * 8, np.uint8).copy(), aux_off=np.array([0, 0], np.uint64),
             aux=np.zeros(8, np.uint8), mate_tid=np.array([-1], np.int32), mate_pos=np.array([-1], np.int32), tlen=np.array([0], np.int32),
             qual=np.full(seq_len + 8, 30, np.uint8), qual_off=np.array([0], np.uint64))
    return segs, batch, x


def _check_long_cigar_goes_to_cg_tag(ctx, octx=None):
    """bam_write1 (htslib sam.c): n_cigar_op = 2, CIGAR field <l_seq>S<ref_len>N, the real ops in CG:B,I behind every other tag."""
    segs, batch, x = _long_cigar_case()
    want = None
    for c in (ctx, octx):
        if c is None:
            continue
        c.set_reference([np.frombuffer(b"A" * 40_000, np.uint8).copy()])
        c.set_contig_segments(segs)
        res = c.lift(batch)
        assert res.n_records == 1 and res.rec_status[0] == 1 and int(res.rec_cigar_begin[1]) == 70_000
        c.set_names(["ctg"], ["chr1"])
        _, (rb, by) = c.assemble_records(x)
        rec = bytes(by[int(rb[0]):int(rb[1])])
        if want is not None:
            assert rec == want
        want = rec
    ops = np.asarray(res.cigar[:70_000], np.uint32)
    block_size, tid, pos, l_name, mapq, bin_, n_cig, flag, l_seq = struct.unpack_from("<iiiBBHHHi", rec, 0)
    assert block_size == len(rec) - 4 and n_cig == 2 and l_seq == 70_000 and l_name == 3
    o = 36 + l_name
    fake = struct.unpack_from("<2I", rec, o)
    ref_len = int(sum(int(c) >> 4 for c in ops if (int(c) & 15) in (0, 2, 3, 7, 8)))
    assert fake == ((l_seq << 4) | 4, (ref_len << 4) | 3)
    o += 8 + (l_seq + 1) // 2 + l_seq
    aux = rec[o:]
    fields = aux_fields(aux)
    assert [t for t, _, _ in fields] == [b"PS", b"ZM", b"CG"]          # CG is the LAST tag
    cg = aux[-(8 + 4 * 70_000):]
    assert cg[:4] == b"CGBI" and struct.unpack_from("<I", cg, 4)[0] == 70_000
    assert np.array_equal(np.frombuffer(cg[8:], np.uint32), ops)
    return rec


def test_oracle_and_emulation_write_long_cigars_as_cg_tag():
    import emul_lib
    import oracle_lib
    _check_long_cigar_goes_to_cg_tag(abi.Context(emul_lib.load(), 0, 1), abi.Context(oracle_lib.load(), 0, 1))


@pytest.mark.gpu
def test_gpu_writes_long_cigars_as_cg_tag(tmp_path):
    """... and the record, framed by the device (both the two-pass and the fused writer), is restored by the BAM reader."""
    import oracle_lib
    from portello_b200 import bamio, lib
    gctx = lib.GpuContext(0, 1)
    rec = _check_long_cigar_goes_to_cg_tag(gctx, abi.Context(oracle_lib.load(), 0, 1))
    segs, batch, x = _long_cigar_case()
    header = helpers.bam_header_bytes("@HD\tVN:1.6\n", ["chr1"], [40_000])
    for fused in (False, True):
        gctx.lift(batch)
        if fused:
            _, stream = gctx.frame_records(x, header, flags=abi.BGZF_EOF)
        else:
            gctx.assemble_records(x)
            _, stream = gctx.bgzf_store_records(header, flags=abi.BGZF_EOF)
        path = tmp_path / f"long_{int(fused)}.bam"
        path.write_bytes(stream)
        d = bamio.BamFile(str(path)).fetch(bamio.FETCH_ALL).arrays()
        assert len(d["tid"]) == 1 and int(d["cigar_begin"][1]) == 70_000
        assert np.array_equal(d["cigar"], np.frombuffer(rec[-4 * 70_000:], np.uint32))
        assert b"CG" not in [t for t, _, _ in aux_fields(d["aux"].tobytes())]     # (the reader moves the tag back into the CIGAR)


@pytest.mark.parametrize("case", ["iupac-odd", "tiny"])
def test_emulated_record_kernels_match_oracle(case):
    """The device bodies of assemble_bam.cuh compiled for the host (tests/emul, one 'thread' at a time) against the oracle:
    the aux walk, record sizes, tag text, destination-aligned chunking and the flipped paths, without a GPU."""
    import emul_lib
    E = emul_lib.load()
    if case == "iupac-odd":
        s, pb = iupac_case(5, n_reads=250)
    else:
        s = synth.make("tiny", seed=8, n_reads=500)
        pb = helpers.pack(s)
    ectx = abi.Context(E, 0, 1)
    ectx.set_reference(helpers.reference_arrays(s))
    ectx.set_contig_records(s.contig_records)
    octx = helpers.oracle_context(s)
    re_ = helpers.lift_c(ectx, pb.c, allow_panic=True)
    ro = helpers.lift_c(octx, pb.c, allow_panic=True)
    assert re_.diff(ro) is None
    x, _, _ = make_extras(s, pb, 2)
    for c in (ectx, octx):
        c.set_names(s.contig_names, s.chrom_names)
    _, (rbo, byo) = octx.assemble_records(x)
    _, (rbe, bye) = ectx.assemble_records(x)
    assert np.array_equal(rbo, rbe)
    if not np.array_equal(byo, bye):
        i = int(np.flatnonzero(byo != bye)[0])
        k = int(np.searchsorted(rbo, i, side="right") - 1)
        raise AssertionError(f"record {k} differs at byte {i - int(rbo[k])} of {int(rbo[k + 1] - rbo[k])}")


def _long_name_rejected(ctx, s, pb):
    helpers.lift_c(ctx, pb.c)
    x, _, _ = make_extras(s, pb, 1)
    n = pb.c.n_reads
    x["name_off"] = np.arange(n + 1, dtype=np.uint64) * np.uint64(300)       # 300-byte qnames: l_read_name is a u8
    x["names"] = np.full(300 * n + 16, ord("q"), np.uint8)
    ctx.set_names(s.contig_names, s.chrom_names)
    with pytest.raises(abi.PtlError, match="254"):
        ctx.assemble_records(x)


def test_read_names_beyond_u8_are_rejected():
    import emul_lib
    s = synth.make("tiny", seed=3, n_reads=20)
    pb = helpers.pack(s)
    _long_name_rejected(helpers.oracle_context(s), s, pb)
    ectx = abi.Context(emul_lib.load(), 0, 1)
    ectx.set_reference(helpers.reference_arrays(s))
    ectx.set_contig_records(s.contig_records)
    _long_name_rejected(ectx, s, pb)


@pytest.mark.gpu
def test_gpu_read_names_beyond_u8_are_rejected():
    s = synth.make("tiny", seed=3, n_reads=20)
    pb = helpers.pack(s)
    _long_name_rejected(helpers.gpu_context(s), s, pb)


@pytest.mark.gpu
def test_gpu_malformed_extras_are_rejected_not_dereferenced():
    """The extras are caller memory: an offset outside its pool is PTL_ERR_INVALID_ARG, and the context survives (no sticky
    CUDA fault) -- the next well-formed call on the same slot succeeds."""
    s = synth.make("tiny", seed=3, n_reads=200)
    pb = helpers.pack(s)
    gctx = helpers.gpu_context(s)
    gctx.set_names(s.contig_names, s.chrom_names)
    helpers.lift_c(gctx, pb.c)
    x, _, _ = make_extras(s, pb, 1)
    n = pb.c.n_reads

    def broken(field, idx, value):
        y = dict(x)
        y[field] = np.array(x[field], copy=True)
        y[field][idx] = value
        return y

    for y in (broken("aux_off", 5, 1 << 40), broken("name_off", 7, int(x["name_off"][8]) + 1), broken("qual_off", n - 1, len(x["qual"]) - 3),
              broken("aux_off", 3, int(x["aux_off"][2]) - 1 if int(x["aux_off"][2]) else 1 << 33)):
        with pytest.raises(abi.PtlError) as e:
            gctx.assemble_records(y)
        assert e.value.code == abi.PTL_ERR_INVALID_ARG, e.value
    with pytest.raises(abi.PtlError):
        gctx.assemble_bases(x["qual"], broken("qual_off", 0, len(x["qual"]))["qual_off"])
    _, (rb, by) = gctx.assemble_records(x)  # still alive
    assert int(rb[-1]) == by.size > 0


@pytest.mark.gpu
def test_gpu_stale_extras_are_not_reused_for_a_new_batch():
    """PTL_ASM_RESIDENT_QUAL after ANOTHER batch on the slot must fail (names / aux / qualities belong to the old reads)."""
    s = synth.make("tiny", seed=4, n_reads=400)
    pa, pb2 = helpers.pack(s, 0, 150), helpers.pack(s, 150, 250)
    gctx = helpers.gpu_context(s)
    gctx.set_names(s.contig_names, s.chrom_names)
    helpers.lift_c(gctx, pa.c)
    x, _, _ = make_extras(s, pa, 1)
    gctx.assemble_records(x)
    gctx.assemble_records(None, flags=abi.ASM_RESIDENT_QUAL)
    helpers.lift_c(gctx, pb2.c)
    for call in (lambda: gctx.assemble_records(None, flags=abi.ASM_RESIDENT_QUAL), lambda: gctx.bgzf_store_records(b""),
                 lambda: gctx.assemble_bases(None, None, flags=abi.ASM_RESIDENT_QUAL)):
        with pytest.raises(abi.PtlError):
            call()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["iupac-odd", "tiny", "config1", "wg-flips"])
def test_gpu_frame_records_equals_assemble_then_store(case):
    """ptl_frame_records (record assembly fused with level-0 BGZF framing: the record stream is never materialised) must be
    byte-identical to ptl_assemble_records + ptl_bgzf_store_records, for every prefix length / EOF choice, and gunzip to
    the records."""
    import gzip
    if case == "iupac-odd":
        s, pb = iupac_case(13, n_reads=2500)   # records of 1..900 bases: many records per BGZF block
    elif case == "wg-flips":
        s = synth.make("tiny", seed=19, n_reads=900, rev_contig_frac=0.9, read_len_mean=9000, read_len_sd=2500, read_len_min=2000, read_len_max=15000)
        pb = helpers.pack(s)
    else:
        s = synth.make(case, **({"n_reads": 600} if case == "config1" else {}))
        pb = helpers.pack(s)
    gctx = helpers.gpu_context(s)
    gctx.set_names(s.contig_names, s.chrom_names)
    helpers.lift_c(gctx, pb.c, allow_panic=True)
    x, _, _ = make_extras(s, pb, 9)
    _, (rb, by) = gctx.assemble_records(x)
    hdr = helpers.bam_header_bytes("@HD\tVN:1.6\tSO:unsorted\n", s.chrom_names, [int(s.chrom_len[i]) for i in range(s.n_chrom)])
    for prefix, flags in ((b"", 0), (hdr, abi.BGZF_EOF), (b"x" * 12345, 0), (b"y" * 7, abi.BGZF_EOF)):
        gctx.assemble_records(x)
        _, want = gctx.bgzf_store_records(prefix, flags=flags)
        helpers.lift_c(gctx, pb.c, allow_panic=True)   # a fresh batch on the slot: nothing of the two-call path is left over
        z, got = gctx.frame_records(x, prefix, flags=flags)
        assert got == want, (case, len(prefix), flags, len(got), len(want), next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), None))
        assert gzip.decompress(got) == prefix + by.tobytes()
        assert z.n_blocks == (len(prefix) + by.size + 0xff00 - 1) // 0xff00 and z.kernel_ms > 0
    # resident inputs, no download: the timing mode of the bench
    z2, none = gctx.frame_records(None, b"", flags=abi.ASM_RESIDENT_QUAL | abi.ASM_NO_DOWNLOAD)
    assert none is None and z2.n_bytes > by.size
