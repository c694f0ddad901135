"""BAM / BGZF / BAI input without htslib (input half) and Phase A on real files (rank 3): a synthetic
data set is written as the three files portello takes (FASTA, contig->reference BAM, read->contig BAM, both indexed), read
back through ptl_bam_* / ptl_fasta_* / ptl_scan_contig_bam and compared with the in-memory records it was written from;
region fetches are compared with a brute-force overlap filter (what IndexedReader::fetch + read returns,
src/read_alignment_scanner.rs:382-393), and the files themselves are parsed by an independent pure-Python reader
(gzip module + struct) so that writer and reader cannot agree on a private dialect."""
import gzip
import os
import struct

import numpy as np
import pytest

import helpers
from portello_b200 import abi, bamio, lib, synth


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    s = synth.make("tiny", seed=17, n_reads=2500)
    paths = bamio.write_dataset(s, str(tmp_path_factory.mktemp("ds")), n_unmapped=25)
    return s, paths


def g(p, k):
    return np.ctypeslib.as_array(p, (k,)) if k else np.zeros(0)


def ref_len_of(cigar):
    return int(sum(int(c) >> 4 for c in cigar if (int(c) & 15) in (0, 2, 3, 7, 8)))


def python_bam_records(path):
    """Independent reader: gzip (RFC 1952 multi-member) + struct, SAM spec 4.2."""
    raw = gzip.open(path, "rb").read()
    assert raw[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<i", raw, 4)[0]
    at = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, at)[0]
    at += 4
    refs = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", raw, at)[0]
        name = raw[at + 4: at + 4 + ln - 1].decode()
        refs.append((name, struct.unpack_from("<i", raw, at + 4 + ln)[0]))
        at += 8 + ln
    recs = []
    while at < len(raw):
        bs = struct.unpack_from("<i", raw, at)[0]
        tid, pos, l_name, mapq, bin_, n_cig, flag, l_seq, mtid, mpos, tlen = struct.unpack_from("<iiBBHHHiiii", raw, at + 4)
        o = at + 36
        name = raw[o: o + l_name - 1]
        o += l_name
        cigar = np.frombuffer(raw, np.uint32, n_cig, o)
        o += 4 * n_cig
        seq = raw[o: o + (l_seq + 1) // 2]
        o += (l_seq + 1) // 2
        qual = raw[o: o + l_seq]
        o += l_seq
        aux = raw[o: at + 4 + bs]
        recs.append(dict(tid=tid, pos=pos, name=name, mapq=mapq, bin=bin_, flag=flag, cigar=cigar, seq=seq, qual=qual, aux=aux, raw=raw[at: at + 4 + bs]))
        at += 4 + bs
    return raw[8: 8 + l_text].decode(), refs, recs


def test_header_and_every_record_roundtrip(dataset):
    s, paths = dataset
    f = bamio.BamFile(paths["reads"])
    assert f.has_index and f.has_eof_marker
    assert f.ref_names == s.contig_names and f.ref_len == [int(x) for x in s.contig_lengths()]
    text, refs, py = python_bam_records(paths["reads"])
    assert [r[0] for r in refs] == f.ref_names and text == f.header_text and text.startswith("@HD")
    d = f.fetch(bamio.FETCH_ALL)
    a = d.arrays()
    rr = s.read_records
    n = rr.n_reads
    assert d.n == n + 25 == len(py)
    for k, ptr in (("tid", rr.tid), ("pos", rr.pos), ("flag", rr.flag), ("mapq", rr.mapq), ("seq_len", rr.seq_len), ("bin", rr.bin)):
        assert np.array_equal(a[k][:n], g(ptr, n)), k
    cb = g(rr.cigar_begin, n + 1)
    assert np.array_equal(a["cigar_begin"][:n + 1], cb) and np.array_equal(a["cigar"][:int(cb[n])], g(rr.cigar, int(cb[n])))
    assert np.array_equal(a["seq4"][:int(rr.seq4_bytes)], g(rr.seq4, int(rr.seq4_bytes)))
    assert a["sa"][:n] == [rr.sa_tag[i] for i in range(n)] and any(a["sa"])
    # against the independent reader: names, qualities, aux blocks, mate fields
    for i in (0, 1, n // 2, n - 1, n, n + 24):
        p = py[i]
        assert bytes(a["names"][int(a["name_off"][i]): int(a["name_off"][i + 1])]) == p["name"]
        assert bytes(a["aux"][int(a["aux_off"][i]): int(a["aux_off"][i + 1])]) == p["aux"]
        assert bytes(a["qual"][int(a["qual_off"][i]): int(a["qual_off"][i]) + int(a["seq_len"][i])]) == p["qual"]
        assert int(a["mate_tid"][i]) == -1 and int(a["tlen"][i]) == 0
    assert all(p["bin"] == (4680 if p["flag"] & 4 else lib.load().reg2bin(p["pos"], p["pos"] + max(ref_len_of(p["cigar"]), 1))) for p in py[::97])


def test_region_fetch_equals_brute_force(dataset):
    s, paths = dataset
    f = bamio.BamFile(paths["reads"])
    a = f.fetch(bamio.FETCH_ALL).arrays()
    n = s.read_records.n_reads
    cb = a["cigar_begin"]
    ends = a["pos"][:n] + np.array([max(ref_len_of(a["cigar"][int(cb[i]): int(cb[i + 1])]), 1) for i in range(n)])
    rng = np.random.default_rng(1)
    seen = 0
    for k in range(60):
        t = int(rng.integers(0, len(f.ref_names)))
        L = f.ref_len[t]
        b = int(rng.integers(0, L))
        e = min(L, b + int(rng.integers(1, 60000))) if k % 5 else L  # some windows reach the contig end
        got = f.fetch(t, b, e).arrays()
        want = np.flatnonzero((a["tid"][:n] == t) & (a["pos"][:n] < e) & (ends > b))
        assert np.array_equal(got["pos"], a["pos"][want]) and np.array_equal(got["flag"], a["flag"][want]), (t, b, e)
        got2 = f.fetch(t, b, e, bamio.START_IN_REGION | bamio.SKIP_SUPPLEMENTARY).arrays()
        want2 = np.flatnonzero((a["tid"][:n] == t) & (a["pos"][:n] < e) & (a["pos"][:n] >= b) & ((a["flag"][:n] & 0x800) == 0))
        assert np.array_equal(got2["pos"], a["pos"][want2])
        seen += len(want)
    assert seen > 1000
    # the reference's windows tile every contig: each record is fetched by exactly one START_IN_REGION window
    L = lib.load()
    total = 0
    for t, ln in enumerate(f.ref_len):
        for b, e in L.region_segments(ln, 100_000):
            total += f.fetch(t, b, e, bamio.START_IN_REGION).n
    assert total == n


def test_unmapped_pass_through_bytes(dataset):
    s, paths = dataset
    f = bamio.BamFile(paths["reads"])
    u = f.fetch(bamio.FETCH_UNMAPPED, flt=bamio.ONLY_UNMAPPED | bamio.KEEP_RAW)
    assert u.n == 25
    raw, off = u.raw()
    _, _, py = python_bam_records(paths["reads"])
    want = b"".join(p["raw"] for p in py if p["flag"] & 4)
    assert raw.tobytes() == want and int(off[-1]) == len(want)
    # framed again by the host container code it is a readable BAM stream
    hdr = helpers.bam_header_bytes("@HD\tVN:1.6\tSO:unsorted\n", s.chrom_names, [int(s.chrom_len[i]) for i in range(s.n_chrom)])
    out = gzip.decompress(bamio.bgzf_compress(hdr + raw.tobytes()))
    assert out == hdr + want


def test_contig_scan_equals_in_memory_records(dataset):
    """ptl_scan_contig_bam (the record loop of scan_contig_bam) on the written file == the generator's contig records,
    and the segments prepared from either are identical (trim + join included)."""
    s, paths = dataset
    fc = bamio.BamFile(paths["contigs"])
    assert fc.ref_names == s.chrom_names
    sc = fc.scan_contigs(s.contig_names, s.contig_lengths())
    c, o = sc.c, s.contig_records
    assert c.n_records == o.n_records and c.n_contigs == o.n_contigs

    def key(R):
        out = []
        for i in range(R.n_records):
            n_c = int(R.cigar_begin[i + 1] - R.cigar_begin[i])
            cig = bytes(np.ctypeslib.as_array(R.cigar, (int(R.cigar_begin[R.n_records]),))[int(R.cigar_begin[i]): int(R.cigar_begin[i]) + n_c])
            seq = bytes(np.ctypeslib.as_array(R.seq[i], (int(R.contig_len[R.contig_id[i]]),))) if R.seq[i] else b""
            out.append((int(R.tid[i]), int(R.pos[i]), int(R.contig_id[i]), int(R.flag[i]), int(R.mapq[i]), cig, R.sa_tag[i] or b"", seq))
        return sorted(out)

    assert key(c) == key(o)
    L = lib.load()
    a, b = L.prepare_contig_records(c), L.prepare_contig_records(o)
    for fld in ("contig_seg_begin", "seg_seq_order_start", "seg_seq_order_end", "seg_chrom_index", "seg_pos", "seg_is_fwd", "seg_mapq", "seg_cigar_begin", "cigar"):
        assert np.array_equal(getattr(a, fld), getattr(b, fld)), fld


def test_fasta_loader(dataset, tmp_path):
    s, paths = dataset
    fa = bamio.load_fasta(paths["ref"])
    assert [x[0] for x in fa] == s.chrom_names
    assert all(np.array_equal(x[1], y) for x, y in zip(fa, s.reference_arrays()))
    p = tmp_path / "mixed.fa"
    p.write_bytes(b">chrA first one\nacgtNNryACGT\r\nACGT\n>chrB\n\n>chrC\tx\nttt")
    got = bamio.load_fasta(str(p), threads=2)
    assert [(n, bytes(q)) for n, q in got] == [("chrA", b"ACGTNNRYACGTACGT"), ("chrB", b""), ("chrC", b"TTT")]


def test_long_cigar_travels_in_cg_tag(tmp_path):
    """A contig->reference record with more than 65535 CIGAR ops is written as the `<l_seq>S<ref_len>N` placeholder +
    CG:B,I (SAM spec 4.2.2; what htslib writes) and restored on decode (what htslib's reader does)."""
    s = synth.make("tiny", seed=5, n_chrom=1, chrom_len=9_000_000, haplotypes=1, contigs_per_chrom=1, contig_snv_rate=6e-3, junction_per_mb=0.0, unmapped_contig_frac=0.0,
                   sv_per_mb=0.0, n_reads=10)
    o = s.contig_records
    n_ops = [int(o.cigar_begin[i + 1] - o.cigar_begin[i]) for i in range(o.n_records)]
    assert max(n_ops) > 65535
    paths = bamio.write_dataset(s, str(tmp_path))
    _, _, py = python_bam_records(paths["contigs"])
    big = [p for p in py if len(p["cigar"]) == 2 and b"CGBI" in p["aux"]]
    assert big and (int(big[0]["cigar"][0]) & 15) == 4 and (int(big[0]["cigar"][1]) & 15) == 3
    sc = bamio.BamFile(paths["contigs"]).scan_contigs(s.contig_names, s.contig_lengths())
    assert sorted(int(sc.c.cigar_begin[i + 1] - sc.c.cigar_begin[i]) for i in range(sc.c.n_records)) == sorted(n_ops)


def test_errors_are_reported_not_fatal(tmp_path, dataset):
    s, paths = dataset
    with pytest.raises(abi.PtlError):
        bamio.BamFile(str(tmp_path / "missing.bam"))
    p = tmp_path / "garbage.bam"
    p.write_bytes(b"this is not a BGZF stream at all" * 10)
    with pytest.raises(abi.PtlError):
        bamio.BamFile(str(p))
    # an unindexed copy: opens, but region fetch is refused (cli.rs:143-170 insists on the index)
    q = tmp_path / "noindex.bam"
    q.write_bytes(open(paths["reads"], "rb").read())
    f = bamio.BamFile(str(q))
    assert not f.has_index
    with pytest.raises(abi.PtlError):
        f.fetch(0, 0, 1000)
    # a truncated file (no EOF marker, last block cut)
    t = tmp_path / "trunc.bam"
    t.write_bytes(open(paths["reads"], "rb").read()[:-100])
    ft = bamio.BamFile(str(t))
    assert not ft.has_eof_marker
    with pytest.raises(abi.PtlError):
        ft.fetch(bamio.FETCH_ALL)
    # a contig name that the read->assembly header does not know (HashMap index panic, mod.rs:219-220)
    fc = bamio.BamFile(paths["contigs"])
    with pytest.raises(abi.PtlError):
        fc.scan_contigs(s.contig_names[:-1] + ["someone_else"], s.contig_lengths())


def _python_csi(path):
    """Independent CSIv1 reader (gzip + struct): geometry, and per reference {bin: (loffset, chunks)}."""
    raw = gzip.open(path, "rb").read()
    assert raw[:4] == b"CSI\x01"
    min_shift, depth, l_aux = struct.unpack_from("<iii", raw, 4)
    at = 16 + l_aux
    n_ref = struct.unpack_from("<i", raw, at)[0]
    at += 4
    refs = []
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", raw, at)[0]
        at += 4
        bins = {}
        for _ in range(n_bin):
            b, loff, n_chunk = struct.unpack_from("<IQi", raw, at)
            at += 16
            bins[b] = (loff, [struct.unpack_from("<QQ", raw, at + 16 * c) for c in range(n_chunk)])
            at += 16 * n_chunk
        refs.append(bins)
    n_no_coor = struct.unpack_from("<Q", raw, at)[0] if at + 8 <= len(raw) else None
    return min_shift, depth, refs, n_no_coor


@pytest.mark.parametrize("min_shift,depth", [(14, 5), (14, 0), (12, 6), (16, 3)], ids=["bai-geometry", "auto-depth", "fine", "coarse"])
def test_csi_index_fetches_what_the_bai_fetches(dataset, tmp_path, min_shift, depth):
    """A .csi (CSIv1: BGZF-compressed, any bin geometry, a lower-bound offset per bin instead of the linear index) built by
    ptl_bam_index_build_csi must drive the same region / unmapped fetches as the .bai, for the BAI's own geometry and others;
    the file itself is read back by an independent parser: every bin is the smallest one of its geometry that holds some
    record, the meta pseudo-bin carries the counts, and each loffset is a valid lower bound."""
    s, paths = dataset
    with_bai = bamio.BamFile(paths["reads"])
    p = tmp_path / "reads.bam"
    p.write_bytes(open(paths["reads"], "rb").read())
    assert not bamio.BamFile(str(p)).has_index
    bamio.index_bam_csi(str(p), None, min_shift, depth)
    f = bamio.BamFile(str(p))
    assert f.has_index
    # ---- the file
    ms, dp, refs, n_no_coor = _python_csi(str(p) + ".csi")
    assert ms == min_shift and (dp == depth if depth > 0 else (1 << (ms + 3 * dp)) >= max(f.ref_len) and (dp == 1 or (1 << (ms + 3 * (dp - 1))) < max(f.ref_len)))
    a = with_bai.fetch(bamio.FETCH_ALL).arrays()
    n = s.read_records.n_reads
    cb = a["cigar_begin"]
    ends = a["pos"][:n] + np.array([max(ref_len_of(a["cigar"][int(cb[i]): int(cb[i + 1])]), 1) for i in range(n)])
    meta = ((1 << (3 * dp + 3)) - 1) // 7 + 1
    assert n_no_coor == 25

    def smallest_bin(b, e):
        e -= 1
        first, sh = ((1 << (3 * dp)) - 1) // 7, ms
        for l in range(dp, 0, -1):
            if b >> sh == e >> sh:
                return first + (b >> sh)
            first -= 1 << (3 * (l - 1))
            sh += 3
        return 0
    for t, bins in enumerate(refs):
        mine = np.flatnonzero(a["tid"][:n] == t)
        want_bins = {smallest_bin(int(a["pos"][i]), int(ends[i])) for i in mine}
        assert set(bins) - {meta} == want_bins
        if len(mine):
            assert bins[meta][1][1] == (int(np.sum((a["flag"][mine] & 4) == 0)), int(np.sum((a["flag"][mine] & 4) != 0)))
            lo = bins[meta][1][0][0]
            assert all(lo <= loff for b, (loff, _) in bins.items() if b != meta)   # (nothing points in front of the reference's first record)
    # ---- the fetches
    rng = np.random.default_rng(2)
    seen = 0
    for k in range(40):
        t = int(rng.integers(0, len(f.ref_names)))
        L = f.ref_len[t]
        b = int(rng.integers(0, L))
        e = min(L, b + int(rng.integers(1, 80000))) if k % 4 else L
        got, want = f.fetch(t, b, e).arrays(), with_bai.fetch(t, b, e).arrays()
        assert np.array_equal(got["pos"], want["pos"]) and np.array_equal(got["flag"], want["flag"]) and got["names"].tobytes() == want["names"].tobytes()
        idx = np.flatnonzero((a["tid"][:n] == t) & (a["pos"][:n] < e) & (ends > b))
        assert np.array_equal(got["pos"], a["pos"][idx])
        seen += len(idx)
    assert seen > 500
    u = f.fetch(bamio.FETCH_UNMAPPED, flt=bamio.ONLY_UNMAPPED | bamio.KEEP_RAW)
    u0 = with_bai.fetch(bamio.FETCH_UNMAPPED, flt=bamio.ONLY_UNMAPPED | bamio.KEEP_RAW)
    assert u.n == 25 and u.raw()[0].tobytes() == u0.raw()[0].tobytes()


def test_csi_geometry_too_small_is_refused(dataset, tmp_path):
    s, paths = dataset
    p = tmp_path / "reads.bam"
    p.write_bytes(open(paths["reads"], "rb").read())
    with pytest.raises(abi.PtlError):
        bamio.index_bam_csi(str(p), None, 10, 1)      # 8 k bases per reference: the records do not fit
    with pytest.raises(abi.PtlError):
        bamio.index_bam_csi(str(p), None, 40, 2)


def test_corrupted_bam_and_index_files_error_out_cleanly(dataset, tmp_path):
    """Random corruption of the BAM (BGZF headers, record lengths, header fields), the .bai and the .csi (bytes flipped,
    files cut short, 32-bit fields set to extreme values): open / region fetch / unmapped fetch / full scan must either work
    or report an error -- never crash or hang.  The mutations run in one child process, so a crash fails this test only."""
    import subprocess
    import sys
    s, paths = dataset
    code = r'''
import gzip, os, random, sys
sys.path.insert(0, sys.argv[1])
from portello_b200 import abi, bamio
bam, work = sys.argv[2], sys.argv[3]
bamio.index_bam_csi(bam, work + ".orig.csi", 14, 5)
orig = [open(bam, "rb").read(), open(bam + ".bai", "rb").read(), open(work + ".orig.csi", "rb").read()]
rng = random.Random(7)
n_err = n_ok = 0
for it in range(60):
    kind = it % 3
    for f in (work, work + ".bai", work + ".csi"):
        if os.path.exists(f):
            os.remove(f)
    tgt = bytearray(gzip.decompress(orig[2])) if kind == 2 else bytearray(orig[kind])
    mode = rng.random()
    if mode < 0.6:
        for _ in range(rng.randint(1, 4)):
            tgt[rng.randrange(len(tgt))] = rng.randrange(256)
    elif mode < 0.8:
        del tgt[rng.randrange(len(tgt)):]
    else:
        p = rng.randrange(max(1, len(tgt) - 8))
        tgt[p:p + 4] = (0xffffffff if rng.random() < 0.5 else 0x7fffffff).to_bytes(4, "little")
    open(work, "wb").write(bytes(tgt) if kind == 0 else orig[0])
    if kind == 2:
        open(work + ".csi", "wb").write(bamio.bgzf_compress(bytes(tgt)))
    else:
        open(work + ".bai", "wb").write(bytes(tgt) if kind == 1 else orig[1])
    print("case", it, kind, flush=True)
    try:
        f = bamio.BamFile(work)
        if f.has_index:
            for t in range(len(f.ref_names)):
                f.fetch(t, 0, f.ref_len[t]).n
                f.fetch(t, 1000, 50000, bamio.START_IN_REGION).n
            f.fetch(bamio.FETCH_UNMAPPED, flt=bamio.ONLY_UNMAPPED).n
        f.fetch(bamio.FETCH_ALL).n
        n_ok += 1
    except abi.PtlError:
        n_err += 1
print("done", n_ok, n_err)
'''
    r = subprocess.run([sys.executable, "-c", code, os.path.dirname(os.path.dirname(os.path.abspath(__file__))), paths["reads"], str(tmp_path / "w.bam")],
                       capture_output=True, text=True, timeout=600)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    assert r.returncode == 0 and tail.startswith("done"), (r.returncode, tail, r.stderr[-1500:])
    n_ok, n_err = int(tail.split()[1]), int(tail.split()[2])
    assert n_ok + n_err == 60 and n_err >= 10      # (corruption is noticed; some mutations land in bytes that do not matter)


def test_contig_record_with_another_base_count_than_its_contig_is_refused(dataset):
    """The contig preparation reads contig_len bases of every primary contig record (rev_contig_seq, mod.rs:113-125): the scan
    must refuse a record whose sequence has another length instead of letting it read past the record."""
    s, paths = dataset
    f = bamio.BamFile(paths["contigs"])
    lens = np.asarray(s.contig_lengths(), np.uint64)
    f.scan_contigs(s.contig_names, lens)                      # consistent: fine
    with pytest.raises(abi.PtlError, match="bases"):
        f.scan_contigs(s.contig_names, lens + np.uint64(7))   # the read-to-assembly header claims longer contigs
