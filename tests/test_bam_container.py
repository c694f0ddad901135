"""BAM container around the assembled records (output half): ptl_bam_header + ptl_bgzf_compress are
host code written from the SAM specification (no htslib).  Checked against Python's own zlib/gzip (independent
implementation of the same RFC 1952 framing) and by parsing the result back."""
import ctypes as C
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

import helpers
from portello_b200 import abi, lib, synth

EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _fns():
    d = lib.load().dll
    d.ptl_bgzf_bound.restype, d.ptl_bgzf_bound.argtypes = C.c_uint64, [C.c_uint64]
    d.ptl_bgzf_compress.restype = C.c_int64
    d.ptl_bgzf_compress.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
    d.ptl_bam_header.restype = C.c_int64
    d.ptl_bam_header.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64]
    return d


def bgzf(data: bytes, level=6, threads=4, eof=True) -> bytes:
    d = _fns()
    src = np.frombuffer(data, np.uint8) if data else np.zeros(1, np.uint8)
    out = np.zeros(int(d.ptl_bgzf_bound(len(data))), np.uint8)
    n = d.ptl_bgzf_compress(src.ctypes.data, len(data), level, threads, int(eof), out.ctypes.data, out.size)
    assert n >= 0, n
    return out[:n].tobytes()


def bam_header(text: str, names, lens) -> bytes:
    d = _fns()
    arr = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
    ln = (C.c_uint64 * max(len(lens), 1))(*lens)
    out = np.zeros(len(text) + 64 + sum(len(n) + 16 for n in names), np.uint8)
    n = d.ptl_bam_header(text.encode(), len(names), arr, ln, out.ctypes.data, out.size)
    assert n >= 0, n
    return out[:n].tobytes()


def blocks(stream: bytes):
    """[(offset, bsize, payload)] of a BGZF stream, checking the fixed header fields of every block (SAM spec 4.1)."""
    out, at = [], 0
    while at < len(stream):
        assert stream[at: at + 4] == b"\x1f\x8b\x08\x04" and stream[at + 10: at + 16] == b"\x06\x00BC\x02\x00"
        bsize = struct.unpack_from("<H", stream, at + 16)[0] + 1
        crc, isize = struct.unpack_from("<II", stream, at + bsize - 8)
        payload = zlib.decompress(stream[at + 18: at + bsize - 8], -15)
        assert len(payload) == isize and zlib.crc32(payload) == crc and isize <= 0xff00
        out.append((at, bsize, payload))
        at += bsize
    assert at == len(stream)
    return out


@pytest.mark.parametrize("n", [0, 1, 0xff00 - 1, 0xff00, 0xff00 + 1, 1_000_003])
@pytest.mark.parametrize("kind", ["text", "random"])
def test_bgzf_round_trip(n, kind):
    rng = np.random.default_rng(n + 7)
    data = (rng.integers(65, 70, n, dtype=np.uint8) if kind == "text" else rng.integers(0, 256, n, dtype=np.uint8)).tobytes()  # random = incompressible
    z = bgzf(data, threads=3)
    assert gzip.decompress(z) == data                     # python's gzip reads concatenated members
    bl = blocks(z)
    assert z.endswith(EOF_BLOCK) and bl[-1][2] == b""     # the EOF marker block
    assert len(bl) == (n + 0xff00 - 1) // 0xff00 + 1
    assert b"".join(p for _, _, p in bl) == data
    assert bgzf(data, threads=1) == z                     # thread count does not change the bytes


def test_bgzf_levels_and_no_eof():
    data = bytes(range(256)) * 999
    assert gzip.decompress(bgzf(data, level=1)) == data and gzip.decompress(bgzf(data, level=9)) == data
    z = bgzf(data, eof=False)
    assert not z.endswith(EOF_BLOCK) and gzip.decompress(z) == data


def test_bam_header_layout():
    text = "@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:chr1\tLN:1000\n@SQ\tSN:chrUn_x\tLN:77\n"
    h = bam_header(text, ["chr1", "chrUn_x"], [1000, 77])
    assert h[:4] == b"BAM\x01"
    (l_text,) = struct.unpack_from("<i", h, 4)
    assert h[8: 8 + l_text].decode() == text
    at = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", h, at)
    at += 4
    got = []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", h, at)
        name = h[at + 4: at + 4 + l_name]
        (l_ref,) = struct.unpack_from("<i", h, at + 4 + l_name)
        got.append((name, l_ref))
        at += 8 + l_name
    assert got == [(b"chr1\0", 1000), (b"chrUn_x\0", 77)] and at == len(h)


def test_oracle_records_to_bam_file(tmp_path):
    """Whole output path on the CPU checker: lift -> records -> header + BGZF -> a .bam file that parses back record by record."""
    from test_assemble_records import make_extras
    s = synth.make("tiny", seed=5, n_reads=400)
    pb = helpers.pack(s)
    octx = helpers.oracle_context(s)
    res = helpers.lift_c(octx, pb.c)
    x, names, _ = make_extras(s, pb, 3)
    octx.set_names(s.contig_names, s.chrom_names)
    _, (rb, by) = octx.assemble_records(x)
    lens = [int(s.chrom_len[i]) for i in range(s.n_chrom)]
    text = "@HD\tVN:1.6\tSO:unsorted\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in zip(s.chrom_names, lens))
    path = tmp_path / "out.bam"
    path.write_bytes(bgzf(bam_header(text, s.chrom_names, lens) + by.tobytes()))
    raw = gzip.open(path, "rb").read()
    assert raw[:4] == b"BAM\x01"
    at = 8 + struct.unpack_from("<i", raw, 4)[0]
    n_ref = struct.unpack_from("<i", raw, at)[0]
    at += 4
    for _ in range(n_ref):
        at += 8 + struct.unpack_from("<i", raw, at)[0]
    k = 0
    while at < len(raw):
        (block_size,) = struct.unpack_from("<I", raw, at)
        tid, pos = struct.unpack_from("<ii", raw, at + 4)
        assert (tid, pos) == (int(res.rec_tid[k]), int(res.rec_pos[k]))
        at += 4 + block_size
        k += 1
    assert k == res.n_records and at == len(raw)


def bgzf_level0_model(data: bytes, eof: bool) -> bytes:
    """Level-0 BGZF as htslib writes it: per <= 0xff00 bytes one gzip member holding ONE stored deflate block."""
    out = []
    for at in range(0, len(data), 0xff00):
        d = data[at: at + 0xff00]
        bsize = len(d) + 31
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize - 1) + b"\x01" + struct.pack("<HH", len(d), ~len(d) & 0xffff)
                   + d + struct.pack("<II", zlib.crc32(d), len(d)))
    return b"".join(out) + (EOF_BLOCK if eof else b"")


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["tiny", "config1", "one-read"])
def test_gpu_bgzf_store_matches_model(case):
    """ptl_bgzf_store_records (CRC32 + framing on the device) against the Python model and python's own gzip reader."""
    from test_assemble_records import make_extras
    s = synth.make("tiny" if case != "config1" else "config1", seed=6, **({"n_reads": 600} if case == "config1" else {"n_reads": 1 if case == "one-read" else 700}))
    pb = helpers.pack(s)
    gctx = helpers.gpu_context(s)
    helpers.lift_c(gctx, pb.c, allow_panic=True)
    x, _, _ = make_extras(s, pb, 4)
    gctx.set_names(s.contig_names, s.chrom_names)
    _, (rb, by) = gctx.assemble_records(x)
    lens = [int(s.chrom_len[i]) for i in range(s.n_chrom)]
    for prefix, eofs in ((b"", (False,)), (bam_header("@HD\tVN:1.6\tSO:unsorted\n", s.chrom_names, lens), (True,)), (b"x" * 12345, (False, True))):
        for eof in eofs:
            o, z = gctx.bgzf_store_records(prefix, flags=abi.BGZF_EOF if eof else 0)
            want = prefix + by.tobytes()
            assert gzip.decompress(z) == want if want else z == (EOF_BLOCK if eof else b"")
            assert z == bgzf_level0_model(want, eof)
            assert o.n_blocks == (len(want) + 0xff00 - 1) // 0xff00 and o.bytes_read == len(want)
    o2, none = gctx.bgzf_store_records(b"", flags=abi.ASM_NO_DOWNLOAD)
    assert none is None and o2.kernel_ms >= 0


def test_bgzf_level0_model_is_valid_gzip():
    data = bytes(np.random.default_rng(3).integers(0, 256, 200_001, dtype=np.uint8))
    assert gzip.decompress(bgzf_level0_model(data, True)) == data
    assert [len(p) for _, _, p in blocks(bgzf_level0_model(data, True))] == [0xff00, 0xff00, 0xff00, 200_001 - 3 * 0xff00, 0]


@pytest.mark.parametrize("n,misalign", [(1, 0), (100, 3), (0xff00 - 1, 0), (0xff00, 0), (0xff00, 5), (0xff00 + 1, 16), (200_001, 0), (200_001, 7)])
def test_emulated_bgzf_kernel_matches_model(n, misalign):
    """bgzf_store_kernel (slice CRCs through per-lane nibble tables, GF(2) shift-table combine tree, destination-aligned copy)
    run by 256 host threads in lock step (tests/emul), against the Python model and python's gzip: CPU coverage of the CRC
    arithmetic; `misalign` moves the stream off its 16-byte alignment (the funnel-shift load path)."""
    import emul_lib
    E = emul_lib.load().dll
    E.ptl_emul_bgzf_store.restype = C.c_int64
    E.ptl_emul_bgzf_store.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64]
    rng = np.random.default_rng(n + misalign)
    buf = np.zeros(n + 64, np.uint8)
    off = (16 - buf.ctypes.data % 16) % 16 + misalign
    data = rng.integers(0, 256, n, dtype=np.uint8)
    buf[off: off + n] = data
    out = np.zeros(n + 31 * ((n + 0xff00 - 1) // 0xff00) + 28 + 64, np.uint8)
    got = E.ptl_emul_bgzf_store(buf.ctypes.data + off, n, 1, out.ctypes.data + 1, out.size - 1)  # (odd destination on purpose)
    z = out[1: 1 + got].tobytes()
    assert z == bgzf_level0_model(data.tobytes(), True)
    assert gzip.decompress(z) == data.tobytes()


def test_container_errors_are_negative():
    """include/portello_b200.h: both calls return the bytes written or a NEGATIVE PTL_ERR_* code (a positive error code would
    read as a 1-byte success)."""
    d = _fns()
    data = np.frombuffer(os.urandom(200_000), np.uint8)
    out = np.zeros(int(d.ptl_bgzf_bound(data.size)), np.uint8)
    assert d.ptl_bgzf_compress(data.ctypes.data, data.size, 6, 2, 1, out.ctypes.data, 1000) < 0          # cap too small
    n_ok = d.ptl_bgzf_compress(data.ctypes.data, data.size, 6, 2, 1, out.ctypes.data, out.size)
    assert n_ok > 28
    assert d.ptl_bgzf_compress(data.ctypes.data, data.size, 6, 2, 1, out.ctypes.data, n_ok - 1) < 0      # no room for the EOF marker
    assert d.ptl_bgzf_compress(data.ctypes.data, data.size, 11, 2, 1, out.ctypes.data, out.size) < 0     # bad level
    names = (C.c_char_p * 1)(b"chr1")
    hb = np.zeros(64, np.uint8)
    assert d.ptl_bam_header(b"@HD\n", 1, names, (C.c_uint64 * 1)(1 << 31), hb.ctypes.data, hb.size) < 0  # l_ref > int32
    assert d.ptl_bam_header(b"@HD\n", 1, names, (C.c_uint64 * 1)(1000), hb.ctypes.data, 8) < 0           # cap too small
    assert d.ptl_bam_header(b"@HD\n", 1, names, (C.c_uint64 * 1)(1000), hb.ctypes.data, hb.size) > 0
