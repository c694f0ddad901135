"""ctypes mirror of include/portello_b200.h.

Plumbing only: struct layouts + a thin `LiftLib` wrapper that drives any shared library exporting the ABI under a
given symbol prefix (`ptl_` = the CUDA product, `ptl_oracle_` = the CPU checker used by tests).  This module loads no
library itself and contains no compute.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

PTL_OK = 0
PTL_ERR_INVALID_ARG = 1
PTL_ERR_CUDA = 2
PTL_ERR_NO_DEVICE = 3
PTL_ERR_STATE = 4
PTL_ERR_LIFT_PANIC = 5
PTL_ERR_INPUT = 6

STAGE_LEFT_SHIFT = 1
STAGE_LIFTOVER = 2
STAGE_SIMPLIFY = 4
STAGE_ALL = 7

OPS = "MIDNSHP=X"

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i8p = C.POINTER(C.c_int8)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)


class ContigSegmentsC(C.Structure):
    _fields_ = [
        ("n_contigs", C.c_uint32),
        ("contig_len", u64p),
        ("contig_seg_begin", u32p),
        ("rev_contig_seq", C.POINTER(u8p)),
        ("n_segments", C.c_uint32),
        ("seg_seq_order_start", u32p),
        ("seg_seq_order_end", u32p),
        ("seg_chrom_index", i32p),
        ("seg_pos", i64p),
        ("seg_is_fwd", u8p),
        ("seg_mapq", u8p),
        ("seg_cigar_begin", u64p),
        ("cigar", u32p),
    ]


class ContigRecordsC(C.Structure):
    _fields_ = [
        ("n_records", C.c_uint32),
        ("contig_id", u32p),
        ("flag", u16p),
        ("tid", i32p),
        ("pos", i64p),
        ("mapq", u8p),
        ("cigar_begin", u64p),
        ("cigar", u32p),
        ("sa_tag", C.POINTER(C.c_char_p)),
        ("seq", C.POINTER(u8p)),
        ("n_contigs", C.c_uint32),
        ("contig_len", u64p),
        ("contig_names", C.POINTER(C.c_char_p)),
        ("n_ref_chrom", C.c_uint32),
        ("ref_chrom_names", C.POINTER(C.c_char_p)),
    ]


class BatchC(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint32),
        ("read_flag", u16p),
        ("read_mapq", u8p),
        ("read_bin", u16p),
        ("read_seq_len", u32p),
        ("read_seq_off", u64p),
        ("read_seg_begin", u32p),
        ("n_read_segments", C.c_uint32),
        ("rseg_contig", u32p),
        ("rseg_pos", i64p),
        ("rseg_is_fwd", u8p),
        ("rseg_cigar_begin", u64p),
        ("rseg_cigar_len", u32p),
        ("cigar", u32p),
        ("n_cigar", C.c_uint64),
        ("seq4", u8p),
        ("seq4_bytes", C.c_uint64),
        ("indel_win", u64p),
        ("rseg_win_begin", u32p),
        ("n_indel_win", C.c_uint64),
    ]


class ResultC(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint32),
        ("read_rec_begin", u32p),
        ("n_records", C.c_uint32),
        ("rec_status", i8p),
        ("rec_read_segment", u32p),
        ("rec_contig_segment", u32p),
        ("rec_tid", i32p),
        ("rec_pos", i64p),
        ("rec_mapq", u8p),
        ("rec_flag", u16p),
        ("rec_bin", u16p),
        ("rec_need_flip", u8p),
        ("rec_cigar_begin", u64p),
        ("cigar", u32p),
        ("n_cigar", C.c_uint64),
        ("n_pairs", C.c_uint64),
        ("n_lifted", C.c_uint64),
        ("n_errors", C.c_uint64),
        ("first_error_read", C.c_int64),
        ("first_error_status", C.c_int32),
    ]


class ReadQualsC(C.Structure):
    _fields_ = [("qual", u8p), ("read_qual_off", u64p), ("qual_bytes", C.c_uint64)]


class RecordBasesC(C.Structure):
    _fields_ = [("n_records", C.c_uint32), ("rec_seq_begin", u64p), ("seq4", u8p), ("rec_qual_begin", u64p), ("qual", u8p),
                ("kernel_ms", C.c_float), ("bytes_read", C.c_uint64), ("bytes_written", C.c_uint64)]


ASM_RESIDENT_QUAL, ASM_NO_DOWNLOAD = 1, 2


class ReadExtrasC(C.Structure):
    _fields_ = [("name_off", u64p), ("names", u8p), ("aux_off", u64p), ("aux", u8p), ("mate_tid", i32p), ("mate_pos", i32p),
                ("tlen", i32p), ("quals", ReadQualsC)]


class BgzfStreamC(C.Structure):
    _fields_ = [("n_bytes", C.c_uint64), ("bytes", u8p), ("n_blocks", C.c_uint64), ("kernel_ms", C.c_float),
                ("bytes_read", C.c_uint64), ("bytes_written", C.c_uint64)]


BGZF_EOF = 4


class BamRecordsC(C.Structure):
    _fields_ = [("n_records", C.c_uint32), ("rec_begin", u64p), ("bytes", u8p), ("kernel_ms", C.c_float),
                ("bytes_read", C.c_uint64), ("bytes_written", C.c_uint64)]


class SplitSegmentsC(C.Structure):
    _fields_ = [
        ("seq_order_start", u32p),
        ("seq_order_end", u32p),
        ("contig", u32p),
        ("pos", i64p),
        ("is_fwd", u8p),
        ("mapq", u8p),
        ("from_primary", u8p),
        ("cigar_begin", u32p),
        ("cigar", u32p),
    ]


# ---------------------------------------------------------------------------------------------------- CIGAR helpers
def cigar_from_string(s: str) -> np.ndarray:
    out, n = [], ""
    for ch in s:
        if ch.isdigit():
            n += ch
        else:
            out.append((int(n) << 4) | OPS.index(ch))
            n = ""
    assert n == "", f"bad CIGAR {s!r}"
    return np.asarray(out, dtype=np.uint32)


def cigar_to_string(ops: Sequence[int]) -> str:
    return "".join(f"{int(v) >> 4}{OPS[int(v) & 0xF]}" for v in ops)


_NT16 = "=ACMGRSVTWYHKDBN"


def pack_seq4(seq: str | bytes) -> np.ndarray:
    """ASCII bases -> BAM 4-bit packed (high nibble first)."""
    if isinstance(seq, bytes):
        seq = seq.decode()
    codes = [_NT16.index(ch) for ch in seq]
    if len(codes) & 1:
        codes.append(0)
    return np.asarray([(codes[i] << 4) | codes[i + 1] for i in range(0, len(codes), 2)], dtype=np.uint8)


def _ptr(a: Optional[np.ndarray], typ):
    if a is None:
        return C.cast(None, typ)
    return a.ctypes.data_as(typ)


def _arr(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=dtype))


# ---------------------------------------------------------------------------------------------------- python-side views
@dataclass
class ContigSegments:
    """Numpy SoA twin of ptl_contig_segments (post trim/join unless stated)."""

    contig_len: np.ndarray
    contig_seg_begin: np.ndarray
    rev_contig_seq: List[Optional[np.ndarray]]
    seg_seq_order_start: np.ndarray
    seg_seq_order_end: np.ndarray
    seg_chrom_index: np.ndarray
    seg_pos: np.ndarray
    seg_is_fwd: np.ndarray
    seg_mapq: np.ndarray
    seg_cigar_begin: np.ndarray
    cigar: np.ndarray
    _keep: list = field(default_factory=list, repr=False)

    def to_c(self) -> ContigSegmentsC:
        self.contig_len = _arr(self.contig_len, np.uint64)
        self.contig_seg_begin = _arr(self.contig_seg_begin, np.uint32)
        self.seg_seq_order_start = _arr(self.seg_seq_order_start, np.uint32)
        self.seg_seq_order_end = _arr(self.seg_seq_order_end, np.uint32)
        self.seg_chrom_index = _arr(self.seg_chrom_index, np.int32)
        self.seg_pos = _arr(self.seg_pos, np.int64)
        self.seg_is_fwd = _arr(self.seg_is_fwd, np.uint8)
        self.seg_mapq = _arr(self.seg_mapq, np.uint8)
        self.seg_cigar_begin = _arr(self.seg_cigar_begin, np.uint64)
        self.cigar = _arr(self.cigar, np.uint32)
        n = len(self.contig_len)
        ptrs = (u8p * max(n, 1))()
        for i, s in enumerate(self.rev_contig_seq):
            if s is not None:
                s = _arr(s, np.uint8)
                self._keep.append(s)
                ptrs[i] = s.ctypes.data_as(u8p)
        self._keep.append(ptrs)
        c = ContigSegmentsC()
        c.n_contigs = n
        c.contig_len = _ptr(self.contig_len, u64p)
        c.contig_seg_begin = _ptr(self.contig_seg_begin, u32p)
        c.rev_contig_seq = C.cast(ptrs, C.POINTER(u8p))
        c.n_segments = len(self.seg_pos)
        c.seg_seq_order_start = _ptr(self.seg_seq_order_start, u32p)
        c.seg_seq_order_end = _ptr(self.seg_seq_order_end, u32p)
        c.seg_chrom_index = _ptr(self.seg_chrom_index, i32p)
        c.seg_pos = _ptr(self.seg_pos, i64p)
        c.seg_is_fwd = _ptr(self.seg_is_fwd, u8p)
        c.seg_mapq = _ptr(self.seg_mapq, u8p)
        c.seg_cigar_begin = _ptr(self.seg_cigar_begin, u64p)
        c.cigar = _ptr(self.cigar, u32p)
        return c

    @staticmethod
    def from_c(c: ContigSegmentsC) -> "ContigSegments":
        n, m = c.n_contigs, c.n_segments
        cl = np.ctypeslib.as_array(c.contig_len, (n,)).copy() if n else np.zeros(0, np.uint64)
        sb = np.ctypeslib.as_array(c.contig_seg_begin, (n + 1,)).copy()
        cb = np.ctypeslib.as_array(c.seg_cigar_begin, (m + 1,)).copy()
        g = lambda p, dt: (np.ctypeslib.as_array(p, (m,)).copy() if m else np.zeros(0, dt))
        rev = []
        for i in range(n):
            p = c.rev_contig_seq[i]
            rev.append(np.ctypeslib.as_array(p, (int(cl[i]),)).copy() if p else None)
        ncig = int(cb[-1])
        return ContigSegments(
            cl, sb, rev, g(c.seg_seq_order_start, np.uint32), g(c.seg_seq_order_end, np.uint32),
            g(c.seg_chrom_index, np.int32), g(c.seg_pos, np.int64), g(c.seg_is_fwd, np.uint8), g(c.seg_mapq, np.uint8),
            cb, np.ctypeslib.as_array(c.cigar, (ncig,)).copy() if ncig else np.zeros(0, np.uint32))

    def segment_cigar(self, k: int) -> np.ndarray:
        return self.cigar[int(self.seg_cigar_begin[k]):int(self.seg_cigar_begin[k + 1])]


@dataclass
class Batch:
    """Numpy SoA twin of ptl_batch."""

    read_flag: np.ndarray
    read_mapq: np.ndarray
    read_bin: np.ndarray
    read_seq_len: np.ndarray
    read_seq_off: np.ndarray
    read_seg_begin: np.ndarray
    rseg_contig: np.ndarray
    rseg_pos: np.ndarray
    rseg_is_fwd: np.ndarray
    rseg_cigar_begin: np.ndarray
    rseg_cigar_len: np.ndarray
    cigar: np.ndarray
    seq4: np.ndarray

    @property
    def n_reads(self) -> int:
        return len(self.read_flag)

    def to_c(self) -> BatchC:
        self.read_flag = _arr(self.read_flag, np.uint16)
        self.read_mapq = _arr(self.read_mapq, np.uint8)
        self.read_bin = _arr(self.read_bin, np.uint16)
        self.read_seq_len = _arr(self.read_seq_len, np.uint32)
        self.read_seq_off = _arr(self.read_seq_off, np.uint64)
        self.read_seg_begin = _arr(self.read_seg_begin, np.uint32)
        self.rseg_contig = _arr(self.rseg_contig, np.uint32)
        self.rseg_pos = _arr(self.rseg_pos, np.int64)
        self.rseg_is_fwd = _arr(self.rseg_is_fwd, np.uint8)
        self.rseg_cigar_begin = _arr(self.rseg_cigar_begin, np.uint64)
        self.rseg_cigar_len = _arr(self.rseg_cigar_len, np.uint32)
        self.cigar = _arr(self.cigar, np.uint32)
        if not (isinstance(self.seq4, np.ndarray) and self.seq4.dtype == np.uint8 and self.seq4.flags.c_contiguous):
            self.seq4 = _arr(self.seq4, np.uint8)
        b = BatchC()
        b.n_reads = len(self.read_flag)
        b.read_flag = _ptr(self.read_flag, u16p)
        b.read_mapq = _ptr(self.read_mapq, u8p)
        b.read_bin = _ptr(self.read_bin, u16p)
        b.read_seq_len = _ptr(self.read_seq_len, u32p)
        b.read_seq_off = _ptr(self.read_seq_off, u64p)
        b.read_seg_begin = _ptr(self.read_seg_begin, u32p)
        b.n_read_segments = len(self.rseg_contig)
        b.rseg_contig = _ptr(self.rseg_contig, u32p)
        b.rseg_pos = _ptr(self.rseg_pos, i64p)
        b.rseg_is_fwd = _ptr(self.rseg_is_fwd, u8p)
        b.rseg_cigar_begin = _ptr(self.rseg_cigar_begin, u64p)
        b.rseg_cigar_len = _ptr(self.rseg_cigar_len, u32p)
        b.cigar = _ptr(self.cigar, u32p)
        b.n_cigar = len(self.cigar)
        b.seq4 = _ptr(self.seq4, u8p)
        b.seq4_bytes = self.seq4.nbytes
        return b


@dataclass
class Result:
    """Owned numpy copy of a ptl_result."""

    read_rec_begin: np.ndarray
    rec_status: np.ndarray
    rec_read_segment: np.ndarray
    rec_contig_segment: np.ndarray
    rec_tid: np.ndarray
    rec_pos: np.ndarray
    rec_mapq: np.ndarray
    rec_flag: np.ndarray
    rec_bin: np.ndarray
    rec_need_flip: np.ndarray
    rec_cigar_begin: np.ndarray
    cigar: np.ndarray
    n_pairs: int
    n_lifted: int
    n_errors: int
    first_error_read: int
    first_error_status: int

    FIELDS = ("read_rec_begin", "rec_status", "rec_read_segment", "rec_contig_segment", "rec_tid", "rec_pos",
              "rec_mapq", "rec_flag", "rec_bin", "rec_need_flip", "rec_cigar_begin", "cigar")

    @property
    def n_records(self) -> int:
        return len(self.rec_status)

    @staticmethod
    def from_c(r: ResultC, copy: bool = True) -> "Result":
        nr, nrec, ncig = r.n_reads, r.n_records, r.n_cigar

        def g(p, n, dt):
            if n == 0:
                return np.zeros(0, dt)
            a = np.ctypeslib.as_array(p, (n,))
            return a.copy() if copy else a

        return Result(
            g(r.read_rec_begin, nr + 1, np.uint32), g(r.rec_status, nrec, np.int8),
            g(r.rec_read_segment, nrec, np.uint32), g(r.rec_contig_segment, nrec, np.uint32),
            g(r.rec_tid, nrec, np.int32), g(r.rec_pos, nrec, np.int64), g(r.rec_mapq, nrec, np.uint8),
            g(r.rec_flag, nrec, np.uint16), g(r.rec_bin, nrec, np.uint16), g(r.rec_need_flip, nrec, np.uint8),
            g(r.rec_cigar_begin, nrec + 1, np.uint64), g(r.cigar, ncig, np.uint32),
            int(r.n_pairs), int(r.n_lifted), int(r.n_errors), int(r.first_error_read), int(r.first_error_status))

    def record_cigar(self, k: int) -> str:
        return cigar_to_string(self.cigar[int(self.rec_cigar_begin[k]):int(self.rec_cigar_begin[k + 1])])

    def diff(self, other: "Result") -> Optional[str]:
        """First difference against another result (None if bit-identical)."""
        for f in self.FIELDS:
            a, b = getattr(self, f), getattr(other, f)
            if a.shape != b.shape:
                return f"{f}: shape {a.shape} != {b.shape}"
            if not np.array_equal(a, b):
                i = int(np.flatnonzero(a != b)[0])
                return f"{f}[{i}]: {a[i]} != {b[i]}"
        for f in ("n_pairs", "n_lifted", "n_errors"):
            if getattr(self, f) != getattr(other, f):
                return f"{f}: {getattr(self, f)} != {getattr(other, f)}"
        return None


class PtlError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ptl status {code}: {msg}")
        self.code = code


class LiftLib:
    """Drives one shared library exporting the portello_b200 ABI under `prefix`."""

    def __init__(self, cdll: C.CDLL, prefix: str):
        self.dll, self.prefix = cdll, prefix
        f = self._fn
        f("create", C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p)])
        f("destroy", None, [C.c_void_p])
        f("last_error", C.c_char_p, [C.c_void_p])
        f("version", C.c_char_p, [])
        f("set_reference", C.c_int, [C.c_void_p, C.c_uint32, u64p, C.POINTER(u8p)])
        f("set_contig_segments", C.c_int, [C.c_void_p, C.POINTER(ContigSegmentsC)])
        f("set_raw_contig_segments", C.c_int, [C.c_void_p, C.POINTER(ContigSegmentsC)])
        f("set_contig_records", C.c_int, [C.c_void_p, C.POINTER(ContigRecordsC)])
        f("get_contig_segments", C.c_int, [C.c_void_p, C.POINTER(ContigSegmentsC)])
        f("get_segment_table", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, u32p, i32p, u32p])
        f("lift_submit", C.c_int, [C.c_void_p, C.c_int, C.POINTER(BatchC)])
        f("lift_submit_ex", C.c_int, [C.c_void_p, C.c_int, C.POINTER(BatchC), C.c_uint32])
        f("lift_wait", C.c_int, [C.c_void_p, C.c_int, C.POINTER(ResultC)])

    def _fn(self, name, restype, argtypes):
        fn = getattr(self.dll, self.prefix + name)
        fn.restype, fn.argtypes = restype, argtypes
        setattr(self, "_" + name, fn)
        return fn

    def version(self) -> str:
        return self._version().decode()


class Context:
    """One ptl_ctx (one GPU for the product). Keeps python references to everything lent to C."""

    def __init__(self, lib: LiftLib, device: int = 0, n_slots: int = 2):
        self.lib = lib
        self.h = C.c_void_p()
        rc = lib._create(device, n_slots, C.byref(self.h))
        if rc != PTL_OK:
            raise PtlError(rc, "ptl_create failed (no CUDA device? the product has no CPU fallback)")
        self._keep = {}

    def close(self):
        if self.h:
            self.lib._destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_long_pair_ops(self, n_ops: int):
        """Pairs with more CIGAR ops than this are lifted warp-cooperatively (0 = every pair); results do not change.
        A library without the knob (the oracle) ignores the call."""
        fn = getattr(self.lib.dll, self.lib.prefix + "set_long_pair_ops", None)
        if fn is not None:
            fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_uint32]
            self._check(fn(self.h, int(n_ops)))

    def _check(self, rc: int, allow=()):
        if rc != PTL_OK and rc not in allow:
            raise PtlError(rc, self.lib._last_error(self.h).decode("utf-8", "replace"))
        return rc

    def set_reference(self, chroms: Sequence[np.ndarray]):
        chroms = [_arr(c, np.uint8) for c in chroms]
        lens = np.asarray([len(c) for c in chroms], dtype=np.uint64)
        ptrs = (u8p * max(len(chroms), 1))(*[c.ctypes.data_as(u8p) for c in chroms])
        self._check(self.lib._set_reference(self.h, len(chroms), _ptr(lens, u64p), C.cast(ptrs, C.POINTER(u8p))))

    def set_contig_segments(self, segs: ContigSegments, raw: bool = False):
        c = segs.to_c()
        fn = self.lib._set_raw_contig_segments if raw else self.lib._set_contig_segments
        self._check(fn(self.h, C.byref(c)))

    def set_contig_records(self, recs_c: ContigRecordsC):
        self._check(self.lib._set_contig_records(self.h, C.byref(recs_c)))

    def get_contig_segments(self) -> ContigSegments:
        c = ContigSegmentsC()
        self._check(self.lib._get_contig_segments(self.h, C.byref(c)))
        return ContigSegments.from_c(c)

    def get_segment_table(self, segment: int, cap: int = 1 << 16):
        keys = np.zeros(cap, np.uint32)
        vals = np.zeros(cap, np.int32)
        n = C.c_uint32()
        rc = self.lib._get_segment_table(self.h, segment, cap, _ptr(keys, u32p), _ptr(vals, i32p), C.byref(n))
        if rc == PTL_ERR_INVALID_ARG and n.value > cap:
            return self.get_segment_table(segment, n.value)
        self._check(rc)
        return keys[: n.value].copy(), vals[: n.value].copy()

    def submit(self, batch: Batch, slot: int = 0, stage_mask: int = STAGE_ALL):
        c = batch.to_c()
        self._keep[slot] = (batch, c)
        if stage_mask == STAGE_ALL:
            self._check(self.lib._lift_submit(self.h, slot, C.byref(c)))
        else:
            self._check(self.lib._lift_submit_ex(self.h, slot, C.byref(c), stage_mask))

    def wait(self, slot: int = 0, allow_panic: bool = False, copy: bool = True) -> Result:
        r = ResultC()
        self._check(self.lib._lift_wait(self.h, slot, C.byref(r)), allow=(PTL_ERR_LIFT_PANIC,) if allow_panic else ())
        return Result.from_c(r, copy=copy)

    def lift(self, batch: Batch, slot: int = 0, stage_mask: int = STAGE_ALL, allow_panic: bool = False) -> Result:
        self.submit(batch, slot, stage_mask)
        return self.wait(slot, allow_panic)

    def assemble_bases(self, qual: Optional[np.ndarray], read_qual_off: Optional[np.ndarray], slot: int = 0, flags: int = 0):
        """ptl_assemble_bases on the slot's last lifted batch.  Returns (RecordBasesC, per-record list of (seq4 bytes, qual bytes))
        -- the list is None with ASM_NO_DOWNLOAD."""
        fn = getattr(self.lib.dll, self.lib.prefix + "assemble_bases")
        fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_int, C.POINTER(ReadQualsC), C.c_uint32, C.POINTER(RecordBasesC)]
        q = ReadQualsC()
        if qual is not None:
            qa, qo = _arr(qual, np.uint8), _arr(read_qual_off, np.uint64)
            self._keep[("qual", slot)] = (qa, qo)
            q.qual, q.read_qual_off, q.qual_bytes = _ptr(qa, u8p), _ptr(qo, u64p), qa.size
        out = RecordBasesC()
        self._check(fn(self.h, slot, C.byref(q) if qual is not None else None, flags, C.byref(out)))
        if flags & ASM_NO_DOWNLOAD:
            return out, None
        n = out.n_records
        sb = np.ctypeslib.as_array(out.rec_seq_begin, (n + 1,)).copy()
        qb = np.ctypeslib.as_array(out.rec_qual_begin, (n + 1,)).copy()
        seq = np.ctypeslib.as_array(out.seq4, (max(int(sb[n]), 1),)).copy() if n else np.zeros(0, np.uint8)
        ql = np.ctypeslib.as_array(out.qual, (max(int(qb[n]), 1),)).copy() if n else np.zeros(0, np.uint8)
        return out, (sb, seq, qb, ql)

    def set_names(self, contig_names, chrom_names):
        """ptl_set_names: contig names (PS:Z) and reference names (SA:Z) for ptl_assemble_records."""
        fn = getattr(self.lib.dll, self.lib.prefix + "set_names")
        fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(C.c_char_p)]
        a = (C.c_char_p * max(len(contig_names), 1))(*[n.encode() for n in contig_names])
        b = (C.c_char_p * max(len(chrom_names), 1))(*[n.encode() for n in chrom_names])
        self._check(fn(self.h, len(contig_names), a, len(chrom_names), b))

    def _extras_c(self, extras: dict, slot: int) -> ReadExtrasC:
        x = ReadExtrasC()
        k = {
            "name_off": _arr(extras["name_off"], np.uint64), "names": _arr(extras["names"], np.uint8),
            "aux_off": _arr(extras["aux_off"], np.uint64), "aux": _arr(extras["aux"], np.uint8),
            "mate_tid": _arr(extras["mate_tid"], np.int32), "mate_pos": _arr(extras["mate_pos"], np.int32),
            "tlen": _arr(extras["tlen"], np.int32), "qual": _arr(extras["qual"], np.uint8), "qual_off": _arr(extras["qual_off"], np.uint64),
        }
        self._keep[("extras", slot)] = k
        x.name_off, x.names, x.aux_off, x.aux = _ptr(k["name_off"], u64p), _ptr(k["names"], u8p), _ptr(k["aux_off"], u64p), _ptr(k["aux"], u8p)
        x.mate_tid, x.mate_pos, x.tlen = _ptr(k["mate_tid"], i32p), _ptr(k["mate_pos"], i32p), _ptr(k["tlen"], i32p)
        x.quals.qual, x.quals.read_qual_off, x.quals.qual_bytes = _ptr(k["qual"], u8p), _ptr(k["qual_off"], u64p), k["qual"].size
        return x

    def assemble_records(self, extras: Optional[dict], slot: int = 0, flags: int = 0):
        """ptl_assemble_records on the slot's last lifted batch.  `extras`: dict with name_off, names, aux_off, aux, mate_tid,
        mate_pos, tlen, qual, qual_off (numpy).  Returns (BamRecordsC, (rec_begin, bytes)) -- the pair is None with ASM_NO_DOWNLOAD."""
        fn = getattr(self.lib.dll, self.lib.prefix + "assemble_records")
        fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_int, C.POINTER(ReadExtrasC), C.c_uint32, C.POINTER(BamRecordsC)]
        x = self._extras_c(extras, slot) if extras is not None else ReadExtrasC()
        out = BamRecordsC()
        self._check(fn(self.h, slot, C.byref(x) if extras is not None else None, flags, C.byref(out)))
        if flags & ASM_NO_DOWNLOAD:
            return out, None
        n = out.n_records
        rb = np.ctypeslib.as_array(out.rec_begin, (n + 1,)).copy()
        by = np.ctypeslib.as_array(out.bytes, (max(int(rb[n]), 1),))[: int(rb[n])].copy() if n else np.zeros(0, np.uint8)
        return out, (rb, by)

    def frame_records(self, extras: Optional[dict], prefix: bytes = b"", slot: int = 0, flags: int = 0, copy: bool = True):
        """ptl_frame_records: assemble + frame in one pass.  Returns (BgzfStreamC, bytes or None)."""
        fn = getattr(self.lib.dll, self.lib.prefix + "frame_records")
        fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_int, C.POINTER(ReadExtrasC), C.c_char_p, C.c_uint64, C.c_uint32, C.POINTER(BgzfStreamC)]
        x = self._extras_c(extras, slot) if extras is not None else None
        out = BgzfStreamC()
        self._check(fn(self.h, slot, C.byref(x) if x is not None else None, prefix if prefix else None, len(prefix), flags, C.byref(out)))
        if (flags & ASM_NO_DOWNLOAD) or not copy:
            return out, None
        return out, bytes(np.ctypeslib.as_array(out.bytes, (max(int(out.n_bytes), 1),))[: int(out.n_bytes)])

    def bgzf_store_records(self, prefix: bytes = b"", slot: int = 0, flags: int = 0, copy: bool = True):
        """ptl_bgzf_store_records: [prefix | records of the slot's last assemble_records] as level-0 BGZF.  Returns
        (BgzfStreamC, bytes) -- bytes is None with ASM_NO_DOWNLOAD, or with copy=False (the stream then stays in the slot's
        pinned buffer, out.bytes)."""
        fn = getattr(self.lib.dll, self.lib.prefix + "bgzf_store_records")
        fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_uint64, C.c_uint32, C.POINTER(BgzfStreamC)]
        out = BgzfStreamC()
        self._check(fn(self.h, slot, prefix if prefix else None, len(prefix), flags, C.byref(out)))
        if (flags & ASM_NO_DOWNLOAD) or not copy:
            return out, None
        return out, bytes(np.ctypeslib.as_array(out.bytes, (max(int(out.n_bytes), 1),))[: int(out.n_bytes)])

