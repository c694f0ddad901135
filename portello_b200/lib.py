"""Loader + ctypes prototypes for the product library portello_b200/csrc/libportello_b200.so.

Fails loudly if the library has not been built: there is no pure-Python or CPU implementation of the liftover path
in this package (the CPU checker lives in oracle/ and is only used by tests / bench baselines).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .abi import (BatchC, ContigRecordsC, ContigSegments, ContigSegmentsC, Context, LiftLib, PtlError, ResultC,
                  SplitSegmentsC, u8p, u16p, u32p, u64p, i32p, i64p)

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
# PORTELLO_B200_LIB: load another build of the same library (kernel tuning A/B runs under gpurun)
SO = os.environ.get("PORTELLO_B200_LIB") or os.path.join(CSRC, "libportello_b200.so")


class ReadRecordsC(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint32),
        ("tid", i32p),
        ("pos", i64p),
        ("flag", u16p),
        ("mapq", u8p),
        ("bin", u16p),
        ("seq_len", u32p),
        ("seq_off", u64p),
        ("seq4", u8p),
        ("seq4_bytes", C.c_uint64),
        ("cigar_begin", u64p),
        ("cigar", u32p),
        ("sa_tag", C.POINTER(C.c_char_p)),
    ]


def build(force: bool = False) -> str:
    """Compile the CUDA/C++ sources in-tree with the Makefile (nvcc cross-compiles sm_100a without a GPU)."""
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libportello_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return SO


class ProductLib(LiftLib):
    def __init__(self, dll: C.CDLL):
        super().__init__(dll, "ptl_")
        d = dll
        d.ptl_lift_upload.restype = C.c_int
        d.ptl_lift_upload.argtypes = [C.c_void_p, C.c_int, C.POINTER(BatchC)]
        d.ptl_lift_run.restype = C.c_int
        d.ptl_lift_run.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        d.ptl_lift_download.restype = C.c_int
        d.ptl_lift_download.argtypes = [C.c_void_p, C.c_int, C.POINTER(ResultC)]
        d.ptl_slot_stream.restype = C.c_void_p
        d.ptl_slot_stream.argtypes = [C.c_void_p, C.c_int]
        d.ptl_slot_kernel_times.restype = C.c_int
        d.ptl_slot_kernel_times.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
        d.ptl_launch_count.restype = C.c_uint64
        d.ptl_launch_count.argtypes = [C.c_void_p]
        d.ptl_slot_counters.restype = C.c_int
        d.ptl_slot_counters.argtypes = [C.c_void_p, C.c_int, u64p]
        d.ptl_set_seq_zero_copy.restype = C.c_int
        d.ptl_set_seq_zero_copy.argtypes = [C.c_void_p, C.c_int]
        d.ptl_host_alloc.restype = C.c_void_p
        d.ptl_host_alloc.argtypes = [C.c_size_t]
        d.ptl_host_free.argtypes = [C.c_void_p]
        d.ptl_pack_split_segments.restype = C.c_int
        d.ptl_pack_split_segments.argtypes = [C.c_uint32, C.POINTER(C.c_char_p), C.c_int32, C.c_int64, C.c_uint16, C.c_uint8, u32p, C.c_uint32,
                                              C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(SplitSegmentsC), u32p, u32p]
        d.ptl_pack_last_error.restype = C.c_char_p
        d.ptl_pack_batch.restype = C.c_int
        d.ptl_pack_batch.argtypes = [C.POINTER(ReadRecordsC), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_void_p)]
        d.ptl_pack_batch_ex.restype = C.c_int
        d.ptl_pack_batch_ex.argtypes = [C.POINTER(ReadRecordsC), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_char_p), C.c_int, C.c_int,
                                        C.POINTER(ContigSegmentsC), C.POINTER(C.c_void_p)]
        d.ptl_pack_batch_into.restype = C.c_int
        d.ptl_pack_batch_into.argtypes = [C.c_void_p, C.POINTER(ReadRecordsC), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_char_p), C.c_int,
                                          C.POINTER(ContigSegmentsC)]
        d.ptl_packed_batch_view.argtypes = [C.c_void_p, C.POINTER(BatchC)]
        d.ptl_packed_batch_record_index.restype = u32p
        d.ptl_packed_batch_record_index.argtypes = [C.c_void_p]
        d.ptl_packed_batch_free.argtypes = [C.c_void_p]
        d.ptl_format_sa_tags.restype = C.c_int
        d.ptl_format_sa_tags.argtypes = [C.POINTER(ResultC), C.c_uint32, C.POINTER(C.c_char_p), C.c_char_p, C.c_uint64, u64p, u64p]
        d.ptl_region_segment_count.restype = C.c_uint32
        d.ptl_region_segment_count.argtypes = [C.c_uint64, C.c_uint64]
        d.ptl_region_segments.argtypes = [C.c_uint64, C.c_uint64, u64p, u64p]
        d.ptl_shard_units.argtypes = [C.c_uint32, u64p, C.c_uint32, u32p]
        d.ptl_reg2bin.restype = C.c_uint16
        d.ptl_reg2bin.argtypes = [C.c_int64, C.c_int64]
        d.ptl_prepare_contig_records.restype = C.c_int
        d.ptl_prepare_contig_records.argtypes = [C.POINTER(ContigRecordsC), C.POINTER(C.c_void_p)]
        d.ptl_prepare_raw_contig_segments.restype = C.c_int
        d.ptl_prepare_raw_contig_segments.argtypes = [C.POINTER(ContigSegmentsC), C.POINTER(C.c_void_p)]
        d.ptl_prepared_contigs_view.argtypes = [C.c_void_p, C.POINTER(ContigSegmentsC)]
        d.ptl_prepared_contigs_free.argtypes = [C.c_void_p]
        d.ptl_prepare_last_error.restype = C.c_char_p

    # ---- host-only helpers (no GPU needed)
    def prepare_contig_records(self, recs_c: ContigRecordsC) -> ContigSegments:
        h = C.c_void_p()
        rc = self.dll.ptl_prepare_contig_records(C.byref(recs_c), C.byref(h))
        if rc != 0:
            raise PtlError(rc, self.dll.ptl_prepare_last_error().decode("utf-8", "replace"))
        try:
            v = ContigSegmentsC()
            self.dll.ptl_prepared_contigs_view(h, C.byref(v))
            return ContigSegments.from_c(v)
        finally:
            self.dll.ptl_prepared_contigs_free(h)

    def prepare_raw_contig_segments(self, segs: ContigSegments) -> ContigSegments:
        h = C.c_void_p()
        c = segs.to_c()
        rc = self.dll.ptl_prepare_raw_contig_segments(C.byref(c), C.byref(h))
        if rc != 0:
            raise PtlError(rc, self.dll.ptl_prepare_last_error().decode("utf-8", "replace"))
        try:
            v = ContigSegmentsC()
            self.dll.ptl_prepared_contigs_view(h, C.byref(v))
            return ContigSegments.from_c(v)
        finally:
            self.dll.ptl_prepared_contigs_free(h)

    def region_segments(self, size: int, segment_size: int):
        n = self.dll.ptl_region_segment_count(size, segment_size)
        b, e = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        self.dll.ptl_region_segments(size, segment_size, b.ctypes.data_as(u64p), e.ctypes.data_as(u64p))
        return [[int(b[i]), int(e[i])] for i in range(n)]

    def shard_units(self, weights, n_ranks: int) -> np.ndarray:
        w = np.ascontiguousarray(np.asarray(weights, dtype=np.uint64))
        owner = np.zeros(len(w), np.uint32)
        self.dll.ptl_shard_units(len(w), w.ctypes.data_as(u64p), n_ranks, owner.ctypes.data_as(u32p))
        return owner

    def reg2bin(self, b: int, e: int) -> int:
        return int(self.dll.ptl_reg2bin(b, e))

    def format_sa_tags(self, res_c: ResultC, chrom_names):
        return format_sa_tags_call(self.dll.ptl_format_sa_tags, res_c, chrom_names)


def format_sa_tags_call(fn, res_c: ResultC, chrom_names):
    names = (C.c_char_p * len(chrom_names))(*[n.encode() for n in chrom_names])
    need = C.c_uint64()
    fn(C.byref(res_c), len(chrom_names), names, None, 0, None, C.byref(need))
    buf = C.create_string_buffer(max(int(need.value), 1))
    begin = np.zeros(res_c.n_records + 1, np.uint64)
    rc = fn(C.byref(res_c), len(chrom_names), names, buf, need.value, begin.ctypes.data_as(u64p), C.byref(need))
    if rc != 0:
        raise PtlError(rc, "ptl_format_sa_tags failed")
    raw = buf.raw
    return [raw[int(begin[i]):int(begin[i + 1]) - 1].decode() for i in range(res_c.n_records)]


class PackedBatch:
    """Owner of a ptl_packed_batch (the host packer's output)."""

    def __init__(self, lib: ProductLib, recs_c: ReadRecordsC, first: int, count: int, contig_names, pinned: bool = False, windows=None):
        """`windows`: None = no indel windows; True = for every read segment; or a ContigSegments (Context.get_contig_segments()):
        only for read segments that pair with a reverse-strand contig segment (the ones that go through left_shift_indels)."""
        self.lib = lib
        self.h = C.c_void_p()
        names = (C.c_char_p * len(contig_names))(*[n.encode() for n in contig_names])
        if windows is None:
            rc = lib.dll.ptl_pack_batch(C.byref(recs_c), first, count, len(contig_names), names, int(pinned), C.byref(self.h))
        elif windows is True:
            rc = lib.dll.ptl_pack_batch_ex(C.byref(recs_c), first, count, len(contig_names), names, int(pinned), 1, None, C.byref(self.h))
        else:
            segs_c = windows.to_c()
            rc = lib.dll.ptl_pack_batch_ex(C.byref(recs_c), first, count, len(contig_names), names, int(pinned), 2, C.byref(segs_c), C.byref(self.h))
        if rc != 0:
            raise PtlError(rc, lib.dll.ptl_pack_last_error().decode("utf-8", "replace"))
        self.c = BatchC()
        lib.dll.ptl_packed_batch_view(self.h, C.byref(self.c))
        self.c._owner = self  # the view must keep the arena alive (`pack(...).c` would otherwise dangle)
        self._recs = recs_c  # the batch borrows seq4
        self._names = names
        self._segs_c = None if windows is None or windows is True else segs_c
        self._mode = 0 if windows is None else 1 if windows is True else 2
        self._windows = windows  # (segs_c points into its arrays)

    def repack(self, first: int, count: int, recs_c: ReadRecordsC = None):
        """ptl_pack_batch_into: pack other records into this batch's arena (reused when large enough).  Callable from any
        thread (the C call releases the GIL); distinct PackedBatch objects may be repacked concurrently."""
        recs_c = recs_c if recs_c is not None else self._recs
        rc = self.lib.dll.ptl_pack_batch_into(self.h, C.byref(recs_c), first, count, len(self._names), self._names, self._mode,
                                              C.byref(self._segs_c) if self._segs_c is not None else None)
        if rc != 0:
            raise PtlError(rc, self.lib.dll.ptl_pack_last_error().decode("utf-8", "replace"))
        self._recs = recs_c
        self.lib.dll.ptl_packed_batch_view(self.h, C.byref(self.c))
        self.c._owner = self
        return self

    def record_index(self) -> np.ndarray:
        p = self.lib.dll.ptl_packed_batch_record_index(self.h)
        return np.ctypeslib.as_array(p, (self.c.n_reads,)).copy() if self.c.n_reads else np.zeros(0, np.uint32)

    def close(self):
        if self.h:
            self.lib.dll.ptl_packed_batch_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_lib = None


def load(build_if_missing: bool = True) -> ProductLib:
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            if not build_if_missing:
                raise RuntimeError(f"{SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
            build()
        _lib = ProductLib(C.CDLL(SO))
    return _lib


class GpuContext(Context):
    """Context of the CUDA product with the split upload/run/download phases and device-side timers."""

    def __init__(self, device: int = 0, n_slots: int = 2):
        super().__init__(load(), device, n_slots)

    def upload(self, batch_c: BatchC, slot: int = 0):
        self._keep[("up", slot)] = batch_c
        self._check(self.lib.dll.ptl_lift_upload(self.h, slot, C.byref(batch_c)))

    def run(self, slot: int = 0, stage_mask: int = 7):
        self._check(self.lib.dll.ptl_lift_run(self.h, slot, stage_mask))

    def download(self, slot: int = 0, copy: bool = True, allow_panic: bool = False):
        from .abi import Result, PTL_ERR_LIFT_PANIC

        r = ResultC()
        self._check(self.lib.dll.ptl_lift_download(self.h, slot, C.byref(r)), allow=(PTL_ERR_LIFT_PANIC,) if allow_panic else ())
        return Result.from_c(r, copy=copy)

    def submit_c(self, batch_c: BatchC, slot: int = 0):
        self._keep[("sub", slot)] = batch_c
        self._check(self.lib.dll.ptl_lift_submit(self.h, slot, C.byref(batch_c)))

    def wait_c(self, slot: int = 0) -> ResultC:
        r = ResultC()
        self._check(self.lib.dll.ptl_lift_wait(self.h, slot, C.byref(r)))
        return r

    def stream(self, slot: int = 0) -> int:
        return int(self.lib.dll.ptl_slot_stream(self.h, slot) or 0)

    def kernel_times(self, slot: int = 0):
        names = (C.c_char_p * 8)()
        ms = (C.c_float * 8)()
        n = self.lib.dll.ptl_slot_kernel_times(self.h, slot, 8, names, ms)
        return {names[i].decode(): float(ms[i]) for i in range(n)}

    def counters(self, slot: int = 0):
        out = np.zeros(6, np.uint64)
        self._check(self.lib.dll.ptl_slot_counters(self.h, slot, out.ctypes.data_as(u64p)))
        keys = ("n_pairs", "n_lifted", "n_in_ops", "n_out_ops", "base_bytes", "scratch_ops")
        return {k: int(v) for k, v in zip(keys, out)}

    def launch_count(self) -> int:
        return int(self.lib.dll.ptl_launch_count(self.h))

    def set_seq_zero_copy(self, on: bool):
        self._check(self.lib.dll.ptl_set_seq_zero_copy(self.h, int(on)))
