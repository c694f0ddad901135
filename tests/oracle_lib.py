"""Loader for the CPU oracle (oracle/_build/libptl_oracle.so). TEST INFRASTRUCTURE: importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only."""
import ctypes as C
import os
import subprocess

import numpy as np

from portello_b200 import abi
from portello_b200.abi import (LiftLib, SplitSegmentsC, ResultC, cigar_from_string, cigar_to_string, u8p, u32p, u64p,
                               i32p, i64p)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "libptl_oracle.so")


def build(force=False):
    srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith((".cpp", ".hpp"))]
    srcs.append(os.path.join(ROOT, "include", "portello_b200.h"))
    stale = force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return SO


class OracleLib(LiftLib):
    def __init__(self, dll):
        super().__init__(dll, "ptl_oracle_")
        d = dll
        d.ptl_oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
        d.ptl_oracle_set_faithful_decode.argtypes = [C.c_void_p, C.c_int]
        d.ptl_oracle_last_stats.argtypes = [C.c_void_p, C.c_int, u64p]
        d.ptl_oracle_fn_liftover.restype = C.c_int64
        d.ptl_oracle_fn_liftover.argtypes = [C.c_int64, u32p, C.c_uint32, C.c_int, C.c_int64, u32p, C.c_uint32, i64p, u32p, C.c_uint32]
        d.ptl_oracle_fn_simplify.restype = C.c_int64
        d.ptl_oracle_fn_simplify.argtypes = [C.c_int64, u32p, C.c_uint32, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, i64p, u32p, C.c_uint32]
        d.ptl_oracle_fn_shift.restype = C.c_int64
        d.ptl_oracle_fn_shift.argtypes = [C.c_int, C.c_int64, u32p, C.c_uint32, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, i64p, u32p, C.c_uint32]
        d.ptl_oracle_fn_compress.restype = C.c_int64
        d.ptl_oracle_fn_compress.argtypes = [u32p, C.c_uint32, u32p, C.c_uint32]
        d.ptl_oracle_fn_cleanup.restype = C.c_int64
        d.ptl_oracle_fn_cleanup.argtypes = [u32p, C.c_uint32]
        d.ptl_oracle_fn_clip_read_edges.restype = C.c_int64
        d.ptl_oracle_fn_clip_read_edges.argtypes = [u32p, C.c_uint32, C.c_uint64, C.c_uint64, i64p, u32p, C.c_uint32]
        d.ptl_oracle_fn_read_clip_positions.argtypes = [u32p, C.c_uint32, C.c_int, u64p]
        d.ptl_oracle_fn_offsets.argtypes = [u32p, C.c_uint32, C.c_int, i64p, u64p]
        d.ptl_oracle_fn_homology.argtypes = [C.c_char_p, C.c_uint64, C.c_int64, C.c_int64, C.c_char_p, C.c_uint64, C.c_int64, C.c_int64, i64p, C.c_char_p, C.c_uint32, u32p]
        d.ptl_oracle_fn_tree_map.restype = C.c_int64
        d.ptl_oracle_fn_tree_map.argtypes = [C.c_int64, u32p, C.c_uint32, C.c_int, u64p, i64p, C.c_uint32]
        d.ptl_oracle_fn_tree_map_ref_pos.argtypes = [C.c_int64, u32p, C.c_uint32, C.c_int, C.c_uint32, i64p]
        d.ptl_oracle_fn_tree_map_ref_range.restype = C.c_int64
        d.ptl_oracle_fn_tree_map_ref_range.argtypes = [C.c_int64, u32p, C.c_uint32, C.c_int, C.c_uint64, C.c_uint64, u64p, i64p, C.c_uint32]
        d.ptl_oracle_fn_clip_seg_isec_range.restype = C.c_int64
        d.ptl_oracle_fn_clip_seg_isec_range.argtypes = [u64p, i64p, C.c_int, u32p, C.c_uint32, C.c_int64, C.c_int64, u32p, C.c_uint32]
        d.ptl_oracle_fn_gci.restype = C.c_double
        d.ptl_oracle_fn_gci.argtypes = [u32p, C.c_uint32]
        d.ptl_oracle_fn_reg2bin.restype = C.c_uint16
        d.ptl_oracle_fn_reg2bin.argtypes = [C.c_int64, C.c_int64]
        d.ptl_oracle_fn_region_segments.restype = C.c_uint32
        d.ptl_oracle_fn_region_segments.argtypes = [C.c_uint64, C.c_uint64, u64p, u64p, C.c_uint32]
        d.ptl_oracle_fn_rev_comp.argtypes = [C.c_char_p, C.c_uint64]
        d.ptl_oracle_fn_strip_clip.restype = C.c_int64
        d.ptl_oracle_fn_strip_clip.argtypes = [C.c_int, u32p, C.c_uint32, u32p, C.c_uint32]
        d.ptl_oracle_fn_parse_sa.restype = C.c_int64
        d.ptl_oracle_fn_parse_sa.argtypes = [C.c_char_p, C.c_uint32, i64p, u8p, u8p, i32p, u32p, C.c_char_p, C.c_uint32]
        d.ptl_oracle_split_segments.restype = C.c_int
        d.ptl_oracle_split_segments.argtypes = [C.c_uint32, C.POINTER(C.c_char_p), C.c_int32, C.c_int64, C.c_uint16, C.c_uint8, u32p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(SplitSegmentsC), u32p, u32p]
        d.ptl_oracle_split_last_error.restype = C.c_char_p
        d.ptl_oracle_pack_batch.restype = C.c_int
        d.ptl_oracle_pack_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_void_p)]
        d.ptl_oracle_packed_view.argtypes = [C.c_void_p, C.POINTER(abi.BatchC)]
        d.ptl_oracle_packed_record_index.restype = C.POINTER(C.c_uint32)
        d.ptl_oracle_packed_record_index.argtypes = [C.c_void_p]
        d.ptl_oracle_packed_free.argtypes = [C.c_void_p]
        d.ptl_oracle_pack_last_error.restype = C.c_char_p
        d.ptl_oracle_format_sa_tags.restype = C.c_int
        d.ptl_oracle_format_sa_tags.argtypes = [C.POINTER(ResultC), C.c_uint32, C.POINTER(C.c_char_p), C.c_char_p, C.c_uint64, u64p, u64p]

    # ---- function-level mirrors of the reference's pure functions (string CIGAR in/out for readability)
    @staticmethod
    def _p(a, t):
        return a.ctypes.data_as(t)

    def _out(self, cap=4096):
        return np.zeros(cap, np.uint32)

    def liftover(self, c2r, c2r_pos, pos, cigar):
        cg = cigar_from_string(cigar)
        c2 = cigar_from_string(c2r) if c2r is not None else np.zeros(0, np.uint32)
        out, op = self._out(), C.c_int64()
        n = self.dll.ptl_oracle_fn_liftover(c2r_pos, self._p(c2, u32p), len(c2), int(c2r is not None), pos, self._p(cg, u32p), len(cg), C.byref(op), self._p(out, u32p), len(out))
        return None if n < 0 else (op.value, cigar_to_string(out[:n]))

    def simplify(self, pos, cigar, ref, read):
        cg = cigar_from_string(cigar)
        out, op = self._out(), C.c_int64()
        n = self.dll.ptl_oracle_fn_simplify(pos, self._p(cg, u32p), len(cg), ref.encode(), len(ref), read.encode(), len(read), C.byref(op), self._p(out, u32p), len(out))
        assert n >= 0, n
        return op.value, cigar_to_string(out[:n])

    def shift(self, direction, pos, cigar, ref, read):
        cg = cigar_from_string(cigar)
        out, op = self._out(), C.c_int64()
        n = self.dll.ptl_oracle_fn_shift(int(direction == "right"), pos, self._p(cg, u32p), len(cg), ref.encode(), len(ref), read.encode(), len(read), C.byref(op), self._p(out, u32p), len(out))
        assert n >= 0, n
        return op.value, cigar_to_string(out[:n])

    def compress(self, cigar):
        cg = cigar_from_string(cigar)
        out = self._out()
        n = self.dll.ptl_oracle_fn_compress(self._p(cg, u32p), len(cg), self._p(out, u32p), len(out))
        return cigar_to_string(out[:n])

    def cleanup(self, cigar):
        cg = cigar_from_string(cigar)
        shift = self.dll.ptl_oracle_fn_cleanup(self._p(cg, u32p), len(cg))
        return shift, cigar_to_string(cg)

    def clip_read_edges(self, cigar, left, right):
        cg = cigar_from_string(cigar)
        out, sh = self._out(), C.c_int64()
        n = self.dll.ptl_oracle_fn_clip_read_edges(self._p(cg, u32p), len(cg), left, right, C.byref(sh), self._p(out, u32p), len(out))
        return cigar_to_string(out[:n]), sh.value

    def read_clip_positions(self, cigar, ignore_hard_clip):
        cg = cigar_from_string(cigar)
        o = np.zeros(3, np.uint64)
        self.dll.ptl_oracle_fn_read_clip_positions(self._p(cg, u32p), len(cg), int(ignore_hard_clip), self._p(o, u64p))
        return [int(x) for x in o]

    def offsets(self, cigar, ref_pos, read_pos, ignore_hard_clip):
        cg = cigar_from_string(cigar)
        r = np.full(len(cg), ref_pos, np.int64)
        q = np.full(len(cg), read_pos, np.uint64)
        self.dll.ptl_oracle_fn_offsets(self._p(cg, u32p), len(cg), int(ignore_hard_clip), self._p(r, i64p), self._p(q, u64p))
        return [int(x) for x in r], [int(x) for x in q]

    def homology(self, ref, ref_range, read, read_range):
        rng = np.zeros(2, np.int64)
        hom = C.create_string_buffer(1024)
        n = C.c_uint32()
        self.dll.ptl_oracle_fn_homology(ref.encode(), len(ref), ref_range[0], ref_range[1], read.encode(), len(read), read_range[0], read_range[1], self._p(rng, i64p), hom, 1024, C.byref(n))
        return [int(rng[0]), int(rng[1])], hom.raw[: n.value].decode()

    def tree_map(self, ref_pos, cigar, ignore_hard_clip):
        cg = cigar_from_string(cigar)
        k, v = np.zeros(4096, np.uint64), np.zeros(4096, np.int64)
        n = self.dll.ptl_oracle_fn_tree_map(ref_pos, self._p(cg, u32p), len(cg), int(ignore_hard_clip), self._p(k, u64p), self._p(v, i64p), 4096)
        return [[int(k[i]), (None if v[i] < 0 else int(v[i]))] for i in range(n)]

    def tree_map_ref_pos(self, ref_pos, cigar, ignore_hard_clip, n_pos):
        cg = cigar_from_string(cigar)
        o = np.zeros(n_pos, np.int64)
        self.dll.ptl_oracle_fn_tree_map_ref_pos(ref_pos, self._p(cg, u32p), len(cg), int(ignore_hard_clip), n_pos, self._p(o, i64p))
        return [None if x == np.iinfo(np.int64).min else int(x) for x in o]

    def tree_map_ref_range(self, ref_pos, cigar, ignore_hard_clip, start, end):
        cg = cigar_from_string(cigar)
        k, v = np.zeros(4096, np.uint64), np.zeros(4096, np.int64)
        n = self.dll.ptl_oracle_fn_tree_map_ref_range(ref_pos, self._p(cg, u32p), len(cg), int(ignore_hard_clip), start, end, self._p(k, u64p), self._p(v, i64p), 4096)
        return [[int(k[i]), (None if v[i] < 0 else int(v[i]))] for i in range(n)]

    def clip_seg_isec_range(self, so, pos, is_fwd, cigar, isec):
        cg = cigar_from_string(cigar)
        so_a = np.asarray(so, np.uint64)
        p = C.c_int64(pos)
        out = self._out()
        n = self.dll.ptl_oracle_fn_clip_seg_isec_range(self._p(so_a, u64p), C.byref(p), int(is_fwd), self._p(cg, u32p), len(cg), isec[0], isec[1], self._p(out, u32p), len(out))
        return dict(eliminated=n < 0, so=[int(so_a[0]), int(so_a[1])], pos=p.value, cigar=cigar_to_string(out[: max(n, 0)]))

    def gci(self, cigar):
        cg = cigar_from_string(cigar)
        return self.dll.ptl_oracle_fn_gci(self._p(cg, u32p), len(cg))

    def reg2bin(self, b, e):
        return int(self.dll.ptl_oracle_fn_reg2bin(b, e))

    def region_segments(self, size, seg):
        b, e = np.zeros(1 << 16, np.uint64), np.zeros(1 << 16, np.uint64)
        n = self.dll.ptl_oracle_fn_region_segments(size, seg, self._p(b, u64p), self._p(e, u64p), 1 << 16)
        return [[int(b[i]), int(e[i])] for i in range(n)]

    def rev_comp(self, s):
        buf = C.create_string_buffer(s.encode(), len(s))
        self.dll.ptl_oracle_fn_rev_comp(buf, len(s))
        return buf.raw.decode()

    def strip_clip(self, trailing, cigar):
        cg = cigar_from_string(cigar)
        out = self._out()
        n = self.dll.ptl_oracle_fn_strip_clip(int(trailing), self._p(cg, u32p), len(cg), self._p(out, u32p), len(out))
        return cigar_to_string(out[:n])

    def parse_sa(self, sa):
        cap = 64
        pos, fwd, mq, nm, nc = np.zeros(cap, np.int64), np.zeros(cap, np.uint8), np.zeros(cap, np.uint8), np.zeros(cap, np.int32), np.zeros(cap, np.uint32)
        names = C.create_string_buffer(4096)
        n = self.dll.ptl_oracle_fn_parse_sa(sa.encode(), cap, self._p(pos, i64p), self._p(fwd, u8p), self._p(mq, u8p), self._p(nm, i32p), self._p(nc, u32p), names, 4096)
        if n < 0:
            return None
        rn = names.value.decode().split("\n")[:n]
        return [dict(rname=rn[i], pos=int(pos[i]), is_fwd=bool(fwd[i]), mapq=int(mq[i]), nm=int(nm[i]), n_cigar=int(nc[i])) for i in range(n)]

    def split_segments(self, names, tid, pos, flag, mapq, cigar, sa):
        return split_segments_call(self.dll.ptl_oracle_split_segments, self.dll.ptl_oracle_split_last_error, names, tid, pos, flag, mapq, cigar, sa)


def split_segments_call(fn, errfn, names, tid, pos, flag, mapq, cigar, sa, cap_seg=64, cap_cig=1 << 16):
    """Shared driver for ptl_pack_split_segments / ptl_oracle_split_segments (identical signatures)."""
    cg = cigar_from_string(cigar) if isinstance(cigar, str) else np.asarray(cigar, np.uint32)
    arr_names = (C.c_char_p * len(names))(*[n.encode() for n in names])
    o = SplitSegmentsC()
    bufs = dict(seq_order_start=np.zeros(cap_seg, np.uint32), seq_order_end=np.zeros(cap_seg, np.uint32), contig=np.zeros(cap_seg, np.uint32),
                pos=np.zeros(cap_seg, np.int64), is_fwd=np.zeros(cap_seg, np.uint8), mapq=np.zeros(cap_seg, np.uint8), from_primary=np.zeros(cap_seg, np.uint8),
                cigar_begin=np.zeros(cap_seg + 1, np.uint32), cigar=np.zeros(cap_cig, np.uint32))
    for k, a in bufs.items():
        setattr(o, k, a.ctypes.data_as(dict(SplitSegmentsC._fields_)[k]))
    ns, nc = C.c_uint32(), C.c_uint32()
    rc = fn(len(names), arr_names, tid, pos, flag, mapq, cg.ctypes.data_as(u32p), len(cg), sa.encode() if sa is not None else None, cap_seg, cap_cig, C.byref(o), C.byref(ns), C.byref(nc))
    if rc != 0:
        return dict(error=rc, message=errfn().decode())
    out = []
    for i in range(ns.value):
        cb, ce = int(bufs["cigar_begin"][i]), int(bufs["cigar_begin"][i + 1])
        out.append(dict(so=[int(bufs["seq_order_start"][i]), int(bufs["seq_order_end"][i])], contig=int(bufs["contig"][i]), pos=int(bufs["pos"][i]),
                        is_fwd=bool(bufs["is_fwd"][i]), cigar=cigar_to_string(bufs["cigar"][cb:ce]), mapq=int(bufs["mapq"][i]), from_primary=bool(bufs["from_primary"][i])))
    return out


_lib = None


class OraclePackedBatch:
    """The oracle's own record-loop head (skip supplementary, get_seq_order_read_split_segments): records -> ptl_batch.
    Lets the CPU arm of the bench run without the product library."""

    def __init__(self, recs_c, first, count, contig_names, threads=1):
        self.O = load()
        self.h = C.c_void_p()
        names = (C.c_char_p * max(len(contig_names), 1))(*[n.encode() for n in contig_names])
        rc = self.O.dll.ptl_oracle_pack_batch(C.byref(recs_c), first, count, len(contig_names), names, threads, C.byref(self.h))
        if rc != 0:
            raise abi.PtlError(rc, self.O.dll.ptl_oracle_pack_last_error().decode())
        self.c = abi.BatchC()
        self.O.dll.ptl_oracle_packed_view(self.h, C.byref(self.c))
        self.c._owner = self
        self._recs = recs_c

    def record_index(self):
        p = self.O.dll.ptl_oracle_packed_record_index(self.h)
        return np.ctypeslib.as_array(p, (self.c.n_reads,)).copy() if self.c.n_reads else np.zeros(0, np.uint32)

    def __del__(self):
        try:
            if self.h:
                self.O.dll.ptl_oracle_packed_free(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass


def load() -> OracleLib:
    global _lib
    if _lib is None:
        _lib = OracleLib(C.CDLL(build()))
    return _lib
