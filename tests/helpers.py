"""Shared helpers for the parity tests: build a synthetic dataset, drive the oracle and the CUDA product through the
same C-ABI structs, compare bit-exactly."""
import ctypes as C

import numpy as np

import oracle_lib
from portello_b200 import abi, lib, synth


def reference_arrays(s):
    return [np.ctypeslib.as_array(s.chrom_seq[i], (int(s.chrom_len[i]),)) for i in range(s.n_chrom)]


def oracle_context(s, threads=4, faithful=False):
    O = oracle_lib.load()
    ctx = abi.Context(O, 0, 1)
    ctx.set_reference(reference_arrays(s))
    ctx.set_contig_records(s.contig_records)
    O.dll.ptl_oracle_set_threads(ctx.h, threads)
    O.dll.ptl_oracle_set_faithful_decode(ctx.h, int(faithful))
    return ctx


def gpu_context(s, n_slots=2):
    ctx = lib.GpuContext(0, n_slots)
    ctx.set_reference(reference_arrays(s))
    ctx.set_contig_records(s.contig_records)
    return ctx


def pack(s, first=0, count=None, pinned=False, windows=None):
    n = s.read_records.n_reads
    count = n - first if count is None else count
    return lib.PackedBatch(lib.load(), s.read_records, first, count, s.contig_names, pinned, windows=windows)


def lift_c(ctx, batch_c, stage_mask=abi.STAGE_ALL, slot=0, allow_panic=False):
    """submit/wait with a raw BatchC (avoids numpy round trips for big batches)."""
    L = ctx.lib
    if stage_mask == abi.STAGE_ALL:
        ctx._check(L._lift_submit(ctx.h, slot, C.byref(batch_c)))
    else:
        ctx._check(L._lift_submit_ex(ctx.h, slot, C.byref(batch_c), stage_mask))
    r = abi.ResultC()
    ctx._check(L._lift_wait(ctx.h, slot, C.byref(r)), allow=(abi.PTL_ERR_LIFT_PANIC,) if allow_panic else ())
    return abi.Result.from_c(r)


def malformed_batches(s):
    """Batches whose indices leave their pools (the reference would panic on the slice index): one per field."""
    pb = pack(s)
    b = pb.c
    out = []
    def clone(field, dtype, n, mutate):
        arr = np.ctypeslib.as_array(getattr(b, field), (n,)).copy()
        mutate(arr)
        b2 = abi.BatchC.from_buffer_copy(b)
        setattr(b2, field, arr.ctypes.data_as(dict(abi.BatchC._fields_)[field]))
        b2._keep = (arr, pb)
        return b2
    ns, n = b.n_read_segments, b.n_reads
    out.append(("contig", clone("rseg_contig", np.uint32, ns, lambda a: a.__setitem__(7, 10**6))))
    out.append(("cigar range", clone("rseg_cigar_len", np.uint32, ns, lambda a: a.__setitem__(ns - 1, 10**8))))
    out.append(("position", clone("rseg_pos", np.int64, ns, lambda a: a.__setitem__(3, -5))))
    out.append(("position int32", clone("rseg_pos", np.int64, ns, lambda a: a.__setitem__(3, 2**31))))
    out.append(("bases", clone("read_seq_off", np.uint64, n, lambda a: a.__setitem__(n - 1, int(b.seq4_bytes)))))
    out.append(("segment csr", clone("read_seg_begin", np.uint32, n + 1, lambda a: a.__setitem__(5, ns + 9))))
    # offsets whose sum with the length wraps around 2^64 (found by tools/fuzz/fuzz_batch_through_device_code.py under ASAN)
    out.append(("cigar range wrap", clone("rseg_cigar_begin", np.uint64, ns, lambda a: a.__setitem__(ns // 2, 2**64 - 1))))
    out.append(("bases wrap", clone("read_seq_off", np.uint64, n, lambda a: a.__setitem__(n // 2, 2**64 - 1))))
    return out


def single_pair_case(c2r_cigar, c2r_pos, c2r_fwd, contig_len, rev_seq, pos, cigar, read_seq4, seq_len, read_flag=0, rseg_fwd=1,
                     mapq=60):
    """One contig with one segment + one read with one segment (golden-vector replay through the full ABI)."""
    cg2 = abi.cigar_from_string(c2r_cigar) if isinstance(c2r_cigar, str) else np.asarray(c2r_cigar, np.uint32)
    segs = abi.ContigSegments(
        contig_len=np.array([contig_len], np.uint64), contig_seg_begin=np.array([0, 1], np.uint32),
        rev_contig_seq=[None if rev_seq is None else np.frombuffer(rev_seq.encode() if isinstance(rev_seq, str) else rev_seq, dtype=np.uint8).copy()],
        seg_seq_order_start=np.array([0], np.uint32), seg_seq_order_end=np.array([contig_len], np.uint32),
        seg_chrom_index=np.array([0], np.int32), seg_pos=np.array([c2r_pos], np.int64), seg_is_fwd=np.array([int(c2r_fwd)], np.uint8),
        seg_mapq=np.array([mapq], np.uint8), seg_cigar_begin=np.array([0, len(cg2)], np.uint64), cigar=cg2)
    cg = abi.cigar_from_string(cigar) if isinstance(cigar, str) else np.asarray(cigar, np.uint32)
    batch = abi.Batch(
        read_flag=np.array([read_flag], np.uint16), read_mapq=np.array([33], np.uint8), read_bin=np.array([4681], np.uint16),
        read_seq_len=np.array([seq_len], np.uint32), read_seq_off=np.array([0], np.uint64), read_seg_begin=np.array([0, 1], np.uint32),
        rseg_contig=np.array([0], np.uint32), rseg_pos=np.array([pos], np.int64), rseg_is_fwd=np.array([rseg_fwd], np.uint8),
        rseg_cigar_begin=np.array([0], np.uint64), rseg_cigar_len=np.array([len(cg)], np.uint32), cigar=cg,
        seq4=np.asarray(read_seq4, np.uint8) if len(read_seq4) else np.zeros(max(8, (seq_len + 1) // 2), np.uint8))
    return segs, batch


def cigar_read_len(cigar: str) -> int:
    return int(sum(int(v) >> 4 for v in abi.cigar_from_string(cigar) if (int(v) & 15) in (0, 1, 4, 5, 7, 8)))


def cigar_ref_len(cigar: str) -> int:
    return int(sum(int(v) >> 4 for v in abi.cigar_from_string(cigar) if (int(v) & 15) in (0, 2, 3, 7, 8)))


# Golden vectors use letters outside the BAM 4-bit alphabet; only equality matters, so relabel bijectively.
LETTER_MAP = str.maketrans({"A": "A", "B": "C", "C": "G", "D": "T", "E": "M", "F": "R", "X": "N"})


def relabel(s: str) -> str:
    return s.translate(LETTER_MAP)


def bam_header_bytes(text: str, names, lens) -> bytes:
    """ptl_bam_header: the uncompressed BAM header block."""
    d = lib.load().dll
    d.ptl_bam_header.restype = C.c_int64
    d.ptl_bam_header.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64]
    arr = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
    ln = (C.c_uint64 * max(len(lens), 1))(*lens)
    out = np.zeros(len(text) + 64 + sum(len(n) + 16 for n in names), np.uint8)
    n = d.ptl_bam_header(text.encode(), len(names), arr, ln, out.ctypes.data, out.size)
    assert n >= 0, n
    return out[:n].tobytes()
