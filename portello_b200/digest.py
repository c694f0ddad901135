"""Order-sensitive, shard-decomposable digest of liftover results (bench + test infrastructure, numpy only).

Every output record is hashed together with its KEY = (index of its read in the coordinate-sorted input, index of the
record within the read), every CIGAR op together with its index within the record; the digest of a result set is the
wrapping 64-bit SUM of the record hashes (plus the record and op counts).  So

  * a swapped, missing, duplicated or altered record / op changes the digest (order-sensitive through the keys);
  * the digest of a whole run is the sum of the digests of ANY partition of it into batches, chunks or per-GPU shards,
    which is what lets `bench.py --gpus N` prove that N disjoint shards gathered in unit order equal the N = 1 run, and
    lets the oracle (CPU) and the CUDA path be compared on 100 % of a 6 M-read set without holding both results at once.

`rec_read_segment` is batch-local, so it enters as the segment's index within its read.
"""
from __future__ import annotations

import numpy as np

U64 = np.uint64
_M1, _M2, _G = U64(0xBF58476D1CE4E5B9), U64(0x94D049BB133111EB), U64(0x9E3779B97F4A7C15)


def _mix(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser, vectorised (wrapping arithmetic)."""
    x = x.astype(U64, copy=True)
    x ^= x >> U64(30)
    x *= _M1
    x ^= x >> U64(27)
    x *= _M2
    x ^= x >> U64(31)
    return x


class _Slice:
    pass


class Digest:
    """Accumulates (sum of record hashes, records, ops) over any number of result batches."""

    SLICE = 1 << 18  # reads per internal slice

    def __init__(self):
        self.acc = 0
        self.n_records = 0
        self.n_ops = 0
        self.n_reads = 0

    def add(self, res, read_seg_begin: np.ndarray, global_read_index: np.ndarray):
        """`res`: an abi.Result (numpy views are fine); `read_seg_begin`: the batch's CSR [n_reads+1];
        `global_read_index[r]`: index of batch read r in the coordinate-sorted input set."""
        n_all = len(read_seg_begin) - 1
        if n_all > self.SLICE:  # the digest is a sum over reads: bound the temporaries (a 6 M-read batch has 1.3e8 ops)
            rrb_all = np.asarray(res.read_rec_begin)
            cb_all = np.asarray(res.rec_cigar_begin)
            for a in range(0, n_all, self.SLICE):
                b = min(n_all, a + self.SLICE)
                k0, k1 = int(rrb_all[a]), int(rrb_all[b])
                c0, c1 = int(cb_all[k0]), int(cb_all[k1])
                part = _Slice()
                part.read_rec_begin = np.asarray(rrb_all[a:b + 1], dtype=np.int64) - k0
                part.rec_cigar_begin = np.asarray(cb_all[k0:k1 + 1], dtype=np.int64) - c0
                part.cigar = res.cigar[c0:c1]
                part.rec_read_segment = np.asarray(res.rec_read_segment[k0:k1], dtype=np.int64) - int(read_seg_begin[a])
                for f in ("rec_status", "rec_contig_segment", "rec_tid", "rec_pos", "rec_mapq", "rec_flag", "rec_bin", "rec_need_flip"):
                    setattr(part, f, getattr(res, f)[k0:k1])
                self.add(part, np.asarray(read_seg_begin[a:b + 1], dtype=np.int64) - int(read_seg_begin[a]), global_read_index[a:b])
            return self
        with np.errstate(over="ignore"):
            n_reads = len(read_seg_begin) - 1
            rrb = np.asarray(res.read_rec_begin, dtype=np.int64)
            assert len(rrb) == n_reads + 1 and len(global_read_index) == n_reads
            n_rec = int(rrb[-1])
            per_read = np.diff(rrb)
            rec_read = np.repeat(np.arange(n_reads, dtype=np.int64), per_read)
            j = np.arange(n_rec, dtype=np.int64) - np.repeat(rrb[:-1], per_read)
            key = np.asarray(global_read_index, dtype=np.int64)[rec_read].astype(U64) * U64(4096) + j.astype(U64)
            seg_in_read = np.asarray(res.rec_read_segment, dtype=np.int64) - np.asarray(read_seg_begin, dtype=np.int64)[rec_read]
            cb = np.asarray(res.rec_cigar_begin, dtype=np.int64)
            n_ops = int(cb[-1]) if n_rec else 0
            cnt = np.diff(cb)
            cig = np.asarray(res.cigar[:n_ops], dtype=U64)
            idx = np.arange(n_ops, dtype=np.int64) - np.repeat(cb[:-1], cnt)
            m = _mix(cig ^ (idx.astype(U64) * _G))
            csum = np.zeros(n_ops + 1, U64)
            np.cumsum(m, out=csum[1:])
            cig_hash = csum[cb[1:]] - csum[cb[:-1]]
            h = _mix(key)
            for salt, f in enumerate((res.rec_status, seg_in_read, res.rec_contig_segment, res.rec_tid, res.rec_pos, res.rec_mapq, res.rec_flag,
                                      res.rec_bin, res.rec_need_flip, cnt, cig_hash), start=1):
                v = np.asarray(f)
                v = v.astype(np.int64).astype(U64) if v.dtype != U64 else v
                h = _mix(h ^ (v + U64(salt) * _G))
            self.acc = (self.acc + int(h.sum(dtype=U64))) & 0xFFFFFFFFFFFFFFFF
            self.n_records += n_rec
            self.n_ops += n_ops
            self.n_reads += n_reads
        return self

    def merge(self, other: "Digest"):
        self.acc = (self.acc + other.acc) & 0xFFFFFFFFFFFFFFFF
        self.n_records += other.n_records
        self.n_ops += other.n_ops
        self.n_reads += other.n_reads
        return self

    def as_tuple(self):
        return (self.acc, self.n_records, self.n_ops, self.n_reads)

    def hex(self) -> str:
        return f"{self.acc:016x}/{self.n_records}r/{self.n_ops}ops"

    def __eq__(self, other):
        return isinstance(other, Digest) and self.as_tuple() == other.as_tuple()
