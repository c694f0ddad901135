"""ctypes wrappers of the file-level entry points of libportello_b200.so (include/portello_b200.h: ptl_bam_*, ptl_decoded_*,
ptl_fasta_*, ptl_scan_contig_bam, ptl_bam_header, ptl_bgzf_compress, ptl_bam_index_build) + helpers that write a synthetic
data set as real files (FASTA, two indexed BAMs).  Host-side only: no GPU needed."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import lib
from .abi import ContigRecordsC, PtlError, ReadExtrasC, u8p, u64p
from .lib import ReadRecordsC

FETCH_ALL, FETCH_UNMAPPED = -2, -1
START_IN_REGION, SKIP_SUPPLEMENTARY, SKIP_UNMAPPED_SECONDARY, ONLY_UNMAPPED, KEEP_RAW = 1, 2, 4, 8, 16

_ready = False


def _dll():
    global _ready
    d = lib.load().dll
    if not _ready:
        d.ptl_bam_open.restype, d.ptl_bam_open.argtypes = C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]
        d.ptl_bam_close.argtypes = [C.c_void_p]
        d.ptl_bam_last_error.restype = C.c_char_p
        d.ptl_bam_n_ref.restype, d.ptl_bam_n_ref.argtypes = C.c_uint32, [C.c_void_p]
        d.ptl_bam_ref_name.restype, d.ptl_bam_ref_name.argtypes = C.c_char_p, [C.c_void_p, C.c_uint32]
        d.ptl_bam_ref_len.restype, d.ptl_bam_ref_len.argtypes = C.c_uint64, [C.c_void_p, C.c_uint32]
        d.ptl_bam_header_text.restype, d.ptl_bam_header_text.argtypes = C.c_char_p, [C.c_void_p]
        d.ptl_bam_has_index.restype, d.ptl_bam_has_index.argtypes = C.c_int, [C.c_void_p]
        d.ptl_bam_has_eof_marker.restype, d.ptl_bam_has_eof_marker.argtypes = C.c_int, [C.c_void_p]
        d.ptl_bam_fetch.restype = C.c_int
        d.ptl_bam_fetch.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_uint32, C.POINTER(C.c_void_p)]
        d.ptl_decoded_view.argtypes = [C.c_void_p, C.POINTER(ReadRecordsC), C.POINTER(ReadExtrasC)]
        d.ptl_decoded_raw.restype, d.ptl_decoded_raw.argtypes = u8p, [C.c_void_p, C.POINTER(u64p), u64p]
        d.ptl_decoded_free.argtypes = [C.c_void_p]
        d.ptl_bam_index_build.restype, d.ptl_bam_index_build.argtypes = C.c_int, [C.c_char_p, C.c_char_p]
        d.ptl_bam_index_build_csi.restype, d.ptl_bam_index_build_csi.argtypes = C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        d.ptl_fasta_load.restype, d.ptl_fasta_load.argtypes = C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        d.ptl_fasta_n.restype, d.ptl_fasta_n.argtypes = C.c_uint32, [C.c_void_p]
        d.ptl_fasta_name.restype, d.ptl_fasta_name.argtypes = C.c_char_p, [C.c_void_p, C.c_uint32]
        d.ptl_fasta_seq.restype, d.ptl_fasta_seq.argtypes = u8p, [C.c_void_p, C.c_uint32, u64p]
        d.ptl_fasta_free.argtypes = [C.c_void_p]
        d.ptl_scan_contig_bam.restype = C.c_int
        d.ptl_scan_contig_bam.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_char_p), u64p, C.c_int, C.POINTER(C.c_void_p)]
        d.ptl_contig_scan_view.argtypes = [C.c_void_p, C.POINTER(ContigRecordsC)]
        d.ptl_contig_scan_free.argtypes = [C.c_void_p]
        d.ptl_bgzf_bound.restype, d.ptl_bgzf_bound.argtypes = C.c_uint64, [C.c_uint64]
        d.ptl_bgzf_compress.restype = C.c_int64
        d.ptl_bgzf_compress.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
        _ready = True
    return d


def _check(rc):
    if rc != 0:
        raise PtlError(rc, _dll().ptl_bam_last_error().decode("utf-8", "replace"))


class Decoded:
    """Owner of a ptl_decoded_batch: `recs` (ReadRecordsC, what the packer takes) and `extras` (ReadExtrasC, what the
    record assembly takes) view its arrays."""

    def __init__(self, h):
        self.h = h
        self.recs, self.extras = ReadRecordsC(), ReadExtrasC()
        _dll().ptl_decoded_view(h, C.byref(self.recs), C.byref(self.extras))

    @property
    def n(self):
        return int(self.recs.n_reads)

    def arrays(self):
        """numpy copies of every decoded field."""
        r, x, n = self.recs, self.extras, self.n
        g = lambda p, k: np.ctypeslib.as_array(p, (k,)).copy() if k else np.zeros(0, np.uint8)
        cb = g(r.cigar_begin, n + 1) if n else np.zeros(1, np.uint64)
        no, ao = (g(x.name_off, n + 1), g(x.aux_off, n + 1)) if n else (np.zeros(1, np.uint64), np.zeros(1, np.uint64))
        out = dict(tid=g(r.tid, n), pos=g(r.pos, n), flag=g(r.flag, n), mapq=g(r.mapq, n), bin=g(r.bin, n), seq_len=g(r.seq_len, n), seq_off=g(r.seq_off, n),
                   seq4=g(r.seq4, int(r.seq4_bytes)), cigar_begin=cb, cigar=g(r.cigar, int(cb[-1])), name_off=no, names=g(x.names, int(no[-1])),
                   aux_off=ao, aux=g(x.aux, int(ao[-1])), mate_tid=g(x.mate_tid, n), mate_pos=g(x.mate_pos, n), tlen=g(x.tlen, n),
                   qual=g(x.quals.qual, int(x.quals.qual_bytes)), qual_off=g(x.quals.read_qual_off, n),
                   sa=[r.sa_tag[i] for i in range(n)])
        return out

    def raw(self):
        off, nb = u64p(), C.c_uint64()
        p = _dll().ptl_decoded_raw(self.h, C.byref(off), C.byref(nb))
        n_off = self.n + 1
        return (np.ctypeslib.as_array(p, (int(nb.value),)).copy() if nb.value else np.zeros(0, np.uint8),
                np.ctypeslib.as_array(off, (n_off,)).copy())

    def close(self):
        if self.h:
            _dll().ptl_decoded_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BamFile:
    def __init__(self, path: str):
        self.h = C.c_void_p()
        _check(_dll().ptl_bam_open(path.encode(), C.byref(self.h)))
        d = _dll()
        n = d.ptl_bam_n_ref(self.h)
        self.ref_names = [d.ptl_bam_ref_name(self.h, i).decode() for i in range(n)]
        self.ref_len = [int(d.ptl_bam_ref_len(self.h, i)) for i in range(n)]
        self.header_text = d.ptl_bam_header_text(self.h).decode()
        self.has_index = bool(d.ptl_bam_has_index(self.h))
        self.has_eof_marker = bool(d.ptl_bam_has_eof_marker(self.h))

    def fetch(self, tid: int, begin: int = 0, end: int = 0, flt: int = 0) -> Decoded:
        h = C.c_void_p()
        _check(_dll().ptl_bam_fetch(self.h, tid, begin, end, flt, C.byref(h)))
        return Decoded(h)

    def scan_contigs(self, contig_names, contig_len, threads: int = 4):
        """ptl_scan_contig_bam on this (contig->reference) file -> (owner handle, ContigRecordsC view)."""
        names = (C.c_char_p * max(len(contig_names), 1))(*[n.encode() for n in contig_names])
        lens = np.ascontiguousarray(contig_len, np.uint64)
        h = C.c_void_p()
        _check(_dll().ptl_scan_contig_bam(self.h, len(contig_names), names, lens.ctypes.data_as(u64p), threads, C.byref(h)))
        return ContigScan(h)

    def close(self):
        if self.h:
            _dll().ptl_bam_close(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ContigScan:
    def __init__(self, h):
        self.h = h
        self.c = ContigRecordsC()
        _dll().ptl_contig_scan_view(h, C.byref(self.c))
        self.c._owner = self

    def __del__(self):
        try:
            if self.h:
                _dll().ptl_contig_scan_free(self.h)
                self.h = None
        except Exception:
            pass


def load_fasta(path: str, threads: int = 4):
    """[(name, uint8 array)] as get_genome_ref_from_fasta yields them (upper-cased)."""
    d = _dll()
    h = C.c_void_p()
    _check(d.ptl_fasta_load(path.encode(), threads, C.byref(h)))
    try:
        out = []
        for i in range(d.ptl_fasta_n(h)):
            n = C.c_uint64()
            p = d.ptl_fasta_seq(h, i, C.byref(n))
            out.append((d.ptl_fasta_name(h, i).decode("utf-8", "surrogateescape"), np.ctypeslib.as_array(p, (int(n.value),)).copy() if n.value else np.zeros(0, np.uint8)))
        return out
    finally:
        d.ptl_fasta_free(h)


def bgzf_compress(data, level: int = 6, threads: int = 4, eof: bool = True) -> bytes:
    d = _dll()
    src = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data
    n = int(src.size)
    out = np.zeros(int(d.ptl_bgzf_bound(n)), np.uint8)
    k = d.ptl_bgzf_compress(src.ctypes.data if n else None, n, level, threads, int(eof), out.ctypes.data, out.size)
    if k < 0:
        raise PtlError(int(-k), "ptl_bgzf_compress failed")
    return out[:k].tobytes()


def index_bam(path: str, bai_path: str | None = None):
    _check(_dll().ptl_bam_index_build(path.encode(), bai_path.encode() if bai_path else None))


def index_bam_csi(path: str, csi_path: str | None = None, min_shift: int = 14, depth: int = 0):
    _check(_dll().ptl_bam_index_build_csi(path.encode(), csi_path.encode() if csi_path else None, min_shift, depth))


def write_fasta(path: str, names, seqs, width: int = 60):
    with open(path, "wb") as f:
        for name, seq in zip(names, seqs):
            f.write(b">" + name.encode() + b" synthetic\n")
            a = np.asarray(seq, np.uint8)
            full = (len(a) // width) * width
            if full:
                body = np.empty((full // width, width + 1), np.uint8)
                body[:, :width] = a[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if full < len(a):
                f.write(a[full:].tobytes() + b"\n")


def write_dataset(s, out_dir: str, n_unmapped: int = 0, level: int = 1, threads: int = 4):
    """A synthetic data set (portello_b200.synth.Synth) as the three input files portello takes:
    ref.fa, contigs_to_ref.bam (+ .bai), reads_to_contigs.bam (+ .bai).  Returns their paths."""
    from . import synth as synth_mod

    os.makedirs(out_dir, exist_ok=True)
    fa = os.path.join(out_dir, "ref.fa")
    write_fasta(fa, s.chrom_names, s.reference_arrays())
    paths = {"ref": fa}
    for which, name in ((0, "contigs_to_ref.bam"), (1, "reads_to_contigs.bam")):
        stream = synth_mod.bam_stream(s, which, n_unmapped if which == 1 else 0)
        p = os.path.join(out_dir, name)
        with open(p, "wb") as f:
            f.write(bgzf_compress(stream, level=level, threads=threads, eof=True))
        index_bam(p)
        paths["contigs" if which == 0 else "reads"] = p
    return paths


if __name__ == "__main__":
    # python -m portello_b200.bamio <workload> <n_reads> <out_dir> [n_unmapped]: a synthetic data set as the three input files
    # of portello (and of portello-b200), e.g. to run the real portello on the same inputs on a machine that has it:
    #   portello --assembly-to-ref <out_dir>/contigs_to_ref.bam --read-to-assembly <out_dir>/reads_to_contigs.bam \
    #            --ref <out_dir>/ref.fa --remapped-read-output remapped.bam --unassembled-read-output unassembled.bam
    import sys

    from . import synth as _synth

    wl, n, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    n_un = int(sys.argv[4]) if len(sys.argv) > 4 else n // 200
    print(write_dataset(_synth.make(wl, n_reads=n), out, n_unmapped=n_un))
