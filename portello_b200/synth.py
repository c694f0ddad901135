"""ctypes wrapper of the synthetic data generator (portello_b200/csrc/libptl_synth.so): bench + test infrastructure."""
from __future__ import annotations

import ctypes as C
import os

from .abi import ContigRecordsC, u8p, u64p
from .lib import CSRC, ReadRecordsC, build

SO = os.path.join(CSRC, "libptl_synth.so")


class SynthParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_chrom", C.c_uint32), ("chrom_len", C.c_uint64), ("haplotypes", C.c_uint32),
        ("contigs_per_chrom", C.c_uint32), ("rev_contig_frac", C.c_double), ("unmapped_contig_frac", C.c_double),
        ("contig_snv_rate", C.c_double), ("contig_indel_rate", C.c_double), ("sv_per_mb", C.c_double),
        ("junction_per_mb", C.c_double), ("n_reads", C.c_uint64), ("read_len_mean", C.c_double), ("read_len_sd", C.c_double),
        ("read_len_min", C.c_uint32), ("read_len_max", C.c_uint32), ("read_sub_rate", C.c_double), ("read_indel_rate", C.c_double),
        ("read_cluster_frac", C.c_double), ("read_clip_frac", C.c_double), ("read_sa_frac", C.c_double), ("n_threads", C.c_uint32),
        ("defer_reads", C.c_uint32),
    ]


_dll = None


def _load():
    global _dll
    if _dll is None:
        if not os.path.exists(SO):
            build()
        d = C.CDLL(SO)
        d.ptl_synth_default_params.argtypes = [C.POINTER(SynthParams)]
        d.ptl_synth_create.restype = C.c_void_p
        d.ptl_synth_create.argtypes = [C.POINTER(SynthParams)]
        d.ptl_synth_create_into.restype = C.c_void_p
        d.ptl_synth_create_into.argtypes = [C.POINTER(SynthParams), C.c_void_p, C.c_void_p]
        d.ptl_synth_destroy.argtypes = [C.c_void_p]
        d.ptl_synth_n_chrom.restype = C.c_uint32
        d.ptl_synth_n_chrom.argtypes = [C.c_void_p]
        d.ptl_synth_chrom_len.restype = u64p
        d.ptl_synth_chrom_len.argtypes = [C.c_void_p]
        d.ptl_synth_chrom_seq.restype = C.POINTER(u8p)
        d.ptl_synth_chrom_seq.argtypes = [C.c_void_p]
        d.ptl_synth_chrom_names.restype = C.POINTER(C.c_char_p)
        d.ptl_synth_chrom_names.argtypes = [C.c_void_p]
        d.ptl_synth_n_contigs.restype = C.c_uint32
        d.ptl_synth_n_contigs.argtypes = [C.c_void_p]
        d.ptl_synth_contig_names.restype = C.POINTER(C.c_char_p)
        d.ptl_synth_contig_names.argtypes = [C.c_void_p]
        d.ptl_synth_contig_records.argtypes = [C.c_void_p, C.POINTER(ContigRecordsC)]
        d.ptl_synth_read_records.argtypes = [C.c_void_p, C.POINTER(ReadRecordsC)]
        d.ptl_synth_n_planned.restype = C.c_uint64
        d.ptl_synth_n_planned.argtypes = [C.c_void_p]
        d.ptl_synth_plan_contig.restype = C.POINTER(C.c_uint32)
        d.ptl_synth_plan_contig.argtypes = [C.c_void_p]
        d.ptl_synth_plan_pos.restype = C.POINTER(C.c_int64)
        d.ptl_synth_plan_pos.argtypes = [C.c_void_p]
        d.ptl_synth_contig_len.restype = u64p
        d.ptl_synth_contig_len.argtypes = [C.c_void_p]
        d.ptl_synth_bam_stream.restype = C.POINTER(C.c_uint8)
        d.ptl_synth_bam_stream.argtypes = [C.c_void_p, C.c_int, C.c_uint32, u64p]
        d.ptl_synth_free_bytes.argtypes = [C.POINTER(C.c_uint8)]
        d.ptl_synth_generate_reads.restype = C.c_int
        d.ptl_synth_generate_reads.argtypes = [C.c_void_p, C.c_uint32, u64p, u64p]
        _dll = d
    return _dll


# Named workloads = BASELINE.json configs, scaled where noted.
WORKLOADS = {
    # configs[0]: 1 Mb reference, 2 contigs (one reverse-strand), 20k 15 kb reads
    # (adjacent I/D clusters are rare in HiFi-vs-own-assembly alignments: 0.5 % of read indels; configs[4] makes them dense)
    "config1": dict(seed=1001, n_chrom=1, chrom_len=1_000_000, haplotypes=2, contigs_per_chrom=1, n_reads=20_000, read_cluster_frac=0.005),
    # configs[1]: chr20-scale: 64 Mb reference, ~40 contig alignments carrying SVs, 1M reads
    "chr20": dict(seed=2002, n_chrom=1, chrom_len=64_000_000, haplotypes=2, contigs_per_chrom=5, junction_per_mb=0.25,
                  sv_per_mb=3.0, n_reads=1_000_000, read_cluster_frac=0.005),
    # configs[2]: whole-genome synthetic diploid assembly (~3.1 Gb reference, ~500 contigs per haplotype), 30x HiFi = ~6M reads
    "wg": dict(seed=3003, n_chrom=24, chrom_len=130_000_000, haplotypes=2, contigs_per_chrom=21, rev_contig_frac=0.45, junction_per_mb=0.25,
               sv_per_mb=3.0, n_reads=6_000_000, read_cluster_frac=0.005),
    # tiny cases for the CPU test-suite
    "tiny": dict(seed=7, n_chrom=2, chrom_len=400_000, haplotypes=2, contigs_per_chrom=2, junction_per_mb=12.0,
                 sv_per_mb=8.0, n_reads=3000, read_len_mean=6000, read_len_sd=1500, read_len_min=1000, read_len_max=12000,
                 read_sa_frac=0.05, read_clip_frac=0.05, unmapped_contig_frac=0.1),
    # configs[4]-like stress: fragmented assembly, long reads with dense clustered indels and SA segments
    "stress": dict(seed=5005, n_chrom=1, chrom_len=8_000_000, haplotypes=2, contigs_per_chrom=40, junction_per_mb=6.0,
                   sv_per_mb=6.0, rev_contig_frac=0.5, n_reads=20_000, read_len_mean=100_000, read_len_sd=15_000,
                   read_len_min=20_000, read_len_max=150_000, read_indel_rate=5e-3, read_cluster_frac=0.3, read_sa_frac=0.1),
    # configs[4] at the size SURVEY.md §8d states: 64 Mb reference, ~1700 contigs (median 46 kb; two haplotypes, random cuts), half of
    # them reverse-strand, 200k x 100 kb reads (cut at contig ends), indel rate 5e-3 with 30 % adjacent I/D clusters, 10 % with SA
    "stress_full": dict(seed=5005, n_chrom=1, chrom_len=64_000_000, haplotypes=2, contigs_per_chrom=1600, junction_per_mb=20.0,
                        sv_per_mb=6.0, rev_contig_frac=0.5, n_reads=200_000, read_len_mean=100_000, read_len_sd=15_000,
                        read_len_min=20_000, read_len_max=150_000, read_indel_rate=5e-3, read_cluster_frac=0.3, read_sa_frac=0.1),
}


def params(**kw) -> SynthParams:
    p = SynthParams()
    _load().ptl_synth_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


class Synth:
    def __init__(self, p: SynthParams, host_alloc=None, host_free=None):
        d = _load()
        self.p = p
        if host_alloc is not None:
            self.h = d.ptl_synth_create_into(C.byref(p), C.cast(host_alloc, C.c_void_p), C.cast(host_free, C.c_void_p))
        else:
            self.h = d.ptl_synth_create(C.byref(p))
        if not self.h:
            raise RuntimeError("ptl_synth_create failed (bad parameters?)")
        self.n_chrom = d.ptl_synth_n_chrom(self.h)
        self.n_contigs = d.ptl_synth_n_contigs(self.h)
        self.chrom_len = d.ptl_synth_chrom_len(self.h)
        self.chrom_seq = d.ptl_synth_chrom_seq(self.h)
        cn = d.ptl_synth_chrom_names(self.h)
        self.chrom_names = [cn[i].decode() for i in range(self.n_chrom)]
        tn = d.ptl_synth_contig_names(self.h)
        self.contig_names = [tn[i].decode() for i in range(self.n_contigs)]
        self.contig_records = ContigRecordsC()
        d.ptl_synth_contig_records(self.h, C.byref(self.contig_records))
        self.read_records = ReadRecordsC()
        d.ptl_synth_read_records(self.h, C.byref(self.read_records))

    # ---- the planned read set (BAM order) and shard-wise generation
    def plan(self):
        """(contig index, record position) of EVERY read of the set in coordinate-sorted BAM order (numpy views)."""
        import numpy as np

        d = _load()
        n = int(d.ptl_synth_n_planned(self.h))
        if n == 0:
            return np.zeros(0, np.uint32), np.zeros(0, np.int64)
        return np.ctypeslib.as_array(d.ptl_synth_plan_contig(self.h), (n,)), np.ctypeslib.as_array(d.ptl_synth_plan_pos(self.h), (n,))

    def contig_lengths(self):
        import numpy as np

        return np.ctypeslib.as_array(_load().ptl_synth_contig_len(self.h), (self.n_contigs,))

    def generate_reads(self, ranges):
        """Generate the reads of `ranges` = [(first, count), ...] of the BAM order (concatenated); self.read_records then
        views them (the previous read records are released)."""
        import numpy as np

        first = np.ascontiguousarray([r[0] for r in ranges], np.uint64)
        count = np.ascontiguousarray([r[1] for r in ranges], np.uint64)
        rc = _load().ptl_synth_generate_reads(self.h, len(ranges), first.ctypes.data_as(u64p), count.ctypes.data_as(u64p))
        if rc != 0:
            raise ValueError("ptl_synth_generate_reads: range outside the planned read set")
        _load().ptl_synth_read_records(self.h, C.byref(self.read_records))

    def reference_arrays(self):
        """The chromosomes as borrowed numpy uint8 views (valid while this object lives)."""
        import numpy as np

        return [np.ctypeslib.as_array(self.chrom_seq[i], (int(self.chrom_len[i]),)) for i in range(self.n_chrom)]

    def close(self):
        if self.h:
            _load().ptl_synth_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bam_stream(s: "Synth", which: int, n_unmapped: int = 0):
    """The data set as an uncompressed BAM byte stream (numpy uint8): which = 0 contig->reference, 1 read->contig."""
    import numpy as np

    d = _load()
    n = C.c_uint64()
    p = d.ptl_synth_bam_stream(s.h, which, n_unmapped, C.byref(n))
    if not p:
        raise MemoryError("ptl_synth_bam_stream failed")
    try:
        return np.ctypeslib.as_array(p, (int(n.value),)).copy()
    finally:
        d.ptl_synth_free_bytes(p)


def make(name_or_kw, host_alloc=None, host_free=None, **override) -> Synth:
    kw = dict(WORKLOADS[name_or_kw]) if isinstance(name_or_kw, str) else dict(name_or_kw)
    kw.update(override)
    return Synth(params(**kw), host_alloc, host_free)
