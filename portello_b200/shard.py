"""Work-unit sharding for multi-GPU runs (host-side plumbing, no compute).

The reference's scheduling grain is (contig x <=20 Mb window), reads assigned by the start position of their primary
record (src/read_alignment_scanner.rs:403,508-534,574-576; get_region_segments lib/rust-vc-utils/src/util.rs:50-67).
Units are weighted by read count and bin-packed onto ranks (greedy LPT, `ptl_shard_units`); every rank computes the same
map, lifts only its units, and the host concatenates unit results in unit order = the reference's record order.
No collective is involved in the data path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

from . import lib

SEGMENT_SIZE = 20_000_000


@dataclass
class Unit:
    contig: int
    begin: int
    end: int
    first_read: int   # index of the unit's first record in the (contig, pos)-sorted record stream
    n_reads: int


def window_units(contig_len, read_tid, read_pos, segment_size: int = SEGMENT_SIZE, region_segments=None) -> List[Unit]:
    """Split every contig into the reference's windows and locate each window's records (sorted input).
    `region_segments(size, segment_size) -> [(begin, end)]` defaults to the product's ptl_region_segments."""
    region_segments = region_segments or lib.load().region_segments
    read_tid = np.asarray(read_tid, dtype=np.int64)
    read_pos = np.asarray(read_pos, dtype=np.int64)
    key = read_tid * (1 << 40) + read_pos
    assert np.all(np.diff(key) >= 0), "records must be sorted by (contig, pos), like a coordinate-sorted BAM"
    units = []
    for c, clen in enumerate(contig_len):
        if int(clen) == 0:
            continue
        for b, e in region_segments(int(clen), segment_size):
            lo = int(np.searchsorted(key, c * (1 << 40) + b, side="left"))
            hi = int(np.searchsorted(key, c * (1 << 40) + e, side="left"))
            units.append(Unit(c, b, e, lo, hi - lo))
    return units


def assign(units: List[Unit], n_ranks: int) -> np.ndarray:
    """owner[i] = rank that lifts unit i (LPT on read counts; deterministic)."""
    return lib.load().shard_units([max(u.n_reads, 0) for u in units], n_ranks)


def rank_ranges(units: List[Unit], owner, rank: int):
    """The record ranges [(first_read, n_reads), ...] of the units `rank` owns, in unit order (= the order in which the
    host concatenates results: the reference's record order), and the global index of every record of the shard."""
    ranges = [(u.first_read, u.n_reads) for ui, u in enumerate(units) if int(owner[ui]) == rank and u.n_reads > 0]
    gri = np.concatenate([np.arange(a, a + n, dtype=np.int64) for a, n in ranges]) if ranges else np.zeros(0, np.int64)
    return ranges, gri
