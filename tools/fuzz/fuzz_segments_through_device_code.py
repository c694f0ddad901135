"""Mutated ptl_contig_segments (a public entry point: ptl_set_contig_segments) through the device code compiled for the host:
the call must reject what leaves a pool, and whatever it accepts must not take the table build or the lift outside their
buffers.  Run against the AddressSanitizer build of tests/emul (README.md).
usage: python tools/fuzz/fuzz_segments_through_device_code.py <seed> <iterations>"""
import os, sys, random, copy
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import emul_lib, helpers
from portello_b200 import abi, synth, lib
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 200
s = synth.make("tiny", seed=29, n_reads=150, rev_contig_frac=0.5, junction_per_mb=15)
pb = helpers.pack(s, windows=True)
base = lib.load().prepare_contig_records(s.contig_records)
fields = ["contig_len", "contig_seg_begin", "seg_seq_order_start", "seg_seq_order_end", "seg_chrom_index", "seg_pos", "seg_is_fwd", "seg_mapq", "seg_cigar_begin", "cigar"]
n_ok = n_rej = 0
for it in range(n_it):
    g = copy.copy(base)
    g._keep = []
    log = []
    for _ in range(rng.randint(1, 3)):
        f = rng.choice(fields)
        arr = np.array(getattr(g, f), copy=True)
        for _ in range(rng.randint(1, 3)):
            i = rng.randrange(len(arr))
            info = np.iinfo(arr.dtype)
            k = rng.random()
            if k < 0.35: arr[i] = rng.randrange(int(info.min), int(info.max) + 1) if info.max < 2**62 else rng.randrange(0, 2**62)
            elif k < 0.65: arr[i] = int(arr[i]) ^ (1 << rng.randrange(0, 8 * arr.dtype.itemsize - (1 if info.min < 0 else 0)))
            elif k < 0.8: arr[i] = 0
            else: arr[i] = info.max
            if f == "contig_len":  # (rev_contig_seq[c] holds contig_len[c] bytes by contract: the claim may shrink, not grow)
                arr[i] = min(int(arr[i]), int(getattr(base, f)[i]))
            log.append((f, i, int(arr[i])))
        setattr(g, f, arr)
    print("case", it, log, flush=True)
    ctx = abi.Context(emul_lib.load(), 0, 1)
    try:
        ctx.set_reference(helpers.reference_arrays(s))
        ctx.set_contig_segments(g)
        helpers.lift_c(ctx, pb.c, allow_panic=True)
        n_ok += 1
    except abi.PtlError:
        n_rej += 1
print("done", n_ok, n_rej)
