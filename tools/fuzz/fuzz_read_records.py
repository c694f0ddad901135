"""Mutated ptl_read_records (the input of ptl_pack_batch*) through the host packer, windows included.  Run against the
AddressSanitizer build of the host objects (README.md).
usage: python tools/fuzz/fuzz_read_records.py <seed> <iterations>"""
import os, sys, random, ctypes as C
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import numpy as np
from portello_b200 import abi, synth, lib
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 200
s = synth.make("tiny", seed=41, n_reads=200, read_sa_frac=0.2, read_cluster_frac=0.3)
L = lib.load()
segs = L.prepare_contig_records(s.contig_records)
r0 = s.read_records
n = r0.n_reads
n_cig = int(r0.cigar_begin[n])
T = dict(lib.ReadRecordsC._fields_)
fields = {"tid": n, "pos": n, "flag": n, "mapq": n, "bin": n, "seq_len": n, "seq_off": n, "cigar_begin": n + 1, "cigar": n_cig}
n_ok = n_rej = 0
for it in range(n_it):
    r = lib.ReadRecordsC.from_buffer_copy(r0)
    keep, log = [], []
    for _ in range(rng.randint(1, 3)):
        f = rng.choice(list(fields))
        arr = np.ctypeslib.as_array(getattr(r0, f), (fields[f],)).copy()
        for _ in range(rng.randint(1, 3)):
            i = rng.randrange(fields[f] - (1 if f == "cigar_begin" else 0))   # (the last cigar_begin is the size of the caller's pool)
            info = np.iinfo(arr.dtype)
            k = rng.random()
            if k < 0.35: arr[i] = rng.randrange(int(info.min), int(info.max) + 1) if info.max < 2**62 else rng.randrange(0, 2**62)
            elif k < 0.65: arr[i] = int(arr[i]) ^ (1 << rng.randrange(0, 8 * arr.dtype.itemsize - (1 if info.min < 0 else 0)))
            elif k < 0.8: arr[i] = 0
            else: arr[i] = info.max
            log.append((f, i, int(arr[i])))
        setattr(r, f, arr.ctypes.data_as(T[f]))
        keep.append(arr)
    print("case", it, log, flush=True)
    for w in (None, True, segs):
        try:
            lib.PackedBatch(L, r, 0, n, s.contig_names, windows=w)
            n_ok += 1
        except abi.PtlError:
            n_rej += 1
print("done", n_ok, n_rej)
