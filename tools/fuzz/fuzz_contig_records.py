"""Mutated ptl_contig_records (the input of ptl_set_contig_records / ptl_prepare_contig_records) through the host contig
preparation.  Run against the AddressSanitizer build of the host objects (README.md).
usage: python tools/fuzz/fuzz_contig_records.py <seed> <iterations>"""
import os, sys, random, ctypes as C
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import numpy as np
from portello_b200 import abi, synth, lib
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 200
s = synth.make("tiny", seed=37, n_reads=10, junction_per_mb=20, rev_contig_frac=0.5)
L = lib.load()
r0 = s.contig_records
n = r0.n_records
n_cig = int(r0.cigar_begin[n])
fields = {"contig_id": (n, abi.u32p), "flag": (n, abi.u16p), "tid": (n, abi.i32p), "pos": (n, abi.i64p), "mapq": (n, abi.u8p),
          "cigar_begin": (n + 1, abi.u64p), "cigar": (n_cig, abi.u32p), "contig_len": (r0.n_contigs, abi.u64p)}
sa0 = [r0.sa_tag[i] for i in range(n)]
n_ok = n_rej = 0
for it in range(n_it):
    r = abi.ContigRecordsC.from_buffer_copy(r0)
    keep, log = [], []
    for _ in range(rng.randint(1, 3)):
        if rng.random() < 0.2:      # an SA tag
            sa = list(sa0)
            i = rng.randrange(n)
            t = bytearray(sa[i] or b"chrA,100,+,50M,60,0;")
            for _ in range(rng.randint(1, 4)):
                k = rng.random()
                if k < 0.5 and t: t[rng.randrange(len(t))] = rng.choice(b",;+-0123456789MIDSH=X*c")
                elif k < 0.7 and t: del t[rng.randrange(len(t))]
                else: t += bytes(t)
            sa[i] = bytes(t).replace(b"\0", b"")
            arr = (C.c_char_p * n)(*sa)
            r.sa_tag = arr
            keep.append(arr)
            log.append(("sa", i, sa[i][:60]))
            continue
        f = rng.choice(list(fields))
        cnt, typ = fields[f]
        arr = np.ctypeslib.as_array(getattr(r0, f), (cnt,)).copy()
        for _ in range(rng.randint(1, 3)):
            i = rng.randrange(cnt - (1 if f == "cigar_begin" else 0))   # (the last cigar_begin is the size of the caller's pool)
            info = np.iinfo(arr.dtype)
            k = rng.random()
            if k < 0.35: arr[i] = rng.randrange(int(info.min), int(info.max) + 1) if info.max < 2**62 else rng.randrange(0, 2**62)
            elif k < 0.65: arr[i] = int(arr[i]) ^ (1 << rng.randrange(0, 8 * arr.dtype.itemsize - (1 if info.min < 0 else 0)))
            elif k < 0.8: arr[i] = 0
            else: arr[i] = info.max
            if f == "contig_len": arr[i] = min(int(arr[i]), int(r0.contig_len[i]))  # (seq[k] holds contig_len bytes by contract)
            if f == "contig_id" and int(arr[i]) < r0.n_contigs and int(r0.contig_len[int(arr[i])]) > int(r0.contig_len[int(r0.contig_id[i])]):
                arr[i] = r0.contig_id[i]                                            # (same contract: not onto a longer contig)
            log.append((f, i, int(arr[i])))
        setattr(r, f, arr.ctypes.data_as(typ))
        keep.append(arr)
    print("case", it, log, flush=True)
    try:
        L.prepare_contig_records(r)
        n_ok += 1
    except abi.PtlError:
        n_rej += 1
print("done", n_ok, n_rej)
