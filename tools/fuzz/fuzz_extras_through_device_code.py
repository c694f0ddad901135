"""Mutated ptl_read_extras (names / aux / qualities / mate fields of ptl_assemble_records) through the record-assembly device
code compiled for the host: offsets outside their pools must be reported (error bit 4 of the assembly), mutated aux bytes
must at worst change which tags are stripped, and nothing may leave a buffer.  Run against the AddressSanitizer build.
usage: python tools/fuzz/fuzz_extras_through_device_code.py <seed> <iterations>"""
import os, sys, random
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import emul_lib, helpers
import test_assemble_records as TA
from portello_b200 import abi, synth
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 200
s = synth.make("tiny", seed=31, n_reads=120, read_sa_frac=0.2, rev_contig_frac=0.5)
pb = helpers.pack(s)
x0, _, _ = TA.make_extras(s, pb, 4)
ctx = abi.Context(emul_lib.load(), 0, 1)
ctx.set_reference(helpers.reference_arrays(s))
ctx.set_contig_records(s.contig_records)
ctx.set_names(s.contig_names, s.chrom_names)
n_ok = n_rej = 0
for it in range(n_it):
    x = {k: np.array(v, copy=True) for k, v in x0.items()}
    log = []
    for _ in range(rng.randint(1, 3)):
        f = rng.choice(list(x))
        arr = x[f]
        for _ in range(rng.randint(1, 4)):
            # (the last entry of name_off / aux_off IS the size of the caller's pool: a claim the library cannot check)
            i = rng.randrange(len(arr) - (1 if f in ("name_off", "aux_off") else 0))
            info = np.iinfo(arr.dtype)
            k = rng.random()
            if k < 0.4: arr[i] = rng.randrange(int(info.min), int(info.max) + 1) if info.max < 2**62 else rng.randrange(0, 2**62)
            elif k < 0.7: arr[i] = int(arr[i]) ^ (1 << rng.randrange(0, 8 * arr.dtype.itemsize - (1 if info.min < 0 else 0)))
            elif k < 0.85: arr[i] = 0
            else: arr[i] = info.max
            log.append((f, i, int(arr[i])))
    print("case", it, log, flush=True)
    try:
        helpers.lift_c(ctx, pb.c, allow_panic=True)
        ctx.assemble_records(x)
        n_ok += 1
    except abi.PtlError:
        n_rej += 1
print("done", n_ok, n_rej)
