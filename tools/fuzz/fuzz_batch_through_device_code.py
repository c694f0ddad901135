"""Mutated ptl_batch arrays through the device code compiled for the host (tests/emul): the device-side validation must
reject what would leave a pool (PTL_ERR_INVALID_ARG) or report a lift panic; a wrong CIGAR op or length may change the
result but must never make the kernels read or write outside their buffers (the process would die here).
usage: python tools/fuzz/fuzz_batch_through_device_code.py <seed> <iterations>"""
import os, sys, random
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import emul_lib, helpers
from portello_b200 import abi, synth
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 200
s = synth.make("tiny", seed=23, n_reads=400, read_sa_frac=0.2, rev_contig_frac=0.5, read_cluster_frac=0.2)
pb = helpers.pack(s, windows=True)
b = pb.c
ctx = abi.Context(emul_lib.load(), 0, 1)
ctx.set_reference(helpers.reference_arrays(s))
ctx.set_contig_records(s.contig_records)
if os.environ.get("FUZZ_LONG_OPS"):  # e.g. 0: every pair down the warp-cooperative path (lift_warp.cuh)
    ctx.set_long_pair_ops(int(os.environ["FUZZ_LONG_OPS"]))
fields = {"read_flag": b.n_reads, "read_seq_len": b.n_reads, "read_seq_off": b.n_reads, "read_seg_begin": b.n_reads + 1, "rseg_contig": b.n_read_segments,
          "rseg_pos": b.n_read_segments, "rseg_is_fwd": b.n_read_segments, "rseg_cigar_begin": b.n_read_segments, "rseg_cigar_len": b.n_read_segments,
          "cigar": int(b.n_cigar), "rseg_win_begin": b.n_read_segments + 1, "indel_win": int(b.n_indel_win)}
types = dict(abi.BatchC._fields_)
n_ok = n_rej = n_panic = 0
for it in range(n_it):
    b2 = abi.BatchC.from_buffer_copy(b)
    keep = [pb]
    log = []
    for _ in range(rng.randint(1, 3)):
        f = rng.choice(list(fields))
        arr = np.ctypeslib.as_array(getattr(b, f), (fields[f],)).copy()
        for _ in range(rng.randint(1, 4)):
            i = rng.randrange(len(arr))
            k = rng.random()
            info = np.iinfo(arr.dtype)
            if k < 0.4: arr[i] = rng.randrange(int(info.min), int(info.max) + 1) if info.max < 2**62 else rng.randrange(0, 2**62)
            elif k < 0.7: arr[i] = int(arr[i]) ^ (1 << rng.randrange(0, 8 * arr.dtype.itemsize - (1 if info.min < 0 else 0)))
            elif k < 0.85: arr[i] = 0
            else: arr[i] = info.max
            log.append((f, i, int(arr[i])))
        setattr(b2, f, arr.ctypes.data_as(types[f]))
        keep.append(arr)
    print("case", it, [(f, i, v) for f, i, v in log], flush=True)
    try:
        helpers.lift_c(ctx, b2, allow_panic=True)
        n_ok += 1
    except abi.PtlError as e:
        if "panic" in str(e).lower(): n_panic += 1
        else: n_rej += 1
print("done", n_ok, n_rej, n_panic)
