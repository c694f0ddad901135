"""Stage times of the device-resident step for the library named by PORTELLO_B200_LIB (kernel tuning A/B, under gpurun)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, helpers
from portello_b200 import lib, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "chr20"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
s = synth.make(wl, n_reads=n)
ctx = lib.GpuContext(0, 1)
ctx.set_reference(helpers.reference_arrays(s)); ctx.set_contig_records(s.contig_records)
pb = lib.PackedBatch(lib.load(), s.read_records, 0, s.read_records.n_reads, s.contig_names, pinned=True, windows=ctx.get_contig_segments())  # as bench.py packs
ctx.upload(pb.c, 0)
for _ in range(3): ctx.run(0)
ctx.counters(0)
acc = {}
for _ in range(10):
    ctx.run(0)
    for k, v in ctx.kernel_times(0).items(): acc.setdefault(k, []).append(v)
c = ctx.counters(0)
print(os.path.basename(os.environ.get("PORTELLO_B200_LIB", "default")), wl, {k: round(float(np.mean(v)), 4) for k, v in acc.items()}, "total", round(sum(float(np.mean(v)) for v in acc.values()), 4), "pairs", c["n_pairs"])
