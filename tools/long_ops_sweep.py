"""Stage times vs the long-pair threshold (ptl_set_long_pair_ops): which pairs take the warp-cooperative liftover.
usage (under gpurun): python tools/long_ops_sweep.py <workload> <n_reads> <threshold>..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, helpers
from portello_b200 import lib, synth
wl, n = sys.argv[1], int(sys.argv[2])
s = synth.make(wl, n_reads=n)
ctx = lib.GpuContext(0, 1)
ctx.set_reference(helpers.reference_arrays(s)); ctx.set_contig_records(s.contig_records)
pb = lib.PackedBatch(lib.load(), s.read_records, 0, s.read_records.n_reads, s.contig_names, pinned=True)
for thr in [int(x) for x in sys.argv[3:]]:
    ctx.set_long_pair_ops(thr)
    ctx.upload(pb.c, 0)
    for _ in range(3): ctx.run(0)
    acc = {}
    for _ in range(10):
        ctx.run(0)
        for k, v in ctx.kernel_times(0).items(): acc.setdefault(k, []).append(v)
    c = ctx.counters(0)
    print(wl, "long_ops", thr, {k: round(float(np.mean(v)), 4) for k, v in acc.items()}, "total", round(sum(float(np.mean(v)) for v in acc.values()), 4), "pairs", c["n_pairs"], "lifted", c["n_lifted"], flush=True)
