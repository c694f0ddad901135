"""Host enqueue time vs device time of ptl_lift_run (is the device-resident step launch bound?), under gpurun."""
import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch, helpers
from portello_b200 import lib, synth
s = synth.make("chr20", n_reads=1_000_000)
ctx = lib.GpuContext(0, 1)
ctx.set_reference(helpers.reference_arrays(s)); ctx.set_contig_records(s.contig_records)
pb = lib.PackedBatch(lib.load(), s.read_records, 0, s.read_records.n_reads, s.contig_names, pinned=True)
ctx.upload(pb.c, 0)
for _ in range(3): ctx.run(0)
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(ctx.stream(0))
for rep in range(3):
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ev0.record(stream)
    for _ in range(10): ctx.run(0)
    t1 = time.perf_counter(); ev1.record(stream); ev1.synchronize(); t2 = time.perf_counter()
    print("host enqueue per run %.3f ms, device per run %.3f ms, wall per run %.3f ms" % ((t1 - t0) * 100, ev0.elapsed_time(ev1) / 10, (t2 - t0) * 100), ctx.kernel_times(0))
