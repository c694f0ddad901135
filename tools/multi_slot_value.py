"""Device-resident step time with the read set split over S slots (streams) that run concurrently, vs one resident batch.
usage: python tools/multi_slot_value.py [workload] [reads] [slots]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, helpers
from portello_b200 import lib, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "chr20"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 3
s = synth.make(wl, n_reads=n)
L = lib.load()
ctx = lib.GpuContext(0, S + 1)
ctx.set_reference(helpers.reference_arrays(s)); ctx.set_contig_records(s.contig_records)
segs = ctx.get_contig_segments()
whole = lib.PackedBatch(L, s.read_records, 0, n, s.contig_names, pinned=True, windows=segs)
parts = [lib.PackedBatch(L, s.read_records, k * (n // S), (n // S) if k < S - 1 else n - (S - 1) * (n // S), s.contig_names, pinned=True, windows=segs) for k in range(S)]
dev = torch.device("cuda", 0)
streams = [torch.cuda.ExternalStream(ctx.stream(k), device=dev) for k in range(S + 1)]
def timed(fn, steps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for _ in range(steps): fn()
    e1.record(streams[0]); e1.synchronize(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
ctx.upload(whole.c, S)
for _ in range(3): ctx.run(S)
ctx.counters(S)
# single batch on slot S, timed on its own stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(streams[S])
for _ in range(20): ctx.run(S)
e1.record(streams[S]); e1.synchronize()
single = e0.elapsed_time(e1) / 20
for k in range(S): ctx.upload(parts[k].c, k)
for _ in range(3):
    for k in range(S): ctx.run(k)
for k in range(S): ctx.counters(k)
def step():
    # fork: every slot stream waits for stream 0's position, runs its part; join: stream 0 waits for all
    ev = torch.cuda.Event(); ev.record(streams[0])
    for k in range(1, S): streams[k].wait_event(ev)
    for k in range(S): ctx.run(k)
    for k in range(1, S):
        e = torch.cuda.Event(); e.record(streams[k]); streams[0].wait_event(e)
multi = timed(step)
pairs = sum(ctx.counters(k)["n_pairs"] for k in range(S))
print(wl, n, "single batch ms", round(single, 4), f"| {S} concurrent slots ms", round(multi, 4), "pairs", pairs, f"-> {pairs / multi * 1e3:.3e} /s vs {pairs / single * 1e3:.3e}")
