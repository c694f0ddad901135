"""Where does the end-to-end time go? (diagnostic, run under gpurun)"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, helpers
from portello_b200 import lib, synth
L = lib.load()
alloc = C.cast(L.dll.ptl_host_alloc, C.c_void_p); free = C.cast(L.dll.ptl_host_free, C.c_void_p)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
s = synth.make("chr20", host_alloc=alloc, host_free=free, n_reads=n)
ctx = lib.GpuContext(0, 3)
ctx.set_reference(helpers.reference_arrays(s)); ctx.set_contig_records(s.contig_records)
pb0 = lib.PackedBatch(L, s.read_records, 0, n, s.contig_names, pinned=True)
pbw = lib.PackedBatch(L, s.read_records, 0, n, s.contig_names, pinned=True, windows=ctx.get_contig_segments())
print(f"reads {n}  rsegs {pb0.c.n_read_segments}  cigar ops {pb0.c.n_cigar}  indel windows {pbw.c.n_indel_win}")
def sync(): torch.cuda.synchronize()
for zc, pb in ((False, pb0), (True, pb0), (True, pbw)):
    ctx.set_seq_zero_copy(zc)
    for rep in range(3):
        sync(); t0 = time.perf_counter()
        ctx.upload(pb.c, 0); sync(); t1 = time.perf_counter()
        ctx.run(0); sync(); t2 = time.perf_counter()
        r = ctx.download(0, copy=False); t3 = time.perf_counter()
    kt = ctx.kernel_times(0)
    print(f"zero_copy={zc} windows={pb is pbw}: upload {1e3*(t1-t0):.2f} ms  run {1e3*(t2-t1):.2f} ms  download {1e3*(t3-t2):.2f} ms   stages {kt}")
    sync(); t0 = time.perf_counter()
    for rep in range(5):
        ctx.submit_c(pb.c, 0); ctx.wait_c(0)
    print(f"   submit+wait avg {1e3*(time.perf_counter()-t0)/5:.2f} ms")
