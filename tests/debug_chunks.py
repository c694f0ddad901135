import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, helpers
from portello_b200 import abi, lib, synth
s = synth.make("tiny", seed=31, n_reads=6000)
gctx = helpers.gpu_context(s, n_slots=3)
octx = helpers.oracle_context(s)
whole = helpers.lift_c(gctx, helpers.pack(s).c)
owhole = helpers.lift_c(octx, helpers.pack(s).c)
print("whole vs oracle:", whole.diff(owhole))
n = s.read_records.n_reads
cuts = [0, n // 3, 2 * n // 3, n]
for trial in range(3):
    packs = [helpers.pack(s, cuts[i], cuts[i + 1] - cuts[i]) for i in range(3)]
    if trial == 0:
        for i, p in enumerate(packs): gctx._check(gctx.lib._lift_submit(gctx.h, i, C.byref(p.c)))
        parts = []
        for i in range(3):
            r = abi.ResultC(); gctx._check(gctx.lib._lift_wait(gctx.h, i, C.byref(r))); parts.append(abi.Result.from_c(r))
    else:
        parts = [helpers.lift_c(gctx, p.c, slot=(i if trial == 1 else 0)) for i, p in enumerate(packs)]
    for i, p in enumerate(packs):
        o = helpers.lift_c(octx, p.c)
        d = parts[i].diff(o)
        print("trial", trial, "part", i, "vs oracle:", d)
        if d:
            a, b = parts[i], o
            nb = min(len(a.rec_cigar_begin), len(b.rec_cigar_begin))
            k = int(np.flatnonzero(a.rec_cigar_begin[:nb] != b.rec_cigar_begin[:nb])[0]) - 1
            print("  first differing record", k, "gpu:", a.record_cigar(k)[-80:], " oracle:", b.record_cigar(k)[-80:], "status", a.rec_status[k], "pos", a.rec_pos[k], b.rec_pos[k])
            rd = int(np.searchsorted(a.read_rec_begin, k, side="right") - 1)
            print("  read", rd, "+", cuts[i], "flag", s.read_records.flag[cuts[i] + rd], "sa", s.read_records.sa_tag[cuts[i] + rd])
