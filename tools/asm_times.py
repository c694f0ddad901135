"""Kernel times of the record assembly on one chunk (library named by PORTELLO_B200_LIB): python tools/asm_times.py [workload] [reads]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import argparse, numpy as np
import bench_records
from portello_b200 import lib, synth
import helpers
wl = sys.argv[1] if len(sys.argv) > 1 else "chr20"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
s = synth.make(wl, n_reads=n)
L = lib.load()
ctx = helpers.gpu_context(s, n_slots=3)
ch = lib.PackedBatch(L, s.read_records, 0, n, s.contig_names, pinned=True)
args = argparse.Namespace(no_assemble=False, warmup=3, steps=10, workload=wl)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
bench_records.measure_e2e_records = lambda *a, **k: None
r = bench_records.measure(args, ctx, s, L, ch, peak)
a = r["assemble_records"]
print(os.path.basename(os.environ.get("PORTELLO_B200_LIB", "default")), wl, "bam_write ms", round(a["kernel_ms"], 4), "frac", round(a["roofline"]["frac"], 3),
      "bgzf ms", round(a["bgzf_store"]["kernel_ms"], 4), "FUSED ms", round(a["frame_records"]["kernel_ms"], 4), "frac", round(a["frame_records"]["roofline"]["frac"], 3), "bases ms", round(r["assemble_bases"]["kernel_ms"], 4), "flipped", r["assemble_bases"]["flipped_records"])
