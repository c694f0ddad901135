#!/usr/bin/env python
"""End-to-end example on a GPU box: synthetic read->contig alignments -> liftover -> BAM records -> BGZF, everything from the
liftover to the framed bytes on the device; the host only writes the file.  The output is uncompressed BAM, the format
the reference writes to stdout for `samtools sort` (src/read_alignment_scanner.rs:66-71).
usage: tools/lift_to_bam.py [workload=tiny] [n_reads=2000] [out=gpurun_out/lifted.bam]"""
import gzip, os, struct, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import helpers
from portello_b200 import abi, lib, synth
from test_assemble_records import make_extras
from test_bam_container import bam_header

wl = sys.argv[1] if len(sys.argv) > 1 else "tiny"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(R, "gpurun_out", "lifted.bam")
s = synth.make(wl, n_reads=n)
ctx = helpers.gpu_context(s)
ctx.set_names(s.contig_names, s.chrom_names)
lens = [int(s.chrom_len[i]) for i in range(s.n_chrom)]
text = "@HD\tVN:1.6\tSO:unsorted\n" + "".join(f"@SQ\tSN:{nm}\tLN:{l}\n" for nm, l in zip(s.chrom_names, lens)) + "@PG\tID:portello_b200\tPN:portello_b200\n"
header = bam_header(text, s.chrom_names, lens)
os.makedirs(os.path.dirname(out), exist_ok=True)
chunk, n_rec, t0 = 50_000, 0, time.perf_counter()
with open(out, "wb") as f:
    for first in range(0, s.read_records.n_reads, chunk):
        pb = lib.PackedBatch(lib.load(), s.read_records, first, min(chunk, s.read_records.n_reads - first), s.contig_names)
        res = helpers.lift_c(ctx, pb.c)
        x, _, _ = make_extras(s, pb, 1)  # synthetic qnames / aux / qualities for the batch
        o, _ = ctx.assemble_records(x, flags=abi.ASM_NO_DOWNLOAD)
        last = first + chunk >= s.read_records.n_reads
        z, data = ctx.bgzf_store_records(header if first == 0 else b"", flags=abi.BGZF_EOF if last else 0)
        f.write(data)
        n_rec += res.n_records
print(f"{out}: {n_rec} records, {os.path.getsize(out)} bytes in {time.perf_counter() - t0:.2f} s (incl. synthetic extras)")
raw = gzip.open(out, "rb").read()   # python's gzip checks every block's CRC32
assert raw[:4] == b"BAM\x01"
at = 8 + struct.unpack_from("<i", raw, 4)[0]
n_ref = struct.unpack_from("<i", raw, at)[0]; at += 4
for _ in range(n_ref): at += 8 + struct.unpack_from("<i", raw, at)[0]
k = 0
while at < len(raw):
    at += 4 + struct.unpack_from("<I", raw, at)[0]; k += 1
assert k == n_rec and at == len(raw), (k, n_rec)
print(f"parsed back: BAM magic, {n_ref} references, {k} records, all BGZF CRCs ok")
