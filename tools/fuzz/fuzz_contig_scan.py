import os, sys, random, tempfile, gzip, struct
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from portello_b200 import bamio, synth, abi, lib
d = tempfile.mkdtemp(prefix="fz_")
s = synth.make("tiny", seed=11, n_reads=50, chrom_len=600_000, contigs_per_chrom=3, junction_per_mb=15)
paths = bamio.write_dataset(s, d)
raw = bytearray(gzip.decompress(open(paths["contigs"], "rb").read()))
# header end
l_text = struct.unpack_from("<i", raw, 4)[0]; at = 8 + l_text
n_ref = struct.unpack_from("<i", raw, at)[0]; at += 4
for _ in range(n_ref):
    ln = struct.unpack_from("<i", raw, at)[0]; at += 8 + ln
hdr_end = at
rec_starts = []
at = hdr_end
while at < len(raw):
    bs = struct.unpack_from("<i", raw, at)[0]; rec_starts.append(at); at += 4 + bs
print("records", len(rec_starts))
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
L = lib.load()
n_ok = n_err = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 200):
    t = bytearray(raw)
    for _ in range(rng.randint(1, 3)):
        p = rng.choice(rec_starts) + rng.randrange(0, 400) if rng.random() < 0.7 else len(t) - 1 - rng.randrange(0, 3000)
        p = min(p, len(t) - 5)
        m = rng.random()
        if m < 0.6: t[p] = rng.randrange(256)
        elif m < 0.8: t[p:p+4] = rng.choice([0, 1, 0x7fffffff, 0xffffffff, 70000]).to_bytes(4, "little")
        else: t[p:p+1] = bytes([rng.choice(b",;+-0123456789MIDS=X")])
    w = os.path.join(d, "c.bam")
    open(w, "wb").write(bamio.bgzf_compress(bytes(t)))
    print("case", it, flush=True)
    try:
        try: os.remove(w + ".bai")
        except OSError: pass
        try:
            bamio.index_bam(w)
        except abi.PtlError:
            pass
        f = bamio.BamFile(w)
        sc = f.scan_contigs(s.contig_names, s.contig_lengths())
        try:
            L.prepare_contig_records(sc.c)
        except abi.PtlError:
            pass
        n_ok += 1
    except abi.PtlError:
        n_err += 1
print("done", n_ok, n_err)
