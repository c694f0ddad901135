import os, sys, random, tempfile, gzip, struct
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from portello_b200 import bamio, synth, abi, lib
d = tempfile.mkdtemp(prefix="fr_")
s = synth.make("tiny", seed=13, n_reads=200, read_sa_frac=0.3)
paths = bamio.write_dataset(s, d, n_unmapped=3)
raw = bytearray(gzip.decompress(open(paths["reads"], "rb").read()))
l_text = struct.unpack_from("<i", raw, 4)[0]; at = 8 + l_text
n_ref = struct.unpack_from("<i", raw, at)[0]; at += 4
for _ in range(n_ref):
    ln = struct.unpack_from("<i", raw, at)[0]; at += 8 + ln
hdr_end = at
rec_starts = []
while at < len(raw):
    bs = struct.unpack_from("<i", raw, at)[0]; rec_starts.append((at, bs)); at += 4 + bs
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
L = lib.load()
n_ok = n_err = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 300):
    t = bytearray(raw)
    for _ in range(rng.randint(1, 3)):
        st, bs = rng.choice(rec_starts)
        m = rng.random()
        if m < 0.5: p = st + rng.randrange(0, min(bs, 200))          # fixed fields, name, CIGAR
        else: p = st + 4 + bs - 1 - rng.randrange(0, min(bs, 300))   # aux block (SA tags live there)
        p = max(hdr_end, min(p, len(t) - 5))
        k = rng.random()
        if k < 0.6: t[p] = rng.randrange(256)
        elif k < 0.8: t[p:p+4] = rng.choice([0, 1, 0x7fffffff, 0xffffffff, 70000]).to_bytes(4, "little")
        else: t[p:p+1] = bytes([rng.choice(b",;+-0123456789MIDS=X")])
    w = os.path.join(d, "r.bam")
    open(w, "wb").write(bamio.bgzf_compress(bytes(t)))
    print("case", it, flush=True)
    try:
        f = bamio.BamFile(w)
        dec = f.fetch(bamio.FETCH_ALL, flt=bamio.SKIP_SUPPLEMENTARY)
        mapped = dec.n
        if mapped:
            try:
                lib.PackedBatch(L, dec.recs, 0, mapped, s.contig_names, windows=True)
            except abi.PtlError:
                pass
        n_ok += 1
    except abi.PtlError:
        n_err += 1
print("done", n_ok, n_err)
